"""Wall-clock breakdown of the reference-facing calls (development tool)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opencl_raytracer_b200 import host, scene as scn, scenes

def t(fn, n=5):
    fn(); ts = []
    for _ in range(n):
        a = time.perf_counter(); fn(); ts.append((time.perf_counter() - a) * 1e3)
    return min(ts), float(np.median(ts))

v, f = scenes.sibenik_standin()
sc = scn.scene_from_mesh(v, f)
for w, h, ss in ((1920, 1080, 4), (3840, 2160, 16)):
    rt = host.RayTracer(host.Options(width=w, height=h, nSuperSamples=ss))
    with host.CudaHost(rt) as hst:
        out = np.empty((rt.totalHeight, rt.totalWidth), np.float32)
        print(w, h, ss, "upload", t(lambda: hst.upload_scene(sc)), "render", t(lambda: hst()),
              "download", t(lambda: hst.download(out)), "download_u8", t(lambda: hst.download_u8()))
