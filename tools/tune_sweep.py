"""Kernel-time sweep over RTX_TUNE settings on one workload (development tool).
usage: python tools/tune_sweep.py c2|c3 "k=v,k=v" "k=v" ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opencl_raytracer_b200 import host, scene as scn, scenes
v, f = scenes.sibenik_standin(); sc = scn.scene_from_mesh(v, f)
w, h, ss = (1920, 1080, 4) if sys.argv[1] == "c2" else (3840, 2160, 16)
rt = host.RayTracer(host.Options(width=w, height=h, nSuperSamples=ss))
for setting in sys.argv[2:]:
    os.environ["RTX_TUNE"] = setting if setting != "default" else ""
    with host.CudaHost(rt) as hst:
        hst.upload_scene(sc)
        best = 1e9
        for _ in range(6):
            hst(); best = min(best, hst.stats()["kernel_ms"])
        print("%-50s %.3f ms  %.0f Mrays/s" % (setting, best, rt.totalWidth * rt.totalHeight / best / 1e3))
