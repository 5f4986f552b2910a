"""Kernel-time sweep over RTX_TUNE settings on C4 (10.16 M triangles, 4K, s=1) and C5 (2^26 of the 2^28
random rays vs the stand-in tree).  Development tool.
usage: python tools/sweep_c4c5.py c4|c5 "k=v,k=v" "default" ..."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opencl_raytracer_b200 import host, scene as scn, scenes  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

which = sys.argv[1]
if which == "c4":
    v, f = po.read_mesh_bin(po.staged_bunny_path())
    v, f = scenes.subdivided(v, f)
    t = time.time()
    sc = scn.scene_from_mesh(v, f, name="bunny_x144")
    print("host bvh %.1f s, %d triangles" % (time.time() - t, sc.num_triangles), flush=True)
    rt = host.RayTracer(host.Options(width=3840, height=2160, nSuperSamples=1))
else:
    v, f = scenes.sibenik_standin()
    sc = scn.scene_from_mesh(v, f, name="sibenik_standin")
    rt = host.RayTracer(host.Options(width=32, height=32, nSuperSamples=1))
ref_sum = None
for setting in sys.argv[2:]:
    os.environ["RTX_TUNE"] = setting if setting != "default" else ""
    try:
        with host.CudaHost(rt) as h:
            h.upload_scene(sc)
            best = 1e9
            if which == "c4":
                for _ in range(5):
                    h()
                    best = min(best, h.stats()["kernel_ms"])
                n = rt.totalWidth * rt.totalHeight
                chk = float(h.download().astype("float64").sum())
            else:
                n = 1 << 26
                h.trace_random_rays(1234, 0, 1 << 22)
                for _ in range(3):
                    hits, idsum, _, _ = h.trace_random_rays(1234, 0, n)
                    best = min(best, h.stats()["kernel_ms"])
                chk = (hits, idsum)
            if ref_sum is None:
                ref_sum = chk
            print("%-60s %8.3f ms  %7.0f Mrays/s  %s" % (setting, best, n / best / 1e3, "same" if chk == ref_sum else "DIFFERENT RESULT"), flush=True)
            st = h.stats()
            if st.get("node_visits"):
                print("    per ray: %.1f box tests, %.2f triangle tests, %.2f leaf-box tests" % (st["node_visits"] / n, st["tri_tests"] / n, st["leafbox_tests"] / n), flush=True)
    except Exception as e:          # a setting the library refuses
        print("%-60s %s" % (setting, e), flush=True)
