"""C4 diagnostic: box tests / triangle tests per ray of the ordered traversal against the literal walk (counters on)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opencl_raytracer_b200 import host, scene as scn, scenes  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

v, f = po.read_mesh_bin(po.staged_bunny_path())
v, f = scenes.subdivided(v, f)
sc = scn.scene_from_mesh(v, f, name="bunny_x144")
w, h = (3840, 2160) if len(sys.argv) < 3 else (int(sys.argv[1]), int(sys.argv[2]))
rt = host.RayTracer(host.Options(width=w, height=h, nSuperSamples=1))
n = rt.totalWidth * rt.totalHeight
for name, tun in (("ordered", {}), ("ordered leaf 2", {host.TUNE_LEAF_SIZE: 2}), ("literal walk", {host.TUNE_KERNEL: host.KERNEL_EXHAUSTIVE})):
    with host.CudaHost(rt) as hh:
        for k, val in tun.items():
            hh.set_tunable(k, val)
        hh.set_tunable(host.TUNE_COUNTERS, 1)
        hh.upload_scene(sc)
        hh()
        st = hh.stats()
        print("%-16s %.3f ms  box tests/ray %.1f  triangle tests/ray %.2f  leafbox %.2f  depth %d" % (
            name, st["kernel_ms"], st["node_visits"] / n, st["tri_tests"] / n, st["leafbox_tests"] / n, st["tree_depth"]), flush=True)
