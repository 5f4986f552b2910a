"""Ambient-occlusion timing on a GPU box (development tool): primary pass alone vs primary + occlusion pass.

usage: python tools/ao_timing.py [bunny|sibenik] [width height ss] [method samples]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opencl_raytracer_b200 import host, scene as scn, scenes  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def main():
    a = sys.argv[1:]
    name = a[0] if a else "sibenik"
    w, h, ss = (int(x) for x in a[1:4]) if len(a) >= 4 else ((1920, 1080, 4) if name == "sibenik" else (600, 600, 4))
    method, samples = (int(x) for x in a[4:6]) if len(a) >= 6 else (0, 3)
    if name == "bunny":
        v, f = po.read_mesh_bin(po.staged_bunny_path())
    else:
        v, f = scenes.sibenik_standin()
    sc = scn.scene_from_mesh(v, f, name=name)
    for ao_on in (False, True):
        rt = host.RayTracer(host.Options(width=w, height=h, nSuperSamples=ss, enableAO=ao_on, aoNumSamples=samples, aoMethod=method))
        with host.CudaHost(rt) as hst:
            hst.upload_scene(sc)
            best = 1e9
            for _ in range(5):
                hst()
                best = min(best, hst.stats()["kernel_ms"])
            st = hst.stats()
            img = hst.download()
            hit = float((img > 0).mean())
            print("%s %dx%d ao=%d method=%d samples=%d: %.3f ms, %d launches, %.1f M primary rays/s, nonzero pixels %.3f" % (
                name, rt.totalWidth, rt.totalHeight, ao_on, method, samples, best, st["kernel_launches"], st["rays"] / best / 1e3, hit))


if __name__ == "__main__":
    main()
