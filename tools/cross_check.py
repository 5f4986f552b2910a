"""Full-size cross-check of the fast kernels against the literal walk compiled for the GPU (development tool).
k_render_exhaustive is the reference's own algorithm (no re-ordering, no culling); every fast path must give
bit-identical hit ids, distances and pixels on every ray of the frame.

usage: python tools/cross_check.py [c1big] [c2] [c3] [c3j] [c4]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opencl_raytracer_b200 import host, scene as scn, scenes  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def run(sc, w, h, ss, settings, jitter=0, focal=1.0):
    rt = host.RayTracer(host.Options(width=w, height=h, nSuperSamples=ss, focalLength=focal))
    base = None
    for name, tun in settings:
        with host.CudaHost(rt, jitter_seed=jitter) as hst:
            for k, v in tun.items():
                hst.set_tunable(k, v)
            hst.set_tunable(host.TUNE_RECORD_HITS, 1)
            hst.upload_scene(sc)
            hst()
            img = hst.download()
            fid, dist = hst.download_hits()
            ms = hst.stats()["kernel_ms"]
        if base is None:
            base = (img, fid, dist)
            print("  %-28s %.2f ms (baseline: literal walk)" % (name, ms), flush=True)
        else:
            print("  %-28s %.2f ms  id mismatches %d, distance %d, pixel %d of %d" % (
                name, ms, int((fid != base[1]).sum()), int((dist != base[2]).sum()), int((img != base[0]).sum()), fid.size), flush=True)
            bad = np.argwhere(fid != base[1])[:5]
            for y, x in bad:
                print("    (%d,%d): %d/%r vs literal %d/%r" % (x, y, fid[y, x], float(dist[y, x]), base[1][y, x], float(base[2][y, x])))


def main():
    which = sys.argv[1:] or ["c1big", "c2", "c3", "c3j", "c4"]
    EX = ("exhaustive", {host.TUNE_KERNEL: host.KERNEL_EXHAUSTIVE})
    fast = [("default", {}), ("frustum off, 1 ray/lane", {host.TUNE_FRUSTUM: 0, host.TUNE_RAYS_PER_THREAD: 1}),
            ("frustum off, 4 rays/lane", {host.TUNE_FRUSTUM: 0, host.TUNE_RAYS_PER_THREAD: 4}),
            ("frustum off, refill", {host.TUNE_FRUSTUM: 0, host.TUNE_RAYS_PER_THREAD: 0}),
            ("frustum forced", {host.TUNE_FRUSTUM: 1}), ("leaf size 4", {host.TUNE_LEAF_SIZE: 4})]
    sib = None
    if any(c in which for c in ("c2", "c3", "c3j")):
        v, f = scenes.sibenik_standin()
        sib = scn.scene_from_mesh(v, f, name="sibenik_standin")
    if "c1big" in which:
        v, f = po.read_mesh_bin(po.staged_bunny_path())
        bunny = scn.scene_from_mesh(v, f, name="bunny")
        print("bunny 4000x4000 (16 M rays), focal 1 and 2.5")
        run(bunny, 2000, 2000, 4, [EX] + fast)
        run(bunny, 2000, 2000, 4, [EX] + fast[:2], focal=2.5)
    if "c2" in which:
        print("C2 3840x2160")
        run(sib, 1920, 1080, 4, [EX] + fast)
    if "c3" in which:
        print("C3 15360x8640")
        run(sib, 3840, 2160, 16, [EX] + fast[:2])
    if "c3j" in which:
        print("C3 jittered")
        run(sib, 3840, 2160, 16, [EX] + fast[:1], jitter=0x5EED)
    if "c4" in which:
        v, f = po.read_mesh_bin(po.staged_bunny_path())
        v2, f2 = scenes.subdivided(v, f)
        big = scn.scene_from_mesh(v2, f2, name="bunny_x144")
        print("C4 bunny x144 3840x2160")
        run(big, 3840, 2160, 1, [EX] + fast[:4])


if __name__ == "__main__":
    main()
