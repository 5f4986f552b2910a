"""Find rays on which the two arbitrary-ray kernels disagree (development tool), and ask the oracle who is right.

usage: python tools/c5_diff.py [log2_total=28] [log2_chunk=24]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opencl_raytracer_b200 import host, scene as scn, scenes  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def main():
    lt = int(sys.argv[1]) if len(sys.argv) > 1 else 28
    lc = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    v, f = scenes.sibenik_standin()
    sib = scn.scene_from_mesh(v, f, name="sibenik_standin")
    lo, hi = sib.root_box()
    rt = host.RayTracer(host.Options(width=32, height=32, nSuperSamples=1))
    n = 1 << lc
    with host.CudaHost(rt) as h:
        h.upload_scene(sib)
        for c in range(1 << (lt - lc)):
            first = c * n
            h.set_tunable(host.TUNE_INCOHERENT_KERNEL, 1)
            a = h.trace_random_rays(1234, first, n)[:2]
            h.set_tunable(host.TUNE_INCOHERENT_KERNEL, 0)
            b = h.trace_random_rays(1234, first, n)[:2]
            if a == b:
                continue
            print("chunk %d differs: refill %s plain %s" % (c, a, b), flush=True)
            h.set_tunable(host.TUNE_INCOHERENT_KERNEL, 1)
            _, _, fa, da = h.trace_random_rays(1234, first, n, want_arrays=True)
            h.set_tunable(host.TUNE_INCOHERENT_KERNEL, 0)
            _, _, fb, db = h.trace_random_rays(1234, first, n, want_arrays=True)
            idx = np.nonzero((fa != fb) | (da != db))[0]
            for i in idx[:8]:
                o, d = po.gen_random_rays(1234, first + int(i), 1, lo, hi)
                ref = po.trace_rays(sib, o, d, 100000.0)
                print("  ray %d: refill (%d, %r) plain (%d, %r) oracle (%d, %r)  o=%r d=%r" % (
                    first + i, fa[i], float(da[i]), fb[i], float(db[i]), ref.face_id[0], float(ref.distance[0]), o[0].tolist(), d[0].tolist()), flush=True)
    print("done")


if __name__ == "__main__":
    main()
