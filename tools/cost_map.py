"""Per-ray cost map of C4 (analysis tool): SM cycles each primary ray spent in the traversal, written by the
one-ray-per-lane kernel into the distance buffer when RTX_EXP_COSTMAP is set.
usage: RTX_EXP_COSTMAP=1 python tools/cost_map.py [out.npz]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opencl_raytracer_b200 import host, scene as scn, scenes  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

v, f = po.read_mesh_bin(po.staged_bunny_path())
v, f = scenes.subdivided(v, f)
sc = scn.scene_from_mesh(v, f, name="bunny_x144")
rt = host.RayTracer(host.Options(width=3840, height=2160, nSuperSamples=1))
with host.CudaHost(rt) as h:
    h.set_tunable(host.TUNE_RECORD_HITS, 1)
    h.set_tunable(host.TUNE_FRUSTUM, 0)
    h.upload_scene(sc)
    h()
    h()
    fid, cost = h.download_hits()
    print("kernel_ms", h.stats()["kernel_ms"])
hit = fid != host.NO_HIT
print("rays %d, hit %.3f" % (cost.size, hit.mean()))
print("cycles per ray: mean %.0f, median %.0f, p90 %.0f, p99 %.0f, p99.9 %.0f, max %.0f" % (
    cost.mean(), np.median(cost), *np.percentile(cost, [90, 99, 99.9]), cost.max()))
print("share of all cycles spent by the most expensive 1%% of rays: %.1f %%" % (100 * np.sort(cost.ravel())[-cost.size // 100:].sum() / cost.sum()))
# coarse map: 27 x 48 blocks of 80 x 80 pixels, mean cycles
blk = cost.reshape(27, 80, 48, 80).mean(axis=(1, 3))
np.set_printoptions(linewidth=250)
print((blk / 1000).astype(int))
# per 8x4 unit (what a warp pulls): max over the unit = the warp's time
unit = cost.reshape(540, 4, 480, 8).max(axis=(1, 3))
print("per-unit (8x4) max cycles: mean %.0f, p99 %.0f, max %.0f" % (unit.mean(), np.percentile(unit, 99), unit.max()))
ys, xs = np.unravel_index(np.argsort(unit.ravel())[-10:], unit.shape)
print("ten most expensive units (tile row, tile col, kcycles):", [(int(y), int(x), int(unit[y, x] / 1000)) for y, x in zip(ys, xs)])
if len(sys.argv) > 1:
    np.savez_compressed(sys.argv[1], cost=cost.astype(np.float32), hit=hit)
