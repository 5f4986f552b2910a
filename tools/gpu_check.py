"""Quick parity + timing check on a GPU box (development tool, not a test).

usage: python tools/gpu_check.py [scene ...]   scenes: bunny sibenik soup
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opencl_raytracer_b200 import host, scene as scn, scenes  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def get_scene(name):
    if name == "bunny":
        v, f = po.read_mesh_bin(po.staged_bunny_path())
        return scn.scene_from_mesh(v, f, name="bunny")
    if name == "sibenik":
        v, f = scenes.sibenik_standin()
        return scn.scene_from_mesh(v, f, name="sibenik_standin")
    v, f = scenes.random_soup(2000, seed=7)
    return scn.scene_from_mesh(v, f, name="soup")


def compare(tag, fid, dist, img, ref):
    ids_bad = int((fid != ref.face_id).sum())
    both = (fid != host.NO_HIT) & (ref.face_id != host.NO_HIT)
    rel = np.abs(dist[both] - ref.distance[both]) / ref.distance[both]
    img_bad = int((img != ref.image).sum())
    print("  %-28s id mismatches %d / %d (%.5f%%)  max rel dist %.2e  image value mismatches %d  maxabs %.3g" % (
        tag, ids_bad, fid.size, 100.0 * ids_bad / fid.size, rel.max() if rel.size else 0.0, img_bad,
        float(np.nanmax(np.abs(img - ref.image)))))
    return ids_bad


def main():
    names = sys.argv[1:] or ["soup", "bunny", "sibenik"]
    host.CudaHost.printInfo()
    for name in names:
        sc = get_scene(name)
        w, h, ss = (600, 600, 4) if name != "sibenik" else (1920, 1080, 4)
        rt = host.RayTracer(host.Options(width=w, height=h, nSuperSamples=ss))
        print("== %s: %d tris, %d nodes, %dx%d rays" % (sc.name, sc.num_triangles, sc.num_nodes, rt.totalWidth, rt.totalHeight))
        t = time.time()
        ref = po.render(sc, rt.totalWidth, rt.totalHeight, 1.0, True, want_counters=True)
        print("  oracle: %.2fs  V=%.2f T=%.3f h=%.4f" % (time.time() - t, ref.counters["V"], ref.counters["T"], ref.counters["h"]))
        for kernel, leaf, top, rpt, fr in ((host.KERNEL_EXHAUSTIVE, 1, 0, 1, 0), (host.KERNEL_PERSISTENT, 1, 0, 1, 0), (host.KERNEL_PERSISTENT, 1, 0, 4, 0),
                                           (host.KERNEL_PERSISTENT, 1, 0, 4, 1), (host.KERNEL_PERSISTENT, 4, 0, 4, 1)):
            with host.CudaHost(rt) as hst:
                hst.set_tunable(host.TUNE_KERNEL, kernel)
                hst.set_tunable(host.TUNE_LEAF_SIZE, leaf)
                hst.set_tunable(host.TUNE_TOP_SMEM, top)
                hst.set_tunable(host.TUNE_RAYS_PER_THREAD, rpt)
                hst.set_tunable(host.TUNE_FRUSTUM, fr)
                hst.set_tunable(host.TUNE_RECORD_HITS, 1)
                hst.set_tunable(host.TUNE_COUNTERS, 1)
                t = time.time(); hst.upload_scene(sc); t_up = time.time() - t
                hst()
                st = hst.stats()
                img = hst.download()
                fid, dist = hst.download_hits()
                tag = "%s leaf=%d rpt=%d frustum=%d" % ("exhaustive" if kernel else "persistent", leaf, rpt, fr)
                compare(tag, fid, dist, img, ref)
                rays = st["rays"]
                print("    counters: visits/ray %.2f tri/ray %.3f leafbox/ray %.3f exact rays %d overflow packets %d depth %d pairs %d upload %.1f ms" % (
                    st["node_visits"] / rays, st["tri_tests"] / rays, st["leafbox_tests"] / rays, st["exact_path_rays"], st["packet_overflows"],
                    st["tree_depth"], st["num_pairs"], t_up * 1e3))
                hst.set_tunable(host.TUNE_RECORD_HITS, 0)
                hst.set_tunable(host.TUNE_COUNTERS, 0)
                best = 1e9
                for _ in range(5):
                    hst()
                    best = min(best, hst.stats()["kernel_ms"])
                print("    kernel %.3f ms -> %.1f Mrays/s" % (best, rays / best / 1e3))
                u8 = hst.download_u8()
                ref_u8 = po.resize(ref.image, w, h, rt.n)
                print("    u8 mismatches %d (max diff %d)" % (int((u8 != ref_u8).sum()), int(np.abs(u8.astype(int) - ref_u8.astype(int)).max())))


if __name__ == "__main__":
    main()
