"""Measure the BASELINE.json configurations that bench.py does not cover (development / reporting tool).

usage: python tools/run_configs.py [c1] [c2] [c4] [c5]   -> one JSON line per config on stdout
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opencl_raytracer_b200 import host, scene as scn, scenes  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def time_render(h, n=8):
    best = 1e9
    for _ in range(n):
        h()
        best = min(best, h.stats()["kernel_ms"])
    return best


def check_rows(h, sc, rt, rows):
    h.set_tunable(host.TUNE_RECORD_HITS, 1)
    h()
    img = h.download()
    fid, dist = h.download_hits()
    h.set_tunable(host.TUNE_RECORD_HITS, 0)
    ref = po.render(sc, rt.totalWidth, rt.totalHeight, 1.0, True, rows=rows, want_counters=True)
    sel = slice(*rows)
    n = fid[sel].size
    return {"rays_checked": int(n), "id_mismatches": int((fid[sel] != ref.face_id[sel]).sum()),
            "distance_mismatches": int((dist[sel] != ref.distance[sel]).sum()),
            "pixel_mismatches": int((img[sel] != ref.image[sel]).sum()),
            "V": ref.counters["V"], "T": ref.counters["T"], "h": ref.counters["h"],
            "B_bytes_per_ray": po.algorithmic_bytes_per_ray(ref.counters["V"], ref.counters["T"], ref.counters["h"])}


def primary(name, sc, w, hgt, ss, rows_step):
    rt = host.RayTracer(host.Options(width=w, height=hgt, nSuperSamples=ss))
    out = {"config": name, "scene": sc.name, "triangles": sc.num_triangles, "rays": rt.totalWidth * rt.totalHeight}
    with host.CudaHost(rt) as h:
        t = time.time(); h.upload_scene(sc); out["upload_ms"] = (time.time() - t) * 1e3
        ms = time_render(h)
        st = h.stats()
        out.update(kernel_ms=ms, mrays_s=out["rays"] / ms / 1e3, launches=st["kernel_launches"], tree_depth=st["tree_depth"])
        out["parity"] = check_rows(h, sc, rt, (rows_step // 2, rt.totalHeight, rows_step))
        for setting, key in (("frustum=0", "per_ray_traversal"), ("frustum=1", "frustum_forced"), ("frustum=0,rays_per_thread=1", "one_ray_per_lane"),
                             ("frustum=0,rays_per_thread=0", "refill_kernel")):
            for kv in setting.split(","):
                k, v = kv.split("=")
                h.set_tunable(getattr(host, "TUNE_" + k.upper()), int(v))
            ms2 = time_render(h, 4)
            out[key] = {"kernel_ms": ms2, "mrays_s": out["rays"] / ms2 / 1e3}
    print(json.dumps(out), flush=True)


def main():
    which = sys.argv[1:] or ["c1", "c2", "c4", "c5"]
    if "c1" in which:
        v, f = po.read_mesh_bin(po.staged_bunny_path())
        primary("C1 bunny 600x600 s=4", scn.scene_from_mesh(v, f, name="bunny"), 600, 600, 4, 1)
    sib = None
    if "c2" in which or "c5" in which:
        v, f = scenes.sibenik_standin()
        sib = scn.scene_from_mesh(v, f, name="sibenik_standin")
    if "c2" in which:
        primary("C2 sibenik-standin 1920x1080 s=4", sib, 1920, 1080, 4, 8)
        primary("C2b sibenik-standin 1920x1080 s=1", sib, 1920, 1080, 1, 4)
    if "c4" in which:
        v, f = po.read_mesh_bin(po.staged_bunny_path())
        t = time.time(); v2, f2 = scenes.subdivided(v, f); t_sub = time.time() - t
        t = time.time(); big = scn.scene_from_mesh(v2, f2, name="bunny_x144"); t_bvh = time.time() - t
        sys.stderr.write("c4: subdivide %.1fs, bvh %.1fs, %d tris, %d nodes\n" % (t_sub, t_bvh, big.num_triangles, big.num_nodes))
        primary("C4 bunny x144 3840x2160 s=1", big, 3840, 2160, 1, 270)
    if "c5" in which:
        rt = host.RayTracer(host.Options(width=32, height=32, nSuperSamples=1))
        lo, hi = sib.root_box()
        with host.CudaHost(rt) as h:
            h.upload_scene(sib)
            n_chk = 1 << 20
            o, d = po.gen_random_rays(1234, 0, n_chk, lo, hi)
            ref = po.trace_rays(sib, o, d, 100000.0, want_counters=True)
            hits, idsum, fid, dist = h.trace_random_rays(1234, 0, n_chk, want_arrays=True)
            par = {"rays_checked": n_chk, "id_mismatches": int((fid != ref.face_id).sum()), "distance_mismatches": int((dist != ref.distance).sum()),
                   "V": ref.counters["V"], "T": ref.counters["T"], "h": ref.counters["h"],
                   "B_bytes_per_ray": po.algorithmic_bytes_per_ray(ref.counters["V"], ref.counters["T"], ref.counters["h"], 8.0)}
            n = 1 << 28
            h.trace_random_rays(1234, 0, 1 << 24)
            best = 1e9
            for _ in range(3):
                hits, idsum, _, _ = h.trace_random_rays(1234, 0, n)
                best = min(best, h.stats()["kernel_ms"])
            h.set_tunable(host.TUNE_INCOHERENT_KERNEL, 0)
            old = 1e9
            for _ in range(2):
                hits0, idsum0, _, _ = h.trace_random_rays(1234, 0, n)
                old = min(old, h.stats()["kernel_ms"])
            print(json.dumps({"config": "C5 2^28 random rays vs sibenik-standin", "rays": n, "kernel_ms": best, "mrays_s": n / best / 1e3,
                              "hits": hits, "face_id_sum": idsum, "parity": par,
                              "plain_while_while": {"kernel_ms": old, "mrays_s": n / old / 1e3, "same_checksums": (hits0, idsum0) == (hits, idsum)}}), flush=True)


if __name__ == "__main__":
    main()
