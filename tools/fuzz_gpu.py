"""Differential fuzzing on a GPU box (development tool): every fast strategy against the literal walk compiled for
the GPU (k_render_exhaustive = the reference's algorithm, no re-ordering, no culling), over random meshes that
stress the parity argument: needle and sliver triangles, shared edges (ties), coplanar duplicates, scenes scaled
by 1e-3 .. 1e3, grids whose edges line up with pixel columns, odd image sizes, extreme focal lengths; plus
arbitrary rays (both kernels vs the literal walk), ambient occlusion against the CPU oracle on small frames, and the
device-built BVH (rtx_upload_mesh) against the host builder's arrays.

usage: python tools/fuzz_gpu.py [first_seed=0] [count=40]
       RTX_B200_LIB=$PWD/opencl_raytracer_b200/lib/librtx_b200_dbg.so python tools/fuzz_gpu.py ...   (after `make -C
       opencl_raytracer_b200/csrc debug`): the same run against the bounds-checked build -- every device-side index into the
       scene arrays, stacks, queues and candidate lists is range-checked (the stand-in for compute-sanitizer, which is
       closed on this pool); a violation is reported with its source line.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opencl_raytracer_b200 import host, scene as scn, scenes  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def make_mesh(rng):
    kind = rng.integers(0, 8)
    n = int(rng.choice([30, 200, 1000, 4000]))
    if os.environ.get("FUZZ_BIG"):                         # deeper trees: FUZZ_BIG=30000 python tools/fuzz_gpu.py ...
        n = int(os.environ["FUZZ_BIG"])
    if kind == 0:        # plain soup
        v, f = scenes.random_soup(n, seed=int(rng.integers(1 << 30)), size=float(rng.choice([0.05, 0.35, 1.0])))
        return v.astype(np.float64), f, "soup"
    if kind == 1:        # needles and slivers: the third vertex 1e-5 .. 1e-2 edge lengths off the first edge
        v, f = scenes.needle_soup(n, seed=int(rng.integers(1 << 30)))
        return v.astype(np.float64), f, "needles"
    if kind == 2:        # a regular grid wall facing the camera: shared edges everywhere, edges on pixel columns
        m = int(rng.choice([4, 16, 64]))
        z = float(rng.uniform(-3, 0.5))
        half = float(rng.choice([1.0, 1.5, 3.0]))
        g = np.linspace(-half, half, m + 1)
        xx, yy = np.meshgrid(g, g, indexing="ij")
        tilt = float(rng.choice([0.0, 0.0, 0.3]))
        v = np.stack([xx, yy, z + tilt * xx], -1).reshape(-1, 3)
        i, j = np.meshgrid(np.arange(m), np.arange(m), indexing="ij")
        a, b, c, d = i * (m + 1) + j, (i + 1) * (m + 1) + j, (i + 1) * (m + 1) + j + 1, i * (m + 1) + j + 1
        f = np.concatenate([np.stack([a, b, c], -1).reshape(-1, 3), np.stack([a, c, d], -1).reshape(-1, 3)])
        return v, f, "grid%d" % m
    if kind == 3:        # coplanar duplicates and near-duplicates (ties, deep chains)
        v, f = scenes.random_soup(max(10, n // 10), seed=int(rng.integers(1 << 30)))
        v = v.astype(np.float64)
        reps = int(rng.choice([2, 5]))
        vs, fs = [v], [f]
        for r in range(1, reps):
            off = 0.0 if rng.random() < 0.5 else 1e-7 * r
            vs.append(v + off)
            fs.append(f + r * v.shape[0])
        return np.concatenate(vs), np.concatenate(fs), "duplicates"
    if kind == 4:        # spheres: fine closed surfaces, silhouettes
        parts = [scenes.icosphere(rng.uniform(-1, 1, 3) * [1, 1, 0.5], float(rng.uniform(0.1, 0.8)), int(rng.integers(1, 5))) for _ in range(int(rng.integers(1, 5)))]
        v, f = scenes._merge(parts)
        return np.asarray(v, np.float64), np.asarray(f), "spheres"
    if kind == 6:        # degenerate triangles: zero area (repeated vertices), collinear vertices, 1e-7-sized specks
        v, f = scenes.degenerate_soup(n, seed=int(rng.integers(1 << 30)))
        return v.astype(np.float64), f, "degenerate"
    if kind == 7:        # far from the origin: large coordinates, small triangles (rounding of P dominates)
        v, f = scenes.random_soup(n, seed=int(rng.integers(1 << 30)), size=0.02, extent=0.6)
        v = v.astype(np.float64)
        v[:, 2] = v[:, 2] * 0.2 - float(rng.choice([50.0, 400.0]))        # a thin slab 50 or 400 units down the -z axis
        v[:, :2] *= float(rng.choice([20.0, 150.0]))
        return v, f, "far"
    v, f = scenes.sibenik_standin(detail=float(rng.choice([0.2, 0.35])))
    return v.astype(np.float64), f, "standin"


def render_all(sc, rt, settings, jitter=0):
    out = []
    for name, tun in settings:
        with host.CudaHost(rt, jitter_seed=jitter) as h:
            for k, v in tun.items():
                h.set_tunable(k, v)
            h.set_tunable(host.TUNE_RECORD_HITS, 1)
            h.upload_scene(sc)
            h()
            out.append((name, h.download(), h.download_hits(), h.stats()["tree_depth"]))
    return out


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    EX = ("literal", {host.TUNE_KERNEL: host.KERNEL_EXHAUSTIVE})
    fast = [("default", {}), ("1 ray/lane", {host.TUNE_FRUSTUM: 0, host.TUNE_RAYS_PER_THREAD: 1}),
            ("4 rays/lane", {host.TUNE_FRUSTUM: 0, host.TUNE_RAYS_PER_THREAD: 4}),
            ("refill", {host.TUNE_FRUSTUM: 0, host.TUNE_RAYS_PER_THREAD: 0}),
            ("frustum forced", {host.TUNE_FRUSTUM: 1}), ("frustum, 4 rays", {host.TUNE_FRUSTUM: 1, host.TUNE_LIST_RAYS_PER_THREAD: 4}),
            ("leaf 4", {host.TUNE_LEAF_SIZE: 4}), ("leaf 8 frustum", {host.TUNE_LEAF_SIZE: 8, host.TUNE_FRUSTUM: 1})]
    total_bad = 0
    for seed in range(first, first + count):
        rng = np.random.default_rng(1000 + seed)
        v, f, kind = make_mesh(rng)
        scale = float(10.0 ** rng.choice([0, 0, 0, -3, -1, 1, 3]))
        # the camera is fixed at (0,0,2): scale the scene about the point the camera looks at, and zoom to keep it in view
        if scale != 1.0:
            v = (v - [0, 0, 2]) * scale + [0, 0, 2]
        sc = scn.scene_from_mesh(v.astype(np.float32), f, name=kind)
        w, hgt = [(64, 48), (97, 61), (200, 120), (33, 200), (256, 256)][int(rng.integers(0, 5))]
        ss = int(rng.choice([1, 4, 9, 16]))
        if seed % 8 == 5:                      # a frame with >= 10 000 tiles: the two-level (super-tile) frustum pass
            w, hgt, ss = [(1920, 1080, 4), (2500, 1300, 4), (1000, 900, 16)][int(rng.integers(0, 3))]
        focal = float(rng.choice([0.3, 1.0, 1.0, 2.5]))
        jitter = int(rng.choice([0, 0, 0x5EED]))
        rt = host.RayTracer(host.Options(width=w, height=hgt, nSuperSamples=ss, focalLength=focal))
        res = render_all(sc, rt, [EX] + fast, jitter)
        base = res[0]
        bad = 0
        # the literal walk on the GPU against the CPU oracle (a row sample when the frame is large)
        step = max(1, (rt.totalWidth * rt.totalHeight) // 150000)
        rows = (int(rng.integers(0, step)), rt.totalHeight, step)
        ref = po.render(sc, rt.totalWidth, rt.totalHeight, po.focal_roundtrip(focal), True, jitter_seed=jitter, rows=rows)
        sel = slice(*rows)
        nb = int((base[2][0][sel] != ref.face_id[sel]).sum()) + int((base[2][1][sel] != ref.distance[sel]).sum()) + \
            int(((base[1][sel] != ref.image[sel]) & ~(np.isnan(base[1][sel]) & np.isnan(ref.image[sel]))).sum())
        if nb:
            bad += nb
            print("  MISMATCH seed %d literal GPU walk vs CPU oracle: %d differences" % (seed, nb), flush=True)
        for name, img, (fid, dist), _ in res[1:]:
            nb = int((fid != base[2][0]).sum()) + int((dist != base[2][1]).sum()) + int(((img != base[1]) & ~(np.isnan(img) & np.isnan(base[1]))).sum())
            if nb:
                bad += nb
                ys, xs = np.nonzero(fid != base[2][0])
                ex = [(int(x), int(y), int(fid[y, x]), int(base[2][0][y, x]), float(dist[y, x]), float(base[2][1][y, x])) for y, x in list(zip(ys, xs))[:3]]
                print("  MISMATCH seed %d %-16s: %d differences, e.g. (x, y, id, literal id, dist, literal dist) %s" % (seed, name, nb, ex), flush=True)
        # arbitrary rays: both kernels vs the literal walk, two max distances
        lo, hi = sc.root_box()
        nr = 1 << 16
        with host.CudaHost(rt) as h:
            h.upload_scene(sc)
            sums = {}
            for md in (100000.0, float(0.3 * np.abs(hi - lo).max())):
                for name, tun in (("refill", {host.TUNE_INCOHERENT_KERNEL: 1}), ("plain", {host.TUNE_INCOHERENT_KERNEL: 0}),
                                  ("literal", {host.TUNE_KERNEL: host.KERNEL_EXHAUSTIVE})):
                    for k, val in tun.items():
                        h.set_tunable(k, val)
                    sums[name] = h.trace_random_rays(77 + seed, 0, nr, md, want_arrays=True)
                    h.set_tunable(host.TUNE_KERNEL, host.KERNEL_PERSISTENT)
                for name in ("refill", "plain"):
                    nb = int((sums[name][2] != sums["literal"][2]).sum()) + int((sums[name][3] != sums["literal"][3]).sum())
                    if nb:
                        bad += nb
                        print("  MISMATCH seed %d rays %-8s max_distance %g: %d differences" % (seed, name, md, nb), flush=True)
        # ambient occlusion against the CPU oracle (small frame)
        if True:
            ao = po.Ao.make(method=seed % 2, samples=int(rng.integers(1, 4)), max_distance=float(0.1 * np.abs(hi - lo).max() * rng.choice([0.1, 0.3, 1.0, 5.0])))
            rta = host.RayTracer(host.Options(width=48, height=32, nSuperSamples=4, focalLength=focal, enableAO=True, aoNumSamples=ao.samples,
                                              aoMethod=ao.method, aoMaxDistance=float(ao.max_distance)))
            with host.CudaHost(rta) as h:
                h.upload_scene(sc)
                h()
                got = h.download()
            ref = po.render(sc, rta.totalWidth, rta.totalHeight, po.focal_roundtrip(focal), True, ao=po.Ao.make(
                method=ao.method, samples=ao.samples, max_distance=po.focal_roundtrip(float(ao.max_distance)))).image
            nb = int(((got != ref) & ~(np.isnan(got) & np.isnan(ref))).sum())
            if nb:
                bad += nb
                print("  MISMATCH seed %d ambient occlusion method %d samples %d: %d pixels" % (seed, ao.method, ao.samples, nb), flush=True)
        # the BVH built on the device from the raw mesh == the host builder's arrays, bit for bit
        with host.CudaHost(rt, jitter_seed=jitter) as h:
            h.upload_mesh(sc.vertices, sc.orig_faces, None if seed % 2 else sc.normals)
            nodes_d, aabbs_d, tri_d, faces_d = h.download_tree()
            if seed % 2:
                vn = h.download_normals()
                nbn = int((vn.view(np.uint32) != np.ascontiguousarray(sc.normals, np.float32).reshape(-1, 4).view(np.uint32)).sum())
                if nbn:
                    bad += nbn
                    print("  MISMATCH seed %d device vertex normals vs host: %d words" % (seed, nbn), flush=True)
            nb = int((nodes_d != sc.nodes).sum()) + int((tri_d != sc.triangles).sum()) + int((faces_d != sc.faces).sum()) + \
                int((aabbs_d.view(np.uint32) != np.ascontiguousarray(sc.aabbs, np.float32).reshape(-1, 4).view(np.uint32)).sum())
            if nb:
                bad += nb
                print("  MISMATCH seed %d device-built tree vs host builder: %d words" % (seed, nb), flush=True)
        # tile partition (emulated ranks on one GPU) and the pipelined download
        if seed % 3 == 0:
            import torch
            world = int(rng.choice([2, 3, 5]))
            tx, ty, tpr = host.tile_layout(rt.totalWidth, rt.totalHeight, world)
            gathered = torch.zeros(world * tpr * 1024, dtype=torch.float32, device="cuda")
            ctxs = [host.CudaHost(rt, jitter_seed=jitter, tile_rank=r, tile_world=world) for r in range(world)]
            for r, c in enumerate(ctxs):
                c.upload_scene(sc)
                c.bind_output(gathered[r * tpr * 1024:(r + 1) * tpr * 1024].data_ptr(), tpr * 1024)
                c()
            torch.cuda.synchronize()
            ctxs[0].deinterleave_async(gathered.data_ptr(), world)
            ctxs[0].synchronize()
            img = ctxs[0].download()
            for c in ctxs:
                c.close()
            nb = int(((img != base[1]) & ~(np.isnan(img) & np.isnan(base[1]))).sum())
            with host.CudaHost(rt, jitter_seed=jitter) as h:
                h.upload_scene(sc)
                img = h.render_download()
            nb += int(((img != base[1]) & ~(np.isnan(img) & np.isnan(base[1]))).sum())
            if nb:
                bad += nb
                print("  MISMATCH seed %d tile partition (world %d) / render_download: %d pixels" % (seed, world, nb), flush=True)
        # the collective-free tile paths: every emulated rank stores its share straight into one image (device and mapped host memory)
        if seed % 3 == 1:
            world = int(rng.choice([2, 3, 4]))
            shared = np.full((rt.totalHeight, rt.totalWidth), -1.0, np.float32)
            alias = host.host_register(shared)
            ctxs = [host.CudaHost(rt, jitter_seed=jitter, tile_rank=r, tile_world=world) for r in range(world)]
            for c in ctxs:
                c.upload_scene(sc)
                if seed % 2:
                    c.set_tunable(host.TUNE_FRUSTUM, 1)          # the packet kernel sends finished tiles itself
                c.render_store(alias)
            nb = int(((shared != base[1]) & ~(np.isnan(shared) & np.isnan(base[1]))).sum())
            if 32 % rt.n == 0:
                d_u8, _ = ctxs[0].peer_alloc(w * hgt)
                for c in ctxs:
                    c.set_tunable(host.TUNE_FRUSTUM, -1)
                    c()
                    c.resize_u8_to_async(d_u8)
                    c.synchronize()
                ctxs[0].adopt_u8(d_u8)
                nb += int((ctxs[0].download_u8() != host.host_resize(base[1], rt)).sum())
                ctxs[0].peer_free(d_u8)
            host.host_unregister(shared)
            for c in ctxs:
                c.close()
            if nb:
                bad += nb
                print("  MISMATCH seed %d direct stores (world %d): %d pixels" % (seed, world, nb), flush=True)
        # bounds-checked debug build (RTX_B200_LIB=.../librtx_b200_dbg.so): no index left its array during this seed
        try:
            with host.CudaHost(rt) as h:
                viol = h.debug_bounds()
            if viol[0]:
                bad += viol[0]
                print("  BOUNDS seed %d: %d out-of-range indices, first at rtx_kernels.cuh:%d (index %d, limit %d)" % (seed, *viol), flush=True)
        except host.RtxError:
            pass                                     # release build
        hitfrac = float((base[2][0] != host.NO_HIT).mean())
        print("seed %3d %-10s scale %-6g %4d tris depth %2d  %dx%d s=%d f=%.1f jitter %d  hit %.2f  %s" % (
            seed, kind, scale, sc.num_triangles, base[3], w, hgt, ss, focal, int(jitter != 0), hitfrac, "ok" if bad == 0 else "BAD (%d)" % bad), flush=True)
        total_bad += bad
    print("fuzz done: %d seeds, %d differences" % (count, total_bad))
    return 1 if total_bad else 0


if __name__ == "__main__":
    sys.exit(main())
