"""Generate the committed golden vectors from the REFERENCE ITSELF.

Run in the build container (needs /root/reference, i.e. oracle/_ref/):

    python tests/golden/make_golden.py

Everything is produced by oracle/_ref/libref_oracle.so -- the reference's own
kernel text (src/intersect_kernel.cl) and host code (mesh.cc, bvh.cc,
ray_tracer.cc, compiler_options.h) compiled for the CPU -- never by the
restatement or the CUDA path.  Outputs (small, committed):

  soup_64x48.npz      float image, hit ids, distances, u8 image of a 300-triangle soup
  soup_rays.npz       4096 arbitrary rays (incl. axis-parallel ones) and their hits
  quad_33x17.npz      two big triangles sharing a diagonal, odd image size
  scenes.json         sha256 digests of the reference builder's arrays + bunny known answers
  soup_sah.npz        the soup's tree from the reference's SAH builder (`-r sah`, bvh.cc:178-236) as upload arrays +
                      its render; `python tests/golden/make_golden.py sah` writes only this file
  soup_special_rays.npz  3072 rays with zero / subnormal / tiny / huge direction components from origins exactly on
                      leaf-box planes (`python tests/golden/make_golden.py special` writes only this file)
  soup_ao.npz         the soup with ambient occlusion (uniform rings 3, random 3, random 1 with a longer reach);
                      `python tests/golden/make_golden.py ao` writes only this file
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from opencl_raytracer_b200 import scenes  # noqa: E402  (mesh generators only)
from opencl_raytracer_b200.scene import Scene  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def ref_scene(v, f):
    r = po.ref_scene_from_mesh(v, f)
    return Scene(r.faces, r.nodes, r.aabbs, r.vertices, r.normals, r.triangles, r.orig_faces)


def render_case(sc, w, h, ss, focal=1.0, shading=True):
    tw, th = po.ref_total_dims(w, h, ss)
    img = po.ref_render(sc, tw, th, po.ref_focal_roundtrip(focal), shading)
    fid, dist = po.ref_primary_hits(sc, tw, th, po.ref_focal_roundtrip(focal))
    u8 = po.ref_resize(img, w, h, ss)
    return dict(image=img, face_id=fid, distance=dist, u8=u8, width=w, height=h, nss=ss, focal=np.float32(focal),
                shading=int(shading))


def make_ao():
    """Ambient occlusion (intersect_kernel.cl:214-277, 305-307) through the reference's kernel text."""
    v, f = scenes.random_soup(300, seed=11)
    sc = ref_scene(v, f)
    tw, th = po.ref_total_dims(32, 24, 4)
    out = dict(verts=v, faces=f, width=32, height=24, nss=4)
    for name, ao in (("uniform3", po.Ao.make(method=0, samples=3)), ("random3", po.Ao.make(method=1, samples=3)),
                     ("random1_far", po.Ao.make(method=1, samples=1, max_distance=1.5)),
                     ("uniform2_a10_60", po.Ao.make(method=0, samples=2, max_distance=0.7, alpha_min=10, alpha_max=60)),
                     ("uniform2_d07", po.Ao.make(method=0, samples=2, max_distance=0.7))):
        img = po.ref_render_ao(sc, tw, th, ao, po.ref_focal_roundtrip(1.0))
        out["image_" + name] = img
        out["u8_" + name] = po.ref_resize(img, 32, 24, 4)
        out["params_" + name] = np.array([ao.method, ao.samples, ao.alpha_min, ao.alpha_max], np.int32)
        out["maxdist_" + name] = np.float32(ao.max_distance)
    np.savez_compressed(os.path.join(OUT, "soup_ao.npz"), **out)
    print("soup_ao.npz written")


def make_sah():
    """A tree of different topology for the same triangles: the reference's O(n^2) SAH builder (chatty on stdout)."""
    v, f = scenes.random_soup(300, seed=11)
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)
    try:
        r = po.ref_scene_from_mesh(v, f, sah=True)
    finally:
        os.dup2(saved, 1)
        os.close(devnull)
        os.close(saved)
    sc = Scene(r.faces, r.nodes, r.aabbs, r.vertices, r.normals, r.triangles, r.orig_faces)
    out = render_case(sc, 32, 24, 4)
    ao = po.Ao.make(method=0, samples=2, max_distance=0.7)
    tw, th = po.ref_total_dims(32, 24, 4)
    img_ao = po.ref_render_ao(sc, tw, th, ao, po.ref_focal_roundtrip(1.0))
    np.savez_compressed(os.path.join(OUT, "soup_sah.npz"), t_faces=sc.faces, t_nodes=sc.nodes, t_aabbs=sc.aabbs,
                        t_vertices=sc.vertices, t_normals=sc.normals, image_ao_uniform2_d07=img_ao, **out)
    print("soup_sah.npz written (%d nodes)" % sc.nodes.size)


def make_special():
    """Rays at the edge of IEEE arithmetic: direction components that are zero, subnormal (1/d overflows to inf below
    2^-128, is finite just above), the smallest and the largest normals, from origins that lie EXACTLY on planes of leaf
    boxes (so (bb - o) * (1/d) is 0 * inf = NaN there) and on vertices.  Hits of the reference's kernel text."""
    v, f = scenes.random_soup(300, seed=11)
    sc = ref_scene(v, f)
    rng = np.random.default_rng(77)
    n = 3072
    leaf = np.flatnonzero(sc.nodes == 1)
    pick = rng.choice(leaf, n)
    lo, hi = sc.aabbs[2 * pick, :3], sc.aabbs[2 * pick + 1, :3]
    o = np.zeros((n, 4), np.float32)
    d = np.zeros((n, 4), np.float32)
    o[:, :3] = lo + (hi - lo) * rng.uniform(-0.5, 1.5, (n, 3)).astype(np.float32)
    d[:, :3] = rng.normal(size=(n, 3)).astype(np.float32)
    ax = rng.integers(0, 3, n)
    rows = np.arange(n)
    side = rng.random(n) < 0.5
    o[rows, ax] = np.where(side, lo[rows, ax], hi[rows, ax])                  # exactly on a plane of the picked leaf box
    special = np.array([0.0, -0.0, 1e-45, -1e-45, 1e-40, -1e-40, 2.9e-39, -2.9e-39, 2.95e-39, 5e-39, -5e-39, 1.17549435e-38,
                        -1.17549435e-38, 1.2e-38, 3e-38, 1e-30, -1e-30, 3.4e38, -3.4e38], np.float32)
    d[rows, ax] = special[rng.integers(0, special.size, n)]
    two = rng.random(n) < 0.25                                                # a second special component
    ax2 = (ax + 1 + rng.integers(0, 2, n)) % 3
    d[rows[two], ax2[two]] = special[rng.integers(0, special.size, int(two.sum()))]
    o[:64, :3] = sc.vertices[sc.faces[:64], :3]                               # origins on vertices
    # (no NaN / inf components: with them the reference's triangle test "hits" without ever updating the closest
    # hit, and the kernel then reads an Intersection it never initialised -- intersect_kernel.cl:106-112, 292-299)
    fid, dist = po.ref_trace_rays(sc, o, d, 100000.0)
    fid2, dist2 = po.ref_trace_rays(sc, o, d, 0.75)
    np.savez_compressed(os.path.join(OUT, "soup_special_rays.npz"), verts=v, faces=f, origins=o, dirs=d, face_id=fid, distance=dist,
                        face_id_d075=fid2, distance_d075=dist2)
    print("soup_special_rays.npz written: %d rays, %d hits" % (n, int((fid != po.NO_HIT).sum())))


def main():
    assert po.ref() is not None, "oracle/_ref/libref_oracle.so missing: run make -C oracle"
    if sys.argv[1:] == ["special"]:
        return make_special()
    if sys.argv[1:] == ["ao"]:
        return make_ao()
    if sys.argv[1:] == ["sah"]:
        return make_sah()
    digests = {}

    v, f = scenes.random_soup(300, seed=11)
    sc = ref_scene(v, f)
    digests["soup300_seed11"] = sc.digest()
    np.savez_compressed(os.path.join(OUT, "soup_64x48.npz"), verts=v, faces=f, **render_case(sc, 32, 24, 4, focal=1.2345678))

    # arbitrary rays: random + axis-parallel (zero direction components -> 0*inf NaN slabs)
    lo, hi = sc.root_box()
    o, d = po.gen_random_rays(1234, 0, 4096 - 64, lo, hi)
    rng = np.random.default_rng(5)
    ao = np.zeros((64, 4), np.float32)
    ad = np.zeros((64, 4), np.float32)
    ao[:, :3] = rng.uniform(lo, hi, (64, 3))
    ax = rng.integers(0, 3, 64)
    ad[np.arange(64), ax] = rng.choice([-1.0, 1.0], 64)
    ad[np.arange(32, 64), (ax[32:] + 1) % 3] = rng.uniform(-1, 1, 32)           # one zero component only
    ao[:8, :3] = sc.vertices[sc.faces[:8], :3]                      # origins exactly on vertices / box planes
    o = np.concatenate([o, ao]); d = np.concatenate([d, ad])
    fid, dist = po.ref_trace_rays(sc, o, d, 100000.0)
    fid2, dist2 = po.ref_trace_rays(sc, o, d, 0.75)
    np.savez_compressed(os.path.join(OUT, "soup_rays.npz"), verts=v, faces=f, origins=o, dirs=d, face_id=fid, distance=dist,
                        face_id_d075=fid2, distance_d075=dist2)

    v, f = scenes.quad_wall()
    sc = ref_scene(v, f)
    digests["quad_wall"] = sc.digest()
    np.savez_compressed(os.path.join(OUT, "quad_33x17.npz"), verts=v, faces=f, **render_case(sc, 33, 17, 1, shading=False))

    v, f = scenes.sibenik_standin()
    sc = ref_scene(v, f)
    digests["sibenik_standin"] = sc.digest()
    digests["sibenik_standin_tris"] = int(sc.num_triangles)

    kat = {}
    bunny = "/root/reference/meshes/bunny.off"
    if os.path.exists(bunny):
        r = po.ref_scene_from_off(bunny)
        sc = Scene(r.faces, r.nodes, r.aabbs, r.vertices, r.normals)
        digests["bunny"] = sc.digest()
        img = po.ref_render(sc, 1200, 1200, 1.0, True)
        fid, dist = po.ref_primary_hits(sc, 1200, 1200, 1.0)
        u8 = po.ref_resize(img, 600, 600, 4)
        import hashlib
        kat = dict(nodes=int(sc.nodes.size), nodes_head=[int(x) for x in sc.nodes[:8]], hit_rays=int((fid != po.NO_HIT).sum()),
                   pgm_mean=float(u8.mean()), pgm_nonzero=int((u8 != 0).sum()),
                   image_sha256=hashlib.sha256(img.tobytes()).hexdigest(),
                   face_id_sha256=hashlib.sha256(fid.tobytes()).hexdigest(),
                   distance_sha256=hashlib.sha256(dist.tobytes()).hexdigest(),
                   u8_sha256=hashlib.sha256(u8.tobytes()).hexdigest())
    with open(os.path.join(OUT, "scenes.json"), "w") as fh:
        json.dump(dict(digests=digests, bunny_c1=kat,
                       focal_roundtrip={str(x): po.ref_focal_roundtrip(x) for x in (1.0, 1.2345678, 0.5, 3.3333333, 0.001, 123456.789)}),
                  fh, indent=1, sort_keys=True)
    make_ao()
    make_sah()
    make_special()
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
