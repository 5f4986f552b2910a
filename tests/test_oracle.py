"""The CPU oracle against the reference's own golden vectors (CPU only).

tests/golden/*.npz were produced by the reference's kernel text and host code
compiled for the CPU (tests/golden/make_golden.py); here the plain-C
restatement (oracle/rt_oracle.c) must reproduce them bit for bit, and -- where
oracle/_ref/ is present -- match the reference-compiled library directly.
"""
import hashlib

import numpy as np
import pytest


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def test_port_matches_golden_render(po, soup_scene, soup_golden, golden_meta):
    g = soup_golden
    assert soup_scene.digest() == golden_meta["digests"]["soup300_seed11"]
    n = int(np.sqrt(int(g["nss"])))
    W, H = int(g["width"]) * n, int(g["height"]) * n
    r = po.render(soup_scene, W, H, po.focal_roundtrip(float(g["focal"])), bool(g["shading"]))
    assert np.array_equal(r.face_id, g["face_id"])
    assert np.array_equal(_bits(r.distance), _bits(g["distance"]))
    assert np.array_equal(_bits(r.image), _bits(g["image"]))
    assert np.array_equal(po.resize(r.image, int(g["width"]), int(g["height"]), n), g["u8"])
    assert (r.face_id != po.NO_HIT).sum() > 100 and (r.face_id == po.NO_HIT).sum() > 100


def test_port_matches_golden_rays(po, soup_scene, rays_golden):
    g = rays_golden
    r = po.trace_rays(soup_scene, g["origins"], g["dirs"], 100000.0)
    assert np.array_equal(r.face_id, g["face_id"])
    assert np.array_equal(_bits(r.distance), _bits(g["distance"]))
    r = po.trace_rays(soup_scene, g["origins"], g["dirs"], 0.75)
    assert np.array_equal(r.face_id, g["face_id_d075"])
    assert np.array_equal(_bits(r.distance), _bits(g["distance_d075"]))
    # the axis-parallel tail really exercises 0*inf slabs and still produces hits
    assert (g["dirs"][-64:, :3] == 0).any(axis=1).all()


def test_port_matches_golden_special_direction_rays(po, soup_scene, special_rays_golden):
    """zero / subnormal / smallest-normal / huge direction components from origins exactly on leaf-box planes"""
    g = special_rays_golden
    for md, key in ((100000.0, ""), (0.75, "_d075")):
        r = po.trace_rays(soup_scene, g["origins"], g["dirs"], md)
        assert np.array_equal(r.face_id, g["face_id" + key])
        assert np.array_equal(_bits(r.distance), _bits(g["distance" + key]))
    d = np.abs(g["dirs"][:, :3])
    assert ((d > 0) & (d < 1.17549435e-38)).any() and (d == 0).any()        # subnormal and zero components are in the set
    assert (g["face_id"] != po.NO_HIT).sum() > 500                            # ... and the set still produces hits


def test_port_matches_golden_quad_odd_size(po, scene_mod, quad_golden, golden_meta):
    g = quad_golden
    sc = scene_mod.scene_from_mesh(g["verts"], g["faces"])
    assert sc.digest() == golden_meta["digests"]["quad_wall"]
    r = po.render(sc, 33, 17, 1.0, False)
    assert np.array_equal(r.face_id, g["face_id"])
    assert np.array_equal(_bits(r.distance), _bits(g["distance"]))
    assert np.array_equal(_bits(r.image), _bits(g["image"]))
    # every ray hits the wall; the shared diagonal is a tie region: first leaf wins
    assert (r.face_id != po.NO_HIT).all()
    assert set(np.unique(r.face_id)) == {0, 3}


def test_focal_roundtrip(po, golden_meta):
    for k, v in golden_meta["focal_roundtrip"].items():
        assert po.focal_roundtrip(float(k)) == np.float32(v)


def test_slab_nan_semantics(po):
    """intersect_kernel.cl:21-61 with a zero direction component and the origin on a box plane: 0*inf = NaN.
    Rejections are `a > b` comparisons (false for NaN) and max(a,b) = a < b ? b : a keeps a NaN only in the
    `a` position: a NaN from the y or z slab is dropped (box accepted), one from the x slab survives to the
    final `t_min < max_distance` and rejects.  Only the literal form reproduces this."""
    bb = np.array([0, 0, 0, 0, 1, 1, 1, 0], np.float32)
    assert po.aabb_intersect(bb, [0.5, 0.0, -1.0, 0], [0.0, 0.0, 1.0, 0], 100000.0)       # y: (0-0)*inf = NaN, dropped
    assert po.aabb_intersect(bb, [0.5, 1.0, -1.0, 0], [0.0, 0.0, 1.0, 0], 100000.0)       # y: NaN in ty_max, dropped
    assert not po.aabb_intersect(bb, [0.0, 0.5, -1.0, 0], [0.0, 0.0, 1.0, 0], 100000.0)   # x: NaN t_min survives
    assert not po.aabb_intersect(bb, [-0.5, 0.5, -1.0, 0], [0.0, 0.0, 1.0, 0], 100000.0)  # outside in x
    assert po.aabb_intersect(bb, [0.5, 0.5, -1.0, 0], [0.0, 0.0, 1.0, 0], 100000.0)
    assert not po.aabb_intersect(bb, [0.5, 0.5, -1.0, 0], [0.0, 0.0, 1.0, 0], 0.5)        # t_min = 1 >= max_distance
    assert not po.aabb_intersect(bb, [0.5, 0.5, 2.0, 0], [0.0, 0.0, 1.0, 0], 100000.0)    # behind the ray


def test_resize_truncates(po):
    tmp = np.array([[0.0, 1.0], [1.0, 1.0]], np.float32)
    assert po.resize(tmp, 1, 1, 2)[0, 0] == 191           # 0.75*255 = 191.25 -> 191
    tmp = np.full((4, 6), 0.999, np.float32)
    assert (po.resize(tmp, 3, 2, 2) == 254).all()


def test_jitter_and_random_rays_are_deterministic(po):
    a = [po.jitter(0x5EED, x, y) for x, y in ((0, 0), (1, 0), (0, 1), (4095, 2159))]
    assert all(0.0 <= v < 1.0 for p in a for v in p) and len(set(a)) == 4
    o, d = po.gen_random_rays(1234, 0, 1000, [-1, -2, -3], [1, 2, 3])
    o2, d2 = po.gen_random_rays(1234, 500, 500, [-1, -2, -3], [1, 2, 3])
    assert np.array_equal(o[500:], o2) and np.array_equal(d[500:], d2)
    assert (np.abs(o[:, :3]) <= [1, 2, 3]).all()
    assert np.allclose(np.linalg.norm(d[:, :3].astype(np.float64), axis=1), 1.0, atol=1e-6)


def test_counters_and_partial_rows(po, soup_scene):
    full = po.render(soup_scene, 64, 48, 1.0, True, want_counters=True)
    c = full.counters
    assert c["rays"] == 64 * 48 and c["node_visits"] >= c["box_hits"] >= c["tri_tests"] >= c["tri_hits"]
    assert c["hit_rays"] == int((full.face_id != po.NO_HIT).sum())
    part = po.render(soup_scene, 64, 48, 1.0, True, rows=(3, 48, 16), want_counters=True)
    assert part.counters["rays"] == 64 * 3
    for y in (3, 19, 35):
        assert np.array_equal(part.image[y], full.image[y])
    assert (part.image[4] == 0).all()
    one = po.render(soup_scene, 64, 48, 1.0, True, nthreads=1)
    assert np.array_equal(one.image, full.image)


def test_port_matches_reference_library_live(po, sibenik_scene):
    """Where oracle/_ref/ exists: restatement == reference kernel text on a 75k-triangle interior."""
    if po.ref() is None:
        pytest.skip("oracle/_ref/libref_oracle.so not present")
    W, H = 240, 136
    r = po.render(sibenik_scene, W, H, 1.0, True)
    img = po.ref_render(sibenik_scene, W, H, 1.0, True)
    fid, dist = po.ref_primary_hits(sibenik_scene, W, H, 1.0)
    assert np.array_equal(_bits(r.image), _bits(img))
    assert np.array_equal(r.face_id, fid) and np.array_equal(_bits(r.distance), _bits(dist))
    flat = po.render(sibenik_scene, W, H, 1.0, False)
    assert np.array_equal(_bits(flat.image), _bits(po.ref_render(sibenik_scene, W, H, 1.0, False)))


def test_bunny_known_answers(po, bunny_scene, golden_meta):
    """Config C1 (render -a 0, 600x600, s=4): SURVEY 8c known answers + reference hashes."""
    k = golden_meta["bunny_c1"]
    assert bunny_scene.digest() == golden_meta["digests"]["bunny"]
    assert bunny_scene.nodes.size == k["nodes"] and list(bunny_scene.nodes[:8]) == k["nodes_head"]
    r = po.render(bunny_scene, 1200, 1200, 1.0, True, want_counters=True)
    assert int((r.face_id != po.NO_HIT).sum()) == k["hit_rays"] == 766118
    assert hashlib.sha256(r.image.tobytes()).hexdigest() == k["image_sha256"]
    assert hashlib.sha256(r.face_id.tobytes()).hexdigest() == k["face_id_sha256"]
    assert hashlib.sha256(r.distance.tobytes()).hexdigest() == k["distance_sha256"]
    u8 = po.resize(r.image, 600, 600, 2)
    assert hashlib.sha256(u8.tobytes()).hexdigest() == k["u8_sha256"]
    assert int((u8 != 0).sum()) == k["pgm_nonzero"] == 191727
    assert abs(r.counters["V"] - 26.98) < 0.01 and abs(r.counters["T"] - 2.014) < 0.001
    assert abs(po.algorithmic_bytes_per_ray(r.counters["V"], r.counters["T"], r.counters["h"]) - 989.5) < 0.5


def _ao_of(po, g, name):
    m, n, amin, amax = (int(x) for x in g["params_" + name])
    return po.Ao.make(method=m, samples=n, max_distance=float(g["maxdist_" + name]), alpha_min=amin, alpha_max=amax)


@pytest.mark.parametrize("name", ["uniform3", "random3", "random1_far", "uniform2_a10_60", "uniform2_d07"])
def test_port_matches_golden_ambient_occlusion(po, soup_scene, ao_golden, name):
    """intersect_kernel.cl:214-277, 305-307 (both samplers) against the reference kernel text's output."""
    g = ao_golden
    n = int(np.sqrt(int(g["nss"])))
    W, H = int(g["width"]) * n, int(g["height"]) * n
    ao = _ao_of(po, g, name)
    r = po.render(soup_scene, W, H, 1.0, True, ao=ao)
    assert np.array_equal(_bits(r.image), _bits(g["image_" + name]))
    assert np.array_equal(po.resize(r.image, int(g["width"]), int(g["height"]), n), g["u8_" + name])
    plain = po.render(soup_scene, W, H, 1.0, True)
    assert (r.image <= plain.image).all() and (r.image < plain.image).sum() > 50      # occlusion only darkens, and does
    assert np.array_equal(r.face_id, plain.face_id)


def test_ambient_occlusion_port_matches_reference_library_live(po, sibenik_scene):
    """Where oracle/_ref/ exists: restatement == reference kernel text with AO_ENABLE on the interior scene
    (the CLI's default options: uniform, 3 rings, 0.2, 4..90 degrees) and with the random sampler."""
    if po.ref() is None:
        pytest.skip("oracle/_ref/libref_oracle.so not present")
    W, H = 96, 64
    for ao in (po.Ao.make(), po.Ao.make(method=1, samples=4), po.Ao.make(method=0, samples=4, max_distance=2.0)):
        r = po.render(sibenik_scene, W, H, 1.0, True, ao=ao)
        img = po.ref_render_ao(sibenik_scene, W, H, ao)
        assert np.array_equal(_bits(r.image), _bits(img))


def test_port_on_the_reference_sah_tree(po, sah_scene, sah_golden, soup_scene):
    """The path consumes whatever tree bvh.cc builds: `-r sah` (bvh.cc:178-236) gives another topology for the same
    triangles.  Arrays and renders come from the reference builder / kernel text (tests/golden/make_golden.py sah)."""
    g = sah_golden
    assert sah_scene.nodes.size == soup_scene.nodes.size and not np.array_equal(sah_scene.nodes, soup_scene.nodes)
    r = po.render(sah_scene, 64, 48, 1.0, True)
    assert np.array_equal(r.face_id, g["face_id"]) and np.array_equal(_bits(r.distance), _bits(g["distance"]))
    assert np.array_equal(_bits(r.image), _bits(g["image"]))
    ao = po.Ao.make(method=0, samples=2, max_distance=0.7)
    assert np.array_equal(_bits(po.render(sah_scene, 64, 48, 1.0, True, ao=ao).image), _bits(g["image_ao_uniform2_d07"]))
    # same triangles, other leaf order: same picture wherever no two triangles tie
    plain = po.render(soup_scene, 64, 48, 1.0, True)
    assert (r.distance == plain.distance).mean() > 0.999


def test_ambient_occlusion_bunny_defaults_against_reference_library(po, bunny_scene):
    """`./render bunny.off out.pgm` with no options (uniform AO, 3 rings, 0.2): restatement == the reference's kernel
    text on a 300x300 frame of the reference's own mesh (thousands of candidate leaves within reach of a surface point)."""
    if po.ref() is None:
        pytest.skip("oracle/_ref/libref_oracle.so not present")
    ao = po.Ao.make()
    r = po.render(bunny_scene, 300, 300, 1.0, True, ao=ao)
    img = po.ref_render_ao(bunny_scene, 300, 300, ao)
    assert np.array_equal(_bits(r.image), _bits(img))
    plain = po.render(bunny_scene, 300, 300, 1.0, True)
    assert (r.image < plain.image).mean() > 0.1
