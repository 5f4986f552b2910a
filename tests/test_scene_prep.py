"""Host scene preparation (include/rtx_scene.h) == the reference's mesh.cc + bvh.cc (CPU only)."""
import os

import numpy as np
import pytest

from opencl_raytracer_b200 import scenes


def _same(a, b):
    for n in ("faces", "nodes", "aabbs", "vertices", "normals", "triangles"):
        x, y = np.ascontiguousarray(getattr(a, n)), np.ascontiguousarray(getattr(b, n))
        assert x.shape == y.shape and np.array_equal(x.view(np.uint32), y.view(np.uint32)), n


def check_invariants(sc):
    """full binary pre-order tree, subtree sizes, exact parent boxes."""
    nodes, aabbs = sc.nodes, sc.aabbs
    n = nodes.size
    assert nodes[0] == n == 2 * sc.num_triangles - 1
    leaves = np.flatnonzero(nodes == 1)
    assert leaves.size == sc.num_triangles
    inner = np.flatnonzero(nodes > 1)
    left = inner + 1
    right = left + nodes[left]
    assert (nodes[inner] == 1 + nodes[left] + nodes[right]).all()
    lo, hi = aabbs[0::2, :3], aabbs[1::2, :3]
    assert np.array_equal(lo[inner], np.minimum(lo[left], lo[right]))
    assert np.array_equal(hi[inner], np.maximum(hi[left], hi[right]))
    tri = sc.vertices[sc.faces.reshape(-1, 3), :3]
    assert np.array_equal(lo[leaves], tri.min(1)) and np.array_equal(hi[leaves], tri.max(1))
    assert (aabbs[:, 3] == 0).all() and (sc.vertices[:, 3] == 0).all() and (sc.normals[:, 3] == 0).all()
    assert np.array_equal(sc.faces.reshape(-1, 3), sc.orig_faces.reshape(-1, 3)[sc.triangles])
    assert sorted(sc.triangles.tolist()) == list(range(sc.num_triangles))


def test_golden_digests(scene_mod, soup_scene, sibenik_scene, golden_meta):
    d = golden_meta["digests"]
    assert soup_scene.digest() == d["soup300_seed11"]
    assert sibenik_scene.digest() == d["sibenik_standin"]
    assert sibenik_scene.num_triangles == d["sibenik_standin_tris"]
    check_invariants(soup_scene)
    check_invariants(sibenik_scene)


def test_thread_count_does_not_change_the_tree(scene_mod):
    v, f = scenes.sibenik_standin()
    a = scene_mod.scene_from_mesh(v, f, nthreads=1)
    b = scene_mod.scene_from_mesh(v, f, nthreads=5)
    _same(a, b)


def test_matches_reference_builder_live(scene_mod, po):
    if po.ref() is None:
        pytest.skip("oracle/_ref/libref_oracle.so not present")
    cases = [scenes.random_soup(1, seed=3), scenes.random_soup(2, seed=4), scenes.random_soup(777, seed=5),
             scenes.quad_wall(), scenes.sibenik_standin(detail=0.4)]
    # duplicates: all centroids equal -> the reference's "left/right empty" fix-ups (bvh.cc:85-93) build a chain
    v, f = scenes.random_soup(1, seed=9)
    cases.append((np.tile(v, (40, 1)), (np.arange(120, dtype=np.uint32)).reshape(-1, 3)))
    for v, f in cases:
        _same(scene_mod.scene_from_mesh(v, f), po.ref_scene_from_mesh(v, f))


def test_bunny_digest(bunny_scene, golden_meta):
    assert bunny_scene.digest() == golden_meta["digests"]["bunny"]
    check_invariants(bunny_scene)


def test_off_loader(tmp_path, scene_mod, soup_golden, soup_scene):
    p = str(tmp_path / "soup.off")
    scene_mod.write_off(p, soup_golden["verts"], soup_golden["faces"])
    sc = scene_mod.scene_from_off(p)
    _same(sc, soup_scene)
    # error behaviour of load_off_mesh (mesh.cc:7-67)
    with pytest.raises(scene_mod.SceneError) as e:
        scene_mod.scene_from_off(str(tmp_path / "missing.off"))
    assert e.value.code == 2
    bad = tmp_path / "bad.off"
    bad.write_text("PLY\n1 1 0\n")
    with pytest.raises(scene_mod.SceneError) as e:
        scene_mod.scene_from_off(str(bad))
    assert e.value.code == 3
    quad = tmp_path / "quad.off"
    quad.write_text("OFF\n4 1 0\n0 0 0\n1 0 0\n1 1 0\n0 1 0\n4 0 1 2 3\n")
    with pytest.raises(scene_mod.SceneError) as e:
        scene_mod.scene_from_off(str(quad))
    assert e.value.code == 3 and "!= 3" in str(e.value)
    # a face naming a vertex that does not exist is skipped (mesh.cc:48-59)
    skip = tmp_path / "skip.off"
    skip.write_text("OFF\n4 2 0\n0 0 0\n1 0 0\n1 1 0\n0 1 0\n3 0 1 2\n3 0 1 9\n")
    assert scene_mod.scene_from_off(str(skip)).num_triangles == 1
    empty = tmp_path / "empty.off"
    empty.write_text("OFF\n3 0 0\n0 0 0\n1 0 0\n1 1 0\n")
    with pytest.raises(scene_mod.SceneError) as e:
        scene_mod.scene_from_off(str(empty))
    assert e.value.code == 4


def test_subdivision_counts():
    v, f = scenes.icosphere((0, 0, 0), 1.0, 1)
    v4, f4 = scenes.subdivide_1to4(v, f)
    v9, f9 = scenes.subdivide_1to9(v, f)
    assert f4.shape[0] == 4 * f.shape[0] and f9.shape[0] == 9 * f.shape[0]
    e = 3 * f.shape[0] // 2                        # closed surface
    assert v4.shape[0] == v.shape[0] + e and v9.shape[0] == v.shape[0] + 2 * e + f.shape[0]

    def area(v, f):
        a, b, c = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
        return np.linalg.norm(np.cross(b - a, c - a), axis=1).sum()

    assert np.isclose(area(v4, f4), area(v, f)) and np.isclose(area(v9, f9), area(v, f))
    vs, fs = scenes.subdivided(v, f)
    assert fs.shape[0] == 144 * f.shape[0] and vs.dtype == np.float32


def test_scene_cache_roundtrip_and_corruption(tmp_path):
    """On-disk cache: a hit returns the builder's arrays byte for byte; a damaged entry is rebuilt."""
    from opencl_raytracer_b200 import scene as scn, scenes
    v, f = scenes.sibenik_standin(detail=0.2)
    built = scn.scene_from_mesh(v, f, name="standin")
    d = str(tmp_path / "cache")
    first = scn.cached_scene_from_mesh(v, f, d, name="standin")
    entries = list((tmp_path / "cache").glob("scene_*.npz"))
    assert len(entries) == 1 and first.digest() == built.digest()
    again = scn.cached_scene_from_mesh(v, f, d, name="standin")          # served from the file
    assert again.digest() == built.digest() and again.name == "standin"
    assert np.array_equal(again.triangles, built.triangles) and np.array_equal(again.orig_faces, built.orig_faces)
    # another mesh -> another key
    assert scn.mesh_key(v * 2.0, f) != scn.mesh_key(v, f)
    # a flipped byte in the payload is caught by the stored sha256 and the entry rebuilt
    raw = bytearray(entries[0].read_bytes())
    sc2 = scn.load_scene(str(entries[0]))
    sc2.aabbs = sc2.aabbs.copy()
    sc2.aabbs[3, 1] += 1.0
    import pytest
    scn.save_scene(str(entries[0]), sc2)                                  # digest recomputed: loads fine, other contents
    assert scn.load_scene(str(entries[0])).digest() != built.digest()
    entries[0].write_bytes(bytes(raw[:len(raw) // 2]))                    # truncated file
    with pytest.raises(scn.SceneError):
        scn.load_scene(str(entries[0]))
    healed = scn.cached_scene_from_mesh(v, f, d)
    assert healed.digest() == built.digest()


def test_scene_cache_keyed_by_off_file(tmp_path):
    from opencl_raytracer_b200 import scene as scn, scenes
    v, f = scenes.sibenik_standin(detail=0.2)
    off = str(tmp_path / "m.off")
    scn.write_off(off, v, f)
    a = scn.cached_scene_from_off(off, str(tmp_path / "c"))
    b = scn.cached_scene_from_off(off, str(tmp_path / "c"))
    assert a.digest() == b.digest() == scn.scene_from_off(off).digest()
    assert len(list((tmp_path / "c").glob("*.npz"))) == 1


def test_sah_builder_emits_the_reference_tree(sah_golden):
    """BVH::Method::SURFACE_AREA_HEURISTIC (bvh.cc:178-236, `-r sah`) on the host: the golden arrays come from the
    reference's own builder (tests/golden/make_golden.py sah)."""
    from opencl_raytracer_b200 import scene as scn, scenes
    v, f = scenes.random_soup(300, seed=11)
    sc = scn.scene_from_mesh(v, f, sah=True)
    g = sah_golden
    for key, got in (("t_faces", sc.faces), ("t_nodes", sc.nodes), ("t_aabbs", sc.aabbs), ("t_vertices", sc.vertices), ("t_normals", sc.normals)):
        assert np.array_equal(g[key], got), key
    assert not np.array_equal(scn.scene_from_mesh(v, f).nodes, sc.nodes)        # another topology than the default builder
    assert sc.digest() == scn.scene_from_mesh(v, f, sah=True, nthreads=1).digest()


def test_sah_builder_against_the_reference_builder_with_ties(po):
    """Regular grids give many equal centroids: the order std::sort leaves them in decides the tree.  Runs where the
    reference's bvh.cc was compiled (oracle/_ref); the reference prints one line per cut candidate, so stdout is muted."""
    import sys
    from opencl_raytracer_b200 import scene as scn, scenes
    try:
        po.ref()
    except Exception as e:
        pytest.skip("oracle/_ref not built: %s" % e)
    for v, f in (scenes.sibenik_standin(detail=0.06), scenes.random_soup(120, seed=9)):
        sys.stdout.flush()
        saved, devnull = os.dup(1), os.open(os.devnull, os.O_WRONLY)
        os.dup2(devnull, 1)
        try:
            r = po.ref_scene_from_mesh(v, f, sah=True)
        finally:
            os.dup2(saved, 1)
            os.close(devnull)
            os.close(saved)
        m = scn.scene_from_mesh(v, f, sah=True)
        for a, b in ((r.faces, m.faces), (r.nodes, m.nodes), (r.aabbs, m.aabbs), (r.triangles, m.triangles)):
            assert np.array_equal(a, b)
