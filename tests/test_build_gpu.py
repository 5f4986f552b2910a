"""rtx_upload_mesh: the reference's longest-axis BVH built on the device (rtx_build.cuh) must emit the very arrays
bvh.cc emits -- `nodes`, `aabbs`, `triangles` and the leaf-ordered faces of render.cc:88-95 -- here compared bit
for bit with the host builder of include/rtx_scene.h, which tests/test_scene_prep.py pins to the reference's own
bvh.cc through the golden digests."""
import numpy as np
import pytest

from conftest import require_gpu

pytestmark = pytest.mark.gpu


def _check(host, po, sc, render=True):
    rt = host.RayTracer(host.Options(width=96, height=64, nSuperSamples=4))
    with host.CudaHost(rt) as h:
        h.upload_mesh(sc.vertices, sc.orig_faces, sc.normals)
        nodes, aabbs, tri, faces = h.download_tree()
        assert np.array_equal(nodes, sc.nodes)
        assert np.array_equal(tri, sc.triangles)
        assert np.array_equal(faces, sc.faces)
        assert np.array_equal(aabbs.view(np.uint32), np.ascontiguousarray(sc.aabbs, np.float32).reshape(-1, 4).view(np.uint32))
        ms, levels = h.build_stats()
        assert levels >= 1 or sc.num_triangles == 1
        # mesh.cc:95-139 on the device: same sums in the same order -> the same bits (NaN-free: zero normals stay zero)
        h.upload_mesh(sc.vertices, sc.orig_faces, None)
        vn = h.download_normals()
        assert np.array_equal(vn.view(np.uint32), np.ascontiguousarray(sc.normals, np.float32).reshape(-1, 4).view(np.uint32))
        assert np.array_equal(h.download_tree()[0], sc.nodes)
        if render:
            h.set_tunable(host.TUNE_RECORD_HITS, 1)
            h()
            ref = po.render(sc, rt.totalWidth, rt.totalHeight, 1.0, True)
            fid, dist = h.download_hits()
            assert np.array_equal(fid, ref.face_id) and np.array_equal(dist, ref.distance)
            img = h.download()
            assert np.array_equal(np.nan_to_num(img), np.nan_to_num(ref.image))
        return ms, levels


def test_device_build_equals_reference_builder(po, scene_mod, soup_scene, sibenik_scene, bunny_scene):
    host = require_gpu()
    for sc in (soup_scene, sibenik_scene, bunny_scene):
        ms, levels = _check(host, po, sc)
        assert levels < 64


@pytest.mark.parametrize("kind", ["needles", "degenerate", "chain", "single", "pair", "grid"])
def test_device_build_edge_cases(po, scene_mod, kind):
    """Empty-side fix-ups (bvh.cc:85-93): coincident centroids peel one triangle per level (a chain as deep as the
    mesh); one and two triangles; centroids exactly on the cut plane (regular grid)."""
    host = require_gpu()
    from opencl_raytracer_b200 import scenes
    if kind == "needles":
        v, f = scenes.needle_soup(1500, seed=4)
    elif kind == "degenerate":
        v, f = scenes.degenerate_soup(1200, seed=9)
    elif kind == "chain":
        v1, _ = scenes.random_soup(1, seed=3, extent=0.5, size=2.0, big=0)
        v = np.tile(v1, (70, 1))
        f = np.arange(210, dtype=np.uint32).reshape(-1, 3)
    elif kind == "single":
        v, f = scenes.random_soup(1, seed=3, extent=0.5, size=2.0, big=0)
    elif kind == "pair":
        v, f = scenes.quad_wall()
    else:
        g = np.linspace(-1.0, 1.0, 33)
        xx, yy = np.meshgrid(g, g, indexing="ij")
        v = np.stack([xx, yy, np.full_like(xx, -1.0)], -1).reshape(-1, 3)
        i, j = np.meshgrid(np.arange(32), np.arange(32), indexing="ij")
        a, b, c, d = i * 33 + j, (i + 1) * 33 + j, (i + 1) * 33 + j + 1, i * 33 + j + 1
        f = np.concatenate([np.stack([a, b, c], -1).reshape(-1, 3), np.stack([a, c, d], -1).reshape(-1, 3)])
    sc = scene_mod.scene_from_mesh(np.asarray(v, np.float32), f, name=kind)
    ms, levels = _check(host, po, sc)
    if kind == "chain":
        assert levels >= 69


def test_device_build_rejects_bad_faces(soup_scene):
    host = require_gpu()
    rt = host.RayTracer(host.Options(width=16, height=16, nSuperSamples=1))
    with host.CudaHost(rt) as h:
        bad = soup_scene.orig_faces.copy()
        bad[7] = soup_scene.vertices.shape[0] + 3
        with pytest.raises(host.RtxError) as e:
            h.upload_mesh(soup_scene.vertices, bad, soup_scene.normals)
        assert e.value.code == host.ERR_ARG and "face index" in str(e.value)
        with pytest.raises(host.RtxError):
            h()
        h.upload_mesh(soup_scene.vertices, soup_scene.orig_faces, soup_scene.normals)
        assert h() is True
