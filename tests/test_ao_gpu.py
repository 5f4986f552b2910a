"""Ambient occlusion (intersect_kernel.cl:128-183, 214-277, 305-307) on the B200 through the C ABI.

The oracle pins the transcendental functions the samplers use (sin, cos, acos, cospi, sinpi: "evaluate in
double precision, round once to float"), so the device can be held to the same bar as the primary path: the
float image is compared bit for bit.  Contractual tolerance (north_star): 8-bit pixels within +-1.
"""
import numpy as np
import pytest

from conftest import AO_CASES, require_gpu

pytestmark = pytest.mark.gpu

PIXEL_ATOL = 1


def _options(host, w, h, ss, ao, **kw):
    return host.Options(width=w, height=h, nSuperSamples=ss, enableAO=True, aoNumSamples=ao.samples, aoMethod=ao.method,
                        aoMaxDistance=float(ao.max_distance), aoAlphaMin=ao.alpha_min, aoAlphaMax=ao.alpha_max, **kw)


def _compare(host, po, sc, rt, h, ao, jitter_seed=0):
    ref = po.render(sc, rt.totalWidth, rt.totalHeight, po.focal_roundtrip(rt.options.focalLength), rt.options.enableShading,
                    jitter_seed=jitter_seed, ao=ao)
    img = h.download()
    u8 = h.download_u8()
    ref_u8 = po.resize(ref.image, rt.options.width, rt.options.height, rt.n)
    assert np.abs(u8.astype(int) - ref_u8.astype(int)).max() <= PIXEL_ATOL
    diff = img != ref.image
    assert not diff.any(), "%d of %d pixels differ, max %.3g" % (int(diff.sum()), diff.size, float(np.abs(img - ref.image).max()))
    assert np.array_equal(u8, ref_u8)
    return ref


@pytest.mark.parametrize("name", AO_CASES)
@pytest.mark.parametrize("kernel", [0, 1])
def test_soup_matches_reference_golden_ao(po, soup_scene, ao_golden, name, kernel):
    """Against the golden vectors the reference's own kernel text produced with AO_ENABLE."""
    host = require_gpu()
    g = ao_golden
    m, n, amin, amax = (int(x) for x in g["params_" + name])
    ao = po.Ao.make(method=m, samples=n, max_distance=float(g["maxdist_" + name]), alpha_min=amin, alpha_max=amax)
    rt = host.RayTracer(_options(host, int(g["width"]), int(g["height"]), int(g["nss"]), ao))
    with host.CudaHost(rt) as h:
        h.set_tunable(host.TUNE_KERNEL, kernel)
        h.upload_scene(soup_scene)
        assert h() is True
        assert np.array_equal(h.download(), g["image_" + name])
        assert np.array_equal(h.download_u8(), g["u8_" + name])
        assert h.stats()["kernel_launches"] >= 2


@pytest.mark.parametrize("method,samples", [(0, 3), (1, 3), (0, 1), (1, 6)])
def test_sibenik_standin_ao(po, sibenik_scene, method, samples):
    """The CLI's default options (uniform, 3 rings, 0.2) and the random sampler on the interior scene; the primary
    pass goes through the frustum front end, the occlusion pass starts from its recorded hits."""
    host = require_gpu()
    ao = po.Ao.make(method=method, samples=samples)
    rt = host.RayTracer(_options(host, 480, 270, 4, ao))
    with host.CudaHost(rt) as h:
        h.upload_scene(sibenik_scene)
        h()
        ref = _compare(host, po, sibenik_scene, rt, h, ao)
        plain = po.render(sibenik_scene, rt.totalWidth, rt.totalHeight, 1.0, True)
        assert (ref.image < plain.image).mean() > 0.005


def test_bunny_default_cli_options(po, bunny_scene):
    """`./render bunny.off out.pgm` with no options: 600x600, s=4, shading, uniform AO with 3 rings."""
    host = require_gpu()
    ao = po.Ao.make()
    rt = host.RayTracer(_options(host, 600, 600, 4, ao))
    with host.CudaHost(rt) as h:
        h.upload_scene(bunny_scene)
        h()
        _compare(host, po, bunny_scene, rt, h, ao)


@pytest.mark.parametrize("w,h_,ss,shading,leaf", [(33, 17, 1, False, 1), (101, 77, 1, True, 4), (50, 31, 9, True, 2)])
def test_ao_odd_sizes_leaf_sizes_and_no_shading(po, soup_scene, w, h_, ss, shading, leaf):
    """Odd sizes (literal-walk primary rays), leaves with several triangles (per-triangle leaf box check in the
    any-hit search), and AO without shading (value = 1 * occlusion, :296-307)."""
    host = require_gpu()
    for ao in (po.Ao.make(method=0, samples=2, max_distance=0.6), po.Ao.make(method=1, samples=2, max_distance=0.6)):
        rt = host.RayTracer(_options(host, w, h_, ss, ao, enableShading=shading))
        with host.CudaHost(rt) as h:
            h.set_tunable(host.TUNE_LEAF_SIZE, leaf)
            h.upload_scene(soup_scene)
            h()
            _compare(host, po, soup_scene, rt, h, ao)


def test_ao_zero_samples_is_off(po, soup_scene):
    """`#if defined(AO_ENABLE) && AO_NUM_SAMPLES > 0` (:305): -a 0 renders without occlusion."""
    host = require_gpu()
    rt = host.RayTracer(host.Options(width=64, height=48, nSuperSamples=4, enableAO=True, aoNumSamples=0))
    with host.CudaHost(rt) as h:
        h.upload_scene(soup_scene)
        h()
        assert np.array_equal(h.download(), po.render(soup_scene, 128, 96, 1.0, True).image)
        assert h.stats()["kernel_launches"] <= 5


@pytest.mark.parametrize("world", [2, 3])
def test_ao_tile_partition(po, soup_scene, world):
    """Tile-partitioned contexts with AO: the random sampler is seeded by the pixel's index in the WHOLE image (:169)."""
    import torch
    host = require_gpu()
    ao = po.Ao.make(method=1, samples=2, max_distance=0.6)
    rt = host.RayTracer(_options(host, 150, 70, 4, ao))
    ref = po.render(soup_scene, rt.totalWidth, rt.totalHeight, 1.0, True, ao=ao)
    tx, ty, tpr = host.tile_layout(rt.totalWidth, rt.totalHeight, world)
    gathered = torch.zeros(world * tpr * 1024, dtype=torch.float32, device="cuda")
    ctxs = [host.CudaHost(rt, tile_rank=r, tile_world=world) for r in range(world)]
    try:
        for r, c in enumerate(ctxs):
            c.upload_scene(soup_scene)
            n = tpr * 1024
            c.bind_output(gathered[r * n:(r + 1) * n].data_ptr(), n)
            c()
        torch.cuda.synchronize()
        ctxs[0].deinterleave_async(gathered.data_ptr(), world)
        ctxs[0].synchronize()
        assert np.array_equal(ctxs[0].download(), ref.image)
    finally:
        for c in ctxs:
            c.close()


def test_ao_on_a_tree_deeper_than_the_stack(po, scene_mod):
    """100 coincident triangles build a chain deeper than the 64-entry traversal stack (bvh.cc:85-93): primary rays and
    occlusion rays both take the literal stackless walk (walk_reference / walk_reference_any)."""
    host = require_gpu()
    from opencl_raytracer_b200 import scenes
    v1, _ = scenes.random_soup(1, seed=3, extent=0.5, size=2.0, big=0)
    vs, fs = scenes.random_soup(60, seed=8, size=0.5)
    v = np.concatenate([np.tile(v1, (100, 1)), vs])
    f = np.concatenate([np.arange(300, dtype=np.uint32).reshape(-1, 3), fs.astype(np.uint32) + 300])
    sc = scene_mod.scene_from_mesh(v, f, name="chain+soup")
    for ao in (po.Ao.make(method=0, samples=2, max_distance=0.8), po.Ao.make(method=1, samples=3, max_distance=0.8)):
        rt = host.RayTracer(_options(host, 96, 64, 4, ao))
        with host.CudaHost(rt) as h:
            h.upload_scene(sc)
            h()
            st = h.stats()
            assert st["tree_depth"] > 64 and st["kernel_variant"] == host.KERNEL_EXHAUSTIVE
            _compare(host, po, sc, rt, h, ao)
