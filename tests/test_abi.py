"""The C-ABI library loads and exports every symbol the headers declare (CPU only; no compute calls)."""
import ctypes
import os
import re

import pytest

from opencl_raytracer_b200 import host, scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rtx_[a-z0-9_]+)\s*\(", text)))


def test_rtx_b200_exports_every_declared_symbol():
    names = declared("rtx_b200.h")
    assert len(names) >= 20 and set(names) == set(host.ABI_SYMBOLS)
    lib = ctypes.CDLL(host.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), n


def test_rtx_scene_exports_every_declared_symbol():
    names = declared("rtx_scene.h")
    lib = scene._load()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), n


def test_product_never_links_the_oracle():
    """The oracle is test infrastructure: the shipped libraries must not depend on it."""
    import subprocess
    for lib in ("librtx_b200.so", "librtx_scene.so"):
        out = subprocess.run(["ldd", os.path.join(ROOT, "opencl_raytracer_b200", "lib", lib)], capture_output=True, text=True).stdout
        assert "oracle" not in out
    for dirpath, _, files in os.walk(os.path.join(ROOT, "opencl_raytracer_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                for pat in (r"#include\s+\"[^\"]*rt_oracle", r"^\s*(from|import)\s+oracle", "pyoracle", "liboracle", "libref_oracle"):
                    assert not re.search(pat, text, flags=re.M), (f, pat)


def test_fails_loudly_without_a_device():
    if host.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(host.RtxError) as e:
        host.CudaHost(host.RayTracer(host.Options()))
    assert e.value.code == host.ERR_NO_DEVICE


def test_argument_errors_need_no_device():
    lib = host.load_library()
    assert lib.rtx_create(None, None) == host.ERR_ARG
    assert lib.rtx_render(None) == host.ERR_ARG
    assert lib.rtx_upload(None, None, 0, None, 0, None, 0, None, 0, None, 0) == host.ERR_ARG
    assert b"null" in lib.rtx_last_error(None)
    assert host.tile_layout(3840, 2160, 8) == (120, 68, 1020)
    assert host.tile_layout(33, 17, 2) == (2, 1, 1)
    rt = host.RayTracer(host.Options(width=1920, height=1080, nSuperSamples=4))
    assert (rt.totalWidth, rt.totalHeight) == (3840, 2160)
    assert host.RayTracer(host.Options(width=10, height=10, nSuperSamples=8)).totalWidth == 20   # (unsigned)sqrt(8) = 2


def test_python_cli_mirror_defaults_match_render_cc():
    """opencl_raytracer_b200.render mirrors src/render.cc:16-47: same option letters, same defaults
    { 600, 600, 1.f, 4, shading, AO on, .2f, 3, UNIFORM, 4, 90, LONGEST }; -h is the height, --help the help."""
    from opencl_raytracer_b200 import render
    a = render.parse(["in.off", "out.pgm"])
    assert (a.width, a.height, a.focal_length, a.supersamples) == (600, 600, 1.0, 4)
    assert (a.ambient_occlusion_samples, a.ambient_occlusion_max_distance, a.ambient_occlusion_method) == (3, 0.2, "uniform")
    assert a.bvh_strategy == "longest" and a.input_mesh == "in.off" and a.output_image == "out.pgm"
    a = render.parse(["-w", "32", "-h", "24", "-a", "0", "-d", "1.5", "-m", "random", "-f", "2.5", "-s", "16", "a", "b"])
    assert (a.width, a.height, a.ambient_occlusion_samples, a.ambient_occlusion_max_distance) == (32, 24, 0, 1.5)
    assert (a.ambient_occlusion_method, a.focal_length, a.supersamples) == ("random", 2.5, 16)
    a = render.parse(["--width=7", "--height=9", "--ambient-occlusion-samples=2", "--supersamples=9", "a", "b"])
    assert (a.width, a.height, a.ambient_occlusion_samples, a.supersamples) == (7, 9, 2, 9)
