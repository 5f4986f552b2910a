"""End-to-end drop-in check: the reference's own `render` CLI (src/render.cc + mesh.cc + bvh.cc + ..., all
unmodified, compiled by oracle/Makefile into oracle/_ref/render_b200) linked against librtx_b200.so through the
shadow opencl_host.h must write the PGM the reference's kernel text produces."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, require_gpu

CLI = os.path.join(ROOT, "oracle", "_ref", "render_b200")


def read_pgm(path):
    data = open(path, "rb").read()
    header, _, rest = data.partition(b"\n")
    magic, w, h, maxv = header.split()
    assert magic == b"P5" and maxv == b"255"
    return np.frombuffer(rest, np.uint8).reshape(int(h), int(w))


@pytest.mark.gpu
def test_reference_cli_writes_the_reference_pgm(tmp_path, scene_mod, soup_golden, ao_golden):
    require_gpu()
    if not os.path.exists(CLI):
        pytest.skip("oracle/_ref/render_b200 not built (reference tree absent at build time)")
    g = soup_golden
    off, pgm = str(tmp_path / "soup.off"), str(tmp_path / "out.pgm")
    scene_mod.write_off(off, g["verts"], g["faces"])
    cmd = [CLI, "-a", "0", "-w", str(int(g["width"])), "-h", str(int(g["height"])), "-s", str(int(g["nss"])),
           "-f", "1.2345678", off, pgm]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "Rendering image" in res.stdout and "Using Device" in res.stdout
    assert np.array_equal(read_pgm(pgm), g["u8"])
    # the CLI's defaults: ambient occlusion on (uniform, 3 rings, 0.2); and `-a`, `-d`.  (`-m` and `-r` cannot be
    # exercised: the reference's own args.h:216-233 returns a reference to a dead temporary from map() and the
    # unmodified parser segfaults on either option under gcc 13 -O2, before any device code runs.)
    a = ao_golden
    size = ["-w", str(int(a["width"])), "-h", str(int(a["height"])), "-s", str(int(a["nss"]))]
    for extra, name in (([], "uniform3"), (["-a", "2", "-d", "0.7"], "uniform2_d07")):
        res = subprocess.run([CLI] + size + extra + [off, pgm], capture_output=True, text=True, timeout=120)
        assert res.returncode == 0, res.stdout + res.stderr
        assert np.array_equal(read_pgm(pgm), a["u8_" + name]), name


@pytest.mark.gpu
def test_reference_cli_on_the_bunny_with_default_options(tmp_path, scene_mod, po, bunny_scene):
    """Config C1 exactly as a user types it: `./render bunny.off out.pgm` -- 600x600, s=4, shading, uniform ambient
    occlusion with 3 rings -- through the unmodified render.cc / mesh.cc / bvh.cc and the shadow opencl_host.h.
    The OFF text is rewritten from the staged reference mesh (9 significant digits round-trip float32)."""
    require_gpu()
    if not os.path.exists(CLI):
        pytest.skip("oracle/_ref/render_b200 not built (reference tree absent at build time)")
    v, f = po.read_mesh_bin(po.staged_bunny_path())
    off, pgm = str(tmp_path / "bunny.off"), str(tmp_path / "bunny.pgm")
    scene_mod.write_off(off, v, f)
    res = subprocess.run([CLI, off, pgm], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    ref = po.render(bunny_scene, 1200, 1200, 1.0, True, ao=po.Ao.make())
    assert np.array_equal(read_pgm(pgm), po.resize(ref.image, 600, 600, 2))
    res = subprocess.run([CLI, "-a", "0", off, pgm], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    ref = po.render(bunny_scene, 1200, 1200, 1.0, True)
    assert np.array_equal(read_pgm(pgm), po.resize(ref.image, 600, 600, 2))


@pytest.mark.gpu
def test_python_cli_mirror_writes_the_same_pgm(tmp_path, scene_mod, soup_golden, ao_golden):
    """`python -m opencl_raytracer_b200.render`: the reference's options and defaults for hosts without the reference
    tree; same bytes as the golden PGM payloads of the reference kernel text, with the tree from the host builder and
    from the device builder; `-m random` works here (the reference's own parser crashes on it)."""
    require_gpu()
    import sys
    g, a = soup_golden, ao_golden
    off, pgm = str(tmp_path / "soup.off"), str(tmp_path / "out.pgm")
    scene_mod.write_off(off, g["verts"], g["faces"])
    base = [sys.executable, "-m", "opencl_raytracer_b200.render"]
    size = ["-w", str(int(a["width"])), "-h", str(int(a["height"])), "-s", str(int(a["nss"]))]
    cases = [(["-a", "0", "-f", "1.2345678", "-w", str(int(g["width"])), "-h", str(int(g["height"])), "-s", str(int(g["nss"]))], g["u8"]),
             (size, a["u8_uniform3"]), (size + ["-m", "random"], a["u8_random3"]),
             (size + ["-m", "random", "-a", "1", "-d", "1.5", "--device-build"], a["u8_random1_far"]),
             (size + ["-a", "2", "-d", "0.7", "--device-build"], a["u8_uniform2_d07"])]
    for extra, want in cases:
        res = subprocess.run(base + extra + [off, pgm], capture_output=True, text=True, timeout=300, cwd=ROOT)
        assert res.returncode == 0, res.stdout + res.stderr
        assert np.array_equal(read_pgm(pgm), want), extra


@pytest.mark.gpu
def test_python_cli_mirror_sah_strategy(tmp_path, scene_mod, sah_golden):
    """`-r sah`: the reference's own parser crashes on the option (see above); the mirror builds the reference's SAH tree
    on the host (rtx_scene_from_off_method) and must write the PGM the reference kernel text renders from that tree."""
    require_gpu()
    import sys
    from opencl_raytracer_b200 import scenes
    g = sah_golden
    v, f = scenes.random_soup(300, seed=11)
    off, pgm = str(tmp_path / "soup.off"), str(tmp_path / "out.pgm")
    scene_mod.write_off(off, v, f)
    cmd = [sys.executable, "-m", "opencl_raytracer_b200.render", "-a", "0", "-r", "sah", "-w", str(int(g["width"])), "-h", str(int(g["height"])),
           "-s", str(int(g["nss"])), "--scene-cache", str(tmp_path / "cache"), off, pgm]
    for _ in range(2):                                   # second run: the tree comes from the on-disk cache
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
        assert res.returncode == 0, res.stdout + res.stderr
        assert np.array_equal(read_pgm(pgm), g["u8"])


def test_reference_cli_fails_loudly_without_a_device(tmp_path, scene_mod, soup_golden):
    from opencl_raytracer_b200 import host
    if host.device_count() > 0:
        pytest.skip("a CUDA device is present")
    if not os.path.exists(CLI):
        pytest.skip("oracle/_ref/render_b200 not built")
    off, pgm = str(tmp_path / "soup.off"), str(tmp_path / "out.pgm")
    scene_mod.write_off(off, soup_golden["verts"][:30], np.arange(30, dtype=np.uint32).reshape(-1, 3))
    res = subprocess.run([CLI, "-a", "0", "-w", "16", "-h", "16", off, pgm], capture_output=True, text=True, timeout=120)
    assert res.returncode != 0 and not os.path.exists(pgm)
    assert "No device found" in (res.stdout + res.stderr)
