import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the in-tree libraries once (incremental; a no-op when they are current)."""
    import __graft_entry__ as g
    g.build(quiet=True)


@pytest.fixture(scope="session")
def po():
    from oracle import pyoracle
    return pyoracle


@pytest.fixture(scope="session")
def golden_meta():
    with open(os.path.join(GOLDEN, "scenes.json")) as f:
        return json.load(f)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


@pytest.fixture(scope="session")
def soup_golden():
    return load_golden("soup_64x48.npz")


@pytest.fixture(scope="session")
def rays_golden():
    return load_golden("soup_rays.npz")


@pytest.fixture(scope="session")
def special_rays_golden():
    return load_golden("soup_special_rays.npz")


@pytest.fixture(scope="session")
def quad_golden():
    return load_golden("quad_33x17.npz")


@pytest.fixture(scope="session")
def scene_mod():
    from opencl_raytracer_b200 import scene
    return scene


@pytest.fixture(scope="session")
def soup_scene(scene_mod, soup_golden):
    return scene_mod.scene_from_mesh(soup_golden["verts"], soup_golden["faces"], name="soup300")


@pytest.fixture(scope="session")
def sibenik_scene(scene_mod):
    from opencl_raytracer_b200 import scenes
    v, f = scenes.sibenik_standin()
    return scene_mod.scene_from_mesh(v, f, name="sibenik_standin")


@pytest.fixture(scope="session")
def bunny_scene(scene_mod, po):
    path = po.staged_bunny_path()
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/bunny_mesh.bin not staged (reference tree absent at build time)")
    v, f = po.read_mesh_bin(path)
    return scene_mod.scene_from_mesh(v, f, name="bunny")


def require_gpu():
    from opencl_raytracer_b200 import host
    if host.device_count() == 0:
        pytest.fail("test marked gpu but no CUDA device is visible")
    return host


@pytest.fixture(scope="session")
def ao_golden():
    return load_golden("soup_ao.npz")


AO_CASES = ("uniform3", "random3", "random1_far", "uniform2_a10_60", "uniform2_d07")


@pytest.fixture(scope="session")
def sah_golden():
    return load_golden("soup_sah.npz")


@pytest.fixture(scope="session")
def sah_scene(scene_mod, sah_golden):
    g = sah_golden
    return scene_mod.Scene(g["t_faces"], g["t_nodes"], g["t_aabbs"], g["t_vertices"], g["t_normals"])
