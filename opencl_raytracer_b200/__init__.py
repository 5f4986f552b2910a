"""opencl_raytracer_b200 -- B200-native closest-hit path for magcks/opencl_raytracer.

Only what the hot path needs: csrc/ (CUDA kernels + the C-ABI of
include/rtx_b200.h, host scene preparation of include/rtx_scene.h), the Python
mirror of the reference's host boundary (host.py), procedural scenes and the
multi-GPU tile partition.  See DESIGN.md.
"""
from .scene import Scene, SceneError, scene_from_mesh, scene_from_off, write_off  # noqa: F401

__all__ = ["Scene", "SceneError", "scene_from_mesh", "scene_from_off", "write_off"]
