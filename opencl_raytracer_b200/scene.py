"""Scene containers and host-side scene preparation (ctypes over librtx_scene.so).

A ``Scene`` holds the five arrays the reference's ``render.cc:86-98`` hands to
``OpenCLHost::upload`` -- leaf-ordered faces, pre-order BVH ``nodes``, ``aabbs``
as (min,max) float4 pairs, float4 ``vertices`` and ``normals`` -- produced by
``include/rtx_scene.h`` (bit-identical to the reference's mesh.cc + bvh.cc
with the longest-axis split; checked in tests/test_scene_prep.py).
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
from dataclasses import dataclass

import numpy as np

_LIBDIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib")
_lib = None


class SceneError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("rtx_scene error %d: %s" % (code, msg))
        self.code = code


def _load():
    global _lib
    if _lib is None:
        path = os.path.join(_LIBDIR, "librtx_scene.so")
        if not os.path.exists(path):
            raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'`" % path)
        lib = C.CDLL(path)
        lib.rtx_scene_from_off.restype = C.c_int
        lib.rtx_scene_from_off.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
        lib.rtx_scene_from_mesh.restype = C.c_int
        lib.rtx_scene_from_mesh.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]
        lib.rtx_scene_free.restype = None
        lib.rtx_scene_free.argtypes = [C.c_void_p]
        lib.rtx_scene_counts.restype = None
        lib.rtx_scene_counts.argtypes = [C.c_void_p, C.POINTER(C.c_size_t * 5)]
        for n in ("faces", "triangles", "orig_faces", "nodes", "aabbs", "vertices", "normals"):
            f = getattr(lib, "rtx_scene_" + n)
            f.restype = C.c_void_p
            f.argtypes = [C.c_void_p]
        lib.rtx_scene_last_error.restype = C.c_char_p
        _lib = lib
    return _lib


@dataclass
class Scene:
    faces: np.ndarray       # u32 [3*T]   leaf-ordered vertex ids        (render.cc:88-95)
    nodes: np.ndarray       # u32 [2T-1]  pre-order subtree sizes        (bvh.cc:115-162)
    aabbs: np.ndarray       # f32 [2*(2T-1), 4]  (min,max) per node
    vertices: np.ndarray    # f32 [V, 4]
    normals: np.ndarray     # f32 [V, 4]
    triangles: np.ndarray | None = None   # u32 [T] leaf index -> input face id (BVH::triangles)
    orig_faces: np.ndarray | None = None  # u32 [3*T] input order
    name: str = ""

    @property
    def num_triangles(self) -> int:
        return self.faces.size // 3

    @property
    def num_nodes(self) -> int:
        return self.nodes.size

    def root_box(self):
        return self.aabbs[0, :3].copy(), self.aabbs[1, :3].copy()

    def upload_bytes(self) -> int:
        """Bytes OpenCLHost::upload moves host->device (opencl_host.cc:120-136)."""
        return int(self.faces.nbytes + self.nodes.nbytes + self.aabbs.nbytes + self.vertices.nbytes + self.normals.nbytes)

    def digest(self) -> str:
        """sha256 over the five upload arrays (golden fixtures pin this)."""
        h = hashlib.sha256()
        for a in (self.faces, self.nodes, self.aabbs, self.vertices, self.normals):
            h.update(np.ascontiguousarray(a).tobytes())
        return h.hexdigest()


def _wrap(lib, handle, name: str) -> Scene:
    cnt = (C.c_size_t * 5)()
    lib.rtx_scene_counts(handle, C.byref(cnt))
    nfi, nn, nab, nv, nnm = [int(c) for c in cnt]

    def arr(fn, ctype, count, shape):
        p = getattr(lib, "rtx_scene_" + fn)(handle)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(ctype)), shape=(count,)).copy().reshape(shape)

    try:
        return Scene(
            faces=arr("faces", C.c_uint32, nfi, (-1,)),
            nodes=arr("nodes", C.c_uint32, nn, (-1,)),
            aabbs=arr("aabbs", C.c_float, nab * 4, (-1, 4)),
            vertices=arr("vertices", C.c_float, nv * 4, (-1, 4)),
            normals=arr("normals", C.c_float, nnm * 4, (-1, 4)),
            triangles=arr("triangles", C.c_uint32, nfi // 3, (-1,)),
            orig_faces=arr("orig_faces", C.c_uint32, nfi, (-1,)),
            name=name,
        )
    finally:
        lib.rtx_scene_free(handle)


def scene_from_off(path: str, nthreads: int = 0) -> Scene:
    """load_off_mesh + compute_vertex_normals + BVH::buildBVH + face sort."""
    lib = _load()
    h = C.c_void_p()
    rc = lib.rtx_scene_from_off(os.fsencode(path), nthreads, C.byref(h))
    if rc != 0:
        raise SceneError(rc, lib.rtx_scene_last_error().decode())
    return _wrap(lib, h, os.path.basename(path))


def scene_from_mesh(verts3, faces, nthreads: int = 0, name: str = "") -> Scene:
    lib = _load()
    verts3 = np.ascontiguousarray(verts3, np.float32).reshape(-1, 3)
    faces = np.ascontiguousarray(faces, np.uint32).reshape(-1, 3)
    h = C.c_void_p()
    rc = lib.rtx_scene_from_mesh(verts3.ctypes.data, verts3.shape[0], faces.ctypes.data, faces.shape[0], nthreads, C.byref(h))
    if rc != 0:
        raise SceneError(rc, lib.rtx_scene_last_error().decode())
    return _wrap(lib, h, name)


def write_off(path: str, verts3, faces) -> None:
    """Write a triangle mesh as OFF text (9 significant digits: float32 round-trips)."""
    verts3 = np.asarray(verts3, np.float32).reshape(-1, 3)
    faces = np.asarray(faces, np.uint32).reshape(-1, 3)
    with open(path, "w") as f:
        f.write("OFF\n%d %d 0\n" % (verts3.shape[0], faces.shape[0]))
        for v in verts3:
            f.write("%.9g %.9g %.9g\n" % (v[0], v[1], v[2]))
        for t in faces:
            f.write("3 %d %d %d\n" % (t[0], t[1], t[2]))
