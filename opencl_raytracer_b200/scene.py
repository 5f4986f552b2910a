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
import zipfile
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
        lib.rtx_scene_from_off_method.restype = C.c_int
        lib.rtx_scene_from_off_method.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        lib.rtx_scene_from_mesh_method.restype = C.c_int
        lib.rtx_scene_from_mesh_method.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
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


def scene_from_off(path: str, nthreads: int = 0, sah: bool = False) -> Scene:
    """load_off_mesh + compute_vertex_normals + BVH::buildBVH + face sort.  sah: the reference's `-r sah` builder."""
    lib = _load()
    h = C.c_void_p()
    rc = lib.rtx_scene_from_off_method(os.fsencode(path), 1 if sah else 0, nthreads, C.byref(h))
    if rc != 0:
        raise SceneError(rc, lib.rtx_scene_last_error().decode())
    return _wrap(lib, h, os.path.basename(path))


def scene_from_mesh(verts3, faces, nthreads: int = 0, name: str = "", sah: bool = False) -> Scene:
    lib = _load()
    verts3 = np.ascontiguousarray(verts3, np.float32).reshape(-1, 3)
    faces = np.ascontiguousarray(faces, np.uint32).reshape(-1, 3)
    h = C.c_void_p()
    rc = lib.rtx_scene_from_mesh_method(verts3.ctypes.data, verts3.shape[0], faces.ctypes.data, faces.shape[0], 1 if sah else 0, nthreads, C.byref(h))
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


# ---------------------------------------------------------------------------------------------------------------------
# On-disk cache of a prepared scene: the five upload arrays of render.cc:88-98 + the two index maps, keyed by
# the input mesh.  The builder is deterministic, so a cache hit returns byte-identical arrays (the file carries their
# sha256 and load_scene() refuses a file whose contents do not match it).
# ---------------------------------------------------------------------------------------------------------------------
CACHE_FORMAT = 1


def save_scene(path: str, sc: Scene) -> None:
    """Write a prepared scene as one .npz (written next to `path` first, then renamed: no half-written cache entries)."""
    tmp = "%s.tmp.%d.npz" % (path, os.getpid())
    np.savez(tmp, format=np.array([CACHE_FORMAT], np.uint32), digest=np.frombuffer(bytes.fromhex(sc.digest()), np.uint8),
             faces=sc.faces, nodes=sc.nodes, aabbs=sc.aabbs, vertices=sc.vertices, normals=sc.normals,
             triangles=sc.triangles if sc.triangles is not None else np.zeros(0, np.uint32),
             orig_faces=sc.orig_faces if sc.orig_faces is not None else np.zeros(0, np.uint32),
             name=np.frombuffer(sc.name.encode(), np.uint8))
    os.replace(tmp, path)


def load_scene(path: str) -> Scene:
    """Read a scene written by save_scene; raises SceneError if the file is of another format or fails its checksum."""
    try:
        with np.load(path) as z:
            if int(z["format"][0]) != CACHE_FORMAT:
                raise SceneError(-1, "%s: cache format %d, this build reads %d" % (path, int(z["format"][0]), CACHE_FORMAT))
            sc = Scene(faces=z["faces"], nodes=z["nodes"], aabbs=z["aabbs"], vertices=z["vertices"], normals=z["normals"],
                       triangles=z["triangles"] if z["triangles"].size else None,
                       orig_faces=z["orig_faces"] if z["orig_faces"].size else None, name=bytes(z["name"]).decode())
            want = bytes(z["digest"]).hex()
    except (OSError, KeyError, ValueError, EOFError, zipfile.BadZipFile) as e:
        raise SceneError(-1, "%s: not a scene cache file (%s)" % (path, e))
    if sc.digest() != want:
        raise SceneError(-1, "%s: contents do not match the stored sha256" % path)
    return sc


def mesh_key(verts3, faces) -> str:
    """Cache key of an input mesh: sha256 over its float32 vertices and uint32 faces (+ the cache format)."""
    h = hashlib.sha256(b"rtx-scene-%d" % CACHE_FORMAT)
    h.update(np.ascontiguousarray(verts3, np.float32).tobytes())
    h.update(np.ascontiguousarray(faces, np.uint32).tobytes())
    return h.hexdigest()[:32]


def cached_scene_from_mesh(verts3, faces, cache_dir: str, nthreads: int = 0, name: str = "", sah: bool = False) -> Scene:
    """scene_from_mesh through an on-disk cache: the 10 M-triangle scene of C4 loads in a fraction of its 2 s build.
    A missing, unreadable or corrupt entry is rebuilt and rewritten."""
    os.makedirs(cache_dir, exist_ok=True)
    path = os.path.join(cache_dir, "scene_%s%s.npz" % (mesh_key(verts3, faces), "_sah" if sah else ""))
    if os.path.exists(path):
        try:
            sc = load_scene(path)
            sc.name = name or sc.name
            return sc
        except SceneError:
            pass
    sc = scene_from_mesh(verts3, faces, nthreads=nthreads, name=name, sah=sah)
    save_scene(path, sc)
    return sc


def cached_scene_from_off(path: str, cache_dir: str, nthreads: int = 0, sah: bool = False) -> Scene:
    """scene_from_off through the same cache, keyed by the sha256 of the OFF file's bytes (and the builder)."""
    os.makedirs(cache_dir, exist_ok=True)
    h = hashlib.sha256(b"rtx-scene-off-%d-%d" % (CACHE_FORMAT, 1 if sah else 0))
    with open(path, "rb") as fh:
        for chunk in iter(lambda: fh.read(1 << 20), b""):
            h.update(chunk)
    entry = os.path.join(cache_dir, "scene_%s.npz" % h.hexdigest()[:32])
    if os.path.exists(entry):
        try:
            return load_scene(entry)
        except SceneError:
            pass
    sc = scene_from_off(path, nthreads=nthreads, sah=sah)
    save_scene(entry, sc)
    return sc
