"""ctypes front end of the CPU oracles.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module (see oracle/rt_oracle.h).

Two checkers live here:

* ``port``  -- oracle/liboracle.so, the plain-C restatement (rt_oracle.c);
* ``ref``   -- oracle/_ref/libref_oracle.so, the reference's own kernel text
  and host scene code compiled for the CPU (oracle/ref_glue.cc); present only
  where oracle/Makefile could see /root/reference (or where the prebuilt file
  travelled with the snapshot).

A *scene* is any object with the five upload arrays of render.cc:98 as numpy
arrays: ``faces`` (u32, 3 per triangle, leaf order), ``nodes`` (u32),
``aabbs`` (f32 [2*nodes,4]), ``vertices`` (f32 [nv,4]), ``normals`` (f32 [nv,4]).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NO_HIT = 0xFFFFFFFF


class _OrcScene(C.Structure):
    _fields_ = [
        ("faces", C.c_void_p), ("nfaceidx", C.c_size_t),
        ("nodes", C.c_void_p), ("nnodes", C.c_size_t),
        ("aabbs", C.c_void_p),
        ("vertices", C.c_void_p), ("nverts", C.c_size_t),
        ("normals", C.c_void_p),
    ]


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("rays", "node_visits", "box_hits", "tri_tests", "tri_hits", "hit_rays", "max_visits")]

    def as_dict(self):
        d = {n: int(getattr(self, n)) for n, _ in self._fields_}
        r = max(d["rays"], 1)
        d["V"] = d["node_visits"] / r
        d["T"] = d["tri_tests"] / r
        d["h"] = d["hit_rays"] / r
        return d


class Ao(C.Structure):
    """Ambient-occlusion options (orc_ao): method 0 = uniform rings, 1 = random hemisphere."""
    _fields_ = [("enable", C.c_int), ("method", C.c_int), ("samples", C.c_uint32), ("max_distance", C.c_float),
                ("alpha_min", C.c_int), ("alpha_max", C.c_int)]

    @classmethod
    def make(cls, method=0, samples=3, max_distance=0.2, alpha_min=4, alpha_max=90):
        return cls(1, method, samples, max_distance, alpha_min, alpha_max)


def algorithmic_bytes_per_ray(V: float, T: float, h: float, extra: float = 0.0) -> float:
    """B = 32 V + 48 T + 48 h + 4 (+extra)."""
    return 32.0 * V + 48.0 * T + 48.0 * h + 4.0 + extra


def build(quiet: bool = True) -> None:
    """Compile the checkers (make is incremental)."""
    subprocess.run(["make", "-C", HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


_port = None
_ref = None


def _keep(scene):
    """C view of a scene + the arrays that must stay alive."""
    arrs = dict(
        faces=np.ascontiguousarray(scene.faces, dtype=np.uint32),
        nodes=np.ascontiguousarray(scene.nodes, dtype=np.uint32),
        aabbs=np.ascontiguousarray(scene.aabbs, dtype=np.float32),
        vertices=np.ascontiguousarray(scene.vertices, dtype=np.float32),
        normals=np.ascontiguousarray(scene.normals, dtype=np.float32),
    )
    assert arrs["aabbs"].size == 8 * arrs["nodes"].size, "aabbs must hold 2 float4 per node"
    assert arrs["vertices"].size == arrs["normals"].size
    s = _OrcScene(arrs["faces"].ctypes.data, arrs["faces"].size,
                  arrs["nodes"].ctypes.data, arrs["nodes"].size,
                  arrs["aabbs"].ctypes.data,
                  arrs["vertices"].ctypes.data, arrs["vertices"].size // 4,
                  arrs["normals"].ctypes.data)
    return s, arrs


def port():
    global _port
    if _port is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        lib = C.CDLL(path)
        lib.orc_focal_roundtrip.restype = C.c_float
        lib.orc_focal_roundtrip.argtypes = [C.c_float]
        lib.orc_render.restype = C.c_int
        lib.orc_render.argtypes = [C.POINTER(_OrcScene), C.c_uint, C.c_uint, C.c_float, C.c_int, C.c_uint32,
                                   C.c_uint, C.c_uint, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.POINTER(Counters), C.c_int]
        lib.orc_render_ao.restype = C.c_int
        lib.orc_render_ao.argtypes = [C.POINTER(_OrcScene), C.c_uint, C.c_uint, C.c_float, C.c_int, C.c_uint32, C.POINTER(Ao),
                                      C.c_uint, C.c_uint, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.POINTER(Counters), C.c_int]
        lib.orc_trace_rays.restype = C.c_int
        lib.orc_trace_rays.argtypes = [C.POINTER(_OrcScene), C.c_void_p, C.c_void_p, C.c_size_t, C.c_float,
                                       C.c_void_p, C.c_void_p, C.POINTER(Counters), C.c_int]
        lib.orc_gen_random_rays.restype = None
        lib.orc_gen_random_rays.argtypes = [C.c_uint32, C.c_uint64, C.c_size_t, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p]
        lib.orc_resize.restype = None
        lib.orc_resize.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_void_p]
        lib.orc_jitter.restype = None
        lib.orc_jitter.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        lib.orc_aabb_intersect.restype = C.c_int
        lib.orc_aabb_intersect.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float]
        lib.orc_online_cpus.restype = C.c_int
        _port = lib
    return _port


def ref():
    """The reference-compiled checker, or None when oracle/_ref/ is absent."""
    global _ref
    if _ref is None:
        path = os.path.join(HERE, "_ref", "libref_oracle.so")
        if not os.path.exists(path):
            return None
        lib = C.CDLL(path)
        lib.ref_scene_from_off.restype = C.c_void_p
        lib.ref_scene_from_off.argtypes = [C.c_char_p, C.c_int]
        lib.ref_scene_from_mesh.restype = C.c_void_p
        lib.ref_scene_from_mesh.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
        lib.ref_scene_free.argtypes = [C.c_void_p]
        lib.ref_scene_counts.argtypes = [C.c_void_p, C.POINTER(C.c_size_t * 5)]
        for n in ("faces", "orig_faces", "triangles", "nodes", "aabbs", "vertices", "normals"):
            f = getattr(lib, "ref_scene_" + n)
            f.restype = C.c_void_p
            f.argtypes = [C.c_void_p]
        lib.ref_render.restype = C.c_int
        lib.ref_render.argtypes = [C.POINTER(_OrcScene), C.c_uint, C.c_uint, C.c_float, C.c_int,
                                   C.c_uint, C.c_uint, C.c_uint, C.c_void_p, C.c_int]
        lib.ref_render_ao.restype = C.c_int
        lib.ref_render_ao.argtypes = [C.POINTER(_OrcScene), C.c_uint, C.c_uint, C.c_float, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int,
                                      C.c_uint, C.c_uint, C.c_uint, C.c_void_p, C.c_int]
        lib.ref_trace_rays.restype = C.c_int
        lib.ref_trace_rays.argtypes = [C.POINTER(_OrcScene), C.c_void_p, C.c_void_p, C.c_size_t, C.c_float,
                                       C.c_void_p, C.c_void_p, C.c_int]
        lib.ref_primary_hits.restype = C.c_int
        lib.ref_primary_hits.argtypes = [C.POINTER(_OrcScene), C.c_uint, C.c_uint, C.c_float,
                                         C.c_void_p, C.c_void_p, C.c_int]
        lib.ref_resize.restype = None
        lib.ref_resize.argtypes = [C.c_uint, C.c_uint, C.c_uint, C.c_void_p, C.c_void_p]
        lib.ref_total_dims.restype = None
        lib.ref_total_dims.argtypes = [C.c_uint, C.c_uint, C.c_uint, C.POINTER(C.c_uint), C.POINTER(C.c_uint)]
        lib.ref_focal_roundtrip.restype = C.c_float
        lib.ref_focal_roundtrip.argtypes = [C.c_float]
        lib.ref_online_cpus.restype = C.c_int
        _ref = lib
    return _ref


# ------------------------------------------------------------------ port ---

def focal_roundtrip(f: float) -> float:
    return float(port().orc_focal_roundtrip(C.c_float(f)))


def render(scene, width: int, height: int, focal: float = 1.0, shading: bool = True, jitter_seed: int = 0,
           rows=None, want_ids: bool = True, want_counters: bool = False, nthreads: int = 0, ao: "Ao | None" = None):
    """intersect_kernel.cl:278-310 over width x height (super-sampled dims).

    rows = (begin, end, step) restricts the rows rendered.  Returns a
    namespace with image (f32 [H,W]), face_id (u32), distance (f32),
    counters (dict or None)."""
    s, keep = _keep(scene)
    b, e, st = rows if rows is not None else (0, height, 1)
    image = np.zeros((height, width), np.float32)
    fid = np.full((height, width), NO_HIT, np.uint32) if want_ids else None
    dist = np.full((height, width), np.inf, np.float32) if want_ids else None
    cnt = Counters() if want_counters else None
    rc = port().orc_render_ao(C.byref(s), width, height, C.c_float(focal), int(bool(shading)), jitter_seed,
                           C.byref(ao) if ao is not None else None,
                           b, e, st, image.ctypes.data,
                           fid.ctypes.data if want_ids else None, dist.ctypes.data if want_ids else None,
                           C.byref(cnt) if cnt is not None else None, nthreads)
    if rc != 0:
        raise RuntimeError("orc_render failed")
    del keep
    return SimpleNamespace(image=image, face_id=fid, distance=dist,
                           counters=cnt.as_dict() if cnt is not None else None)


def trace_rays(scene, origins, dirs, max_distance: float = 100000.0, want_counters: bool = False, nthreads: int = 0):
    s, keep = _keep(scene)
    origins = np.ascontiguousarray(origins, np.float32).reshape(-1, 4)
    dirs = np.ascontiguousarray(dirs, np.float32).reshape(-1, 4)
    n = origins.shape[0]
    fid = np.full(n, NO_HIT, np.uint32)
    dist = np.full(n, np.inf, np.float32)
    cnt = Counters() if want_counters else None
    rc = port().orc_trace_rays(C.byref(s), origins.ctypes.data, dirs.ctypes.data, n, C.c_float(max_distance),
                               fid.ctypes.data, dist.ctypes.data, C.byref(cnt) if cnt is not None else None, nthreads)
    if rc != 0:
        raise RuntimeError("orc_trace_rays failed")
    del keep
    return SimpleNamespace(face_id=fid, distance=dist, counters=cnt.as_dict() if cnt is not None else None)


def gen_random_rays(seed: int, first: int, n: int, bbmin, bbmax):
    bbmin = np.ascontiguousarray(bbmin, np.float32)
    bbmax = np.ascontiguousarray(bbmax, np.float32)
    o = np.zeros((n, 4), np.float32)
    d = np.zeros((n, 4), np.float32)
    port().orc_gen_random_rays(seed, first, n, bbmin.ctypes.data, bbmax.ctypes.data, o.ctypes.data, d.ctypes.data)
    return o, d


def jitter(seed: int, x: int, y: int):
    jx, jy = C.c_float(), C.c_float()
    port().orc_jitter(seed, x, y, C.byref(jx), C.byref(jy))
    return jx.value, jy.value


def resize(tmp, width: int, height: int, n: int):
    """src/ray_tracer.cc:3-15 on a [height*n, width*n] float image."""
    tmp = np.ascontiguousarray(tmp, np.float32)
    assert tmp.shape == (height * n, width * n)
    out = np.zeros((height, width), np.uint8)
    port().orc_resize(tmp.ctypes.data, width * n, width, height, n, out.ctypes.data)
    return out


def aabb_intersect(bb8, pos4, dir4, max_distance: float) -> bool:
    bb8 = np.ascontiguousarray(bb8, np.float32)
    pos4 = np.ascontiguousarray(pos4, np.float32)
    dir4 = np.ascontiguousarray(dir4, np.float32)
    return bool(port().orc_aabb_intersect(bb8.ctypes.data, pos4.ctypes.data, dir4.ctypes.data, C.c_float(max_distance)))


# ------------------------------------------------------------------- ref ---

def _wrap_ref_scene(lib, h):
    cnt = (C.c_size_t * 5)()
    lib.ref_scene_counts(h, C.byref(cnt))
    nfi, nn, nab, nv, nnm = [int(c) for c in cnt]

    def arr(name, ctype, count, shape):
        p = getattr(lib, "ref_scene_" + name)(h)
        a = np.ctypeslib.as_array(C.cast(p, C.POINTER(ctype)), shape=(count,)).copy()
        return a.reshape(shape)

    sc = SimpleNamespace(
        faces=arr("faces", C.c_uint32, nfi, (-1,)),
        orig_faces=arr("orig_faces", C.c_uint32, nfi, (-1,)),
        triangles=arr("triangles", C.c_uint32, nfi // 3, (-1,)),
        nodes=arr("nodes", C.c_uint32, nn, (-1,)),
        aabbs=arr("aabbs", C.c_float, nab * 4, (-1, 4)),
        vertices=arr("vertices", C.c_float, nv * 4, (-1, 4)),
        normals=arr("normals", C.c_float, nnm * 4, (-1, 4)),
    )
    lib.ref_scene_free(h)
    return sc


def ref_scene_from_off(path: str, sah: bool = False):
    lib = ref()
    h = lib.ref_scene_from_off(path.encode(), int(sah))
    if not h:
        raise RuntimeError("reference loader rejected %s" % path)
    return _wrap_ref_scene(lib, h)


def ref_scene_from_mesh(verts3, faces, sah: bool = False):
    lib = ref()
    verts3 = np.ascontiguousarray(verts3, np.float32).reshape(-1, 3)
    faces = np.ascontiguousarray(faces, np.uint32).reshape(-1, 3)
    h = lib.ref_scene_from_mesh(verts3.ctypes.data, verts3.shape[0], faces.ctypes.data, faces.shape[0], int(sah))
    if not h:
        raise RuntimeError("reference builder rejected the mesh")
    return _wrap_ref_scene(lib, h)


def ref_render(scene, width: int, height: int, focal: float = 1.0, shading: bool = True, rows=None, nthreads: int = 0):
    s, keep = _keep(scene)
    b, e, st = rows if rows is not None else (0, height, 1)
    image = np.zeros((height, width), np.float32)
    rc = ref().ref_render(C.byref(s), width, height, C.c_float(focal), int(bool(shading)), b, e, st,
                          image.ctypes.data, nthreads)
    if rc != 0:
        raise RuntimeError("ref_render failed")
    del keep
    return image


def ref_render_ao(scene, width: int, height: int, ao: Ao, focal: float = 1.0, rows=None, nthreads: int = 0):
    """The reference kernel text with AO_ENABLE (shading on); (method, samples) must be one of the compiled copies."""
    s, keep = _keep(scene)
    b, e, st = rows if rows is not None else (0, height, 1)
    image = np.zeros((height, width), np.float32)
    rc = ref().ref_render_ao(C.byref(s), width, height, C.c_float(focal), ao.method, ao.samples, C.c_float(ao.max_distance),
                             ao.alpha_min, ao.alpha_max, b, e, st, image.ctypes.data, nthreads)
    if rc != 0:
        raise RuntimeError("ref_render_ao failed (%d)" % rc)
    del keep
    return image


def ref_primary_hits(scene, width: int, height: int, focal: float = 1.0, nthreads: int = 0):
    s, keep = _keep(scene)
    fid = np.full((height, width), NO_HIT, np.uint32)
    dist = np.full((height, width), np.inf, np.float32)
    rc = ref().ref_primary_hits(C.byref(s), width, height, C.c_float(focal), fid.ctypes.data, dist.ctypes.data, nthreads)
    if rc != 0:
        raise RuntimeError("ref_primary_hits failed")
    del keep
    return fid, dist


def ref_trace_rays(scene, origins, dirs, max_distance: float = 100000.0, nthreads: int = 0):
    s, keep = _keep(scene)
    origins = np.ascontiguousarray(origins, np.float32).reshape(-1, 4)
    dirs = np.ascontiguousarray(dirs, np.float32).reshape(-1, 4)
    n = origins.shape[0]
    fid = np.full(n, NO_HIT, np.uint32)
    dist = np.full(n, np.inf, np.float32)
    rc = ref().ref_trace_rays(C.byref(s), origins.ctypes.data, dirs.ctypes.data, n, C.c_float(max_distance),
                              fid.ctypes.data, dist.ctypes.data, nthreads)
    if rc != 0:
        raise RuntimeError("ref_trace_rays failed")
    del keep
    return fid, dist


def ref_resize(tmp, width: int, height: int, n_super_samples: int):
    tmp = np.ascontiguousarray(tmp, np.float32)
    out = np.zeros((height, width), np.uint8)
    ref().ref_resize(width, height, n_super_samples, tmp.ctypes.data, out.ctypes.data)
    return out


def ref_total_dims(width: int, height: int, n_super_samples: int):
    tw, th = C.c_uint(), C.c_uint()
    ref().ref_total_dims(width, height, n_super_samples, C.byref(tw), C.byref(th))
    return tw.value, th.value


def ref_focal_roundtrip(f: float) -> float:
    return float(ref().ref_focal_roundtrip(C.c_float(f)))


# ------------------------------------------------------- staged ref mesh ---

MESH_MAGIC = b"RTXMESH1"


def write_mesh_bin(path: str, verts3, faces) -> None:
    verts3 = np.ascontiguousarray(verts3, np.float32).reshape(-1, 3)
    faces = np.ascontiguousarray(faces, np.uint32).reshape(-1, 3)
    with open(path, "wb") as f:
        f.write(MESH_MAGIC)
        f.write(np.array([verts3.shape[0], faces.shape[0]], np.uint64).tobytes())
        f.write(verts3.tobytes())
        f.write(faces.tobytes())


def read_mesh_bin(path: str):
    with open(path, "rb") as f:
        if f.read(8) != MESH_MAGIC:
            raise ValueError("not a staged mesh: " + path)
        nv, nf = np.frombuffer(f.read(16), np.uint64)
        verts = np.frombuffer(f.read(int(nv) * 12), np.float32).reshape(-1, 3).copy()
        faces = np.frombuffer(f.read(int(nf) * 12), np.uint32).reshape(-1, 3).copy()
    return verts, faces


def staged_bunny_path() -> str:
    return os.path.join(HERE, "_ref", "bunny_mesh.bin")
