"""Python mirror of the reference's host boundary, over the C ABI (ctypes).

``RayTracer`` mirrors include/ray_tracer.h:3-39 (Options, totalWidth /
totalHeight, resize) and ``CudaHost`` mirrors ``OpenCLHost``
(include/opencl_host.h:127-131): ``printInfo()``, ``upload(faces, nodes, aabbs,
vertices, vnormals)``, ``__call__()`` (the reference's ``bool operator()()``)
and ``download()``.  Every compute call goes through
opencl_raytracer_b200/lib/librtx_b200.so (include/rtx_b200.h); there is no
Python or CPU fallback -- a missing library or device raises.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass

import numpy as np

from .scene import Scene

_LIBDIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib")
LIB_PATH = os.environ.get("RTX_B200_LIB") or os.path.join(_LIBDIR, "librtx_b200.so")     # override: A/B builds of the kernels

NO_HIT = 0xFFFFFFFF

OK, ERR_ARG, ERR_NO_DEVICE, ERR_CUDA, ERR_UNSUPPORTED, ERR_STATE, ERR_NOMEM = range(7)

TUNE_KERNEL, TUNE_LEAF_SIZE, TUNE_RECORD_HITS, TUNE_COUNTERS, TUNE_TOP_SMEM, TUNE_BLOCKS_PER_SM, TUNE_FLATTEN_ON_DEVICE, TUNE_RAYS_PER_THREAD, TUNE_FRUSTUM, TUNE_LIST_RAYS_PER_THREAD, TUNE_INCOHERENT_KERNEL, TUNE_RAY_TABLES, TUNE_PHASE_TIMING = range(1, 14)
PHASES = ("tables", "collect_super", "collect", "traversal", "overflow", "ao")     # RTX_PHASE_* of include/rtx_b200.h
KERNEL_PERSISTENT, KERNEL_EXHAUSTIVE = 0, 1

# every symbol include/rtx_b200.h declares (tests check the library exports them all)
ABI_SYMBOLS = (
    "rtx_print_info", "rtx_device_count", "rtx_device_info", "rtx_create", "rtx_upload", "rtx_render",
    "rtx_download", "rtx_destroy", "rtx_last_error", "rtx_set_tunable", "rtx_get_stats", "rtx_render_async",
    "rtx_synchronize", "rtx_download_hits", "rtx_download_u8", "rtx_device_image", "rtx_trace_rays",
    "rtx_trace_rays_device", "rtx_trace_random_rays", "rtx_tile_layout", "rtx_deinterleave_async", "rtx_bind_output",
    "rtx_probe_bandwidth", "rtx_resize_u8_async", "rtx_deinterleave_u8_async", "rtx_render_download",
    "rtx_upload_mesh", "rtx_download_tree", "rtx_build_stats", "rtx_download_normals",
    "rtx_resize_u8_to_async", "rtx_store_tiles_async", "rtx_adopt_u8", "rtx_peer_alloc", "rtx_peer_open", "rtx_peer_close",
    "rtx_peer_free", "rtx_host_register", "rtx_host_unregister", "rtx_phase_ms", "rtx_copy_to_host", "rtx_bind_output_image", "rtx_render_store", "rtx_render_store_async", "rtx_debug_bounds", "rtx_peer_signal_async", "rtx_peer_wait_async",
)


class RtxError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("rtx error %d: %s" % (code, msg))
        self.code = code


class _Options(C.Structure):
    _fields_ = [
        ("width", C.c_uint32), ("height", C.c_uint32), ("focal_length", C.c_float), ("n_super_samples", C.c_uint32),
        ("enable_shading", C.c_int32), ("enable_ao", C.c_int32), ("ao_max_distance", C.c_float),
        ("ao_num_samples", C.c_uint32), ("ao_method", C.c_int32), ("ao_alpha_min", C.c_int32), ("ao_alpha_max", C.c_int32),
        ("bvh_method", C.c_int32), ("total_width", C.c_uint32), ("total_height", C.c_uint32),
        ("device", C.c_int32), ("jitter_seed", C.c_uint32), ("tile_rank", C.c_uint32), ("tile_world", C.c_uint32),
    ]


class DeviceInfo(C.Structure):
    _fields_ = [
        ("name", C.c_char * 256), ("cc_major", C.c_int32), ("cc_minor", C.c_int32), ("sm_count", C.c_int32),
        ("clock_khz", C.c_int32), ("mem_clock_khz", C.c_int32), ("mem_bus_bits", C.c_int32),
        ("global_mem_bytes", C.c_uint64), ("l2_bytes", C.c_uint64), ("smem_per_sm_bytes", C.c_uint64),
        ("smem_per_block_optin_bytes", C.c_uint64), ("max_threads_per_sm", C.c_int32), ("regs_per_sm", C.c_int32),
        ("driver_version", C.c_int32), ("runtime_version", C.c_int32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("rays", C.c_uint64), ("kernel_ms", C.c_double), ("kernel_launches", C.c_uint32), ("kernel_variant", C.c_uint32),
        ("node_visits", C.c_uint64), ("tri_tests", C.c_uint64), ("leafbox_tests", C.c_uint64),
        ("exact_path_rays", C.c_uint64), ("tree_depth", C.c_uint32), ("num_pairs", C.c_uint32),
        ("packet_overflows", C.c_uint64),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


_lib = None


def load_library():
    """dlopen the C-ABI library and declare its prototypes.  Raises ImportError when it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'`" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, u32p, fp = C.c_void_p, C.c_void_p, C.c_void_p
    lib.rtx_print_info.restype = C.c_int
    lib.rtx_device_count.restype = C.c_int
    lib.rtx_device_count.argtypes = [C.POINTER(C.c_int)]
    lib.rtx_device_info.restype = C.c_int
    lib.rtx_device_info.argtypes = [C.c_int, C.POINTER(DeviceInfo)]
    lib.rtx_create.restype = C.c_int
    lib.rtx_create.argtypes = [C.POINTER(vp), C.POINTER(_Options)]
    lib.rtx_upload.restype = C.c_int
    lib.rtx_upload.argtypes = [vp, u32p, C.c_size_t, u32p, C.c_size_t, fp, C.c_size_t, fp, C.c_size_t, fp, C.c_size_t]
    lib.rtx_upload_mesh.restype = C.c_int
    lib.rtx_upload_mesh.argtypes = [vp, fp, C.c_size_t, u32p, C.c_size_t, fp]
    lib.rtx_download_tree.restype = C.c_int
    lib.rtx_download_tree.argtypes = [vp, u32p, fp, u32p, u32p]
    lib.rtx_download_normals.restype = C.c_int
    lib.rtx_download_normals.argtypes = [vp, fp, C.c_size_t]
    lib.rtx_build_stats.restype = C.c_int
    lib.rtx_build_stats.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_uint32)]
    lib.rtx_render.restype = C.c_int
    lib.rtx_render.argtypes = [vp]
    lib.rtx_download.restype = C.c_int
    lib.rtx_download.argtypes = [vp, fp]
    lib.rtx_destroy.restype = None
    lib.rtx_destroy.argtypes = [vp]
    lib.rtx_last_error.restype = C.c_char_p
    lib.rtx_last_error.argtypes = [vp]
    lib.rtx_set_tunable.restype = C.c_int
    lib.rtx_set_tunable.argtypes = [vp, C.c_int, C.c_int64]
    lib.rtx_get_stats.restype = C.c_int
    lib.rtx_get_stats.argtypes = [vp, C.POINTER(Stats)]
    lib.rtx_render_async.restype = C.c_int
    lib.rtx_render_async.argtypes = [vp, vp]
    lib.rtx_synchronize.restype = C.c_int
    lib.rtx_synchronize.argtypes = [vp]
    lib.rtx_render_download.restype = C.c_int
    lib.rtx_render_download.argtypes = [vp, fp]
    lib.rtx_download_hits.restype = C.c_int
    lib.rtx_download_hits.argtypes = [vp, u32p, fp]
    lib.rtx_download_u8.restype = C.c_int
    lib.rtx_download_u8.argtypes = [vp, vp]
    lib.rtx_device_image.restype = C.c_int
    lib.rtx_device_image.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    lib.rtx_bind_output.restype = C.c_int
    lib.rtx_bind_output.argtypes = [vp, vp, C.c_size_t]
    lib.rtx_resize_u8_async.restype = C.c_int
    lib.rtx_resize_u8_async.argtypes = [vp, vp, C.c_size_t, vp]
    lib.rtx_deinterleave_u8_async.restype = C.c_int
    lib.rtx_deinterleave_u8_async.argtypes = [vp, vp, C.c_uint32, vp]
    lib.rtx_probe_bandwidth.restype = C.c_int
    lib.rtx_probe_bandwidth.argtypes = [vp, C.c_int, C.c_size_t, C.c_int, C.POINTER(C.c_double)]
    lib.rtx_trace_rays.restype = C.c_int
    lib.rtx_trace_rays.argtypes = [vp, fp, fp, C.c_size_t, C.c_float, u32p, fp]
    lib.rtx_trace_rays_device.restype = C.c_int
    lib.rtx_trace_rays_device.argtypes = [vp, vp, vp, C.c_size_t, C.c_float, vp, vp, vp]
    lib.rtx_trace_random_rays.restype = C.c_int
    lib.rtx_trace_random_rays.argtypes = [vp, C.c_uint32, C.c_uint64, C.c_size_t, C.c_float, u32p, fp,
                                          C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.rtx_tile_layout.restype = C.c_int
    lib.rtx_tile_layout.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    lib.rtx_deinterleave_async.restype = C.c_int
    lib.rtx_deinterleave_async.argtypes = [vp, vp, C.c_uint32, vp]
    lib.rtx_resize_u8_to_async.restype = C.c_int
    lib.rtx_resize_u8_to_async.argtypes = [vp, vp, vp]
    lib.rtx_store_tiles_async.restype = C.c_int
    lib.rtx_store_tiles_async.argtypes = [vp, vp, vp]
    lib.rtx_adopt_u8.restype = C.c_int
    lib.rtx_adopt_u8.argtypes = [vp, vp]
    lib.rtx_peer_alloc.restype = C.c_int
    lib.rtx_peer_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp), C.c_char_p]
    lib.rtx_peer_open.restype = C.c_int
    lib.rtx_peer_open.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    lib.rtx_peer_close.restype = C.c_int
    lib.rtx_peer_close.argtypes = [vp, vp]
    lib.rtx_peer_free.restype = C.c_int
    lib.rtx_peer_free.argtypes = [vp, vp]
    lib.rtx_host_register.restype = C.c_int
    lib.rtx_host_register.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    lib.rtx_host_unregister.restype = C.c_int
    lib.rtx_host_unregister.argtypes = [vp]
    lib.rtx_bind_output_image.restype = C.c_int
    lib.rtx_bind_output_image.argtypes = [vp, vp]
    lib.rtx_render_store.restype = C.c_int
    lib.rtx_render_store.argtypes = [vp, vp]
    lib.rtx_render_store_async.restype = C.c_int
    lib.rtx_render_store_async.argtypes = [vp, vp, vp]
    lib.rtx_debug_bounds.restype = C.c_int
    lib.rtx_debug_bounds.argtypes = [vp, C.POINTER(C.c_uint * 4)]
    lib.rtx_peer_signal_async.restype = C.c_int
    lib.rtx_peer_signal_async.argtypes = [vp, vp, C.c_uint32, vp]
    lib.rtx_peer_wait_async.restype = C.c_int
    lib.rtx_peer_wait_async.argtypes = [vp, vp, C.c_uint32, C.c_uint32, vp, vp]
    lib.rtx_copy_to_host.restype = C.c_int
    lib.rtx_copy_to_host.argtypes = [vp, vp, vp, C.c_size_t]
    lib.rtx_phase_ms.restype = C.c_int
    lib.rtx_phase_ms.argtypes = [vp, C.POINTER(C.c_double)]
    _lib = lib
    return lib


def _check(lib, ctx, rc):
    if rc != OK:
        msg = lib.rtx_last_error(ctx)
        raise RtxError(rc, msg.decode() if msg else "?")


def device_count() -> int:
    lib = load_library()
    n = C.c_int(0)
    rc = lib.rtx_device_count(C.byref(n))
    if rc not in (OK, ERR_NO_DEVICE):
        _check(lib, None, rc)
    return n.value


def device_info(device: int = 0) -> DeviceInfo:
    lib = load_library()
    info = DeviceInfo()
    _check(lib, None, lib.rtx_device_info(device, C.byref(info)))
    return info


def tile_layout(total_width: int, total_height: int, world: int):
    lib = load_library()
    tx, ty, tpr = C.c_uint32(), C.c_uint32(), C.c_uint32()
    lib.rtx_tile_layout(total_width, total_height, world, C.byref(tx), C.byref(ty), C.byref(tpr))
    return tx.value, ty.value, tpr.value


@dataclass
class Options:
    """RayTracer::Options (include/ray_tracer.h:17-30) with render.cc:17's defaults, AO off."""
    width: int = 600
    height: int = 600
    focalLength: float = 1.0
    nSuperSamples: int = 4
    enableShading: bool = True
    enableAO: bool = False
    aoMaxDistance: float = 0.2
    aoNumSamples: int = 0
    aoMethod: int = 0
    aoAlphaMin: int = 4
    aoAlphaMax: int = 90
    bvhMethod: int = 0


class RayTracer:
    """include/ray_tracer.h:31-38."""

    def __init__(self, options: Options):
        self.options = options
        n = int(math.sqrt(options.nSuperSamples))   # (unsigned int) sqrt(nSuperSamples)
        self.totalWidth = options.width * n
        self.totalHeight = options.height * n

    @property
    def n(self) -> int:
        return int(math.sqrt(self.options.nSuperSamples))


class CudaHost:
    """Drop-in for ``OpenCLHost`` (include/opencl_host.h:127-131)."""

    def __init__(self, rt: RayTracer, device: int = 0, jitter_seed: int = 0, tile_rank: int = 0, tile_world: int = 1):
        self._lib = load_library()
        self.rt = rt
        o = rt.options
        self._opt = _Options(o.width, o.height, o.focalLength, o.nSuperSamples, int(o.enableShading), int(o.enableAO),
                             o.aoMaxDistance, o.aoNumSamples, o.aoMethod, o.aoAlphaMin, o.aoAlphaMax, o.bvhMethod,
                             rt.totalWidth, rt.totalHeight, device, jitter_seed, tile_rank, tile_world)
        self._ctx = C.c_void_p()
        _check(self._lib, None, self._lib.rtx_create(C.byref(self._ctx), C.byref(self._opt)))
        self.tile_world = max(1, tile_world)
        self.tile_rank = tile_rank if tile_world > 1 else 0
        # experiment hook: RTX_TUNE="rays_per_thread=2,leaf_size=4" overrides the library defaults
        for kv in filter(None, os.environ.get("RTX_TUNE", "").split(",")):
            k, v = kv.split("=")
            self.set_tunable(globals()["TUNE_" + k.strip().upper()], int(v))

    # -- lifetime --
    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx:
            self._lib.rtx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        _check(self._lib, self._ctx, rc)

    # -- the reference's surface --
    @staticmethod
    def printInfo():
        return load_library().rtx_print_info()

    def upload(self, faces, nodes, aabbs, vertices, vnormals):
        faces = np.ascontiguousarray(faces, np.uint32)
        nodes = np.ascontiguousarray(nodes, np.uint32)
        aabbs = np.ascontiguousarray(aabbs, np.float32)
        vertices = np.ascontiguousarray(vertices, np.float32)
        vnormals = np.ascontiguousarray(vnormals, np.float32)
        self._ck(self._lib.rtx_upload(self._ctx, faces.ctypes.data, faces.size, nodes.ctypes.data, nodes.size,
                                      aabbs.ctypes.data, aabbs.size // 4, vertices.ctypes.data, vertices.size // 4,
                                      vnormals.ctypes.data, vnormals.size // 4))

    def upload_scene(self, scene: Scene):
        self.upload(scene.faces, scene.nodes, scene.aabbs, scene.vertices, scene.normals)
        self._ntris = scene.num_triangles

    def upload_mesh(self, vertices, faces, vnormals=None):
        """Raw mesh in (vertices / vnormals with the 16-byte Vec3f stride, faces in input order): the reference's
        longest-axis BVH is built on the device (rtx_upload_mesh) -- no host tree.  vnormals=None: the vertex normals
        of mesh.cc:95-139 are computed on the device as well."""
        vertices = np.ascontiguousarray(vertices, np.float32)
        if vertices.ndim == 2 and vertices.shape[1] == 3:                    # accept plain xyz too
            vertices = np.concatenate([vertices, np.zeros((vertices.shape[0], 1), np.float32)], 1)
        if vnormals is not None:
            vnormals = np.ascontiguousarray(vnormals, np.float32)
        faces = np.ascontiguousarray(faces, np.uint32)
        self._ck(self._lib.rtx_upload_mesh(self._ctx, vertices.ctypes.data, vertices.size // 4, faces.ctypes.data, faces.size // 3,
                                           vnormals.ctypes.data if vnormals is not None else None))
        self._ntris = faces.size // 3
        self._nverts = vertices.size // 4

    def download_normals(self) -> np.ndarray:
        out = np.empty((self._nverts, 4), np.float32)
        self._ck(self._lib.rtx_download_normals(self._ctx, out.ctypes.data, self._nverts))
        return out

    def download_tree(self, want_triangles: bool = True):
        """(nodes, aabbs, triangles, sorted_faces) of the last upload, in the formats of bvh.h:15-17 / render.cc:88-95."""
        s = Stats()
        self._ck(self._lib.rtx_get_stats(self._ctx, C.byref(s)))
        t = getattr(self, "_ntris", None) or (s.num_pairs + 1)
        n = 2 * t - 1
        nodes = np.empty(n, np.uint32)
        aabbs = np.empty((2 * n, 4), np.float32)
        tri = np.empty(t, np.uint32) if want_triangles else None
        faces = np.empty(3 * t, np.uint32)
        self._ck(self._lib.rtx_download_tree(self._ctx, nodes.ctypes.data, aabbs.ctypes.data,
                                             tri.ctypes.data if want_triangles else None, faces.ctypes.data))
        return nodes, aabbs, tri, faces

    def build_stats(self):
        ms, lv = C.c_double(), C.c_uint32()
        self._ck(self._lib.rtx_build_stats(self._ctx, C.byref(ms), C.byref(lv)))
        return ms.value, lv.value

    def __call__(self) -> bool:
        self._ck(self._lib.rtx_render(self._ctx))
        return True

    def download(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.rt.totalHeight, self.rt.totalWidth), np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.size == self.rt.totalWidth * self.rt.totalHeight
        self._ck(self._lib.rtx_download(self._ctx, out.ctypes.data))
        return out

    # -- extensions --
    def render_download(self, out: np.ndarray | None = None) -> np.ndarray:
        """``__call__`` + ``download`` in one call, the device->host copy overlapped with the tracing band by band."""
        if out is None:
            out = np.empty((self.rt.totalHeight, self.rt.totalWidth), np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.size == self.rt.totalWidth * self.rt.totalHeight
        self._ck(self._lib.rtx_render_download(self._ctx, out.ctypes.data))
        return out

    def set_tunable(self, which: int, value: int):
        self._ck(self._lib.rtx_set_tunable(self._ctx, which, value))

    def stats(self) -> dict:
        s = Stats()
        self._ck(self._lib.rtx_get_stats(self._ctx, C.byref(s)))
        return s.as_dict()

    def last_launches(self) -> int:
        """Kernels launched by the last render / trace call (rtx_stats.kernel_launches)."""
        s = Stats()
        self._ck(self._lib.rtx_get_stats(self._ctx, C.byref(s)))
        return int(s.kernel_launches)

    def render_async(self, stream: int = 0):
        self._ck(self._lib.rtx_render_async(self._ctx, C.c_void_p(stream)))

    def synchronize(self):
        self._ck(self._lib.rtx_synchronize(self._ctx))

    def download_hits(self):
        n = (self.rt.totalHeight, self.rt.totalWidth)
        fid = np.empty(n, np.uint32)
        dist = np.empty(n, np.float32)
        self._ck(self._lib.rtx_download_hits(self._ctx, fid.ctypes.data, dist.ctypes.data))
        return fid, dist

    def download_u8(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.rt.options.height, self.rt.options.width), np.uint8)
        assert out.dtype == np.uint8 and out.flags.c_contiguous and out.size == self.rt.options.width * self.rt.options.height
        self._ck(self._lib.rtx_download_u8(self._ctx, out.ctypes.data))
        return out

    def device_image(self):
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self._lib.rtx_device_image(self._ctx, C.byref(p), C.byref(n)))
        return p.value, n.value

    def probe_bandwidth(self, level: int, nbytes: int, iters: int) -> float:
        g = C.c_double()
        self._ck(self._lib.rtx_probe_bandwidth(self._ctx, level, nbytes, iters, C.byref(g)))
        return g.value

    def bind_output(self, device_ptr: int, count: int):
        self._ck(self._lib.rtx_bind_output(self._ctx, C.c_void_p(device_ptr), count))

    def bind_output_image(self, image_f32: int):
        """tile_world > 1: render this rank's tiles straight into the whole row-major float image (peer / mapped host memory)."""
        self._ck(self._lib.rtx_bind_output_image(self._ctx, C.c_void_p(image_f32)))

    def trace_rays(self, origins, dirs, max_distance: float = 100000.0, out_face_id=None, out_distance=None):
        """Closest hits of host rays (4 floats per origin / direction).  Page-locked inputs AND outputs let the chunks
        of a large batch overlap their copies with the tracing (rtx_trace_rays)."""
        origins = np.ascontiguousarray(origins, np.float32).reshape(-1, 4)
        dirs = np.ascontiguousarray(dirs, np.float32).reshape(-1, 4)
        n = origins.shape[0]
        fid = out_face_id if out_face_id is not None else np.empty(n, np.uint32)
        dist = out_distance if out_distance is not None else np.empty(n, np.float32)
        assert fid.dtype == np.uint32 and dist.dtype == np.float32 and fid.size == n and dist.size == n
        self._ck(self._lib.rtx_trace_rays(self._ctx, origins.ctypes.data, dirs.ctypes.data, n, C.c_float(max_distance),
                                          fid.ctypes.data, dist.ctypes.data))
        return fid, dist

    def trace_rays_device(self, d_origins: int, d_dirs: int, nrays: int, max_distance: float, d_face_id: int, d_distance: int, stream: int = 0):
        self._ck(self._lib.rtx_trace_rays_device(self._ctx, C.c_void_p(d_origins), C.c_void_p(d_dirs), nrays,
                                                 C.c_float(max_distance), C.c_void_p(d_face_id), C.c_void_p(d_distance),
                                                 C.c_void_p(stream)))

    def trace_random_rays(self, seed: int, first: int, nrays: int, max_distance: float = 100000.0, want_arrays: bool = False):
        fid = np.empty(nrays, np.uint32) if want_arrays else None
        dist = np.empty(nrays, np.float32) if want_arrays else None
        hits, idsum = C.c_uint64(), C.c_uint64()
        self._ck(self._lib.rtx_trace_random_rays(self._ctx, seed, first, nrays, C.c_float(max_distance),
                                                 fid.ctypes.data if want_arrays else None,
                                                 dist.ctypes.data if want_arrays else None,
                                                 C.byref(hits), C.byref(idsum)))
        return hits.value, idsum.value, fid, dist

    def resize_u8_async(self, d_tiles_u8: int = 0, count: int = 0, stream: int = 0):
        self._ck(self._lib.rtx_resize_u8_async(self._ctx, C.c_void_p(d_tiles_u8), count, C.c_void_p(stream)))

    def deinterleave_u8_async(self, d_gathered: int, world: int, stream: int = 0):
        self._ck(self._lib.rtx_deinterleave_u8_async(self._ctx, C.c_void_p(d_gathered), world, C.c_void_p(stream)))

    def deinterleave_async(self, d_gathered: int, world: int, stream: int = 0):
        self._ck(self._lib.rtx_deinterleave_async(self._ctx, C.c_void_p(d_gathered), world, C.c_void_p(stream)))

    # -- direct stores: this rank's share straight into the final image (own / peer / mapped host memory) --
    def resize_u8_to_async(self, image_u8: int, stream: int = 0):
        self._ck(self._lib.rtx_resize_u8_to_async(self._ctx, C.c_void_p(image_u8), C.c_void_p(stream)))

    def store_tiles_async(self, image_f32: int, stream: int = 0):
        self._ck(self._lib.rtx_store_tiles_async(self._ctx, C.c_void_p(image_f32), C.c_void_p(stream)))

    def render_store(self, image_f32: int):
        """``__call__`` + ``store_tiles_async`` in one blocking call, band by band (the stores overlap the tracing)."""
        self._ck(self._lib.rtx_render_store(self._ctx, C.c_void_p(image_f32)))

    def render_store_async(self, image_f32: int, stream: int = 0):
        self._ck(self._lib.rtx_render_store_async(self._ctx, C.c_void_p(image_f32), C.c_void_p(stream)))

    def adopt_u8(self, d_image_u8: int):
        self._ck(self._lib.rtx_adopt_u8(self._ctx, C.c_void_p(d_image_u8)))

    def peer_alloc(self, nbytes: int):
        """(device pointer, 64-byte IPC handle) of fresh device memory other rank processes can map."""
        p, h = C.c_void_p(), C.create_string_buffer(64)
        self._ck(self._lib.rtx_peer_alloc(self._ctx, nbytes, C.byref(p), h))
        return p.value, h.raw

    def peer_open(self, handle: bytes) -> int:
        p = C.c_void_p()
        self._ck(self._lib.rtx_peer_open(self._ctx, C.create_string_buffer(handle, 64), C.byref(p)))
        return p.value

    def peer_signal_async(self, flag: int, value: int, stream: int = 0):
        self._ck(self._lib.rtx_peer_signal_async(self._ctx, C.c_void_p(flag), value & 0xffffffff, C.c_void_p(stream)))

    def peer_wait_async(self, flags: int, count: int, value: int, timed_out: int, stream: int = 0):
        self._ck(self._lib.rtx_peer_wait_async(self._ctx, C.c_void_p(flags), count, value & 0xffffffff, C.c_void_p(timed_out), C.c_void_p(stream)))

    def peer_close(self, ptr: int):
        self._ck(self._lib.rtx_peer_close(self._ctx, C.c_void_p(ptr)))

    def peer_free(self, ptr: int):
        self._ck(self._lib.rtx_peer_free(self._ctx, C.c_void_p(ptr)))

    def copy_to_host(self, out: np.ndarray, device_ptr: int) -> np.ndarray:
        assert out.flags.c_contiguous
        self._ck(self._lib.rtx_copy_to_host(self._ctx, C.c_void_p(out.ctypes.data), C.c_void_p(device_ptr), out.nbytes))
        return out

    def debug_bounds(self):
        """Bounds-checked debug build only: (violations, line, index, limit) since the last call."""
        out = (C.c_uint * 4)()
        self._ck(self._lib.rtx_debug_bounds(self._ctx, C.byref(out)))
        return tuple(int(x) for x in out)

    def phase_ms(self) -> dict:
        """Device time per launch group of the last frame (needs TUNE_PHASE_TIMING)."""
        ms = (C.c_double * len(PHASES))()
        self._ck(self._lib.rtx_phase_ms(self._ctx, ms))
        return dict(zip(PHASES, [float(x) for x in ms]))


def host_register(array: np.ndarray) -> int:
    """Page-lock a host array (e.g. a shared mapping) and map it into the device; returns the pointer kernels may write."""
    lib = load_library()
    d = C.c_void_p()
    _check(lib, None, lib.rtx_host_register(C.c_void_p(array.ctypes.data), array.nbytes, C.byref(d)))
    return d.value


def host_unregister(array: np.ndarray) -> None:
    lib = load_library()
    _check(lib, None, lib.rtx_host_unregister(C.c_void_p(array.ctypes.data)))


def write_pgm(path: str, image_u8: np.ndarray) -> None:
    """src/render.cc:130-137: "P5 <w> <h> 255\\n" + raw bytes."""
    h, w = image_u8.shape
    with open(path, "wb") as f:
        f.write(b"P5 %d %d 255\n" % (w, h))
        f.write(np.ascontiguousarray(image_u8, np.uint8).tobytes())


def host_resize(tmp: np.ndarray, rt: RayTracer) -> np.ndarray:
    """RayTracer::resize (src/ray_tracer.cc:3-15) on the host, vectorised with
    the same summation order: ssY outer, ssX inner, float32 accumulation."""
    n = rt.n
    h, w = rt.options.height, rt.options.width
    t = np.ascontiguousarray(tmp, np.float32).reshape(rt.totalHeight, rt.totalWidth)
    total = np.zeros((h, w), np.float32)
    for sy in range(n):
        for sx in range(n):
            total = (total + t[sy:h * n:n, sx:w * n:n]).astype(np.float32)
    v = (total / np.float32(n * n)).astype(np.float32) * np.float32(255)
    return v.astype(np.float32).astype(np.int32).astype(np.uint8)
