/*
 * rtx_b200.h -- C ABI of the B200-native closest-hit path.
 * Library: opencl_raytracer_b200/lib/librtx_b200.so (CUDA runtime linked
 * statically; sm_100a only; no CPU fallback -- every entry point that needs a
 * device fails with RTX_ERR_NO_DEVICE / RTX_ERR_CUDA when there is none).
 *
 * Drop-in boundary: the reference's device host layer, include/opencl_host.h
 * :127-131 / src/opencl_host.cc, as called by src/render.cc:82-117.
 *
 *   reference (C++)                                          this ABI
 *   -------------------------------------------------------  -------------------
 *   static void OpenCLHost::printInfo()   opencl_host.cc:76  rtx_print_info, rtx_device_info
 *   OpenCLHost::OpenCLHost(const RayTracer&)  .cc:15-75      rtx_create
 *   void upload(faces,nodes,aabbs,vertices,vnormals) :120    rtx_upload
 *   bool operator()()                         .cc:137-149    rtx_render
 *   void download(float *image)               .cc:150-153    rtx_download
 *   ~OpenCLHost (implicit)                                   rtx_destroy
 *   check()/getErrorString  opencl_host.h:21-126             return codes + rtx_last_error
 *
 * The C++ adapter that restores the reference's class (and its print-and-exit
 * error behaviour) on top of this ABI is opencl_raytracer_b200/shadow/
 * opencl_host.h; INTEGRATION.md shows the build line.
 *
 * Conventions: plain pointers and sizes; all vector arrays use the
 * reference's 16-byte Vec3f stride (include/vec3.h:93-94) and the pad lane is
 * ignored (defined as 0); every call is blocking unless its name ends in
 * _async; nothing throws; 0 = success.  One context = one CUDA device.  A
 * context is not thread-safe; different contexts are independent.
 */
#ifndef RTX_B200_H
#define RTX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTX_OK               0
#define RTX_ERR_ARG          1  /* null pointer, inconsistent sizes, malformed BVH arrays */
#define RTX_ERR_NO_DEVICE    2  /* "No device found" (opencl_host.cc:30-31) */
#define RTX_ERR_CUDA         3  /* a CUDA runtime call failed; see rtx_last_error */
#define RTX_ERR_UNSUPPORTED  4  /* combination this library does not offer (see the call) */
#define RTX_ERR_STATE        5  /* call order: render before upload, ... */
#define RTX_ERR_NOMEM        6

#define RTX_NO_HIT 0xffffffffu

typedef struct rtx_ctx rtx_ctx;

/* Mirrors RayTracer::Options + totalWidth/totalHeight (include/ray_tracer.h
 * :17-39).  What opencl_host.cc:42-53 baked into the kernel as -D macros are
 * run-time parameters here. */
typedef struct rtx_options {
	uint32_t width, height;         /* output image */
	float    focal_length;          /* as given; the 6-significant-digit round trip of
	                                   compiler_options.h:13-19 is applied inside rtx_create */
	uint32_t n_super_samples;
	int32_t  enable_shading;        /* SHADING_ENABLE */
	int32_t  enable_ao;             /* AO_ENABLE: ambient-occlusion rays (intersect_kernel.cl:214-277, 305-307) */
	float    ao_max_distance;       /* AO_MAX_DISTANCE; same 6-digit round trip as focal_length */
	uint32_t ao_num_samples;        /* AO_NUM_SAMPLES: rings (uniform) or random rays; 0 disables AO like the kernel's #if */
	int32_t  ao_method;             /* AO_METHOD: 0 uniform rings, 1 random hemisphere (ray_tracer.h:10-13) */
	int32_t  ao_alpha_min, ao_alpha_max; /* degrees (uniform method) */
	int32_t  bvh_method;            /* informational */
	uint32_t total_width, total_height; /* width/height * (unsigned)sqrt(n_super_samples); 0 = derive */
	/* ---- extensions; all-zero reproduces the reference ---- */
	int32_t  device;                /* CUDA device ordinal */
	uint32_t jitter_seed;           /* 0 = regular grid (+0.5f); else hash jitter (config C3) */
	uint32_t tile_rank, tile_world; /* interleaved 32x32-pixel tile partition; world 0/1 = whole image */
} rtx_options;

typedef struct rtx_device_info_t {
	char     name[256];
	int32_t  cc_major, cc_minor;
	int32_t  sm_count;
	int32_t  clock_khz, mem_clock_khz, mem_bus_bits;
	uint64_t global_mem_bytes;
	uint64_t l2_bytes;
	uint64_t smem_per_sm_bytes, smem_per_block_optin_bytes;
	int32_t  max_threads_per_sm, regs_per_sm;
	int32_t  driver_version, runtime_version;
} rtx_device_info_t;

/* Statistics of the last render / trace call on the context. */
typedef struct rtx_stats {
	uint64_t rays;              /* rays traced by the last call (this rank's share) */
	double   kernel_ms;         /* CUDA-event time of the traversal kernel(s) */
	uint32_t kernel_launches;   /* kernels launched by the last call */
	uint32_t kernel_variant;    /* RTX_KERNEL_* actually used */
	/* device-side counters; only filled when RTX_TUNE_COUNTERS is set */
	uint64_t node_visits;       /* boxes slab-tested */
	uint64_t tri_tests;
	uint64_t leafbox_tests;
	uint64_t exact_path_rays;   /* rays routed to the reference-order walk */
	uint32_t tree_depth;        /* depth of the flattened tree (stack need) */
	uint32_t num_pairs;         /* internal nodes of the flattened tree */
	uint64_t packet_overflows;  /* packets whose frustum lists overflowed (fell back to per-ray traversal) */
} rtx_stats;

/* Tunables (rtx_set_tunable).  Defaults are the measured best. */
#define RTX_TUNE_KERNEL        1  /* RTX_KERNEL_* */
#define RTX_TUNE_LEAF_SIZE     2  /* max triangles per flattened leaf, 1..8 (takes effect at next upload) */
#define RTX_TUNE_RECORD_HITS   3  /* 1: keep per-ray face id + distance for rtx_download_hits */
#define RTX_TUNE_COUNTERS      4  /* 1: count node visits / triangle tests on the device (slower) */
#define RTX_TUNE_TOP_SMEM      5  /* node pairs of the top levels staged in shared memory, 0 = off */
#define RTX_TUNE_BLOCKS_PER_SM 6
#define RTX_TUNE_FLATTEN_ON_DEVICE 7 /* 1: build the GPU layout with kernels, 0: on the host */
#define RTX_TUNE_RAYS_PER_THREAD 8 /* primary rays per lane in the traversal kernel: 1, 2 (2x1 pixels), 4 (2x2 pixels), or 0 = the refill kernel */
#define RTX_TUNE_FRUSTUM       9  /* frustum front end for 16x8-pixel packets: 0 off, 1 on, -1 auto */
#define RTX_TUNE_LIST_RAYS_PER_THREAD 10 /* rays per lane in the candidate-list kernel: 1, 2 or 4 */
#define RTX_TUNE_RAY_TABLES    12  /* 1 (default): per-column/row tables of the pixel terms of the primary ray direction */
#define RTX_TUNE_PHASE_TIMING  13  /* 1: CUDA events around every launch group of a frame, read back with rtx_phase_ms */
#define RTX_TUNE_INCOHERENT_KERNEL 11 /* arbitrary rays: 1 persistent refill + parked leaves (default), 0 plain while-while */

#define RTX_KERNEL_PERSISTENT  0  /* persistent warps, ordered stack traversal, distance culling */
#define RTX_KERNEL_EXHAUSTIVE  1  /* one thread per ray, the reference's stackless pre-order walk */

/* ---- the reference's five calls ---- */

int rtx_print_info(void);
int rtx_device_count(int *count);
int rtx_device_info(int device, rtx_device_info_t *info);

int rtx_create(rtx_ctx **ctx, const rtx_options *options);

/* All host reads finish before this returns (render.cc:96-103 clears the
 * vectors right after).  faces: 3 vertex ids per triangle in BVH leaf order;
 * nodes: pre-order subtree sizes, nodes[0] == nnodes; aabbs16: 2*nnodes
 * (min,max) vectors; verts16/vnormals16: nverts == nnormals vectors. */
int rtx_upload(rtx_ctx *ctx,
               const uint32_t *faces, size_t nfaceidx,
               const uint32_t *nodes, size_t nnodes,
               const float *aabbs16, size_t naabbvec,
               const float *verts16, size_t nverts,
               const float *vnormals16, size_t nnormals);

/* The same from the raw mesh (what mesh.cc hands to bvh.cc): faces are 3 vertex ids per triangle in INPUT order.
 * The reference's longest-axis BVH (src/bvh.cc:59-162: centroid-box midpoint cut, stable partition, one triangle per
 * leaf) is built on the device, level by level, and yields the arrays bvh.cc would -- rtx_download_tree returns them
 * -- so everything downstream is unchanged.  Replaces BVH::buildBVH + the face sort of render.cc:88-95 + rtx_upload
 * for callers that are not bound to the reference's host code.  vnormals16 == NULL: the vertex normals are computed
 * on the device too (compute_vertex_normals, src/mesh.cc:95-139, same additions in the same order);
 * rtx_download_normals returns them. */
int rtx_upload_mesh(rtx_ctx *ctx, const float *verts16, size_t nverts, const uint32_t *faces, size_t nfaces,
                    const float *vnormals16);
int rtx_download_tree(rtx_ctx *ctx, uint32_t *nodes, float *aabbs16, uint32_t *triangles, uint32_t *sorted_faces);
int rtx_build_stats(const rtx_ctx *ctx, double *build_ms, uint32_t *levels);
int rtx_download_normals(rtx_ctx *ctx, float *vnormals16, size_t nverts);   /* the normals of the last rtx_upload_mesh */

/* Launch the traversal over this context's share of the image and wait. */
int rtx_render(rtx_ctx *ctx);

/* total_width*total_height floats, row-major (whole image only: tile_world <= 1). */
int rtx_download(rtx_ctx *ctx, float *image);

void rtx_destroy(rtx_ctx *ctx);

/* ctx may be NULL: message of the last failed call without a context. */
const char *rtx_last_error(const rtx_ctx *ctx);

/* ---- extensions needed by the configurations (not in the reference) ---- */

int rtx_set_tunable(rtx_ctx *ctx, int which, int64_t value);
int rtx_get_stats(const rtx_ctx *ctx, rtx_stats *stats);

/* Enqueue the render on the caller's CUDA stream (cudaStream_t as void*; NULL
 * is CUDA's legacy default stream) without waiting.  The blocking calls use a
 * private non-blocking stream of the context; rtx_synchronize waits for the
 * whole device. */
int rtx_render_async(rtx_ctx *ctx, void *stream);
int rtx_synchronize(rtx_ctx *ctx);

/* rtx_render + rtx_download in one blocking call (whole image only): the frame is traced in bands of tile
 * rows and every finished band is copied to `image` on a second stream while the next one is traced.  `image`
 * should be page-locked for the copies to overlap (pageable memory works, without the overlap). */
int rtx_render_download(rtx_ctx *ctx, float *image);

/* Per-ray hit triangle (3 * leaf index, the kernel's face_id; RTX_NO_HIT on
 * miss) and hit distance (+inf on miss) of the last render; needs
 * RTX_TUNE_RECORD_HITS.  Whole image only. */
int rtx_download_hits(rtx_ctx *ctx, uint32_t *face_id, float *distance);

/* RayTracer::resize (src/ray_tracer.cc:3-15) on the device, then a
 * width*height byte download: the bytes render.cc:130-137 writes after the
 * PGM header. */
int rtx_download_u8(rtx_ctx *ctx, unsigned char *image);

/* Device pointer to this context's float output: the row-major image, or
 * with tile_world > 1 the compact [local_tile][32][32] buffer.  count =
 * number of floats. */
int rtx_device_image(rtx_ctx *ctx, void **device_ptr, size_t *count);

/* Render into caller-owned device memory (e.g. a torch tensor that NCCL will
 * gather) instead of the context's own buffer: `count` floats, at least what
 * rtx_device_image reports.  device_ptr == NULL restores the internal buffer. */
int rtx_bind_output(rtx_ctx *ctx, void *device_ptr, size_t count);

/* tile_world > 1: render this rank's tiles straight into the WHOLE row-major total_width*total_height float image at
 * `image_f32` -- rank 0's device memory mapped over NVLink (rtx_peer_open) or page-locked host memory mapped into the
 * device (rtx_host_register).  Each pixel leaves the SM as it is shaded: the transfer overlaps the tracing and needs
 * no kernel, gather or copy of its own.  Float image only (no hit records, no ambient occlusion).  NULL unbinds. */
int rtx_bind_output_image(rtx_ctx *ctx, void *image_f32);

/* Closest hit for arbitrary rays (config C5), reference scene_intersect
 * semantics with the given max_distance.  origins/dirs: 4 floats per ray.
 * Host-pointer and device-pointer forms.  The host form cuts a large batch into
 * chunks of 4 Mi rays and overlaps upload, tracing and download of consecutive
 * chunks on three streams (page-locked host arrays make the copies asynchronous). */
int rtx_trace_rays(rtx_ctx *ctx, const float *origins, const float *dirs, size_t nrays, float max_distance,
                   uint32_t *face_id, float *distance);
int rtx_trace_rays_device(rtx_ctx *ctx, const void *d_origins, const void *d_dirs, size_t nrays, float max_distance,
                          void *d_face_id, void *d_distance, void *stream);
/* Rays [first, first+nrays) of the counter-based generator,
 * generated on the device; results stay on the device unless host pointers
 * are given (either may be NULL).  sum_face_id / hit_count: checksums. */
int rtx_trace_random_rays(rtx_ctx *ctx, uint32_t seed, uint64_t first, size_t nrays, float max_distance,
                          uint32_t *face_id, float *distance, uint64_t *hit_count, uint64_t *sum_face_id);

/* Cache bandwidth probe for the roofline denominators of cache-resident scenes
 * (MEASURED_PEAKS.json only holds HBM): 128-bit loads over a scratch buffer of
 * `bytes`, `iters` sweeps.  level 0: whole buffer from every CTA, L1 bypassed
 * (L2 when bytes fits L2, HBM when not); level 1: 64 KB per CTA through L1. */
int rtx_probe_bandwidth(rtx_ctx *ctx, int level, size_t bytes, int iters, double *gbps);

/* ---- multi-GPU tile partition (one context per rank) ---- */

/* Layout of the interleaved partition for an image: tiles are 32x32 pixels,
 * tile t belongs to rank t % world.  tiles_per_rank is the padded (equal)
 * count every rank's compact buffer holds. */
int rtx_tile_layout(uint32_t total_width, uint32_t total_height, uint32_t world,
                    uint32_t *tiles_x, uint32_t *tiles_y, uint32_t *tiles_per_rank);

/* Byte path (needs sqrt(n_super_samples) to divide 32 when tile_world > 1): RayTracer::resize on this
 * context's share.  tile_world <= 1: into the context's byte image (d_tiles_u8 ignored).  Otherwise into
 * the caller's compact buffer of tiles_per_rank * (32/n)^2 bytes, to be gathered and handed to
 * rtx_deinterleave_u8_async on rank 0.  rtx_download_u8 then returns the width*height bytes. */
int rtx_resize_u8_async(rtx_ctx *ctx, void *d_tiles_u8, size_t count, void *stream);
int rtx_deinterleave_u8_async(rtx_ctx *ctx, const void *d_gathered_u8, uint32_t world, void *stream);

/* Scatter `world` gathered compact buffers (rank-major, tiles_per_rank*1024
 * floats each) into the row-major image of this context (rank 0). */
int rtx_deinterleave_async(rtx_ctx *ctx, const void *d_gathered, uint32_t world, void *stream);

/* ---- direct stores: no collective, every rank writes its share into the final image itself ----
 *
 * The final image may live in this rank's device memory, in rank 0's (mapped with rtx_peer_open: the stores cross
 * NVLink / NVSwitch, resize + gather + de-interleave become one kernel per rank) or in page-locked host memory mapped
 * into the device (rtx_host_register: every rank's tiles leave over its OWN PCIe link instead of funnelling through
 * rank 0's).  The caller orders the ranks (e.g. a one-element all-reduce on the same stream after the store: it
 * completes on rank 0 only when every rank's store kernel has finished).
 *   rtx_resize_u8_to_async   RayTracer::resize (ray_tracer.cc:3-15) of this rank's tiles -> row-major width*height bytes
 *   rtx_store_tiles_async    this rank's float tiles -> row-major total_width*total_height floats
 *   rtx_adopt_u8             rtx_download_u8 of this context reads the finished bytes from d_image_u8 */
int rtx_resize_u8_to_async(rtx_ctx *ctx, void *image_u8, void *stream);
int rtx_store_tiles_async(rtx_ctx *ctx, void *image_f32, void *stream);
int rtx_adopt_u8(rtx_ctx *ctx, const void *d_image_u8);
/* rtx_render + rtx_store_tiles_async as one blocking call: rtx_render_download for a tile partition whose host image is
 * shared by the rank processes.  tile_world <= 1: same as rtx_render_download (bands traced and copied on two streams). */
int rtx_render_store(rtx_ctx *ctx, void *image_f32);
int rtx_render_store_async(rtx_ctx *ctx, void *image_f32, void *stream);    /* tile partitions only; enqueues and returns */

/* Device memory that other rank processes can map (CUDA IPC; 64-byte handle, any transport). */
int rtx_peer_alloc(rtx_ctx *ctx, size_t bytes, void **device_ptr, unsigned char handle[64]);
int rtx_peer_open(rtx_ctx *ctx, const unsigned char handle[64], void **device_ptr);
int rtx_peer_close(rtx_ctx *ctx, void *device_ptr);
int rtx_peer_free(rtx_ctx *ctx, void *device_ptr);
int rtx_copy_to_host(rtx_ctx *ctx, void *host_dst, const void *device_src, size_t bytes);   /* blocking, after all queued work */

/* Ordering the ranks of the peer-memory paths without a collective: monotonic 32-bit frame counters in a (zeroed)
 * rtx_peer_alloc buffer every rank maps.  rtx_peer_signal_async stores `value` behind everything queued before it on
 * `stream` (system-scope release); rtx_peer_wait_async holds `stream` until flags[0..count) have all reached `value`.
 * A wait gives up after ~2 s of device time and sets *timed_out (device memory, e.g. a word of the same buffer) to 1
 * instead of hanging.  The waiter and the signaller must run on different GPUs. */
int rtx_peer_signal_async(rtx_ctx *ctx, void *flag, uint32_t value, void *stream);
int rtx_peer_wait_async(rtx_ctx *ctx, const void *flags, uint32_t count, uint32_t value, void *timed_out, void *stream);

/* Page-lock and device-map caller-owned host memory (a shared mapping several rank processes opened). */
int rtx_host_register(void *p, size_t bytes, void **device_alias);
int rtx_host_unregister(void *p);

/* Device time of each launch group of the last frame rendered with RTX_TUNE_PHASE_TIMING set (ms[RTX_NUM_PHASES]). */
#define RTX_PHASE_TABLES        0   /* per-column / per-row ray tables */
#define RTX_PHASE_COLLECT_SUPER 1   /* frustum front end, 128x128-pixel super-tiles */
#define RTX_PHASE_COLLECT       2   /* frustum front end, 32x32-pixel tiles */
#define RTX_PHASE_TRAVERSAL     3   /* candidate-list kernel, or the per-ray traversal kernel */
#define RTX_PHASE_OVERFLOW      4   /* per-ray traversal of the tiles whose list overflowed */
#define RTX_PHASE_AO            5   /* ambient-occlusion pass */
#define RTX_NUM_PHASES          6
int rtx_phase_ms(const rtx_ctx *ctx, double *ms);

/* Only in the bounds-checked debug build (lib/librtx_b200_dbg.so, -DRTX_DEBUG_BOUNDS): every device-side index into the
 * scene arrays, stacks, queues and candidate lists is range-checked; out = { violations since the last call, source
 * line, index, limit of the first one }.  The release build returns RTX_ERR_UNSUPPORTED. */
int rtx_debug_bounds(rtx_ctx *ctx, unsigned int out[4]);

#ifdef __cplusplus
}
#endif
#endif
