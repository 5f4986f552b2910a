/*
 * rtx_api.cu -- the C ABI of include/rtx_b200.h over the CUDA runtime.
 *
 * Replaces the reference's device host layer (src/opencl_host.cc): device
 * pick + info dump (:15-31, :76-119), buffer creation and blocking upload
 * (:120-136), kernel launch + finish (:137-149), blocking download
 * (:150-153).  What the reference passed to the OpenCL JIT as -D macros
 * (:42-53) are kernel arguments here; the kernels are compiled ahead of time
 * for sm_100a.  There is no CPU path: without a CUDA device every call that
 * needs one returns RTX_ERR_NO_DEVICE.
 */
#include "rtx_b200.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <vector>

#include "rtx_kernels.cuh"
#include "rtx_build.cuh"

namespace {

thread_local std::string g_error;

struct DevBuf {
	void *p = nullptr;
	size_t bytes = 0;
	cudaError_t alloc(size_t n)
	{
		if (p && bytes >= n && bytes <= 2 * n + 4096) return cudaSuccess;
		release();
		if (n == 0) n = 16;
		cudaError_t e = cudaMalloc(&p, n);
		if (e == cudaSuccess) bytes = n; else p = nullptr;
		return e;
	}
	void release()
	{
		if (p) cudaFree(p);
		p = nullptr;
		bytes = 0;
	}
	template <typename T> T *as() const { return static_cast<T *>(p); }
};

} /* namespace */

#define RTX_MAX_BANDS 16

struct rtx_ctx {
	rtx_options opt;
	float focal;                 /* after the compiler_options.h round trip */
	int device = 0;
	int sm_count = 0;
	cudaStream_t stream = nullptr;
	cudaStream_t copy_stream = nullptr;                 /* rtx_render_download: device->host copies of finished bands */
	cudaEvent_t band_ev[RTX_MAX_BANDS] = {}, copy_done = nullptr;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	bool ev_pending = false;
	std::string error;
	/* tunables */
	int kernel = RTX_KERNEL_PERSISTENT;
	int leaf_size = 1;
	int record_hits = 0;
	int counters = 0;
	int top_smem = 0;
	int blocks_per_sm = 0;       /* 0 = default of the variant */
	int flatten_on_device = 1;
	int rays_per_thread = 1;     /* traversal kernel: 1, 2 (2x1) or 4 (2x2) pixels per lane, 0 = refill kernel */
	int list_rays_per_thread = 2; /* rays per lane in the candidate-list kernel: 1, 2 or 4 */
	int ray_tables = 1;          /* per-column / per-row tables of the pixel terms of the primary ray */
	int incoherent_kernel = 1;   /* 1: persistent refill + parked leaves for arbitrary rays, 0: plain while-while */
	int frustum = -1;            /* frustum front end: 0 off, 1 on, -1 auto (rays per triangle >= 24) */
	/* scene */
	bool uploaded = false;
	SceneDev sc{};
	DevBuf d_pairs, d_tris, d_leafbox, d_pleafbox, d_tnormals, d_ref_nodes, d_ref_aabbs;
	DevBuf t_faces, t_verts, t_vnormals, t_scan;   /* upload staging, kept between uploads */
	DevBuf t_build, d_triangles;                   /* rtx_upload_mesh: builder work space; leaf order -> input face id */
	DevBuf t_parent;                               /* device flatten: parent link of every node pair (k_slack_leaves) */
	DevBuf r_o[2], r_d[2], r_f[2], r_t[2];         /* rtx_trace_rays: two sets of chunk buffers, kept between calls */
	cudaStream_t r_in = nullptr, r_out = nullptr;
	cudaEvent_t r_ev_in[2] = {}, r_ev_done[2] = {}, r_ev_out[2] = {}, r_ev_a = nullptr, r_ev_b = nullptr;
	uint32_t build_levels = 0;
	double build_ms = 0.0;
	bool tree_on_device = false;                   /* d_ref_* / t_faces / d_triangles hold a complete reference tree */
	size_t tree_nodes = 0, tree_tris = 0;
	TreeResult h_tree{};         /* result words of the device-side tree check (k_tree_*) */
	f3 bbmin{}, bbmax{};
	uint32_t tree_depth = 0;
	bool boxes_nested = true;    /* every child box inside its parent's (k_tree_check); false: literal walk for every ray */
	/* image */
	uint32_t W = 0, H = 0, tiles_x = 0, tiles_y = 0, tiles_per_rank = 0, local_tiles = 0;
	uint32_t rank = 0, world = 1;
	float *ext_image = nullptr;  /* caller-owned output (rtx_bind_output) */
	bool ext_rowmajor = false;   /* ... and it is the whole row-major image although tile_world > 1 (rtx_bind_output_image) */
	DevBuf d_image, d_image_full, d_face_id, d_dist, d_u8, d_counter, d_counters, d_sums, d_lists, d_slists, d_raytab;
	cudaEvent_t raytab_ev = nullptr;   /* set once the ray tables are filled (enqueue_render) */
	DevBuf d_tile_done;          /* rtx_render_store: units finished per local tile (fused store of the packet kernels) */
	DevBuf d_hit_st, d_ao_ring;  /* ambient occlusion: (s, t) of the primary hits; sample table of the uniform method */
	bool ao = false;
	float ao_max_distance = 0.f; /* after the compiler_options.h round trip */
	uint32_t ao_ring_cap = 0;
	bool rendered = false, full_valid = false, u8_valid = false;
	const unsigned char *ext_u8 = nullptr;   /* rtx_adopt_u8: the finished byte image lives in caller-owned device memory */
	/* per-phase device times of one frame (RTX_TUNE_PHASE_TIMING): marks around each launch group */
	int phase_timing = 0;
	cudaEvent_t ph_ev[RTX_NUM_PHASES + 1] = {};
	bool ph_rec[RTX_NUM_PHASES + 1] = {};
	double phase_ms[RTX_NUM_PHASES] = {};
	/* stats */
	rtx_stats stats{};
};

namespace {

int fail(rtx_ctx *ctx, int code, const std::string &msg)
{
	if (ctx) ctx->error = msg;
	g_error = msg;
	return code;
}

int cuda_fail(rtx_ctx *ctx, cudaError_t e, const char *what)
{
	const bool nodev = e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice;
	return fail(ctx, nodev ? RTX_ERR_NO_DEVICE : RTX_ERR_CUDA,
	            std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
}

#define CU(ctx, call)                                                   \
	do {                                                                \
		cudaError_t e_ = (call);                                        \
		if (e_ != cudaSuccess) return cuda_fail(ctx, e_, #call);        \
	} while (0)

/* include/compiler_options.h:13-19: `ss << v` (6 significant digits) + 'f',
 * re-read by the OpenCL front end. */
float focal_roundtrip(float f)
{
	char buf[64];
	std::snprintf(buf, sizeof buf, "%g", (double)f);
	return std::strtof(buf, nullptr);
}

/* ----------------------------- flatten ---------------------------------- */

struct Flat {
	std::vector<float4> pairs;
	std::vector<uint32_t> leaf_node;   /* leaf index -> reference node index */
	uint32_t top_pairs = 0;
	uint32_t depth = 0;
};

inline uint32_t leaves_of(uint32_t subtree_size) { return (subtree_size + 1) >> 1; }

/* The invariants of SURVEY 3.3 that the traversal relies on. */
bool validate_tree(const uint32_t *nodes, size_t n, size_t ntris, std::string &why)
{
	if (n == 0 || nodes[0] != n) { why = "nodes[0] must equal the node count"; return false; }
	if ((n & 1) == 0 || leaves_of((uint32_t)n) != ntris) { why = "node count must be 2*triangles-1"; return false; }
	for (size_t i = 0; i < n; ++i) {
		const size_t s = nodes[i];
		if ((s & 1) == 0 || i + s > n) { why = "subtree size out of range at node " + std::to_string(i); return false; }
		if (s > 1) {
			const size_t l = nodes[i + 1];
			if ((l & 1) == 0 || l + 2 > s) { why = "left subtree size invalid at node " + std::to_string(i); return false; }
			if (nodes[i + 1 + l] != s - 1 - l) { why = "children do not tile node " + std::to_string(i); return false; }
		}
	}
	return true;
}

inline int leaf_ref(uint32_t first_tri, uint32_t count) { return (int)~((first_tri << 3) | (count - 1)); }

void flatten(const uint32_t *nodes, const float *aabbs16, size_t n, int leaf_size, uint32_t top_target, Flat &out)
{
	const size_t ntris = leaves_of((uint32_t)n);
	out.leaf_node.resize(ntris);
	for (size_t i = 0, t = 0; i < n; ++i)
		if (nodes[i] == 1) out.leaf_node[t++] = (uint32_t)i;

	struct Item { uint32_t node, first_leaf, depth; int64_t patch; /* float index of the parent's ref slot, -1 = none */ };
	auto box = [&](uint32_t node, float4 &a, float4 &b, int ref) {
		const float *lo = aabbs16 + 8 * (size_t)node, *hi = lo + 4;
		a = make_float4(lo[0], lo[1], lo[2], hi[0]);
		b = make_float4(hi[1], hi[2], __builtin_bit_cast(float, ref), box_slack(lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]));
	};
	out.pairs.clear();
	out.pairs.reserve(4 * (ntris / (size_t)std::max(1, leaf_size) + 16));
	out.depth = 0;

	if (n == 1) {   /* a single triangle: pair 0 = {the leaf, an unreachable far box} */
		float4 a, b;
		box(0, a, b, leaf_ref(0, 1));
		out.pairs.push_back(a); out.pairs.push_back(b);
		out.pairs.push_back(make_float4(3e38f, 3e38f, 3e38f, 3e38f));
		out.pairs.push_back(make_float4(3e38f, 3e38f, __builtin_bit_cast(float, leaf_ref(0, 1)), 0.0f));
		out.top_pairs = 1;
		out.depth = 1;
		return;
	}

	std::deque<Item> bfs;
	std::vector<Item> dfs;
	bfs.push_back(Item{ 0, 0, 1, -1 });
	bool in_bfs = true;
	auto expand = [&](const Item &it) {
		const uint32_t p = (uint32_t)(out.pairs.size() / 4);
		if (it.patch >= 0) reinterpret_cast<float *>(out.pairs.data())[it.patch] = __builtin_bit_cast(float, (int)p);
		if (it.depth > out.depth) out.depth = it.depth;
		out.pairs.resize(out.pairs.size() + 4);
		const uint32_t iL = it.node + 1, iR = it.node + 1 + nodes[it.node + 1];
		const uint32_t child[2] = { iL, iR };
		const uint32_t first[2] = { it.first_leaf, it.first_leaf + leaves_of(nodes[iL]) };
		Item kids[2];
		int nk = 0;
		for (int s = 0; s < 2; ++s) {
			const uint32_t lv = leaves_of(nodes[child[s]]);
			float4 a, b;
			if (lv <= (uint32_t)leaf_size) {
				box(child[s], a, b, leaf_ref(first[s], lv));
			} else {
				box(child[s], a, b, 0);
				kids[nk++] = Item{ child[s], first[s], it.depth + 1, (int64_t)(4 * (4 * (size_t)p + 2 * s + 1) + 2) };
			}
			out.pairs[4 * (size_t)p + 2 * s] = a;
			out.pairs[4 * (size_t)p + 2 * s + 1] = b;
		}
		if (in_bfs) { for (int k = 0; k < nk; ++k) bfs.push_back(kids[k]); }
		else { for (int k = nk - 1; k >= 0; --k) dfs.push_back(kids[k]); }
	};
	while (!bfs.empty() && out.pairs.size() / 4 < top_target) {
		const Item it = bfs.front();
		bfs.pop_front();
		expand(it);
	}
	out.top_pairs = (uint32_t)(out.pairs.size() / 4);
	in_bfs = false;
	while (!bfs.empty()) {
		dfs.push_back(bfs.front());
		bfs.pop_front();
		while (!dfs.empty()) {
			const Item it = dfs.back();
			dfs.pop_back();
			expand(it);
		}
	}
}

/* ------------------------------ launches -------------------------------- */

/* Resident CTAs per SM of a kernel, asked once per (kernel, device, shared-memory size): the query costs host time
 * on every launch otherwise.  Only the occupancy is cached.  The dynamic shared-memory opt-in is a property of the
 * kernel, not of the cache entry: it is tracked per (kernel, device) and only ever raised, so a launch with a size
 * that was seen before can never meet an attribute a smaller launch lowered in between. */
struct OccEntry { const void *kernel; int device; size_t smem; int occ; };
struct SmemEntry { const void *kernel; int device; size_t max_smem; };
static std::vector<OccEntry> g_occ;
static std::vector<SmemEntry> g_smem;
static std::mutex g_occ_mutex;     /* contexts on different host threads share the tables */

template <typename K>
cudaError_t resident_blocks(K kernel, int block, size_t smem, int device, int *occ_out)
{
	const void *key = reinterpret_cast<const void *>(kernel);
	std::lock_guard<std::mutex> lock(g_occ_mutex);
	SmemEntry *se = nullptr;
	for (SmemEntry &x : g_smem)
		if (x.kernel == key && x.device == device) { se = &x; break; }
	if (!se || smem > se->max_smem) {
		cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (e != cudaSuccess) return e;
		if (se) se->max_smem = smem; else g_smem.push_back(SmemEntry{ key, device, smem });
	}
	for (const OccEntry &x : g_occ)
		if (x.kernel == key && x.device == device && x.smem == smem) { *occ_out = x.occ; return cudaSuccess; }
	int occ = 0;
	cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, block, smem);
	if (e != cudaSuccess) return e;
	if (occ < 1) occ = 1;
	/* the traversal kernels live on L1 hits of node pairs and triangles: ask for the smallest shared-memory
	 * carve-out that still holds the resident CTAs' stacks and lists (+1 KB per CTA the runtime reserves); the rest
	 * of the SM's 256 KB stays L1: C3 4.14 -> 4.08 ms.  (cudaSharedmemCarveoutMaxL1 itself costs the occupancy: 2.5x slower.) */
	const char *pct = std::getenv("RTX_CARVEOUT_HINT");          /* experiment hook: a percentage, or "off" */
	if (!(pct && pct[0] == 'o')) {
		int want = pct ? std::atoi(pct) : 0;
		if (want <= 0) want = (int)(((size_t)occ * (smem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024)) + 1;
		if (want > 100) want = 100;
		cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, want);   /* a hint: errors ignored */
	}
	g_occ.push_back(OccEntry{ key, device, smem, occ });
	*occ_out = occ;
	return cudaSuccess;
}

/* Phase timing (RTX_TUNE_PHASE_TIMING): mark k is recorded right before phase k is enqueued, mark RTX_NUM_PHASES at the
 * end of the frame; a phase lasts from its mark to the next recorded one (finish_stats). */
cudaError_t phase_mark(rtx_ctx *c, cudaStream_t st, int k)
{
	if (!c->phase_timing) return cudaSuccess;
	if (!c->ph_ev[k]) { cudaError_t e = cudaEventCreate(&c->ph_ev[k]); if (e != cudaSuccess) return e; }
	c->ph_rec[k] = true;
	return cudaEventRecord(c->ph_ev[k], st);
}

template <int BLOCK, int MINB, int SST, bool TOP, bool COUNT, bool RECORD>
cudaError_t launch_render_t(rtx_ctx *c, const Work &w, cudaStream_t st, int blocks_per_sm)
{
	auto k = k_render_persistent<BLOCK, MINB, SST, TOP, COUNT, RECORD>;
	const size_t smem = (size_t)SST * BLOCK * sizeof(uint2) + (TOP ? (size_t)c->sc.top_pairs * 64 : 0);
	int occ = 0;
	cudaError_t e = resident_blocks(k, BLOCK, smem, c->device, &occ);
	if (e != cudaSuccess) return e;
	if (blocks_per_sm > 0 && blocks_per_sm < occ) occ = blocks_per_sm;
	const unsigned grid = (unsigned)(c->sm_count * occ);
	k<<<grid, BLOCK, smem, st>>>(c->sc, w, c->d_counters.as<Counters>());
	return cudaGetLastError();
}

template <bool COUNT, bool RECORD, int SOURCE>
cudaError_t launch_pt(rtx_ctx *c, const RayWork &rw, const Work &pw, cudaStream_t st)
{
	/* 4 CTAs of 256 threads per SM (64 registers): the kernel waits on L2 for most of its life (ncu: 6 warps per
	 * scheduler, 1.1 eligible, long-scoreboard 3.5 per issue at 3 CTAs); C5 25.7 -> 23.0 ms */
	constexpr int BLOCK = 256, MINB = 4, SST = 8;
	auto k = k_trace_persistent<BLOCK, MINB, SST, COUNT, RECORD, SOURCE>;
	const size_t smem = (size_t)SST * BLOCK * sizeof(uint2);
	int occ = 0;
	cudaError_t e = resident_blocks(k, BLOCK, smem, c->device, &occ);
	if (e != cudaSuccess) return e;
	if (c->blocks_per_sm > 0 && c->blocks_per_sm < occ) occ = c->blocks_per_sm;
	k<<<(unsigned)(c->sm_count * occ), BLOCK, smem, st>>>(c->sc, rw, pw, c->d_counters.as<Counters>());
	return cudaGetLastError();
}

template <int BLOCK, int MINB, int SST, bool COUNT, bool RECORD, int RX, int RY, int MODE, bool STORE = false>
cudaError_t launch_packet_t(rtx_ctx *c, const Work &w, cudaStream_t st)
{
	auto k = k_render_packet<BLOCK, MINB, SST, COUNT, RECORD, RX, RY, MODE, STORE>;
	const size_t smem = (size_t)SST * BLOCK * sizeof(uint2) + (MODE == 1 ? (size_t)(BLOCK / 32) * (2 * RTX_CCAP) * 4 : 0);
	int occ = 0;
	cudaError_t e = resident_blocks(k, BLOCK, smem, c->device, &occ);
	if (e != cudaSuccess) return e;
	if (c->blocks_per_sm > 0 && c->blocks_per_sm < occ) occ = c->blocks_per_sm;
	k<<<(unsigned)(c->sm_count * occ), BLOCK, smem, st>>>(c->sc, w, c->d_counters.as<Counters>());
	return cudaGetLastError();
}

/* STORE: the fused tile store of rtx_render_store (default tunables, no counters, no hit records) */
template <bool COUNT, bool RECORD, bool STORE = false>
cudaError_t launch_packet(rtx_ctx *c, const Work &w, cudaStream_t st)
{
	if (c->rays_per_thread == 2 && !w.frustum) return launch_packet_t<256, 3, 8, COUNT, RECORD, 2, 1, 0, STORE>(c, w, st);
	if (w.frustum) {
		/* listed tiles first (no traversal code in that kernel), then the overflowed ones; both pull
		 * units from the same kind of counter, so it is re-zeroed in between */
		cudaError_t e = STORE ? launch_packet_t<256, 4, 0, COUNT, RECORD, 2, 1, 1, STORE>(c, w, st)
		              : c->list_rays_per_thread == 1 ? launch_packet_t<256, 5, 0, COUNT, RECORD, 1, 1, 1>(c, w, st)
		              : c->list_rays_per_thread == 2 ? launch_packet_t<256, 4, 0, COUNT, RECORD, 2, 1, 1>(c, w, st)
		                                             : launch_packet_t<256, 3, 0, COUNT, RECORD, 2, 2, 1>(c, w, st);
		if (e != cudaSuccess) return e;
		if ((e = phase_mark(c, st, RTX_PHASE_OVERFLOW)) != cudaSuccess) return e;
		Work w2 = w;
		w2.counter = w.counter + 1;            /* its own work counter, zeroed together with the first at the start of the frame / band */
		return launch_packet_t<256, 2, 8, COUNT, RECORD, 2, 2, 2, STORE>(c, w2, st);
	}
	return launch_packet_t<256, 2, 8, COUNT, RECORD, 2, 2, 0, STORE>(c, w, st);
}

template <bool TOP, bool COUNT, bool RECORD>
cudaError_t launch_render_v(rtx_ctx *c, const Work &w, cudaStream_t st)
{
	return launch_render_t<256, 3, 8, TOP, COUNT, RECORD>(c, w, st, c->blocks_per_sm);
}

cudaError_t launch_render(rtx_ctx *c, const Work &w, cudaStream_t st)
{
	const bool top = c->top_smem > 0 && c->sc.top_pairs > 0, cnt = c->counters != 0, rec = w.face_id != nullptr;
	if (c->kernel == RTX_KERNEL_EXHAUSTIVE || !w.ordered_ok) {
		const unsigned warps_per_block = 8, grid = (w.num_units + warps_per_block - 1) / warps_per_block;
		if (grid == 0) return cudaSuccess;
		if (cnt) { if (rec) k_render_exhaustive<true, true><<<grid, 256, 0, st>>>(c->sc, w, c->d_counters.as<Counters>());
		           else k_render_exhaustive<true, false><<<grid, 256, 0, st>>>(c->sc, w, c->d_counters.as<Counters>()); }
		else     { if (rec) k_render_exhaustive<false, true><<<grid, 256, 0, st>>>(c->sc, w, c->d_counters.as<Counters>());
		           else k_render_exhaustive<false, false><<<grid, 256, 0, st>>>(c->sc, w, c->d_counters.as<Counters>()); }
		return cudaGetLastError();
	}
	if (c->rays_per_thread == 0 && !top && !w.frustum && (uint64_t)w.num_units * 32ull < 0xfff00000ull) {      /* persistent refill kernel on primary rays */
		RayWork none{};
		if (cnt) return rec ? launch_pt<true, true, 1>(c, none, w, st) : launch_pt<true, false, 1>(c, none, w, st);
		return rec ? launch_pt<false, true, 1>(c, none, w, st) : launch_pt<false, false, 1>(c, none, w, st);
	}
	if ((w.frustum || c->rays_per_thread > 1) && !top) {
		if (w.store_image) return launch_packet<false, false, true>(c, w, st);        /* rtx_render_store checked cnt / rec */
		if (cnt) return rec ? launch_packet<true, true>(c, w, st) : launch_packet<true, false>(c, w, st);
		return rec ? launch_packet<false, true>(c, w, st) : launch_packet<false, false>(c, w, st);
	}
#define RTX_DISPATCH(T, C, R) if (top == T && cnt == C && rec == R) return launch_render_v<T, C, R>(c, w, st)
	RTX_DISPATCH(false, false, false); RTX_DISPATCH(false, false, true);
	RTX_DISPATCH(false, true, false);  RTX_DISPATCH(false, true, true);
	RTX_DISPATCH(true, false, false);  RTX_DISPATCH(true, false, true);
	RTX_DISPATCH(true, true, false);   RTX_DISPATCH(true, true, true);
#undef RTX_DISPATCH
	return cudaErrorInvalidValue;
}

template <bool TOP, bool COUNT>
cudaError_t launch_rays_t(rtx_ctx *c, const RayWork &w, cudaStream_t st)
{
	constexpr int BLOCK = 256, MINB = 3, SST = 8;
	auto k = k_trace_rays<BLOCK, MINB, SST, TOP, COUNT>;
	const size_t smem = (size_t)SST * BLOCK * sizeof(uint2) + (TOP ? (size_t)c->sc.top_pairs * 64 : 0);
	int occ = 0;
	cudaError_t e = resident_blocks(k, BLOCK, smem, c->device, &occ);
	if (e != cudaSuccess) return e;
	if (c->blocks_per_sm > 0 && c->blocks_per_sm < occ) occ = c->blocks_per_sm;
	k<<<(unsigned)(c->sm_count * occ), BLOCK, smem, st>>>(c->sc, w, c->d_counters.as<Counters>());
	return cudaGetLastError();
}

cudaError_t launch_rays(rtx_ctx *c, const RayWork &w, cudaStream_t st)
{
	const bool top = c->top_smem > 0 && c->sc.top_pairs > 0, cnt = c->counters != 0;
	if (c->incoherent_kernel && !top) {
		Work none{};
		return cnt ? launch_pt<true, false, 0>(c, w, none, st) : launch_pt<false, false, 0>(c, w, none, st);
	}
	if (top) return cnt ? launch_rays_t<true, true>(c, w, st) : launch_rays_t<true, false>(c, w, st);
	return cnt ? launch_rays_t<false, true>(c, w, st) : launch_rays_t<false, false>(c, w, st);
}

void launch_resize_u8(const float *src, uint32_t W, uint32_t w, uint32_t h, uint32_t n, unsigned char *out, dim3 grid, dim3 block, cudaStream_t st)
{
	const bool aligned = W == w * n && (reinterpret_cast<uintptr_t>(src) & 15u) == 0;
	if (n == 4 && aligned) k_resize_u8_vec<4><<<grid, block, 0, st>>>(src, W, w, h, out);
	else if (n == 2 && aligned) k_resize_u8_vec<2><<<grid, block, 0, st>>>(src, W, w, h, out);
	else k_resize_u8<<<grid, block, 0, st>>>(src, W, w, h, n, out);
}

int finish_stats(rtx_ctx *c)
{
	if (c->phase_timing && c->ph_rec[0]) {
		/* phase k lasts from the last mark recorded at or before k to the next recorded mark */
		int prev = 0;
		for (int k = 0; k < RTX_NUM_PHASES; ++k) c->phase_ms[k] = 0.0;
		for (int k = 1; k <= RTX_NUM_PHASES; ++k) {
			if (!c->ph_rec[k]) continue;
			float ms = 0.f;
			if (cudaEventElapsedTime(&ms, c->ph_ev[prev], c->ph_ev[k]) == cudaSuccess) c->phase_ms[prev] = ms;
			prev = k;
		}
		for (int k = 0; k <= RTX_NUM_PHASES; ++k) c->ph_rec[k] = false;
	}
	if (c->ev_pending) {
		float ms = 0.f;
		if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess) c->stats.kernel_ms = ms;
		c->ev_pending = false;
	}
	if (c->counters) {
		Counters h{};
		CU(c, cudaMemcpy(&h, c->d_counters.p, sizeof h, cudaMemcpyDeviceToHost));
		c->stats.node_visits = h.node_visits;
		c->stats.tri_tests = h.tri_tests;
		c->stats.leafbox_tests = h.leafbox_tests;
		c->stats.exact_path_rays = h.exact_rays;
		c->stats.packet_overflows = h.overflow_packets;
	}
	return RTX_OK;
}

} /* namespace */

extern "C" {

const char *rtx_last_error(const rtx_ctx *ctx) { return ctx ? ctx->error.c_str() : g_error.c_str(); }

int rtx_device_count(int *count)
{
	if (!count) return fail(nullptr, RTX_ERR_ARG, "null count");
	*count = 0;
	cudaError_t e = cudaGetDeviceCount(count);
	if (e != cudaSuccess) { *count = 0; return cuda_fail(nullptr, e, "cudaGetDeviceCount"); }
	return *count > 0 ? RTX_OK : fail(nullptr, RTX_ERR_NO_DEVICE, "No device found");
}

int rtx_device_info(int device, rtx_device_info_t *info)
{
	if (!info) return fail(nullptr, RTX_ERR_ARG, "null info");
	std::memset(info, 0, sizeof *info);
	cudaDeviceProp p;
	CU(nullptr, cudaGetDeviceProperties(&p, device));
	std::snprintf(info->name, sizeof info->name, "%s", p.name);
	info->cc_major = p.major;
	info->cc_minor = p.minor;
	info->sm_count = p.multiProcessorCount;
	cudaDeviceGetAttribute(&info->clock_khz, cudaDevAttrClockRate, device);
	cudaDeviceGetAttribute(&info->mem_clock_khz, cudaDevAttrMemoryClockRate, device);
	info->mem_bus_bits = p.memoryBusWidth;
	info->global_mem_bytes = p.totalGlobalMem;
	info->l2_bytes = (uint64_t)p.l2CacheSize;
	info->smem_per_sm_bytes = p.sharedMemPerMultiprocessor;
	info->smem_per_block_optin_bytes = p.sharedMemPerBlockOptin;
	info->max_threads_per_sm = p.maxThreadsPerMultiProcessor;
	info->regs_per_sm = p.regsPerMultiprocessor;
	cudaDriverGetVersion(&info->driver_version);
	cudaRuntimeGetVersion(&info->runtime_version);
	return RTX_OK;
}

/* opencl_host.cc:76-119 prints a tree of platforms and devices; the shadow
 * OpenCLHost::printInfo() formats rtx_device_info through the reference's own
 * Info class.  This plain-text form is for C callers. */
int rtx_print_info(void)
{
	int n = 0;
	const int rc = rtx_device_count(&n);
	if (rc != RTX_OK) { std::printf("Hardware information: no CUDA device (%s)\n", g_error.c_str()); return rc; }
	std::printf("*** Hardware information ***\n");
	for (int d = 0; d < n; ++d) {
		rtx_device_info_t i;
		if (rtx_device_info(d, &i) != RTX_OK) continue;
		std::printf("Device #%d\n  Name................... %s\n  Compute capability..... %d.%d\n  Max compute units...... %d\n"
		            "  Global memory (MiB).... %llu\n  L2 cache (MiB)......... %llu\n  Local memory size (B).. %llu\n"
		            "  Driver / runtime....... %d / %d\n",
		            d, i.name, i.cc_major, i.cc_minor, i.sm_count, (unsigned long long)(i.global_mem_bytes >> 20),
		            (unsigned long long)(i.l2_bytes >> 20), (unsigned long long)i.smem_per_block_optin_bytes,
		            i.driver_version, i.runtime_version);
	}
	return RTX_OK;
}

int rtx_tile_layout(uint32_t total_width, uint32_t total_height, uint32_t world, uint32_t *tiles_x, uint32_t *tiles_y,
                    uint32_t *tiles_per_rank)
{
	if (world == 0) world = 1;
	const uint32_t tx = (total_width + RTX_TILE - 1) / RTX_TILE, ty = (total_height + RTX_TILE - 1) / RTX_TILE;
	if (tiles_x) *tiles_x = tx;
	if (tiles_y) *tiles_y = ty;
	if (tiles_per_rank) *tiles_per_rank = (uint32_t)(((uint64_t)tx * ty + world - 1) / world);
	return RTX_OK;
}

int rtx_create(rtx_ctx **out, const rtx_options *options)
{
	if (!out || !options) return fail(nullptr, RTX_ERR_ARG, "null argument");
	*out = nullptr;
	const bool ao = options->enable_ao && options->ao_num_samples > 0;       /* intersect_kernel.cl:305 */
	uint64_t ring_cap = 0;
	if (ao) {
		if (options->ao_method != 0 && options->ao_method != 1)
			return fail(nullptr, RTX_ERR_ARG, "ambient occlusion method must be 0 (uniform) or 1 (random)");
		if (options->ao_method == 0) {
			/* rays per ring <= 2 pi / step + 1 with step = alpha_max / samples (intersect_kernel.cl:238-242) */
			if (options->ao_alpha_max <= 0 || options->ao_alpha_max > 360)
				return fail(nullptr, RTX_ERR_ARG, "uniform ambient occlusion needs 0 < alpha_max <= 360 degrees");
			const double step = (double)options->ao_alpha_max * 3.14159265358979323846 / 180.0 / (double)options->ao_num_samples;
			/* ray_count = (uint)(2 pi cos(angle) / step) (intersect_kernel.cl:240) must be >= 1 on every ring: at
			 * angle >= 90 degrees the cosine is <= 0, the conversion of a negative double to uint is undefined in
			 * OpenCL C as in C, and phi = 0/0 (:243).  The outermost ring has the largest angle. */
			const double worst = ((double)options->ao_alpha_max * (double)(options->ao_num_samples - 1) / (double)options->ao_num_samples
			                      + (double)options->ao_alpha_min) * 3.14159265358979323846 / 180.0;
			if (options->ao_alpha_min < 0 || !(worst < 1.5707) || !(6.2831853071795865 * std::cos(worst) / step >= 1.001))
				return fail(nullptr, RTX_ERR_ARG, "uniform ambient occlusion: the outermost ring (alpha_max*(samples-1)/samples + alpha_min) "
				                                  "must stay below 90 degrees and hold at least one ray");
			ring_cap = (uint64_t)options->ao_num_samples * ((uint64_t)(6.2831853071795865 / step) + 2);
			if (ring_cap > (1u << 22)) return fail(nullptr, RTX_ERR_ARG, "too many ambient occlusion rings");
		}
	}
	if (options->width == 0 || options->height == 0) return fail(nullptr, RTX_ERR_ARG, "empty image");
	int ndev = 0;
	const int rc = rtx_device_count(&ndev);
	if (rc != RTX_OK) return rc;
	if (options->device < 0 || options->device >= ndev) return fail(nullptr, RTX_ERR_NO_DEVICE, "device ordinal out of range");

	rtx_ctx *c = new rtx_ctx;
	c->opt = *options;
	c->device = options->device;
	c->focal = focal_roundtrip(options->focal_length);
	c->ao = ao;
	c->ao_max_distance = focal_roundtrip(options->ao_max_distance);              /* same -D round trip (opencl_host.cc:49) */
	c->ao_ring_cap = (uint32_t)ring_cap;
	const unsigned n = (unsigned)std::sqrt((double)options->n_super_samples);   /* ray_tracer.h:33-34 */
	c->W = options->total_width ? options->total_width : options->width * n;
	c->H = options->total_height ? options->total_height : options->height * n;
	c->world = options->tile_world > 1 ? options->tile_world : 1;
	c->rank = c->world > 1 ? options->tile_rank : 0;
	if (c->rank >= c->world || c->W == 0 || c->H == 0) { delete c; return fail(nullptr, RTX_ERR_ARG, "bad tile partition or image size"); }
	rtx_tile_layout(c->W, c->H, c->world, &c->tiles_x, &c->tiles_y, &c->tiles_per_rank);
	const uint64_t ntiles = (uint64_t)c->tiles_x * c->tiles_y;
	c->local_tiles = (uint32_t)((ntiles + c->world - 1 - c->rank) / c->world);
	if (ntiles * 32 >= 0xffffffffull) { delete c; return fail(nullptr, RTX_ERR_ARG, "image too large"); }

	auto bail = [&](cudaError_t e, const char *what) { const int r = cuda_fail(nullptr, e, what); rtx_destroy(c); return r; };
	cudaError_t e;
	if ((e = cudaSetDevice(c->device)) != cudaSuccess) return bail(e, "cudaSetDevice");
	cudaDeviceProp p;
	if ((e = cudaGetDeviceProperties(&p, c->device)) != cudaSuccess) return bail(e, "cudaGetDeviceProperties");
	c->sm_count = p.multiProcessorCount;
	if (p.major < 10) {
		rtx_destroy(c);
		return fail(nullptr, RTX_ERR_NO_DEVICE, std::string("kernels are built for sm_100a only; device is ") + p.name);
	}
	if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
	if ((e = cudaEventCreate(&c->ev0)) != cudaSuccess) return bail(e, "cudaEventCreate");
	if ((e = cudaEventCreate(&c->ev1)) != cudaSuccess) return bail(e, "cudaEventCreate");
	const size_t out_floats = c->world > 1 ? (size_t)c->tiles_per_rank * RTX_TILE * RTX_TILE : (size_t)c->W * c->H;
	if ((e = c->d_image.alloc(out_floats * sizeof(float))) != cudaSuccess) return bail(e, "cudaMalloc(image)");
	if ((e = c->d_counter.alloc(4 * sizeof(unsigned int))) != cudaSuccess) return bail(e, "cudaMalloc(counter)");
	if ((e = c->d_counters.alloc(sizeof(Counters))) != cudaSuccess) return bail(e, "cudaMalloc(counters)");
	if ((e = c->d_sums.alloc(2 * sizeof(unsigned long long))) != cudaSuccess) return bail(e, "cudaMalloc(sums)");
	if ((e = cudaMemsetAsync(c->d_image.p, 0, out_floats * sizeof(float), c->stream)) != cudaSuccess) return bail(e, "cudaMemset");
	*out = c;
	return RTX_OK;
}

void rtx_destroy(rtx_ctx *c)
{
	if (!c) return;
	cudaSetDevice(c->device);
	if (c->stream) cudaStreamSynchronize(c->stream);
	DevBuf *bufs[] = { &c->t_parent, &c->t_build, &c->d_triangles, &c->t_faces, &c->t_verts, &c->t_vnormals, &c->t_scan, &c->d_pairs, &c->d_pleafbox, &c->d_tris, &c->d_leafbox, &c->d_tnormals, &c->d_ref_nodes, &c->d_ref_aabbs,
	                   &c->d_image, &c->d_image_full, &c->d_face_id, &c->d_dist, &c->d_u8, &c->d_counter, &c->d_counters, &c->d_sums, &c->d_lists, &c->d_slists, &c->d_raytab,
	                   &c->d_hit_st, &c->d_ao_ring, &c->d_tile_done };
	for (DevBuf *b : bufs) b->release();
	if (c->ev0) cudaEventDestroy(c->ev0);
	if (c->ev1) cudaEventDestroy(c->ev1);
	for (cudaEvent_t e : c->band_ev) if (e) cudaEventDestroy(e);
	for (cudaEvent_t e : c->ph_ev) if (e) cudaEventDestroy(e);
	if (c->raytab_ev) cudaEventDestroy(c->raytab_ev);
	for (int k = 0; k < 2; ++k) {
		c->r_o[k].release(); c->r_d[k].release(); c->r_f[k].release(); c->r_t[k].release();
		if (c->r_ev_in[k]) cudaEventDestroy(c->r_ev_in[k]);
		if (c->r_ev_done[k]) cudaEventDestroy(c->r_ev_done[k]);
		if (c->r_ev_out[k]) cudaEventDestroy(c->r_ev_out[k]);
	}
	if (c->r_ev_a) cudaEventDestroy(c->r_ev_a);
	if (c->r_ev_b) cudaEventDestroy(c->r_ev_b);
	if (c->r_in) cudaStreamDestroy(c->r_in);
	if (c->r_out) cudaStreamDestroy(c->r_out);
	if (c->copy_done) cudaEventDestroy(c->copy_done);
	if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
	if (c->stream) cudaStreamDestroy(c->stream);
	delete c;
}

int rtx_set_tunable(rtx_ctx *c, int which, int64_t v)
{
	if (!c) return fail(nullptr, RTX_ERR_ARG, "null context");
	switch (which) {
	case RTX_TUNE_KERNEL:
		if (v != RTX_KERNEL_PERSISTENT && v != RTX_KERNEL_EXHAUSTIVE) return fail(c, RTX_ERR_ARG, "unknown kernel");
		c->kernel = (int)v; break;
	case RTX_TUNE_LEAF_SIZE:
		if (v < 1 || v > 8) return fail(c, RTX_ERR_ARG, "leaf size must be 1..8");
		c->leaf_size = (int)v; break;
	case RTX_TUNE_RECORD_HITS: c->record_hits = v != 0; break;
	case RTX_TUNE_COUNTERS: c->counters = v != 0; break;
	case RTX_TUNE_TOP_SMEM:
		if (v < 0 || v > 1024) return fail(c, RTX_ERR_ARG, "top_smem must be 0..1024 pairs");
		c->top_smem = (int)v; break;
	case RTX_TUNE_BLOCKS_PER_SM: c->blocks_per_sm = (int)v; break;
	case RTX_TUNE_FLATTEN_ON_DEVICE: c->flatten_on_device = v != 0; break;
	case RTX_TUNE_LIST_RAYS_PER_THREAD:
		if (v != 1 && v != 2 && v != 4) return fail(c, RTX_ERR_ARG, "list rays per thread must be 1, 2 or 4");
		c->list_rays_per_thread = (int)v; break;
	case RTX_TUNE_INCOHERENT_KERNEL: c->incoherent_kernel = v != 0; break;
	case RTX_TUNE_RAY_TABLES: c->ray_tables = v != 0; break;
	case RTX_TUNE_FRUSTUM:
		if (v < -1 || v > 1) return fail(c, RTX_ERR_ARG, "frustum must be -1, 0 or 1");
		c->frustum = (int)v; break;
	case RTX_TUNE_PHASE_TIMING: c->phase_timing = v != 0; break;
	case RTX_TUNE_RAYS_PER_THREAD:
		if (v != 0 && v != 1 && v != 2 && v != 4) return fail(c, RTX_ERR_ARG, "rays per thread must be 0 (refill kernel), 1, 2 or 4");
		c->rays_per_thread = (int)v; break;
	default: return fail(c, RTX_ERR_ARG, "unknown tunable");
	}
	return RTX_OK;
}

/* Everything after the five reference arrays are in device memory (copied by rtx_upload or built by
 * rtx_upload_mesh): validation, prefix counts, flatten, culling slack.  faces/nodes/aabbs16 are the host copies the
 * host flatten reads; NULL when the tree only exists on the device.  nodes0 = nodes[0], root_box8 = the root's
 * (min, max) vectors. */
static int upload_finish(rtx_ctx *c, size_t nfaceidx, size_t nnodes, size_t nverts, uint32_t nodes0, const float *root_box8,
                         const uint32_t *faces, const uint32_t *nodes, const float *aabbs16)
{
#define CUU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(c, e_, #call); } while (0)
	const bool device_flatten = c->flatten_on_device && c->top_smem == 0;
	const size_t ntris = nfaceidx / 3;
	cudaStream_t st = c->stream;
	if (!device_flatten && !nodes) return fail(c, RTX_ERR_UNSUPPORTED, "the host flatten (RTX_TUNE_TOP_SMEM / RTX_TUNE_FLATTEN_ON_DEVICE = 0) needs the tree on the host: use rtx_upload");
	std::string why;
	size_t num_pairs = 0, pair_stride = 0;
	uint32_t depth = 0, top_pairs = 0;
	if (device_flatten) {
		/* The tree's invariants, the two prefix counts k_flatten_nodes needs and the depth of the
		 * flattened tree are all computed on the device from the raw arrays (k_tree_*): no host pass over the nodes. */
		if (nnodes == 0 || nodes0 != nnodes || (nnodes & 1) == 0 || leaves_of((uint32_t)nnodes) != ntris)
			return cudaStreamSynchronize(st), fail(c, RTX_ERR_ARG, "malformed BVH: node count must be 2*triangles-1 and nodes[0] must equal it");
		const uint32_t K = (uint32_t)c->leaf_size, n = (uint32_t)nnodes;
		pair_stride = ntris > 1 ? ntris - 1 : 1;                       /* internal nodes of a full binary tree: the bound for any K */
		const uint32_t per_block = RTX_SCAN_BLOCK * RTX_SCAN_ITEMS, nblocks = (n + per_block - 1) / per_block;
		CUU(c->t_scan.alloc(((size_t)3 * n + (size_t)3 * nblocks) * 4 + sizeof(TreeResult) + 16));
		uint32_t *first_leaf = c->t_scan.as<uint32_t>(), *pair_idx = first_leaf + n;
		int *delta = reinterpret_cast<int *>(pair_idx + n), *partials = delta + n;
		TreeResult *res = reinterpret_cast<TreeResult *>(partials + (size_t)3 * nblocks);
		CUU(c->d_pairs.alloc(4 * pair_stride * 64));
		CUU(c->t_parent.alloc(pair_stride * 4));
		c->h_tree = TreeResult{ 0xffffffffu, 0u, n == 1 ? 1u : 0u, n == 1 ? 1u : 0u, 0u };
		CUU(cudaMemcpyAsync(res, &c->h_tree, sizeof(TreeResult), cudaMemcpyHostToDevice, st));
		if (n > 1) {
			CUU(cudaMemsetAsync(delta, 0, (size_t)n * 4, st));
			k_tree_check<<<(n + 255) / 256, 256, 0, st>>>(c->d_ref_nodes.as<uint32_t>(), c->d_ref_aabbs.as<float4>(), n, K, delta, res);
			k_tree_partials<<<nblocks, RTX_SCAN_BLOCK, 0, st>>>(c->d_ref_nodes.as<uint32_t>(), delta, n, K, partials);
			k_tree_spine<<<1, 96, 0, st>>>(partials, nblocks, res);
			k_tree_scan<<<nblocks, RTX_SCAN_BLOCK, 0, st>>>(c->d_ref_nodes.as<uint32_t>(), delta, n, K, partials, first_leaf, pair_idx, res);
		} else {
			CUU(cudaMemsetAsync(first_leaf, 0, 2 * sizeof(uint32_t), st));
		}
		k_flatten_nodes<<<(unsigned)((nnodes + 255) / 256), 256, 0, st>>>(
			c->d_ref_nodes.as<uint32_t>(), c->d_ref_aabbs.as<float4>(), first_leaf, pair_idx,
			c->t_faces.as<uint32_t>(), c->t_verts.as<float4>(), c->t_vnormals.as<float4>(), n, (uint32_t)pair_stride, K,
			c->d_pairs.as<float4>(), c->d_tris.as<float4>(), c->d_leafbox.as<float4>(), c->d_tnormals.as<float4>(), (uint32_t)nverts, res,
			c->t_parent.as<uint32_t>(), c->d_pleafbox.as<float4>());
		CUU(cudaGetLastError());
		CUU(cudaMemcpyAsync(&c->h_tree, res, sizeof(TreeResult), cudaMemcpyDeviceToHost, st));
		CUU(cudaStreamSynchronize(st));
		if (c->h_tree.bad_node != 0xffffffffu)
			return fail(c, RTX_ERR_ARG, "malformed BVH: subtree sizes do not form a pre-order binary tree at node " + std::to_string(c->h_tree.bad_node));
		if (c->h_tree.bad_face) return fail(c, RTX_ERR_ARG, "face index out of range");
		num_pairs = c->h_tree.num_pairs;
		depth = c->h_tree.depth;
		c->boxes_nested = c->h_tree.loose == 0;
	} else {
		/* host flatten (needed for the breadth-first prefix that RTX_TUNE_TOP_SMEM stages in shared memory) */
		if (!validate_tree(nodes, nnodes, ntris, why)) return cudaStreamSynchronize(st), fail(c, RTX_ERR_ARG, "malformed BVH: " + why);
		uint32_t bad = 0;                          /* branch-free range check of the vertex indices */
		for (size_t i = 0; i < nfaceidx; ++i) bad |= (uint32_t)(faces[i] >= nverts);
		if (bad) return cudaStreamSynchronize(st), fail(c, RTX_ERR_ARG, "face index out of range");
		c->boxes_nested = true;                    /* same containment check as k_tree_check */
		for (size_t i = 0; i < nnodes && c->boxes_nested; ++i) {
			if (nodes[i] == 1) continue;
			const float *p = aabbs16 + 8 * i;
			const size_t kids[2] = { i + 1, i + 1 + nodes[i + 1] };
			for (size_t ch : kids) {
				const float *q = aabbs16 + 8 * ch;
				if (!(q[0] >= p[0] && q[1] >= p[1] && q[2] >= p[2] && q[4] <= p[4] && q[5] <= p[5] && q[6] <= p[6])) c->boxes_nested = false;
			}
		}
		Flat flat;
		flatten(nodes, aabbs16, nnodes, c->leaf_size, c->top_smem > 0 ? (uint32_t)c->top_smem : 0u, flat);
		const size_t pair_vecs = flat.pairs.size();
		flat.pairs.resize(4 * pair_vecs);
		for (int v = 1; v < 4; ++v) {       /* octant copies: swap lo/hi of x (bit 0) and of y (bit 1) */
			float4 *dst = flat.pairs.data() + (size_t)v * pair_vecs;
			for (size_t i = 0; i < pair_vecs; i += 2) {
				float4 a = flat.pairs[i], b = flat.pairs[i + 1];     /* a = (lo.x lo.y lo.z hi.x), b = (hi.y hi.z ref 0) */
				if (v & 1) { const float t = a.x; a.x = a.w; a.w = t; }
				if (v & 2) { const float t = a.y; a.y = b.x; b.x = t; }
				dst[i] = a;
				dst[i + 1] = b;
			}
		}
		num_pairs = pair_stride = pair_vecs / 4;
		depth = flat.depth;
		top_pairs = flat.top_pairs;
		CUU(c->t_scan.alloc(ntris * 4));
		CUU(c->d_pairs.alloc(flat.pairs.size() * 16));
		CUU(cudaMemcpyAsync(c->t_scan.p, flat.leaf_node.data(), ntris * 4, cudaMemcpyHostToDevice, st));
		CUU(cudaMemcpyAsync(c->d_pairs.p, flat.pairs.data(), flat.pairs.size() * 16, cudaMemcpyHostToDevice, st));
		k_build_triangles<<<(unsigned)((ntris + 255) / 256), 256, 0, st>>>(
			c->t_faces.as<uint32_t>(), c->t_verts.as<float4>(), c->t_vnormals.as<float4>(), c->t_scan.as<uint32_t>(),
			c->d_ref_aabbs.as<float4>(), (uint32_t)ntris, c->d_tris.as<float4>(), c->d_leafbox.as<float4>(), c->d_tnormals.as<float4>(),
			c->d_pleafbox.as<float4>());
		CUU(cudaGetLastError());
		CUU(cudaStreamSynchronize(st));     /* flat's vectors die at the end of this block */
	}
	CUU(cudaStreamSynchronize(st));         /* the caller frees its arrays right after (render.cc:96-103) */
	{
		/* culling slack of the leaves from the conditioning of their triangles (box_slack); ill-shaped triangles
		 * are rare, and only then is the slack pushed up the tree, one level per pass */
		const unsigned grid = (unsigned)((num_pairs + 255) / 256);
		if (device_flatten) {
			/* parent links from the flatten: every fat leaf raises its ancestors itself, one launch, no round trip */
			k_slack_leaves<<<grid, 256, 0, st>>>(c->d_pairs.as<float4>(), c->d_tris.as<float4>(), (uint32_t)num_pairs, (uint32_t)pair_stride,
			                                      c->t_parent.as<uint32_t>(), nullptr);
			CUU(cudaGetLastError());
			CUU(cudaStreamSynchronize(st));
		} else {
			unsigned int *fat = c->d_counter.as<unsigned int>() + 3;
			CUU(cudaMemsetAsync(fat, 0, sizeof(unsigned int), st));
			k_slack_leaves<<<grid, 256, 0, st>>>(c->d_pairs.as<float4>(), c->d_tris.as<float4>(), (uint32_t)num_pairs, (uint32_t)pair_stride, nullptr, fat);
			CUU(cudaGetLastError());
			unsigned int h_fat = 0;
			CUU(cudaMemcpyAsync(&h_fat, fat, sizeof h_fat, cudaMemcpyDeviceToHost, st));
			CUU(cudaStreamSynchronize(st));
			if (h_fat) {                 /* no parent links in the host flatten: one level per pass */
				for (uint32_t pass = 0; pass < depth; ++pass)
					k_slack_relax<<<grid, 256, 0, st>>>(c->d_pairs.as<float4>(), (uint32_t)num_pairs, (uint32_t)pair_stride);
				CUU(cudaGetLastError());
				CUU(cudaStreamSynchronize(st));
			}
		}
	}
#undef CUU

	c->sc.pairs = c->d_pairs.as<float4>();
	c->sc.tris = c->d_tris.as<float4>();
	c->sc.leafbox = c->d_leafbox.as<float4>();
	c->sc.pleafbox = c->d_pleafbox.as<float4>();
	c->sc.tnormals = c->d_tnormals.as<float4>();
	c->sc.ref_nodes = c->d_ref_nodes.as<uint32_t>();
	c->sc.ref_aabbs = c->d_ref_aabbs.as<float4>();
	c->sc.num_pairs = (uint32_t)pair_stride;            /* distance between the octant copies */
	c->sc.pair_count = (uint32_t)num_pairs;
	c->sc.top_pairs = c->top_smem > 0 ? top_pairs : 0;
	c->sc.num_tris = (uint32_t)ntris;
	c->sc.verify_leafbox = c->leaf_size > 1 ? 1u : 0u;
	c->bbmin = make_f3(root_box8[0], root_box8[1], root_box8[2]);
	c->bbmax = make_f3(root_box8[4], root_box8[5], root_box8[6]);
	float scale = 0.f;
	for (int k = 0; k < 3; ++k) scale = std::fmax(scale, std::fmax(std::fabs(root_box8[k]), std::fabs(root_box8[4 + k])));
	c->sc.scene_scale = scale;
	c->tree_depth = depth;
	c->stats.tree_depth = depth;
	c->stats.num_pairs = (uint32_t)num_pairs;
	c->uploaded = true;
	c->rendered = false;
	return RTX_OK;
}

int rtx_upload(rtx_ctx *c, const uint32_t *faces, size_t nfaceidx, const uint32_t *nodes, size_t nnodes,
               const float *aabbs16, size_t naabbvec, const float *verts16, size_t nverts,
               const float *vnormals16, size_t nnormals)
{
	if (!c) return fail(nullptr, RTX_ERR_ARG, "null context");
	if (!faces || !nodes || !aabbs16 || !verts16 || !vnormals16) return fail(c, RTX_ERR_ARG, "null array");
	if (nfaceidx == 0 || nfaceidx % 3 != 0) return fail(c, RTX_ERR_ARG, "faces must hold 3 indices per triangle");
	const size_t ntris = nfaceidx / 3;
	if (naabbvec != 2 * nnodes) return fail(c, RTX_ERR_ARG, "aabbs must hold 2 vectors per node");
	if (nverts != nnormals || nverts == 0) return fail(c, RTX_ERR_ARG, "one normal per vertex required");
	if (ntris >= (1u << 28)) return fail(c, RTX_ERR_ARG, "too many triangles (limit 2^28)");
	if (nnodes == 0) return fail(c, RTX_ERR_ARG, "malformed BVH: no nodes");       /* nodes[0] / aabbs16[0..7] are read below */
	CU(c, cudaSetDevice(c->device));
	c->uploaded = false;
	c->tree_on_device = false;
	cudaStream_t st = c->stream;
#define CUU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(c, e_, #call); } while (0)
	CUU(c->t_faces.alloc(nfaceidx * 4));
	CUU(c->t_verts.alloc(nverts * 16));
	CUU(c->t_vnormals.alloc(nverts * 16));
	CUU(c->d_ref_nodes.alloc(nnodes * 4));
	CUU(c->d_ref_aabbs.alloc(nnodes * 32));
	CUU(c->d_tris.alloc(ntris * 64));
	CUU(c->d_leafbox.alloc(ntris * 32));
	CUU(c->d_pleafbox.alloc(ntris * 32));
	CUU(c->d_tnormals.alloc(ntris * 48));
	CUU(cudaMemcpyAsync(c->t_faces.p, faces, nfaceidx * 4, cudaMemcpyHostToDevice, st));
	CUU(cudaMemcpyAsync(c->t_verts.p, verts16, nverts * 16, cudaMemcpyHostToDevice, st));
	CUU(cudaMemcpyAsync(c->t_vnormals.p, vnormals16, nverts * 16, cudaMemcpyHostToDevice, st));
	CUU(cudaMemcpyAsync(c->d_ref_nodes.p, nodes, nnodes * 4, cudaMemcpyHostToDevice, st));
	CUU(cudaMemcpyAsync(c->d_ref_aabbs.p, aabbs16, nnodes * 32, cudaMemcpyHostToDevice, st));
	/* the copies above are in flight while the device validates and scans */
	return upload_finish(c, nfaceidx, nnodes, nverts, nodes[0], aabbs16, faces, nodes, aabbs16);
#undef CUU
}

/* Mesh in, everything else on the device: the reference's longest-axis builder (bvh.cc:59-162) level by level
 * (rtx_build.cuh), then the same validation / flatten as rtx_upload.  Emits the arrays bvh.cc would (rtx_download_tree). */
int rtx_upload_mesh(rtx_ctx *c, const float *verts16, size_t nverts, const uint32_t *faces, size_t nfaces, const float *vnormals16)
{
	if (!c) return fail(nullptr, RTX_ERR_ARG, "null context");
	if (!verts16 || !faces) return fail(c, RTX_ERR_ARG, "null array");
	if (nfaces == 0 || nverts == 0) return fail(c, RTX_ERR_ARG, "empty mesh");
	if (nverts >= 0xfffffff0ull) return fail(c, RTX_ERR_ARG, "too many vertices");
	if (nfaces >= (1u << 28)) return fail(c, RTX_ERR_ARG, "too many triangles (limit 2^28)");
	if (!(c->flatten_on_device && c->top_smem == 0)) return fail(c, RTX_ERR_UNSUPPORTED, "rtx_upload_mesh needs the device flatten (no RTX_TUNE_TOP_SMEM)");
	CU(c, cudaSetDevice(c->device));
	c->uploaded = false;
	c->tree_on_device = false;
	cudaStream_t st = c->stream;
	const uint32_t N = (uint32_t)nfaces;
	const size_t nnodes = 2 * (size_t)N - 1;
#define CUU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(c, e_, #call); } while (0)
	CUU(c->t_faces.alloc((size_t)N * 12));
	CUU(c->t_verts.alloc(nverts * 16));
	CUU(c->t_vnormals.alloc(nverts * 16));
	CUU(c->d_ref_nodes.alloc(nnodes * 4));
	CUU(c->d_ref_aabbs.alloc(nnodes * 32));
	CUU(c->d_triangles.alloc((size_t)N * 4));
	CUU(c->d_tris.alloc((size_t)N * 64));
	CUU(c->d_leafbox.alloc((size_t)N * 32));
	CUU(c->d_pleafbox.alloc((size_t)N * 32));
	CUU(c->d_tnormals.alloc((size_t)N * 48));
	/* work space: input faces, per-triangle centroid / lo / hi, two copies of (ids, segments, accumulators), scan, partials, flags */
	const uint32_t per_block = RTX_BVH_BLOCK * RTX_BVH_ITEMS, nblocks = (N + per_block - 1) / per_block;
	const size_t words = (size_t)N * (3 + 9 + 2 * (1 + 3 + RTX_BVH_ACC) + 1) + 1 + nblocks + 8;
	CUU(c->t_build.alloc(words * 4));
	uint32_t *p = c->t_build.as<uint32_t>();
	uint32_t *in_faces = p;                      p += (size_t)N * 3;
	float *tc = reinterpret_cast<float *>(p);    p += (size_t)N * 3;
	float *tlo = reinterpret_cast<float *>(p);   p += (size_t)N * 3;
	float *thi = reinterpret_cast<float *>(p);   p += (size_t)N * 3;
	uint32_t *ids[2]; BvhSeg *seg[2]; uint32_t *acc[2];
	for (int k = 0; k < 2; ++k) {
		ids[k] = p;                              p += N;
		seg[k] = reinterpret_cast<BvhSeg *>(p);  p += (size_t)N * 3;
		acc[k] = p;                              p += (size_t)N * RTX_BVH_ACC;
	}
	uint32_t *scan = p;                          p += (size_t)N + 1;
	uint32_t *partials = p;                      p += nblocks;
	unsigned int *live = p, *bad_face = p + 1;
	CUU(cudaMemcpyAsync(in_faces, faces, (size_t)N * 12, cudaMemcpyHostToDevice, st));
	CUU(cudaMemcpyAsync(c->t_verts.p, verts16, nverts * 16, cudaMemcpyHostToDevice, st));
	CUU(cudaMemsetAsync(live, 0, 2 * sizeof(unsigned int), st));
	CUU(cudaEventRecord(c->ev0, st));
	const unsigned grid = (N + 255) / 256;
	if (vnormals16) {
		CUU(cudaMemcpyAsync(c->t_vnormals.p, vnormals16, nverts * 16, cudaMemcpyHostToDevice, st));
	} else {
		/* compute_vertex_normals (mesh.cc:95-139) on the device, same additions in the same (face) order */
		const uint32_t V = (uint32_t)nverts, vblocks = (V + per_block - 1) / per_block;
		const size_t need = (size_t)N * 4 + (size_t)N * 3 + 3 * ((size_t)V + 1) + vblocks + 8;
		CUU(c->t_scan.alloc(need * 4));                    /* upload_finish re-uses t_scan afterwards */
		uint32_t *w = c->t_scan.as<uint32_t>();
		float4 *fnormal = reinterpret_cast<float4 *>(w);   w += (size_t)N * 4;
		uint32_t *list = w;                                w += (size_t)N * 3;
		uint32_t *count = w;                               w += (size_t)V + 1;
		uint32_t *offset = w;                              w += (size_t)V + 1;
		uint32_t *cursor = w;                              w += (size_t)V + 1;
		uint32_t *vpart = w;
		CUU(cudaMemsetAsync(count, 0, ((size_t)V + 1) * 4, st));
		CUU(cudaMemsetAsync(cursor, 0, ((size_t)V + 1) * 4, st));
		k_vn_faces<<<grid, 256, 0, st>>>(in_faces, c->t_verts.as<float4>(), N, V, fnormal, count);
		k_scan_partials<<<vblocks, RTX_BVH_BLOCK, 0, st>>>(count, V, vpart);
		k_bvh_spine<<<1, 32, 0, st>>>(vpart, vblocks);
		k_scan_apply<<<vblocks, RTX_BVH_BLOCK, 0, st>>>(count, V, vpart, offset);
		k_vn_fill<<<grid, 256, 0, st>>>(in_faces, N, V, offset, cursor, list);
		k_vn_vertices<<<(V + 127) / 128, 128, 0, st>>>(offset, list, fnormal, V, c->t_vnormals.as<float4>());
		CUU(cudaGetLastError());
	}
	k_bvh_prepare<<<grid, 256, 0, st>>>(in_faces, c->t_verts.as<float4>(), N, (uint32_t)nverts, tc, tlo, thi, ids[0], seg[0], acc[0], bad_face);
	CUU(cudaGetLastError());
	uint32_t levels = 0;
	int cur = 0;
	unsigned int h_flags[2] = { N > 1 ? 1u : 0u, 0u };
	while (h_flags[0] != 0) {
		if (++levels > N) return cudaStreamSynchronize(st), fail(c, RTX_ERR_CUDA, "BVH build did not terminate");
		/* the host looks at the live-segment count every 4th level only: a level without live segments is a no-op
		 * (every kernel returns at once), a round trip to the host costs as much as a level of a small mesh */
		const bool look = (levels & 3u) == 0 || N < 4096 || N > (1u << 20);      /* a wasted level of a big mesh costs more than the round trip */
		k_bvh_accumulate<<<(N + 256 * RTX_BVH_ACC_ITEMS - 1) / (256 * RTX_BVH_ACC_ITEMS), 256, 0, st>>>(ids[cur], seg[cur], N, tc, tlo, thi, acc[cur]);
		k_bvh_split<<<grid, 256, 0, st>>>(seg[cur], N, acc[cur], c->d_ref_nodes.as<uint32_t>(), c->d_ref_aabbs.as<float4>());
		k_bvh_flag_partials<<<nblocks, RTX_BVH_BLOCK, 0, st>>>(ids[cur], seg[cur], acc[cur], tc, N, partials);
		k_bvh_spine<<<1, 32, 0, st>>>(partials, nblocks);
		k_bvh_flag_scan<<<nblocks, RTX_BVH_BLOCK, 0, st>>>(ids[cur], seg[cur], acc[cur], tc, N, partials, scan);
		CUU(cudaMemsetAsync(live, 0, sizeof(unsigned int), st));
		k_bvh_scatter<<<grid, 256, 0, st>>>(ids[cur], seg[cur], acc[cur], tc, scan, N, ids[cur ^ 1], seg[cur ^ 1], acc[cur ^ 1], live);
		CUU(cudaGetLastError());
		if (look) {
			CUU(cudaMemcpyAsync(h_flags, live, sizeof h_flags, cudaMemcpyDeviceToHost, st));
			CUU(cudaStreamSynchronize(st));
		}
		cur ^= 1;
	}
	k_bvh_leaves<<<grid, 256, 0, st>>>(ids[cur], seg[cur], N, tlo, thi, in_faces, c->d_ref_nodes.as<uint32_t>(), c->d_ref_aabbs.as<float4>(),
	                                   c->d_triangles.as<uint32_t>(), c->t_faces.as<uint32_t>());
	CUU(cudaGetLastError());
	CUU(cudaEventRecord(c->ev1, st));
	float root_box[8];
	uint32_t nodes0 = 0;
	CUU(cudaMemcpyAsync(root_box, c->d_ref_aabbs.p, sizeof root_box, cudaMemcpyDeviceToHost, st));
	CUU(cudaMemcpyAsync(&nodes0, c->d_ref_nodes.p, sizeof nodes0, cudaMemcpyDeviceToHost, st));
	CUU(cudaMemcpyAsync(h_flags, live, sizeof h_flags, cudaMemcpyDeviceToHost, st));
	CUU(cudaStreamSynchronize(st));
	if (h_flags[1]) return fail(c, RTX_ERR_ARG, "face index out of range");
	float ms = 0.f;
	if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess) c->build_ms = ms;
	c->build_levels = levels;
	const int rc = upload_finish(c, (size_t)N * 3, nnodes, nverts, nodes0, root_box, nullptr, nullptr, nullptr);
	if (rc == RTX_OK) { c->tree_on_device = true; c->tree_nodes = nnodes; c->tree_tris = N; }
	return rc;
#undef CUU
}

/* The tree of the last upload in the reference's own formats (bvh.h:15-17 + the leaf-ordered faces of
 * render.cc:88-95); any pointer may be NULL.  nodes: 2T-1 words, aabbs16: 2(2T-1) vectors of 4 floats,
 * triangles: T words (leaf order -> input face id; only after rtx_upload_mesh), sorted_faces: 3T words. */
int rtx_download_tree(rtx_ctx *c, uint32_t *nodes, float *aabbs16, uint32_t *triangles, uint32_t *sorted_faces)
{
	if (!c) return fail(nullptr, RTX_ERR_ARG, "null context");
	if (!c->uploaded) return fail(c, RTX_ERR_STATE, "no scene uploaded");
	if (triangles && !c->tree_on_device) return fail(c, RTX_ERR_STATE, "the leaf-order triangle ids exist only after rtx_upload_mesh");
	CU(c, cudaSetDevice(c->device));
	CU(c, cudaStreamSynchronize(c->stream));
	const size_t T = c->sc.num_tris, n = 2 * T - 1;
	if (nodes) CU(c, cudaMemcpy(nodes, c->d_ref_nodes.p, n * 4, cudaMemcpyDeviceToHost));
	if (aabbs16) CU(c, cudaMemcpy(aabbs16, c->d_ref_aabbs.p, n * 32, cudaMemcpyDeviceToHost));
	if (triangles) CU(c, cudaMemcpy(triangles, c->d_triangles.p, T * 4, cudaMemcpyDeviceToHost));
	if (sorted_faces) CU(c, cudaMemcpy(sorted_faces, c->t_faces.p, T * 12, cudaMemcpyDeviceToHost));
	return RTX_OK;
}

int rtx_download_normals(rtx_ctx *c, float *vnormals16, size_t nverts)
{
	if (!c || !vnormals16) return fail(c, RTX_ERR_ARG, "null argument");
	if (!c->tree_on_device) return fail(c, RTX_ERR_STATE, "normals are kept only after rtx_upload_mesh");
	if (nverts * 16 > c->t_vnormals.bytes) return fail(c, RTX_ERR_ARG, "more vertices than were uploaded");
	CU(c, cudaSetDevice(c->device));
	CU(c, cudaStreamSynchronize(c->stream));
	CU(c, cudaMemcpy(vnormals16, c->t_vnormals.p, nverts * 16, cudaMemcpyDeviceToHost));
	return RTX_OK;
}

/* device time of the last rtx_upload_mesh build (prepare .. leaves) and the number of levels it took */
int rtx_build_stats(const rtx_ctx *c, double *build_ms, uint32_t *levels)
{
	if (!c) return fail(nullptr, RTX_ERR_ARG, "null context");
	if (build_ms) *build_ms = c->build_ms;
	if (levels) *levels = c->build_levels;
	return RTX_OK;
}

/* host_dst != NULL (rtx_render_download, whole image only): the frame is rendered in bands of tile rows and each
 * finished band -- a contiguous row range of the row-major image -- is copied to the host on a second stream
 * while the next band renders, so the device->host copy (the longer of the two at PCIe rates) hides the tracing. */
static int enqueue_render(rtx_ctx *c, cudaStream_t st, float *host_dst = nullptr)
{
	if (!c) return fail(nullptr, RTX_ERR_ARG, "null context");
	if (!c->uploaded) return fail(c, RTX_ERR_STATE, "render before upload");
	CU(c, cudaSetDevice(c->device));
	const size_t out_n = c->world > 1 ? (size_t)c->tiles_per_rank * RTX_TILE * RTX_TILE : (size_t)c->W * c->H;
	const bool rec = c->record_hits || c->ao;      /* the occlusion pass starts from the recorded hits */
	if (rec) {
		CU(c, c->d_face_id.alloc(out_n * 4));
		CU(c, c->d_dist.alloc(out_n * 4));
	}
	if (c->ao) CU(c, c->d_hit_st.alloc(out_n * sizeof(float2)));
	Work w{};
	w.cam.W = c->W;
	w.cam.H = c->H;
	w.cam.a = c->focal * (float)(c->W < c->H ? c->H : c->W);        /* intersect_kernel.cl:285 */
	w.cam.w_over_2a = (float)c->W / (2.0f * w.cam.a);               /* :287 */
	w.cam.h_over_2a = (float)c->H / (2.0f * w.cam.a);               /* :288 */
	w.cam.jitter_seed = c->opt.jitter_seed;
	w.cam.shading = c->opt.enable_shading ? 1 : 0;
	w.cam.ux = w.cam.vy = nullptr;
	const bool banded = host_dst != nullptr && c->world == 1;
	const int timing_was = c->phase_timing;
	if (banded) c->phase_timing = 0;                   /* several passes per frame: the marks would be re-recorded */
	CU(c, phase_mark(c, st, RTX_PHASE_TABLES));
	uint32_t table_launch = 0;
	if (c->ray_tables && c->opt.jitter_seed == 0) {
		/* the tables depend on the camera only, which is fixed for the context: filled by the first frame, then reused
		 * (later frames on another stream wait for the event of that first launch) */
		float *ux = c->d_raytab.as<float>(), *vy = ux ? ux + c->W : nullptr;
		if (!c->raytab_ev) {
			CU(c, c->d_raytab.alloc(((size_t)c->W + c->H) * sizeof(float)));
			ux = c->d_raytab.as<float>();
			vy = ux + c->W;
			const uint32_t m = c->W > c->H ? c->W : c->H;
			k_ray_tables<<<(m + 255) / 256, 256, 0, st>>>(w.cam, ux, vy);
			CU(c, cudaGetLastError());
			CU(c, cudaEventCreateWithFlags(&c->raytab_ev, cudaEventDisableTiming));
			CU(c, cudaEventRecord(c->raytab_ev, st));
			table_launch = 1;
		} else {
			CU(c, cudaStreamWaitEvent(st, c->raytab_ev, 0));
		}
		w.cam.ux = ux;
		w.cam.vy = vy;
	}
	w.tiles_x = c->tiles_x;
	w.tiles_y = c->tiles_y;
	w.rank = c->rank;
	w.world = c->world;
	w.local_tiles = c->local_tiles;
	w.tile_begin = 0;
	w.tile_count = c->local_tiles;
	w.num_units = c->local_tiles * 32u;
	w.counter = c->d_counter.as<unsigned int>();
	w.image = c->ext_image ? c->ext_image : c->d_image.as<float>();
	w.rowmajor = c->ext_image && c->ext_rowmajor ? 1 : 0;
	if (w.rowmajor && c->world > 1 && rec) return fail(c, RTX_ERR_UNSUPPORTED, "rtx_bind_output_image carries the float image only (no hit records, no ambient occlusion)");
	w.face_id = rec ? c->d_face_id.as<uint32_t>() : nullptr;
	w.dist = rec ? c->d_dist.as<float>() : nullptr;
	w.hit_st = c->ao ? c->d_hit_st.as<float2>() : nullptr;
	w.ordered_ok = (c->tree_depth <= RTX_STACK_MAX && c->boxes_nested) ? 1 : 0;
	w.cost_map = std::getenv("RTX_EXP_COSTMAP") ? 1 : 0;
	/* frustum front end pays off when packets see few triangles: many rays per triangle */
	const double rays_per_tri = (double)c->W * c->H / (double)c->sc.num_tris;
	w.frustum = (c->frustum == 1 || (c->frustum < 0 && rays_per_tri >= 24.0)) && w.ordered_ok ? 1 : 0;
	const bool persistent = !(c->kernel == RTX_KERNEL_EXHAUSTIVE || !w.ordered_ok);
	if (!persistent || c->top_smem > 0) w.frustum = 0;
	if (w.frustum) {
		CU(c, c->d_lists.alloc((size_t)c->local_tiles * RTX_LIST_STRIDE * 4));
		CU(c, c->d_slists.alloc((size_t)((c->tiles_x + RTX_SUPER - 1) / RTX_SUPER) * ((c->tiles_y + RTX_SUPER - 1) / RTX_SUPER) * RTX_SLIST_STRIDE * 4));
		w.lists = c->d_lists.as<uint32_t>();
		w.overflow_tiles = c->d_counter.as<unsigned int>() + 2;      /* words: [0] work counter, [1] work counter of the overflow launch, [2] overflowed tiles, [3] upload scratch */
	}
	CU(c, cudaMemsetAsync(c->d_counter.p, 0, 4 * sizeof(unsigned int), st));
	if (c->counters) CU(c, cudaMemsetAsync(c->d_counters.p, 0, sizeof(Counters), st));
	CU(c, cudaEventRecord(c->ev0, st));
	if (w.frustum && c->local_tiles > 0) {
		/* Two levels pay off once there are many tiles: the super-tile pass has the latency of one
		 * breadth-first walk (~40 us) however few super-tiles there are. */
		const uint32_t nsuper = ((c->tiles_x + RTX_SUPER - 1) / RTX_SUPER) * ((c->tiles_y + RTX_SUPER - 1) / RTX_SUPER);
		const bool two_level = c->local_tiles >= 10000;
		if (two_level) {
			CU(c, phase_mark(c, st, RTX_PHASE_COLLECT_SUPER));
			k_frustum_collect_super<<<(nsuper + 3) / 4, 128, 0, st>>>(c->sc, w, c->d_slists.as<uint32_t>());
			CU(c, cudaGetLastError());
		}
		CU(c, phase_mark(c, st, RTX_PHASE_COLLECT));
		k_frustum_collect<<<(c->local_tiles + 7) / 8, 256, 0, st>>>(c->sc, w, two_level ? c->d_slists.as<uint32_t>() : nullptr, c->d_lists.as<uint32_t>());
		CU(c, cudaGetLastError());
		c->stats.kernel_launches = (two_level ? 2 : 1) + table_launch;
	} else {
		c->stats.kernel_launches = table_launch;
	}
	const uint32_t launches_per_pass = w.frustum ? 2u : 1u;
	const size_t frame_bytes = (size_t)c->W * c->H * sizeof(float);
	uint32_t nbands = 1, passes = 1;
	if (host_dst && !c->ao && c->local_tiles > 0) {
		nbands = (uint32_t)(frame_bytes / c->world / (8u << 20));  /* >= 8 MB per copy keeps PCIe near its rate */
		if (nbands > RTX_MAX_BANDS) nbands = RTX_MAX_BANDS;
		if (nbands > c->tiles_y) nbands = c->tiles_y;
		if (nbands < 1) nbands = 1;
	}
	if (host_dst && c->world > 1) {
		/* rtx_render_store on a tile partition: trace, then store this rank's tiles into the whole image at host_dst
		 * (mapped host or peer memory).  Measured: running the store kernel for finished bands next to the tracing of
		 * the following ones is SLOWER (2 GPUs, C3: 10.0 ms against 7.7 ms) -- the persistent traversal CTAs own every
		 * SM, so the store CTAs only ever run between bands, in pieces too small to fill the PCIe link. */
		/* The packet kernels (frustum path) do the store themselves: the warp that finishes a tile's last unit sends the
		 * tile, so the transfer runs inside the traversal kernel.  The other kernels are followed by k_store_tiles. */
		const bool packet_kernel = persistent && c->top_smem == 0 && (w.frustum || c->rays_per_thread > 1) && !w.rowmajor &&
		                           !c->counters && !rec && (!w.frustum || c->list_rays_per_thread == 2);
		if (packet_kernel && c->local_tiles > 0) {
			CU(c, c->d_tile_done.alloc((size_t)c->local_tiles * sizeof(unsigned int)));
			CU(c, cudaMemsetAsync(c->d_tile_done.p, 0, (size_t)c->local_tiles * sizeof(unsigned int), st));
			w.store_image = host_dst;
			w.tile_done = c->d_tile_done.as<unsigned int>();
		}
		CU(c, phase_mark(c, st, RTX_PHASE_TRAVERSAL));
		CU(c, launch_render(c, w, st));
		if (!w.store_image && c->local_tiles > 0) {
			k_store_tiles<<<c->local_tiles, 256, 0, st>>>(w.image, 0u, c->local_tiles, c->rank, c->world, c->tiles_x, c->W, c->H, host_dst);
			CU(c, cudaGetLastError());
		}
		w.store_image = nullptr;
		host_dst = nullptr;                                            /* nothing left to copy at the end of the frame */
	} else if (host_dst && nbands > 1) {
		if (!c->copy_stream) CU(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
		if (!c->copy_done) CU(c, cudaEventCreateWithFlags(&c->copy_done, cudaEventDisableTiming));
		for (uint32_t b = 0; b < nbands; ++b)
			if (!c->band_ev[b]) CU(c, cudaEventCreateWithFlags(&c->band_ev[b], cudaEventDisableTiming));
		const float *src = w.image;
		passes = 0;
		for (uint32_t b = 0; b < nbands; ++b) {
			/* the first band is a quarter of the others: the copy engine starts early and never waits afterwards
			 * (a band's copy takes longer than the next band's tracing) */
			const uint64_t den = 4ull * nbands - 3;
			const uint32_t r0 = b == 0 ? 0u : (uint32_t)((uint64_t)c->tiles_y * (4ull * b - 3) / den);
			const uint32_t r1 = (uint32_t)((uint64_t)c->tiles_y * (4ull * (b + 1) - 3) / den);
			if (r1 == r0) continue;
			++passes;
			w.tile_begin = r0 * c->tiles_x;
			w.tile_count = (r1 - r0) * c->tiles_x;
			w.num_units = w.tile_count * 32u;
			if (b > 0) CU(c, cudaMemsetAsync(w.counter, 0, 2 * sizeof(unsigned int), st));   /* the two work counters only */
			CU(c, launch_render(c, w, st));
			CU(c, cudaEventRecord(c->band_ev[b], st));
			CU(c, cudaStreamWaitEvent(c->copy_stream, c->band_ev[b], 0));
			const size_t y0 = (size_t)r0 * RTX_TILE, y1 = std::min<size_t>((size_t)r1 * RTX_TILE, c->H);
			CU(c, cudaMemcpyAsync(host_dst + y0 * c->W, src + y0 * c->W, (y1 - y0) * c->W * sizeof(float), cudaMemcpyDeviceToHost, c->copy_stream));
		}
		CU(c, cudaEventRecord(c->copy_done, c->copy_stream));
		w.tile_begin = 0;
		w.tile_count = c->local_tiles;
		w.num_units = c->local_tiles * 32u;
	} else {
		CU(c, phase_mark(c, st, RTX_PHASE_TRAVERSAL));
		CU(c, launch_render(c, w, st));
	}
	c->stats.kernel_launches += launches_per_pass * passes;
	uint32_t ao_launches = 0;
	if (c->ao && w.num_units > 0) {
		/* second pass over the hit pixels (intersect_kernel.cl:305-307) */
		AoParams ao{};
		ao.method = c->opt.ao_method;
		ao.samples = c->opt.ao_num_samples;
		ao.max_distance = c->ao_max_distance;
		ao.alpha_min = c->opt.ao_alpha_min;
		ao.alpha_max = c->opt.ao_alpha_max;
		if (ao.method == 0) {
			CU(c, c->d_ao_ring.alloc(((size_t)c->ao_ring_cap + 1) * sizeof(float4)));
			ao.ring = c->d_ao_ring.as<float4>();
			ao.ring_cap = c->ao_ring_cap;
			k_ao_ring_table<<<1, 32, 0, st>>>(ao, c->d_ao_ring.as<float4>());
			CU(c, cudaGetLastError());
			++ao_launches;
		}
		CU(c, phase_mark(c, st, RTX_PHASE_AO));
		k_ambient_occlusion<<<(w.num_units + 3) / 4, 128, 0, st>>>(c->sc, w, ao);
		CU(c, cudaGetLastError());
		++ao_launches;
	}
	CU(c, cudaEventRecord(c->ev1, st));
	CU(c, phase_mark(c, st, RTX_NUM_PHASES));
	c->phase_timing = timing_was;
	if (host_dst) {
		if (nbands > 1) CU(c, cudaStreamWaitEvent(st, c->copy_done, 0));     /* `st` is done when the copies are */
		else CU(c, cudaMemcpyAsync(host_dst, w.image, frame_bytes, cudaMemcpyDeviceToHost, st));
	}
	c->ev_pending = true;
	c->stats.rays = (uint64_t)c->local_tiles * RTX_TILE * RTX_TILE;
	if (c->world == 1) c->stats.rays = (uint64_t)c->W * c->H;
	c->stats.kernel_launches += ao_launches;
	c->stats.kernel_variant = (c->kernel == RTX_KERNEL_EXHAUSTIVE || !w.ordered_ok) ? RTX_KERNEL_EXHAUSTIVE : RTX_KERNEL_PERSISTENT;
	c->rendered = true;
	c->full_valid = false;
	c->u8_valid = false;
	c->ext_u8 = nullptr;
	return RTX_OK;
}

int rtx_render_async(rtx_ctx *c, void *stream)
{
	return enqueue_render(c, static_cast<cudaStream_t>(stream));
}

int rtx_synchronize(rtx_ctx *c)
{
	if (!c) return fail(nullptr, RTX_ERR_ARG, "null context");
	CU(c, cudaSetDevice(c->device));
	CU(c, cudaDeviceSynchronize());
	return finish_stats(c);
}

int rtx_render(rtx_ctx *c)
{
	const int rc = enqueue_render(c, c ? c->stream : nullptr);
	if (rc != RTX_OK) return rc;
	CU(c, cudaStreamSynchronize(c->stream));   /* queue.finish(), opencl_host.cc:147 */
	return finish_stats(c);
}

/* operator()() + download() (opencl_host.cc:137-153) as one call, the copy overlapped with the tracing. */
int rtx_render_download(rtx_ctx *c, float *image)
{
	if (!c || !image) return fail(c, RTX_ERR_ARG, "null argument");
	if (c->world > 1) return fail(c, RTX_ERR_STATE, "this rank holds a tile partition; gather and de-interleave first");
	const int rc = enqueue_render(c, c->stream, image);
	if (rc != RTX_OK) return rc;
	CU(c, cudaStreamSynchronize(c->stream));
	return finish_stats(c);
}

/* Tile partition (tile_world > 1): trace this context's share and store it into the WHOLE row-major float image at
 * `image_f32` -- page-locked host memory mapped into the device (rtx_host_register; every rank over its own PCIe link)
 * or a peer's device memory.  tile_world <= 1: rtx_render_download.  Blocking.  Float image only. */
int rtx_render_store_async(rtx_ctx *c, void *image_f32, void *stream)
{
	if (!c || !image_f32) return fail(c, RTX_ERR_ARG, "null argument");
	if (c->ao || c->record_hits) return fail(c, RTX_ERR_UNSUPPORTED, "rtx_render_store carries the float image only");
	if (c->world > 1 && c->ext_image && c->ext_rowmajor) return fail(c, RTX_ERR_STATE, "unbind rtx_bind_output_image first");
	if (c->world <= 1) return fail(c, RTX_ERR_STATE, "rtx_render_store_async is for tile partitions; use rtx_render_download / rtx_render_async");
	return enqueue_render(c, static_cast<cudaStream_t>(stream), static_cast<float *>(image_f32));
}

int rtx_render_store(rtx_ctx *c, void *image_f32)
{
	if (!c || !image_f32) return fail(c, RTX_ERR_ARG, "null argument");
	if (c->ao || c->record_hits) return fail(c, RTX_ERR_UNSUPPORTED, "rtx_render_store carries the float image only");
	if (c->world > 1 && c->ext_image && c->ext_rowmajor) return fail(c, RTX_ERR_STATE, "unbind rtx_bind_output_image first");
	const int rc = enqueue_render(c, c->stream, static_cast<float *>(image_f32));
	if (rc != RTX_OK) return rc;
	CU(c, cudaStreamSynchronize(c->stream));
	return finish_stats(c);
}

int rtx_get_stats(const rtx_ctx *c, rtx_stats *stats)
{
	if (!c || !stats) return fail(nullptr, RTX_ERR_ARG, "null argument");
	rtx_ctx *m = const_cast<rtx_ctx *>(c);
	if (m->ev_pending && cudaEventQuery(m->ev1) == cudaSuccess) finish_stats(m);
	*stats = c->stats;
	return RTX_OK;
}

int rtx_device_image(rtx_ctx *c, void **device_ptr, size_t *count)
{
	if (!c || !device_ptr) return fail(c, RTX_ERR_ARG, "null argument");
	if (c->world > 1 && c->full_valid) {
		*device_ptr = c->d_image_full.p;
		if (count) *count = (size_t)c->W * c->H;
		return RTX_OK;
	}
	*device_ptr = c->ext_image ? (void *)c->ext_image : c->d_image.p;
	if (count) *count = c->world > 1 ? (size_t)c->tiles_per_rank * RTX_TILE * RTX_TILE : (size_t)c->W * c->H;
	return RTX_OK;
}

int rtx_bind_output(rtx_ctx *c, void *device_ptr, size_t count)
{
	if (!c) return fail(nullptr, RTX_ERR_ARG, "null context");
	const size_t need = c->world > 1 ? (size_t)c->tiles_per_rank * RTX_TILE * RTX_TILE : (size_t)c->W * c->H;
	if (device_ptr && count < need) return fail(c, RTX_ERR_ARG, "bound output is smaller than this context's share of the image");
	c->ext_image = static_cast<float *>(device_ptr);
	c->ext_rowmajor = false;
	c->rendered = false;
	return RTX_OK;
}

/* Render this rank's tiles straight into the WHOLE row-major image, wherever the kernel can address it: rank 0's
 * device memory mapped over NVLink (rtx_peer_open) or page-locked host memory (rtx_host_register).  The pixels leave
 * the SM as they are shaded, so the transfer overlaps the tracing tile by tile and needs no kernel of its own. */
int rtx_bind_output_image(rtx_ctx *c, void *image_f32)
{
	if (!c) return fail(nullptr, RTX_ERR_ARG, "null context");
	c->ext_image = static_cast<float *>(image_f32);
	c->ext_rowmajor = image_f32 != nullptr;
	c->rendered = false;
	return RTX_OK;
}

static const float *full_image(rtx_ctx *c)
{
	if (c->world == 1) return c->ext_image ? c->ext_image : c->d_image.as<float>();
	return c->full_valid ? c->d_image_full.as<float>() : nullptr;
}

int rtx_download(rtx_ctx *c, float *image)
{
	if (!c || !image) return fail(c, RTX_ERR_ARG, "null argument");
	if (!c->rendered && !c->full_valid) return fail(c, RTX_ERR_STATE, "download before render");
	const float *src = full_image(c);
	if (!src) return fail(c, RTX_ERR_STATE, "this rank holds a tile partition; gather and de-interleave first");
	CU(c, cudaSetDevice(c->device));
	CU(c, cudaStreamSynchronize(c->stream));
	CU(c, cudaMemcpy(image, src, (size_t)c->W * c->H * sizeof(float), cudaMemcpyDeviceToHost)); /* opencl_host.cc:151 */
	return RTX_OK;
}

int rtx_download_hits(rtx_ctx *c, uint32_t *face_id, float *distance)
{
	if (!c) return fail(nullptr, RTX_ERR_ARG, "null context");
	if (!c->rendered || !c->record_hits || !c->d_face_id.p) return fail(c, RTX_ERR_STATE, "set RTX_TUNE_RECORD_HITS and render first");
	if (c->world > 1) return fail(c, RTX_ERR_STATE, "hit download needs the whole image on one context");
	CU(c, cudaSetDevice(c->device));
	CU(c, cudaStreamSynchronize(c->stream));
	const size_t n = (size_t)c->W * c->H;
	if (face_id) CU(c, cudaMemcpy(face_id, c->d_face_id.p, n * 4, cudaMemcpyDeviceToHost));
	if (distance) CU(c, cudaMemcpy(distance, c->d_dist.p, n * 4, cudaMemcpyDeviceToHost));
	return RTX_OK;
}

static uint32_t super_n(const rtx_ctx *c) { return (uint32_t)std::sqrt((double)c->opt.n_super_samples); }

int rtx_resize_u8_async(rtx_ctx *c, void *d_tiles_u8, size_t count, void *stream)
{
	if (!c) return fail(nullptr, RTX_ERR_ARG, "null context");
	if (!c->rendered) return fail(c, RTX_ERR_STATE, "resize before render");
	if (c->world > 1 && c->ext_image && c->ext_rowmajor) return fail(c, RTX_ERR_STATE, "this rank's tiles went straight into a whole-image binding (rtx_bind_output_image); there is no compact tile buffer to read");
	const uint32_t n = super_n(c), w = c->opt.width, h = c->opt.height;
	if (n == 0 || (uint64_t)w * n > c->W || (uint64_t)h * n > c->H) return fail(c, RTX_ERR_ARG, "image smaller than width*n x height*n");
	CU(c, cudaSetDevice(c->device));
	cudaStream_t st = static_cast<cudaStream_t>(stream);
	if (c->world == 1) {
		CU(c, c->d_u8.alloc((size_t)w * h));
		const float *src = c->ext_image ? c->ext_image : c->d_image.as<float>();
		const dim3 block(32, 8), grid((w + 31) / 32, (h + 7) / 8);
		launch_resize_u8(src, c->W, w, h, n, c->d_u8.as<unsigned char>(), grid, block, st);
		CU(c, cudaGetLastError());
		c->u8_valid = true;
		return RTX_OK;
	}
	if (RTX_TILE % n != 0) return fail(c, RTX_ERR_UNSUPPORTED, "per-rank resize needs sqrt(n_super_samples) to divide 32; gather floats instead");
	const uint32_t m = RTX_TILE / n;
	const size_t need = (size_t)c->tiles_per_rank * m * m;
	if (!d_tiles_u8 || count < need) return fail(c, RTX_ERR_ARG, "u8 tile buffer too small");
	const float *src = c->ext_image ? c->ext_image : c->d_image.as<float>();
	k_resize_tiles_u8<<<(unsigned)((need + 255) / 256), 256, 0, st>>>(src, c->tiles_per_rank, n, static_cast<unsigned char *>(d_tiles_u8));
	CU(c, cudaGetLastError());
	return RTX_OK;
}

int rtx_deinterleave_u8_async(rtx_ctx *c, const void *d_gathered, uint32_t world, void *stream)
{
	if (!c || !d_gathered || world == 0) return fail(c, RTX_ERR_ARG, "null argument");
	const uint32_t n = super_n(c);
	if (n == 0 || RTX_TILE % n != 0) return fail(c, RTX_ERR_UNSUPPORTED, "sqrt(n_super_samples) must divide 32");
	CU(c, cudaSetDevice(c->device));
	CU(c, c->d_u8.alloc((size_t)c->opt.width * c->opt.height));
	uint32_t tx, ty, tpr;
	rtx_tile_layout(c->W, c->H, world, &tx, &ty, &tpr);
	const size_t rows = (size_t)tx * ty * (RTX_TILE / n);
	k_deinterleave_u8<<<(unsigned)((rows + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const unsigned char *>(d_gathered), world, tpr, tx, ty, n,
	                                                                          c->opt.width, c->opt.height, c->d_u8.as<unsigned char>());
	CU(c, cudaGetLastError());
	c->u8_valid = true;
	return RTX_OK;
}

int rtx_download_u8(rtx_ctx *c, unsigned char *image)
{
	if (!c || !image) return fail(c, RTX_ERR_ARG, "null argument");
	if (c->u8_valid) {          /* already resized on the device (rtx_resize_u8_async / rtx_deinterleave_u8_async / rtx_adopt_u8) */
		CU(c, cudaSetDevice(c->device));
		CU(c, cudaDeviceSynchronize());
		CU(c, cudaMemcpy(image, c->ext_u8 ? (const void *)c->ext_u8 : c->d_u8.p, (size_t)c->opt.width * c->opt.height, cudaMemcpyDeviceToHost));
		return RTX_OK;
	}
	const float *src = full_image(c);
	if (!src || (!c->rendered && !c->full_valid)) return fail(c, RTX_ERR_STATE, "no complete image on this context");
	const uint32_t n = (uint32_t)std::sqrt((double)c->opt.n_super_samples);
	const uint32_t w = c->opt.width, h = c->opt.height;
	if (n == 0 || (uint64_t)w * n > c->W || (uint64_t)h * n > c->H) return fail(c, RTX_ERR_ARG, "image smaller than width*n x height*n");
	CU(c, cudaSetDevice(c->device));
	CU(c, c->d_u8.alloc((size_t)w * h));
	const dim3 block(32, 8), grid((w + 31) / 32, (h + 7) / 8);
	launch_resize_u8(src, c->W, w, h, n, c->d_u8.as<unsigned char>(), grid, block, c->stream);
	CU(c, cudaGetLastError());
	CU(c, cudaMemcpyAsync(image, c->d_u8.p, (size_t)w * h, cudaMemcpyDeviceToHost, c->stream));
	CU(c, cudaStreamSynchronize(c->stream));
	return RTX_OK;
}

/* ---- direct stores: a rank's share written straight into the final image, wherever it lives ---- */

int rtx_resize_u8_to_async(rtx_ctx *c, void *image_u8, void *stream)
{
	if (!c || !image_u8) return fail(c, RTX_ERR_ARG, "null argument");
	if (!c->rendered) return fail(c, RTX_ERR_STATE, "resize before render");
	if (c->world > 1 && c->ext_image && c->ext_rowmajor) return fail(c, RTX_ERR_STATE, "this rank's tiles went straight into a whole-image binding (rtx_bind_output_image); there is no compact tile buffer to read");
	const uint32_t n = super_n(c), w = c->opt.width, h = c->opt.height;
	if (n == 0 || RTX_TILE % n != 0) return fail(c, RTX_ERR_UNSUPPORTED, "sqrt(n_super_samples) must divide 32");
	if ((uint64_t)w * n > c->W || (uint64_t)h * n > c->H) return fail(c, RTX_ERR_ARG, "image smaller than width*n x height*n");
	CU(c, cudaSetDevice(c->device));
	cudaStream_t st = static_cast<cudaStream_t>(stream);
	if (c->world == 1) {          /* the whole frame is here, row-major */
		const float *src = c->ext_image ? c->ext_image : c->d_image.as<float>();
		const dim3 block(32, 8), grid((w + 31) / 32, (h + 7) / 8);
		launch_resize_u8(src, c->W, w, h, n, static_cast<unsigned char *>(image_u8), grid, block, st);
		CU(c, cudaGetLastError());
		return RTX_OK;
	}
	const uint32_t m = RTX_TILE / n;
	const bool quads = (m & 3u) == 0 && (w & 3u) == 0;
	const size_t work = (size_t)c->local_tiles * (quads ? m / 4 : m) * m;
	if (work == 0) return RTX_OK;
	const float *src = c->ext_image ? c->ext_image : c->d_image.as<float>();
	k_resize_tiles_u8_to<<<(unsigned)((work + 255) / 256), 256, 0, st>>>(src, c->local_tiles, n, c->rank, c->world, c->tiles_x, w, h,
	                                                                       static_cast<unsigned char *>(image_u8));
	CU(c, cudaGetLastError());
	return RTX_OK;
}

int rtx_store_tiles_async(rtx_ctx *c, void *image_f32, void *stream)
{
	if (!c || !image_f32) return fail(c, RTX_ERR_ARG, "null argument");
	if (!c->rendered) return fail(c, RTX_ERR_STATE, "store before render");
	if (c->world > 1 && c->ext_image && c->ext_rowmajor) return fail(c, RTX_ERR_STATE, "this rank's tiles went straight into a whole-image binding (rtx_bind_output_image); there is no compact tile buffer to read");
	CU(c, cudaSetDevice(c->device));
	cudaStream_t st = static_cast<cudaStream_t>(stream);
	const float *src = c->ext_image ? c->ext_image : c->d_image.as<float>();
	if (c->world == 1) {
		CU(c, cudaMemcpyAsync(image_f32, src, (size_t)c->W * c->H * sizeof(float), cudaMemcpyDefault, st));
		return RTX_OK;
	}
	if (c->local_tiles == 0) return RTX_OK;
	k_store_tiles<<<c->local_tiles, 256, 0, st>>>(src, 0u, c->local_tiles, c->rank, c->world, c->tiles_x, c->W, c->H, static_cast<float *>(image_f32));
	CU(c, cudaGetLastError());
	return RTX_OK;
}

int rtx_adopt_u8(rtx_ctx *c, const void *d_image_u8)
{
	if (!c || !d_image_u8) return fail(c, RTX_ERR_ARG, "null argument");
	c->ext_u8 = static_cast<const unsigned char *>(d_image_u8);
	c->u8_valid = true;
	return RTX_OK;
}

/* Device memory another process can map (CUDA IPC): rank 0 allocates the final image and hands the 64-byte handle to
 * the other ranks (any transport: torch.distributed, a pipe, a file); they open it and pass the pointer to
 * rtx_resize_u8_to_async / rtx_store_tiles_async.  Peer access rides NVLink / NVSwitch. */
int rtx_peer_alloc(rtx_ctx *c, size_t bytes, void **device_ptr, unsigned char handle[64])
{
	if (!c || !device_ptr || !handle || bytes == 0) return fail(c, RTX_ERR_ARG, "bad argument");
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
	CU(c, cudaSetDevice(c->device));
	void *p = nullptr;
	CU(c, cudaMalloc(&p, bytes));
	cudaError_t e = cudaMemset(p, 0, bytes);
	cudaIpcMemHandle_t h;
	if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
	if (e != cudaSuccess) { cudaFree(p); return cuda_fail(c, e, "cudaIpcGetMemHandle"); }
	std::memcpy(handle, &h, 64);
	*device_ptr = p;
	return RTX_OK;
}

int rtx_peer_open(rtx_ctx *c, const unsigned char handle[64], void **device_ptr)
{
	if (!c || !device_ptr || !handle) return fail(c, RTX_ERR_ARG, "bad argument");
	CU(c, cudaSetDevice(c->device));
	cudaIpcMemHandle_t h;
	std::memcpy(&h, handle, 64);
	void *p = nullptr;
	CU(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
	*device_ptr = p;
	return RTX_OK;
}

int rtx_peer_close(rtx_ctx *c, void *device_ptr)
{
	if (!c || !device_ptr) return fail(c, RTX_ERR_ARG, "bad argument");
	CU(c, cudaSetDevice(c->device));
	CU(c, cudaIpcCloseMemHandle(device_ptr));
	return RTX_OK;
}

int rtx_peer_free(rtx_ctx *c, void *device_ptr)
{
	if (!c || !device_ptr) return fail(c, RTX_ERR_ARG, "bad argument");
	CU(c, cudaSetDevice(c->device));
	if (c->ext_u8 == device_ptr) { c->ext_u8 = nullptr; c->u8_valid = false; }
	CU(c, cudaFree(device_ptr));
	return RTX_OK;
}

/* Frame counters in peer memory instead of a collective (k_peer_signal / k_peer_wait): `flag` / `flags` point into a
 * buffer of rtx_peer_alloc (zero-initialised) mapped by the ranks.  Waits give up after ~2 s of device time and set
 * flags[count_or_64 ...]: see rtx_peer_wait_async. */
int rtx_peer_signal_async(rtx_ctx *c, void *flag, uint32_t value, void *stream)
{
	if (!c || !flag) return fail(c, RTX_ERR_ARG, "null argument");
	CU(c, cudaSetDevice(c->device));
	k_peer_signal<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<unsigned int *>(flag), value);
	CU(c, cudaGetLastError());
	return RTX_OK;
}

/* Wait until flags[0 .. count) have all reached `value` (monotonic counters, wrap-safe).  `timed_out` (one word, may
 * live in the same buffer) is set to 1 if a counter did not arrive within ~2 s: the stream goes on, nothing hangs. */
int rtx_peer_wait_async(rtx_ctx *c, const void *flags, uint32_t count, uint32_t value, void *timed_out, void *stream)
{
	if (!c || !flags || !timed_out || count == 0 || count > 1024) return fail(c, RTX_ERR_ARG, "bad argument");
	CU(c, cudaSetDevice(c->device));
	k_peer_wait<<<1, (count + 31) / 32 * 32, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const unsigned int *>(flags), count, value,
	                                                                               static_cast<unsigned int *>(timed_out), 4000000000ll);
	CU(c, cudaGetLastError());
	return RTX_OK;
}

/* Page-lock caller-owned host memory (e.g. a shared mapping that several rank processes opened) and map it into the
 * device's address space; *device_alias is what kernels of this process may write (== p under unified addressing). */
int rtx_host_register(void *p, size_t bytes, void **device_alias)
{
	if (!p || bytes == 0) return fail(nullptr, RTX_ERR_ARG, "bad argument");
	cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
	if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaHostRegister");
	void *d = nullptr;
	e = cudaHostGetDevicePointer(&d, p, 0);
	if (e != cudaSuccess) { cudaHostUnregister(p); return cuda_fail(nullptr, e, "cudaHostGetDevicePointer"); }
	if (device_alias) *device_alias = d;
	return RTX_OK;
}

int rtx_host_unregister(void *p)
{
	if (!p) return fail(nullptr, RTX_ERR_ARG, "null pointer");
	cudaError_t e = cudaHostUnregister(p);
	if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaHostUnregister");
	return RTX_OK;
}

/* Blocking copy of raw device memory (e.g. a buffer of rtx_peer_alloc) to the host, after everything queued on the device. */
int rtx_copy_to_host(rtx_ctx *c, void *host_dst, const void *device_src, size_t bytes)
{
	if (!c || !host_dst || !device_src) return fail(c, RTX_ERR_ARG, "null argument");
	CU(c, cudaSetDevice(c->device));
	CU(c, cudaDeviceSynchronize());
	CU(c, cudaMemcpy(host_dst, device_src, bytes, cudaMemcpyDeviceToHost));
	return RTX_OK;
}

/* Debug build only (RTX_DEBUG_BOUNDS): out-of-range indices the kernels caught since the last call
 * (count, then source line / index / limit of the first one); the counters are reset. */
int rtx_debug_bounds(rtx_ctx *c, unsigned int out[4])
{
	if (!c || !out) return fail(c, RTX_ERR_ARG, "null argument");
#ifdef RTX_DEBUG_BOUNDS
	CU(c, cudaSetDevice(c->device));
	CU(c, cudaDeviceSynchronize());
	CU(c, cudaMemcpyFromSymbol(out, g_rtx_bounds, 4 * sizeof(unsigned int)));
	const unsigned int zero[4] = { 0u, 0u, 0u, 0u };
	CU(c, cudaMemcpyToSymbol(g_rtx_bounds, zero, sizeof zero));
	return RTX_OK;
#else
	out[0] = out[1] = out[2] = out[3] = 0u;
	return fail(c, RTX_ERR_UNSUPPORTED, "this library was built without RTX_DEBUG_BOUNDS (make -C opencl_raytracer_b200/csrc debug)");
#endif
}

int rtx_phase_ms(const rtx_ctx *c, double *ms)
{
	if (!c || !ms) return fail(nullptr, RTX_ERR_ARG, "null argument");
	rtx_ctx *m = const_cast<rtx_ctx *>(c);
	if (m->ev_pending && cudaEventQuery(m->ev1) == cudaSuccess) finish_stats(m);
	for (int k = 0; k < RTX_NUM_PHASES; ++k) ms[k] = c->phase_ms[k];
	return RTX_OK;
}

int rtx_deinterleave_async(rtx_ctx *c, const void *d_gathered, uint32_t world, void *stream)
{
	if (!c || !d_gathered || world == 0) return fail(c, RTX_ERR_ARG, "null argument");
	CU(c, cudaSetDevice(c->device));
	CU(c, c->d_image_full.alloc((size_t)c->W * c->H * sizeof(float)));
	uint32_t tx, ty, tpr;
	rtx_tile_layout(c->W, c->H, world, &tx, &ty, &tpr);
	cudaStream_t st = static_cast<cudaStream_t>(stream);
	k_deinterleave<<<tx * ty, 256, 0, st>>>(static_cast<const float *>(d_gathered), world, tpr, tx, ty, c->W, c->H,
	                                         c->d_image_full.as<float>());
	CU(c, cudaGetLastError());
	c->full_valid = true;
	return RTX_OK;
}

int rtx_trace_rays_device(rtx_ctx *c, const void *d_origins, const void *d_dirs, size_t nrays, float max_distance,
                          void *d_face_id, void *d_distance, void *stream)
{
	if (!c) return fail(nullptr, RTX_ERR_ARG, "null context");
	if (!c->uploaded) return fail(c, RTX_ERR_STATE, "trace before upload");
	if (nrays == 0) return RTX_OK;
	if (!d_origins || !d_dirs) return fail(c, RTX_ERR_ARG, "null ray arrays");
	if (nrays >= 0xfff00000ull) return fail(c, RTX_ERR_ARG, "too many rays in one call (limit 2^32 - 2^20)");
	CU(c, cudaSetDevice(c->device));
	cudaStream_t st = static_cast<cudaStream_t>(stream);
	RayWork w{};
	w.origins = static_cast<const float4 *>(d_origins);
	w.dirs = static_cast<const float4 *>(d_dirs);
	w.nrays = nrays;
	w.max_distance = max_distance;
	w.counter = c->d_counter.as<unsigned int>();
	w.face_id = static_cast<uint32_t *>(d_face_id);
	w.dist = static_cast<float *>(d_distance);
	w.ordered_ok = (c->tree_depth <= RTX_STACK_MAX && c->boxes_nested) ? 1 : 0;
	w.exhaustive = c->kernel == RTX_KERNEL_EXHAUSTIVE;
	CU(c, cudaMemsetAsync(c->d_counter.p, 0, sizeof(unsigned int), st));
	if (c->counters) CU(c, cudaMemsetAsync(c->d_counters.p, 0, sizeof(Counters), st));
	CU(c, cudaEventRecord(c->ev0, st));
	CU(c, launch_rays(c, w, st));
	CU(c, cudaEventRecord(c->ev1, st));
	c->ev_pending = true;
	c->stats.rays = nrays;
	c->stats.kernel_launches = 1;
	c->stats.kernel_variant = (w.exhaustive || !w.ordered_ok) ? RTX_KERNEL_EXHAUSTIVE : RTX_KERNEL_PERSISTENT;
	return RTX_OK;
}

/* Host arrays in, host arrays out.  Batches larger than one chunk are pipelined over three streams with two sets
 * of device buffers: while chunk k is traced, chunk k+1 is on its way up and chunk k-1 on its way down, so a big
 * batch costs about its longest stage (the 32 B/ray upload at PCIe rates) instead of the sum of the three. */
#define RTX_RAY_CHUNK (4u << 20)
int rtx_trace_rays(rtx_ctx *c, const float *origins, const float *dirs, size_t nrays, float max_distance,
                   uint32_t *face_id, float *distance)
{
	if (!c) return fail(nullptr, RTX_ERR_ARG, "null context");
	if (!c->uploaded) return fail(c, RTX_ERR_STATE, "trace before upload");
	if (nrays == 0) return RTX_OK;
	if (!origins || !dirs) return fail(c, RTX_ERR_ARG, "null ray arrays");
	CU(c, cudaSetDevice(c->device));
	const size_t chunk = nrays < RTX_RAY_CHUNK ? nrays : RTX_RAY_CHUNK;
	const size_t nchunks = (nrays + chunk - 1) / chunk;
	const int nbuf = nchunks > 1 ? 2 : 1;
	DevBuf *d_o = c->r_o, *d_d = c->r_d, *d_f = c->r_f, *d_t = c->r_t;
	cudaEvent_t *ev_in = c->r_ev_in, *ev_done = c->r_ev_done, *ev_out = c->r_ev_out;
	auto cleanup = [&] {};                                  /* buffers, streams and events stay with the context */
	cudaError_t e = cudaSuccess;
	auto ok = [&](cudaError_t r) { if (e == cudaSuccess) e = r; return e == cudaSuccess; };
	for (int k = 0; k < nbuf; ++k) {
		ok(d_o[k].alloc(chunk * 16)); ok(d_d[k].alloc(chunk * 16)); ok(d_f[k].alloc(chunk * 4)); ok(d_t[k].alloc(chunk * 4));
		if (!ev_in[k]) ok(cudaEventCreateWithFlags(&ev_in[k], cudaEventDisableTiming));
		if (!ev_done[k]) ok(cudaEventCreateWithFlags(&ev_done[k], cudaEventDisableTiming));
		if (!ev_out[k]) ok(cudaEventCreateWithFlags(&ev_out[k], cudaEventDisableTiming));
	}
	if (!c->r_ev_a) ok(cudaEventCreate(&c->r_ev_a));
	if (!c->r_ev_b) ok(cudaEventCreate(&c->r_ev_b));
	if (nchunks > 1) {
		if (!c->r_in) ok(cudaStreamCreateWithFlags(&c->r_in, cudaStreamNonBlocking));
		if (!c->r_out) ok(cudaStreamCreateWithFlags(&c->r_out, cudaStreamNonBlocking));
	}
	if (e != cudaSuccess) return cuda_fail(c, e, "rtx_trace_rays: buffers / streams");
	cudaStream_t s_in = c->r_in, s_out = c->r_out;
	cudaEvent_t ev_a = c->r_ev_a, ev_b = c->r_ev_b;
	cudaStream_t up = nchunks > 1 ? s_in : c->stream, down = nchunks > 1 ? s_out : c->stream;
	int rc = RTX_OK;
	ok(cudaEventRecord(ev_a, c->stream));
	for (size_t k = 0; k < nchunks && rc == RTX_OK && e == cudaSuccess; ++k) {
		const int b = (int)(k & 1);
		const size_t first = k * chunk, n = nrays - first < chunk ? nrays - first : chunk;
		if (k >= 2) ok(cudaStreamWaitEvent(up, ev_done[b], 0));                     /* chunk k-2 no longer reads these inputs */
		ok(cudaMemcpyAsync(d_o[b].p, origins + 4 * first, n * 16, cudaMemcpyHostToDevice, up));
		ok(cudaMemcpyAsync(d_d[b].p, dirs + 4 * first, n * 16, cudaMemcpyHostToDevice, up));
		ok(cudaEventRecord(ev_in[b], up));
		ok(cudaStreamWaitEvent(c->stream, ev_in[b], 0));
		if (k >= 2) ok(cudaStreamWaitEvent(c->stream, ev_out[b], 0));               /* chunk k-2's results have left */
		if (e != cudaSuccess) break;
		rc = rtx_trace_rays_device(c, d_o[b].p, d_d[b].p, n, max_distance, d_f[b].p, d_t[b].p, c->stream);
		if (rc != RTX_OK) break;
		ok(cudaEventRecord(ev_done[b], c->stream));
		ok(cudaStreamWaitEvent(down, ev_done[b], 0));
		if (face_id) ok(cudaMemcpyAsync(face_id + first, d_f[b].p, n * 4, cudaMemcpyDeviceToHost, down));
		if (distance) ok(cudaMemcpyAsync(distance + first, d_t[b].p, n * 4, cudaMemcpyDeviceToHost, down));
		ok(cudaEventRecord(ev_out[b], down));
	}
	ok(cudaEventRecord(ev_b, c->stream));
	cudaError_t es = cudaStreamSynchronize(c->stream);
	if (s_in) { const cudaError_t e2 = cudaStreamSynchronize(s_in); if (es == cudaSuccess) es = e2; }
	if (s_out) { const cudaError_t e2 = cudaStreamSynchronize(s_out); if (es == cudaSuccess) es = e2; }
	if (rc == RTX_OK && e != cudaSuccess) rc = cuda_fail(c, e, "rtx_trace_rays");
	if (rc == RTX_OK && es != cudaSuccess) rc = cuda_fail(c, es, "cudaStreamSynchronize");
	if (rc == RTX_OK) {
		rc = finish_stats(c);
		float ms = 0.f;
		if (cudaEventElapsedTime(&ms, ev_a, ev_b) == cudaSuccess) c->stats.kernel_ms = ms;   /* the whole pipelined batch on the compute stream */
		c->stats.rays = nrays;
		c->stats.kernel_launches = (uint32_t)nchunks;
	}
	cleanup();
	return rc;
}

int rtx_trace_random_rays(rtx_ctx *c, uint32_t seed, uint64_t first, size_t nrays, float max_distance,
                          uint32_t *face_id, float *distance, uint64_t *hit_count, uint64_t *sum_face_id)
{
	if (!c) return fail(nullptr, RTX_ERR_ARG, "null context");
	if (!c->uploaded) return fail(c, RTX_ERR_STATE, "trace before upload");
	if (hit_count) *hit_count = 0;
	if (sum_face_id) *sum_face_id = 0;
	if (nrays == 0) return RTX_OK;
	if (nrays >= 0xfff00000ull) return fail(c, RTX_ERR_ARG, "too many rays in one call (limit 2^32 - 2^20)");
	CU(c, cudaSetDevice(c->device));
	if (face_id) CU(c, c->d_face_id.alloc(nrays * 4));
	if (distance) CU(c, c->d_dist.alloc(nrays * 4));
	cudaStream_t st = c->stream;
	RayWork w{};
	w.seed = seed;
	w.first = first;
	w.nrays = nrays;
	w.bbmin = c->bbmin;
	w.bbmax = c->bbmax;
	w.max_distance = max_distance;
	w.counter = c->d_counter.as<unsigned int>();
	w.face_id = face_id ? c->d_face_id.as<uint32_t>() : nullptr;
	w.dist = distance ? c->d_dist.as<float>() : nullptr;
	w.hit_count = c->d_sums.as<unsigned long long>();
	w.sum_face_id = c->d_sums.as<unsigned long long>() + 1;
	w.ordered_ok = (c->tree_depth <= RTX_STACK_MAX && c->boxes_nested) ? 1 : 0;
	w.exhaustive = c->kernel == RTX_KERNEL_EXHAUSTIVE;
	CU(c, cudaMemsetAsync(c->d_counter.p, 0, sizeof(unsigned int), st));
	CU(c, cudaMemsetAsync(c->d_sums.p, 0, 2 * sizeof(unsigned long long), st));
	if (c->counters) CU(c, cudaMemsetAsync(c->d_counters.p, 0, sizeof(Counters), st));
	CU(c, cudaEventRecord(c->ev0, st));
	CU(c, launch_rays(c, w, st));
	CU(c, cudaEventRecord(c->ev1, st));
	c->ev_pending = true;
	unsigned long long sums[2] = { 0, 0 };
	CU(c, cudaMemcpyAsync(sums, c->d_sums.p, sizeof sums, cudaMemcpyDeviceToHost, st));
	if (face_id) CU(c, cudaMemcpyAsync(face_id, c->d_face_id.p, nrays * 4, cudaMemcpyDeviceToHost, st));
	if (distance) CU(c, cudaMemcpyAsync(distance, c->d_dist.p, nrays * 4, cudaMemcpyDeviceToHost, st));
	CU(c, cudaStreamSynchronize(st));
	if (hit_count) *hit_count = sums[0];
	if (sum_face_id) *sum_face_id = sums[1];
	c->stats.rays = nrays;
	c->stats.kernel_launches = 1;
	c->stats.kernel_variant = (w.exhaustive || !w.ordered_ok) ? RTX_KERNEL_EXHAUSTIVE : RTX_KERNEL_PERSISTENT;
	c->rendered = false;
	return finish_stats(c);
}

/* level 0 = L2 (or HBM when bytes exceeds L2), level 1 = L1 (per-CTA 64 KB slices). */
int rtx_probe_bandwidth(rtx_ctx *c, int level, size_t bytes, int iters, double *gbps)
{
	if (!c || !gbps || iters < 1 || bytes < (1u << 20)) return fail(c, RTX_ERR_ARG, "bad probe arguments");
	CU(c, cudaSetDevice(c->device));
	DevBuf buf, sink;
	cudaError_t e;
	if ((e = buf.alloc(bytes)) != cudaSuccess || (e = sink.alloc(16)) != cudaSuccess) { buf.release(); return cuda_fail(c, e, "cudaMalloc(probe)"); }
	cudaMemsetAsync(buf.p, 0, bytes, c->stream);
	const size_t n_vecs = bytes / 16, per_cta = level == 1 ? (64u << 10) / 16 : 0;
	const unsigned grid = (unsigned)c->sm_count * 4;
	float best = 1e30f;
	for (int rep = 0; rep < 4; ++rep) {      /* rep 0 warms the cache */
		cudaEventRecord(c->ev0, c->stream);
		k_probe_bw<<<grid, 512, 0, c->stream>>>(buf.as<float4>(), n_vecs, iters, per_cta, sink.as<float>());
		cudaEventRecord(c->ev1, c->stream);
		if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) { buf.release(); sink.release(); return cuda_fail(c, e, "probe kernel"); }
		float ms = 0.f;
		cudaEventElapsedTime(&ms, c->ev0, c->ev1);
		if (rep > 0 && ms < best) best = ms;
	}
	const double moved = level == 1 ? (double)grid * per_cta * 16.0 * iters : (double)n_vecs * 16.0 * iters;
	*gbps = moved / (best * 1e-3) / 1e9;
	buf.release();
	sink.release();
	c->ev_pending = false;
	return RTX_OK;
}

} /* extern "C" */
