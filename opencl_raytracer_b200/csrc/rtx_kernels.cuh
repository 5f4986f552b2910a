/*
 * rtx_kernels.cuh -- traversal kernels (sm_100a).
 *
 * Device scene layout (built by rtx_api.cu::flatten from the reference's
 * upload arrays; DESIGN.md "Data layout"):
 *
 *   pairs     4 x float4 per INTERNAL node of the flattened tree = its two
 *             children as two 32-byte nodes {lo.xyz, hi.xyz, ref, pad}:
 *               q0 = (L.lo.x, L.lo.y, L.lo.z, L.hi.x)
 *               q1 = (L.hi.y, L.hi.z, bits(L.ref), 0)
 *               q2, q3 = the same for R
 *             ref >= 0: index of the child's own pair; ref < 0: leaf,
 *             ~ref = (first_triangle << 3) | (count - 1), count <= 8.
 *             Pairs [0, top_pairs) are the top levels in breadth-first order
 *             (staged in shared memory), the rest depth-first.
 *   tris      4 x float4 per triangle in leaf order (rtx_device.cuh)
 *   leafbox   2 x float4 per triangle: its reference leaf AABB (min, max)
 *   pleafbox  the same with the ray-independent terms of a PRIMARY ray's tests folded in (candidate-list kernel)
 *   tnormals  3 x float4 per triangle: the three corner normals, pre-gathered
 *   ref_nodes / ref_aabbs   the reference's own arrays, for the literal
 *             stackless walk (exhaustive kernel, NaN-slab rays, deep trees)
 */
#pragma once
#include "rtx_device.cuh"

struct SceneDev {
	const float4 *pairs;      /* 4 octant copies, each num_pairs * 4 float4 (copy 0 = unswapped) */
	const float4 *tris;
	const float4 *leafbox;
	const float4 *pleafbox;   /* the leaf boxes again, prepared for PRIMARY rays (camera at (0,0,2)): (lo.x, lo.y, lo.z - 2, A),
	                             (hi.x, hi.y, hi.z - 2, 0) with A = -dot(n, o - a) of the triangle (triangle_test_primary) */
	const float4 *tnormals;
	const uint32_t *ref_nodes;
	const float4 *ref_aabbs;
	uint32_t num_pairs;       /* distance (in pairs) between the octant copies */
	uint32_t pair_count;      /* pairs actually in the tree (<= num_pairs) */
	uint32_t top_pairs;       /* pairs staged in shared memory */
	uint32_t num_tris;
	uint32_t verify_leafbox;  /* leaves hold > 1 triangle: per-triangle leaf box must be checked */
	float scene_scale;        /* max |coordinate| of the root box */
};

struct Counters {
	unsigned long long node_visits, tri_tests, leafbox_tests, exact_rays, overflow_packets;
};

struct HitRec {
	float dist;      /* +inf = miss */
	uint32_t tri;    /* leaf index */
	float s, t;
};

#define RTX_STACK_MAX 64          /* deeper flattened trees use the stackless walk */
#define RTX_TILE 32               /* partition / scheduling tile: 32 x 32 pixels */

/* --------------------------------------------------------------------------
 * The reference's walk, literally (intersect_kernel.cl:184-213): pre-order
 * array, subtree skip on box miss, no culling, strict `>` update so the first
 * triangle in leaf order wins ties.
 * ------------------------------------------------------------------------ */
template <bool COUNT>
__device__ __noinline__ void walk_reference(const SceneDev &sc, f3 o, f3 d, float max_distance, HitRec &best, Counters *cnt)
{
	const uint32_t n = __ldg(sc.ref_nodes);
	uint32_t tri = 0;
	unsigned long long visits = 0, tests = 0;
	for (uint32_t i = 0; i < n;) {
		const uint32_t node_count = __ldg(sc.ref_nodes + i);
		const float4 lo = __ldg(sc.ref_aabbs + 2 * (size_t)i);
		const float4 hi = __ldg(sc.ref_aabbs + 2 * (size_t)i + 1);
		if (COUNT) ++visits;
		if (!aabb_exact(make_f3(lo.x, lo.y, lo.z), make_f3(hi.x, hi.y, hi.z), o, d, max_distance)) {
			tri += (node_count + 1) >> 1;
			i += node_count;
		} else {
			if (node_count == 1) {
				const float4 *q = sc.tris + 4 * RTX_IDX(tri, sc.num_tris);
				TriHit h;
				if (COUNT) ++tests;
				if (triangle_test(__ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3), o, d, __int_as_float(0x7f800000), h)) {
					if (best.dist > h.dist) { best.dist = h.dist; best.tri = tri; best.s = h.s; best.t = h.t; }
				}
				++tri;
			}
			++i;
		}
	}
	if (COUNT) {
		atomicAdd(&cnt->node_visits, visits);
		atomicAdd(&cnt->tri_tests, tests);
	}
}

/* --------------------------------------------------------------------------
 * Ordered stack traversal with distance culling.  Same result as the walk
 * above (DESIGN.md "Why re-ordering is exact"):
 *   - a triangle is tested by the reference iff its leaf box passes the slab
 *     test (ancestor boxes contain it, and the test is monotone); interior
 *     boxes here are those same boxes, tested in min/max form, which decides
 *     identically whenever no 0*inf occurs -- rays with a zero direction
 *     component never enter this function;
 *   - with > 1 triangle per leaf the per-triangle leaf box is tested, with
 *     the literal form, before a hit is accepted;
 *   - closest hit = min distance, ties -> smallest leaf index;
 *   - a subtree/triangle is culled only if its entry parameter exceeds the
 *     best hit's parameter by a margin far larger than any rounding.
 * ------------------------------------------------------------------------ */
struct Slab { float tmin, tmax; };

/* Culling slack of a box, stored in the pad lane of its 32-byte node.  The reference tests a triangle whenever
 * its leaf box passes the slab test and accepts the hit wherever the plane point P lies, as long as the computed
 * s, t are within 1e-5 of the triangle (intersect_kernel.cl:96,101).  So P can lie outside the triangle's own leaf
 * box: by ~1e-5 of the triangle's extent for a well-shaped triangle, by more when s and t are ill-conditioned
 * (s = (uv wv - vv wu) / D loses ~eps * kappa, kappa = uu vv / |D| = 1 / sin^2 of the angle between the edges), and
 * anywhere in the plane once D is rounding noise.  The ray then hits the triangle at parameter r but enters the
 * triangle's box only at r + overshoot / |d_k| along the axis k it overshoots.  A box may therefore be skipped
 * against the best hit so far only if  t_min - slack * max_k |1 / d_k|  lies beyond the culling bound, with
 * slack >= the overshoot of any triangle inside the box:
 *   interior box   2.5e-4 of its largest extent + 4e-6 of its largest coordinate (25 x the nominal tolerance; the
 *                  second term covers the rounding of P itself), raised to the largest slack among its leaves;
 *   leaf           max(2.5e-4, 1e-5 kappa) of the extent, and +inf (never skipped) for kappa > 1e4 or D == 0.
 * tools/fuzz_gpu.py (needle meshes) and tests/test_parity_gpu.py hold this against the literal walk. */
#define RTX_SLACK_REL 2.5e-4f
__host__ __device__ __forceinline__ float box_slack_rel(float lx, float ly, float lz, float hx, float hy, float hz, float rel)
{
	const float ext = fmaxf(fmaxf(hx - lx, hy - ly), hz - lz);
	const float mag = fmaxf(fmaxf(fmaxf(fabsf(lx), fabsf(hx)), fmaxf(fabsf(ly), fabsf(hy))), fmaxf(fabsf(lz), fabsf(hz)));
	if (!(rel <= 3e38f)) return rel;                          /* +inf stays +inf for a box of zero extent too (inf * 0) */
	return rel * ext + 4e-6f * mag;
}
__host__ __device__ __forceinline__ float box_slack(float lx, float ly, float lz, float hx, float hy, float hz)
{
	return box_slack_rel(lx, ly, lz, hx, hy, hz, RTX_SLACK_REL);
}
/* relative slack of one triangle from its record q3 = (uu, uv, vv, D) */
RTX_DEV float tri_slack_rel(float4 q3)
{
	const float kappa = (q3.x * q3.z) / fabsf(q3.w);
	if (!(kappa <= 1e4f)) return __int_as_float(0x7f800000);            /* ill-conditioned, D == 0 or NaN: never cull */
	return fmaxf(RTX_SLACK_REL, 1e-5f * kappa);
}

/* Slab interval of one box.  Same products as the literal test ((bb - o) * (1/d), one rounding each);
 * which of the two products per axis is the entry is decided
 *   OCT = 0 : by the data -- the warp reads the copy of the pair array whose x/y slots were swapped at
 *             upload for its octant, so slot "lo" is the entry plane and "hi" the exit plane; d.z < 0;
 *   OCT = 4 : by min/max, which picks the same product as `div >= 0` whenever no 0*inf occurs
 *             (symmetric in lo/hi, so it works on any copy).
 * PRIMARY: the origin is the reference's fixed camera (0,0,2) (intersect_kernel.cl:284): bb.x - 0 and
 * bb.y - 0 are exact identities and are skipped. */
template <bool PRIMARY, int OCT>
RTX_DEV Slab slab_interval(float lx, float ly, float lz, float hx, float hy, float hz, f3 o, f3 id)
{
	const float ax = PRIMARY ? rn_mul(lx, id.x) : rn_mul(rn_sub(lx, o.x), id.x);
	const float bx = PRIMARY ? rn_mul(hx, id.x) : rn_mul(rn_sub(hx, o.x), id.x);
	const float ay = PRIMARY ? rn_mul(ly, id.y) : rn_mul(rn_sub(ly, o.y), id.y);
	const float by = PRIMARY ? rn_mul(hy, id.y) : rn_mul(rn_sub(hy, o.y), id.y);
	const float az = rn_mul(rn_sub(lz, PRIMARY ? 2.0f : o.z), id.z), bz = rn_mul(rn_sub(hz, PRIMARY ? 2.0f : o.z), id.z);
	Slab s;
	if (OCT < 4) {
		s.tmin = fmaxf(fmaxf(ax, ay), bz);
		s.tmax = fminf(fminf(bx, by), az);
	} else if (PRIMARY) {
		s.tmin = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), bz);
		s.tmax = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), az);
	} else {
		s.tmin = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
		s.tmax = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
	}
	return s;
}

template <int SMEM_STACK, bool TOP_SMEM, bool COUNT, bool PRIMARY, int OCT>
RTX_DEV void traverse_ordered(const SceneDev &sc, const float4 *__restrict__ pairs, const float4 *__restrict__ s_top,
                              uint2 *__restrict__ s_stack, f3 o, f3 d, float max_distance, HitRec &best, Counters *cnt)
{
	const f3 id = make_f3(rn_div(1.0f, d.x), rn_div(1.0f, d.y), rn_div(1.0f, d.z));
	/* culling margins in ray-parameter units */
	const float inv_len = PRIMARY ? 1.0f : rsqrtf(fmaxf(d.x * d.x + d.y * d.y + d.z * d.z, 1e-30f));
	const float scale = fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), sc.scene_scale));
	const float abs_margin = 1e-5f * scale * inv_len;
	/* `cull`: ray parameter beyond which nothing can beat the best hit so far.  Boxes also obey the
	 * reference's max_distance (intersect_kernel.cl:60); triangles do not -- the reference accepts a
	 * hit at any distance once its leaf box passed. */
	float cull = __int_as_float(0x7f800000);
	const float maxid = fminf(fmaxf(fmaxf(fabsf(id.x), fabsf(id.y)), fabsf(id.z)), 1e30f);   /* see box_slack */
	uint2 l_stack[RTX_STACK_MAX - SMEM_STACK];
	int sp = 0;
	int cur = 0;                         /* root pair */
	unsigned long long visits = 0, tests = 0, lbtests = 0;
	const int stride = blockDim.x;

	for (;;) {
		/* ---- interior: test both children of pair `cur` ---- */
		while (cur >= 0) {
			float4 q0, q1, q2, q3;
			if (TOP_SMEM && OCT == 4 && (uint32_t)cur < sc.top_pairs) {
				const float4 *q = s_top + 4 * cur;
				q0 = q[0]; q1 = q[1]; q2 = q[2]; q3 = q[3];
			} else {
				const float4 *q = pairs + 4 * RTX_IDX(cur, sc.pair_count);
				q0 = __ldg(q); q1 = __ldg(q + 1); q2 = __ldg(q + 2); q3 = __ldg(q + 3);
			}
			if (COUNT) visits += 2;
			const Slab L = slab_interval<PRIMARY, OCT>(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, o, id);
			const Slab R = slab_interval<PRIMARY, OCT>(q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, o, id);
			const float cL = __fmaf_rn(-q1.w, maxid, L.tmin), cR = __fmaf_rn(-q3.w, maxid, R.tmin);   /* entry, less the slack */
			const bool hitL = L.tmin <= L.tmax && L.tmin < max_distance && cL < cull && L.tmax > 0.0f;
			const bool hitR = R.tmin <= R.tmax && R.tmin < max_distance && cR < cull && R.tmax > 0.0f;
			if (!(hitL || hitR)) goto pop;
			const int refL = __float_as_int(q1.z), refR = __float_as_int(q3.z);
			const bool r_first = hitR && (!hitL || R.tmin < L.tmin);
			if (hitL && hitR) {
				const uint2 e = make_uint2((uint32_t)(r_first ? refL : refR), __float_as_uint(r_first ? cL : cR));
				if (SMEM_STACK > 0 && sp < SMEM_STACK) s_stack[sp * stride] = e; else l_stack[RTX_IDX(sp - SMEM_STACK, RTX_STACK_MAX - SMEM_STACK)] = e;
				++sp;
			}
			cur = r_first ? refR : refL;
		}
		/* ---- leaf ---- */
		{
			const uint32_t enc = ~(uint32_t)cur;
			const uint32_t first = enc >> 3, count = (enc & 7u) + 1u;
			for (uint32_t k = 0; k < count; ++k) {
				const uint32_t tri = first + k;
				const float4 *q = sc.tris + 4 * RTX_IDX(tri, sc.num_tris);
				TriHit h;
				if (COUNT) ++tests;
				if (!triangle_test(__ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3), o, d, cull, h)) continue;
				if (!(h.dist < best.dist || (h.dist == best.dist && tri < best.tri))) continue;
				if (sc.verify_leafbox) {
					const float4 lo = __ldg(sc.leafbox + 2 * RTX_IDX(tri, sc.num_tris)), hi = __ldg(sc.leafbox + 2 * RTX_IDX(tri, sc.num_tris) + 1);
					if (COUNT) ++lbtests;
					if (!aabb_exact(make_f3(lo.x, lo.y, lo.z), make_f3(hi.x, hi.y, hi.z), o, d, max_distance)) continue;
				}
				best.dist = h.dist; best.tri = tri; best.s = h.s; best.t = h.t;
				/* hits are accepted in decreasing distance, so this is the smallest so far; its ray
				 * parameter is recovered from the distance (|P-o| = r|d| up to rounding) */
				cull = (h.dist * inv_len) * 1.0001f + abs_margin;
			}
		}
pop:
		for (;;) {
			if (sp == 0) goto done;
			--sp;
			const uint2 e = (SMEM_STACK > 0 && sp < SMEM_STACK) ? s_stack[sp * stride] : l_stack[sp - SMEM_STACK];
			if (__uint_as_float(e.y) < cull) { cur = (int)e.x; break; }
		}
	}
done:
	if (COUNT) {
		atomicAdd(&cnt->node_visits, visits);
		atomicAdd(&cnt->tri_tests, tests);
		atomicAdd(&cnt->leafbox_tests, lbtests);
	}
}

/* One ray through the right path.  Rays with a zero direction component make the reference's slab
 * test produce 0*inf = NaN, whose propagation only the literal walk reproduces.  For primary rays
 * (fixed camera, d.z < 0) a warp whose lanes all share the signs of d.x and d.y runs a traversal
 * specialised for that octant; mixed warps (the image's centre row/column) use the min/max form. */
template <int SMEM_STACK, bool TOP_SMEM, bool COUNT, bool PRIMARY>
RTX_DEV void closest_hit(const SceneDev &sc, const float4 *s_top, uint2 *s_stack, bool ordered_ok,
                         f3 o, f3 d, float max_distance, HitRec &best, Counters *cnt)
{
	best.dist = __int_as_float(0x7f800000);
	best.tri = 0xffffffffu;
	best.s = best.t = 0.f;
	const bool plain = ray_is_plain(o, d);
	if (ordered_ok && plain) {
		if (PRIMARY) {
			const unsigned mask = __activemask();
			const int oct = (d.x < 0.0f ? 1 : 0) | (d.y < 0.0f ? 2 : 0);
			const int oct0 = __shfl_sync(mask, oct, __ffs(mask) - 1);
			const bool uniform = __all_sync(mask, oct == oct0) && d.z < 0.0f;
			if (uniform)
				traverse_ordered<SMEM_STACK, TOP_SMEM, COUNT, true, 0>(sc, sc.pairs + (size_t)oct0 * sc.num_pairs * 4, s_top, s_stack, o, d, max_distance, best, cnt);
			else
				traverse_ordered<SMEM_STACK, TOP_SMEM, COUNT, false, 4>(sc, sc.pairs, s_top, s_stack, o, d, max_distance, best, cnt);
		} else {
			traverse_ordered<SMEM_STACK, TOP_SMEM, COUNT, false, 4>(sc, sc.pairs, s_top, s_stack, o, d, max_distance, best, cnt);
		}
	} else {
		if (COUNT) atomicAdd(&cnt->exact_rays, 1ull);
		walk_reference<COUNT>(sc, o, d, max_distance, best, cnt);
	}
}

/* --------------------------------------------------------------------------
 * Pixel <-> work mapping.  The image is cut into 32x32-pixel tiles (row-major
 * tile ids); tile t belongs to rank t % world.  A warp's unit of work is one
 * 8x4-pixel block; 32 consecutive units cover one tile.
 * ------------------------------------------------------------------------ */
struct Work {
	Camera cam;
	uint32_t tiles_x, tiles_y;
	uint32_t rank, world;
	uint32_t local_tiles;        /* tiles this rank renders */
	uint32_t tile_begin, tile_count; /* the local tiles THIS launch renders (a band of tile rows when the download is
	                                    pipelined, rtx_render_download; otherwise 0, local_tiles) */
	uint32_t num_units;          /* tile_count * 32 */
	unsigned int *counter;       /* persistent-kernel work counter (zeroed before launch) */
	float *image;                /* world == 1: row-major W x H; else compact [local_tile][32][32] */
	float *store_image;          /* world > 1, packet kernels: the whole row-major image in mapped host / peer memory; the warp that
	                                finishes a tile's last unit copies the tile there (128-byte rows), so the transfer runs while
	                                the other tiles are still being traced (rtx_render_store) */
	unsigned int *tile_done;     /* ... units finished per local tile (zeroed before the frame) */
	int rowmajor;                /* world > 1: `image` is the whole row-major W x H image (a peer's or mapped host memory,
	                                rtx_bind_output_image) and this rank writes only the pixels of its own tiles */
	uint32_t *face_id;           /* optional (record mode), same indexing as image */
	float *dist;
	float2 *hit_st;              /* optional (ambient occlusion): parametric coordinates of the hit */
	int ordered_ok;
	int cost_map;                /* analysis hook: record mode stores the SM cycles a ray's warp took instead of the distance */
	int frustum;                 /* use the frustum front end for packets */
	const uint32_t *lists;       /* per-tile candidate lists written by k_frustum_collect */
	unsigned int *overflow_tiles; /* number of tiles whose list overflowed (zeroed before launch) */
};

RTX_DEV bool unit_pixel(const Work &w, uint32_t unit, uint32_t lane, uint32_t &x, uint32_t &y, size_t &out)
{
	const uint32_t ltile = unit >> 5, sub = unit & 31u;
	const uint32_t tile = ltile * w.world + w.rank;
	const uint32_t tx = tile % w.tiles_x, ty = tile / w.tiles_x;
	const uint32_t px = ((sub & 3u) << 3) + (lane & 7u), py = ((sub >> 2) << 2) + (lane >> 3);
	x = tx * RTX_TILE + px;
	y = ty * RTX_TILE + py;
	out = (w.world > 1 && !w.rowmajor) ? (size_t)ltile * (RTX_TILE * RTX_TILE) + py * RTX_TILE + px : (size_t)y * w.cam.W + x;
	return ty < w.tiles_y && x < w.cam.W && y < w.cam.H;
}

template <int SMEM_STACK, bool TOP_SMEM, bool COUNT, bool RECORD>
RTX_DEV void trace_pixel(const SceneDev &sc, const Work &w, const float4 *s_top, uint2 *s_stack,
                         uint32_t x, uint32_t y, size_t out, Counters *cnt)
{
	const f3 o = make_f3(0.0f, 0.0f, 2.0f);                           /* :284 */
	const f3 d = primary_dir(w.cam, x, y);
	HitRec best;
	const long long t0 = (RECORD && w.cost_map) ? clock64() : 0ll;
	closest_hit<SMEM_STACK, TOP_SMEM, COUNT, true>(sc, s_top, s_stack, w.ordered_ok != 0, o, d, 100000.0f, best, cnt); /* :292-295 */
	float value = 0.0f;                                               /* :297-299 */
	if (best.tri != 0xffffffffu) value = shade_hit(sc.tnormals, best.tri, best.s, best.t, d, w.cam.shading);
	w.image[out] = value;                                             /* :309 */
	if (RECORD) {
		w.face_id[out] = best.tri != 0xffffffffu ? best.tri * 3u : 0xffffffffu;
		w.dist[out] = w.cost_map ? (float)(clock64() - t0) : best.dist;     /* cost_map: analysis hook (tools/cost_map.py) */
		if (w.hit_st) w.hit_st[out] = make_float2(best.s, best.t);
	}
}

/* Persistent kernel: grid = SMs x resident CTAs, every warp pulls 8x4-pixel
 * units from one atomic counter until the image is done. */
template <int BLOCK, int MIN_BLOCKS, int SMEM_STACK, bool TOP_SMEM, bool COUNT, bool RECORD>
__global__ void __launch_bounds__(BLOCK, MIN_BLOCKS)
k_render_persistent(const SceneDev sc, const Work w, Counters *cnt)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	uint2 *s_stack_all = reinterpret_cast<uint2 *>(smem_raw);
	float4 *s_top = reinterpret_cast<float4 *>(smem_raw + (size_t)SMEM_STACK * BLOCK * sizeof(uint2));
	if (TOP_SMEM) {
		for (uint32_t i = threadIdx.x; i < sc.top_pairs * 4u; i += BLOCK) s_top[i] = __ldg(sc.pairs + i);
		__syncthreads();
	}
	uint2 *s_stack = s_stack_all + threadIdx.x;
	const uint32_t lane = threadIdx.x & 31u;
	for (;;) {
		uint32_t unit = 0;
		if (lane == 0) unit = atomicAdd(w.counter, 1u);
		unit = __shfl_sync(0xffffffffu, unit, 0);
		if (unit >= w.num_units) break;
		uint32_t x, y;
		size_t out;
		if (unit_pixel(w, unit + w.tile_begin * 32u, lane, x, y, out))
			trace_pixel<SMEM_STACK, TOP_SMEM, COUNT, RECORD>(sc, w, s_top, s_stack, x, y, out, cnt);
		__syncwarp();
	}
}

/* --------------------------------------------------------------------------
 * Thread-level ray packets.  ncu on the one-ray-per-thread kernel (profiles/)
 * shows the LSU register write-back saturated (~90 %): every lane pulls the
 * same 64-byte node pair through L1 for one ray's 26 flops.  Here each lane
 * carries NR = RX x RY neighbouring pixels and walks the tree ONCE for them:
 * a pair is loaded once per lane and slab-tested against all NR rays, a child
 * is entered if any ray enters it, a triangle is loaded once and tested by the
 * rays whose own test of its leaf box passed.  Per ray nothing changes: child
 * boxes lie inside parent boxes and the slab test is monotone, so a ray that
 * failed a box fails everything below it -- its own candidate set, order of
 * acceptance (min distance, ties to the smaller leaf index) and culling bound
 * are exactly those of the single-ray traversal.
 * ------------------------------------------------------------------------ */
template <int NR, int SMEM_STACK, bool COUNT, int OCT>
RTX_DEV void traverse_packet(const SceneDev &sc, const float4 *__restrict__ pairs, uint2 *__restrict__ s_stack,
                             const f3 (&d)[NR], uint32_t active, HitRec (&best)[NR], Counters *cnt)
{
	const float max_distance = 100000.0f;                              /* intersect_kernel.cl:292 */
	const f3 o = make_f3(0.0f, 0.0f, 2.0f);                            /* :284 */
	f3 id[NR];
	float cull[NR], maxid[NR];
#pragma unroll
	for (int r = 0; r < NR; ++r) {
		id[r] = make_f3(rn_div(1.0f, d[r].x), rn_div(1.0f, d[r].y), rn_div(1.0f, d[r].z));
		cull[r] = (active >> r) & 1u ? __int_as_float(0x7f800000) : __int_as_float(0xff800000);   /* -inf: enters nothing */
		maxid[r] = fminf(fmaxf(fmaxf(fabsf(id[r].x), fabsf(id[r].y)), fabsf(id[r].z)), 1e30f);       /* see box_slack */
	}
	const float abs_margin = 1e-5f * fmaxf(2.0f, sc.scene_scale);     /* |d| = 1 */
	uint2 l_stack[RTX_STACK_MAX - SMEM_STACK];
	int sp = 0;
	int cur = 0;
	uint32_t mask = 0;                   /* rays that entered the box of `cur` */
	unsigned long long visits = 0, tests = 0, lbtests = 0;
	const int stride = blockDim.x;
	constexpr uint32_t MBITS = (1u << NR) - 1u;

	for (;;) {
		while (cur >= 0) {
			const float4 *q = pairs + 4 * RTX_IDX(cur, sc.pair_count);
			const float4 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2), q3 = __ldg(q + 3);
			if (COUNT) visits += 2 * NR;
			uint32_t mL = 0, mR = 0;
			float tL = __int_as_float(0x7f800000), tR = __int_as_float(0x7f800000);
#pragma unroll
			for (int r = 0; r < NR; ++r) {
				const Slab L = slab_interval<true, OCT>(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, o, id[r]);
				const Slab R = slab_interval<true, OCT>(q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, o, id[r]);
				const float cl = __fmaf_rn(-q1.w, maxid[r], L.tmin), cr = __fmaf_rn(-q3.w, maxid[r], R.tmin);
				const bool hl = L.tmin <= L.tmax && L.tmin < max_distance && cl < cull[r] && L.tmax > 0.0f;
				const bool hr = R.tmin <= R.tmax && R.tmin < max_distance && cr < cull[r] && R.tmax > 0.0f;
				if (hl) { mL |= 1u << r; tL = fminf(tL, cl); }
				if (hr) { mR |= 1u << r; tR = fminf(tR, cr); }
			}
			if (!(mL | mR)) goto pop;
			const int refL = __float_as_int(q1.z), refR = __float_as_int(q3.z);
			const bool r_first = mR && (!mL || tR < tL);
			if (mL && mR) {
				/* entry = (ref, entry parameter with the ray mask in its 4 low mantissa bits) */
				const uint32_t tbits = (__float_as_uint(r_first ? tL : tR) & ~0xFu) | (r_first ? mL : mR);
				const uint2 e = make_uint2((uint32_t)(r_first ? refL : refR), tbits);
				if (SMEM_STACK > 0 && sp < SMEM_STACK) s_stack[sp * stride] = e; else l_stack[RTX_IDX(sp - SMEM_STACK, RTX_STACK_MAX - SMEM_STACK)] = e;
				++sp;
			}
			cur = r_first ? refR : refL;
			mask = r_first ? mR : mL;
		}
		{
			const uint32_t enc = ~(uint32_t)cur;
			const uint32_t first = enc >> 3, count = (enc & 7u) + 1u;
			for (uint32_t k = 0; k < count; ++k) {
				const uint32_t tri = first + k;
				const float4 *q = sc.tris + 4 * RTX_IDX(tri, sc.num_tris);
				const float4 t0 = __ldg(q), t1 = __ldg(q + 1), t2 = __ldg(q + 2), t3 = __ldg(q + 3);
#pragma unroll
				for (int r = 0; r < NR; ++r) {
					if (!((mask >> r) & 1u)) continue;
					TriHit h;
					if (COUNT) ++tests;
					if (!triangle_test(t0, t1, t2, t3, o, d[r], cull[r], h)) continue;
					if (!(h.dist < best[r].dist || (h.dist == best[r].dist && tri < best[r].tri))) continue;
					if (sc.verify_leafbox) {
						const float4 lo = __ldg(sc.leafbox + 2 * RTX_IDX(tri, sc.num_tris)), hi = __ldg(sc.leafbox + 2 * RTX_IDX(tri, sc.num_tris) + 1);
						if (COUNT) ++lbtests;
						if (!aabb_exact(make_f3(lo.x, lo.y, lo.z), make_f3(hi.x, hi.y, hi.z), o, d[r], max_distance)) continue;
					}
					best[r].dist = h.dist; best[r].tri = tri; best[r].s = h.s; best[r].t = h.t;
					cull[r] = h.dist * 1.0001f + abs_margin;
				}
			}
		}
pop:
		for (;;) {
			if (sp == 0) goto done;
			--sp;
			const uint2 e = (SMEM_STACK > 0 && sp < SMEM_STACK) ? s_stack[sp * stride] : l_stack[sp - SMEM_STACK];
			const float t = __uint_as_float(e.y & ~0xFu);
			uint32_t m = 0;
#pragma unroll
			for (int r = 0; r < NR; ++r)
				if (t < cull[r]) m |= 1u << r;
			m &= e.y & MBITS;
			if (m) { cur = (int)e.x; mask = m; break; }
		}
	}
done:
	if (COUNT) {
		atomicAdd(&cnt->node_visits, visits);
		atomicAdd(&cnt->tri_tests, tests);
		atomicAdd(&cnt->leafbox_tests, lbtests);
	}
}

/* --------------------------------------------------------------------------
 * Frustum front end for coherent primary-ray packets.
 *
 * A warp's 16x8-pixel packet (128 rays from the fixed camera) is a thin frustum: on the sibenik stand-in
 * at 4K the union over the packet of all leaf boxes any of its rays enters is ~11 triangles, barely more
 * than one ray's own ~9 (tools/packet_stats.py).  So the tree is walked ONCE per packet against the
 * frustum (level 1, lanes in parallel over a breadth-first queue), giving a short depth-sorted list of
 * candidate leaves, and each ray then runs the reference's own two tests over that list (level 2): the
 * literal-equivalent slab test of the triangle's leaf box and the triangle test.  Per ray this is exactly
 * the reference's candidate rule (point 1): interior boxes only steer the search, so a
 * conservative frustum test may replace 128 per-ray tests.  The frustum test is widened by 1e-4 of the
 * scene scale, far above float rounding, so no leaf a ray's own slab test would pass is ever dropped.
 * Queue or list overflow (a packet grazing hundreds of triangles) falls back to the per-ray traversal.
 * ------------------------------------------------------------------------ */
#define RTX_QCAP 128     /* breadth-first queue of node pairs, per warp */
#define RTX_CCAP 192     /* candidate leaves per 32x32-pixel tile */
#define RTX_LIST_STRIDE (2 * RTX_CCAP + 4)   /* words per tile list: [count, tile x, tile y, pad, enc[CCAP], key[CCAP]] (16-byte aligned) */

struct Frustum { float u0, u1, v0, v1, margin; };   /* ray = (0,0,2) + s * (u, v, -1), s >= 0 */

/* Frustum of the pixel rectangle [X0, X0+PW) x [Y0, Y0+PH) (covers any sub-pixel offset, jittered or not). */
RTX_DEV Frustum make_frustum(const Camera &cam, float X0, float Y0, float PW, float PH, float scene_scale)
{
	Frustum f;
	/* a conservative bound, not a deciding value: a reciprocal and four products instead of four IEEE divisions (2 % of
	 * the list kernel's instructions); the 1e-5 widening below is two orders above the difference */
	const float inv_a = __frcp_rn(cam.a);
	const float ua = X0 * inv_a - cam.w_over_2a, ub = (X0 + PW) * inv_a - cam.w_over_2a;
	const float va = -(Y0 * inv_a - cam.h_over_2a), vb = -((Y0 + PH) * inv_a - cam.h_over_2a);
	f.margin = 1e-4f * fmaxf(2.0f, scene_scale);
	f.u0 = fminf(ua, ub) - 1e-5f; f.u1 = fmaxf(ua, ub) + 1e-5f;
	f.v0 = fminf(va, vb) - 1e-5f; f.v1 = fmaxf(va, vb) + 1e-5f;
	return f;
}

/* Does the box reach into the frustum?  s = depth along -z.  Returns the (conservative) entry depth. */
RTX_DEV bool frustum_box(float lx, float ly, float lz, float hx, float hy, float hz, const Frustum &f, float &entry)
{
	const float s1 = (2.0f - lz) + f.margin;                 /* far depth */
	const float s0 = fmaxf((2.0f - hz) - f.margin, 0.0f);    /* near depth, clamped to the camera plane */
	entry = s0;
	const float xa = f.u0 * s0, xb = f.u0 * s1, xc = f.u1 * s0, xd = f.u1 * s1;
	const float ya = f.v0 * s0, yb = f.v0 * s1, yc = f.v1 * s0, yd = f.v1 * s1;
	return s1 > 0.0f &&
	       hx >= fminf(xa, xb) - f.margin && lx <= fmaxf(xc, xd) + f.margin &&
	       hy >= fminf(ya, yb) - f.margin && ly <= fmaxf(yc, yd) + f.margin;
}

/* Level 1: a warp walks the tree breadth-first against a frustum, lanes in parallel over the queue, and
 * writes the candidate leaves sorted by entry depth to out[4..] (codes) and out[4+cap..] (depths).
 * out[0] = count, or -1 on queue/list overflow. */
RTX_DEV void collect_frustum(const SceneDev &sc, const Frustum &f, int *__restrict__ s_queue, uint32_t *__restrict__ s_tmp,
                             float *__restrict__ s_tkey, int cap, uint32_t *__restrict__ out, uint32_t lane, unsigned int *overflow_counter)
{
	const unsigned full = 0xffffffffu, lt = (1u << lane) - 1u;
	int head = 0, tail = 1, ncand = 0;
	if (lane == 0) s_queue[0] = 0;
	__syncwarp();
	while (head < tail) {
		const int n = min(32, tail - head);
		bool iL = false, iR = false, fL = false, fR = false;
		int refL = 0, refR = 0;
		float eL = 0.f, eR = 0.f;
		if ((int)lane < n) {
			const int pair = s_queue[(head + lane) % RTX_QCAP];
			const float4 *q = sc.pairs + 4 * RTX_IDX(pair, sc.pair_count);
			const float4 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2), q3 = __ldg(q + 3);
			const bool pL = frustum_box(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, f, eL);
			const bool pR = frustum_box(q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, f, eR);
			eL -= q1.w;                 /* a hit may lie up to the box's slack in front of the box (box_slack) */
			eR -= q3.w;
			if (!(eL == eL)) eL = __int_as_float(0xff800000);          /* keys must stay ordered: no NaN */
			if (!(eR == eR)) eR = __int_as_float(0xff800000);
			refL = __float_as_int(q1.z); refR = __float_as_int(q3.z);
			iL = pL && refL >= 0; fL = pL && refL < 0;
			iR = pR && refR >= 0; fR = pR && refR < 0;
		}
		head += n;
		const unsigned bIL = __ballot_sync(full, iL), bIR = __ballot_sync(full, iR);
		const unsigned bFL = __ballot_sync(full, fL), bFR = __ballot_sync(full, fR);
		const int nI = __popc(bIL) + __popc(bIR), nF = __popc(bFL) + __popc(bFR);
		if (tail - head + nI > RTX_QCAP || ncand + nF > cap) { ncand = -1; break; }
		if (iL) s_queue[(tail + __popc(bIL & lt)) % RTX_QCAP] = refL;
		if (iR) s_queue[(tail + __popc(bIL) + __popc(bIR & lt)) % RTX_QCAP] = refR;
		if (fL) { const int k = (int)RTX_IDX(ncand + __popc(bFL & lt), cap); s_tmp[k] = ~(uint32_t)refL; s_tkey[k] = eL; }
		if (fR) { const int k = (int)RTX_IDX(ncand + __popc(bFL) + __popc(bFR & lt), cap); s_tmp[k] = ~(uint32_t)refR; s_tkey[k] = eR; }
		tail += nI;
		ncand += nF;
		__syncwarp();
	}
	if (lane == 0) {
		out[0] = (uint32_t)ncand;
		if (ncand < 0 && overflow_counter) atomicAdd(overflow_counter, 1u);
	}
	/* rank sort by entry depth (ties by position): the list is short */
	for (int i = lane; i < ncand; i += 32) {
		const float ki = s_tkey[i];
		int rank = 0;
		for (int j = 0; j < ncand; ++j) {
			const float kj = s_tkey[j];
			rank += (kj < ki || (kj == ki && j < i)) ? 1 : 0;
		}
		out[4 + rank] = s_tmp[i];
		out[4 + cap + rank] = __float_as_uint(ki);
	}
}

#define RTX_SUPER 4                      /* a super-tile is 4x4 tiles = 128x128 pixels */
#define RTX_SCAP 384                     /* candidate leaves per super-tile */
#define RTX_SLIST_STRIDE (2 * RTX_SCAP + 4)

/* Level 1a: one warp per 128x128-pixel super-tile (indexed over the whole image, whatever the rank). */
__global__ void __launch_bounds__(128)
k_frustum_collect_super(const SceneDev sc, const Work w, uint32_t *__restrict__ slists)
{
	__shared__ int s_queue_all[4][RTX_QCAP];
	__shared__ uint32_t s_tmp_all[4][RTX_SCAP];
	__shared__ float s_tkey_all[4][RTX_SCAP];
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
	const uint32_t stiles_x = (w.tiles_x + RTX_SUPER - 1) / RTX_SUPER, stiles_y = (w.tiles_y + RTX_SUPER - 1) / RTX_SUPER;
	const uint32_t st = blockIdx.x * 4 + warp;
	if (st >= stiles_x * stiles_y) return;
	const uint32_t sx = st % stiles_x, sy = st / stiles_x;
	if (w.world > 1) {           /* a rank only needs the super-tiles that hold one of its tiles (warp-uniform test) */
		bool mine = false;
		for (uint32_t j = 0; j < RTX_SUPER * RTX_SUPER; ++j) {
			const uint32_t tx = sx * RTX_SUPER + (j % RTX_SUPER), ty = sy * RTX_SUPER + (j / RTX_SUPER);
			mine = mine || (tx < w.tiles_x && ty < w.tiles_y && (ty * w.tiles_x + tx) % w.world == w.rank);
		}
		if (!mine) return;
	}
	const float px = (float)(RTX_TILE * RTX_SUPER);
	const Frustum f = make_frustum(w.cam, (float)sx * px, (float)sy * px, px, px, sc.scene_scale);
	collect_frustum(sc, f, s_queue_all[warp], s_tmp_all[warp], s_tkey_all[warp], RTX_SCAP, slists + (size_t)st * RTX_SLIST_STRIDE, lane, nullptr);
}

/* Level 1b: one warp per 32x32-pixel tile of this rank: filter the super-tile's list with the tile's own
 * frustum (order preserved); if the super-tile overflowed, walk the tree for the tile itself.
 * count = -1: the tile's packets use the per-ray traversal. */
__global__ void __launch_bounds__(256)
k_frustum_collect(const SceneDev sc, const Work w, const uint32_t *__restrict__ slists, uint32_t *__restrict__ lists)
{
	__shared__ int s_queue_all[8][RTX_QCAP];
	__shared__ uint32_t s_tmp_all[8][RTX_CCAP];
	__shared__ float s_tkey_all[8][RTX_CCAP];
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
	const uint32_t ltile = blockIdx.x * 8 + warp;
	if (ltile >= w.local_tiles) return;
	const uint32_t tile = ltile * w.world + w.rank;
	const uint32_t tx = tile % w.tiles_x, ty = tile / w.tiles_x;
	const Frustum f = make_frustum(w.cam, (float)(tx * RTX_TILE), (float)(ty * RTX_TILE), (float)RTX_TILE, (float)RTX_TILE, sc.scene_scale);
	uint32_t *out = lists + (size_t)ltile * RTX_LIST_STRIDE;
	if (lane == 0) { out[1] = tx; out[2] = ty; }        /* the list kernel reads them with the count: no division per packet */
	const uint32_t stiles_x = (w.tiles_x + RTX_SUPER - 1) / RTX_SUPER;
	const uint32_t *sl = slists ? slists + ((size_t)(ty / RTX_SUPER) * stiles_x + tx / RTX_SUPER) * RTX_SLIST_STRIDE : nullptr;
	const int n = sl ? (int)__ldg(sl) : -1;             /* no super-tile pass (few tiles): walk the tree per tile */
	if (n < 0) {
		collect_frustum(sc, f, s_queue_all[warp], s_tmp_all[warp], s_tkey_all[warp], RTX_CCAP, out, lane, w.overflow_tiles);
		return;
	}
	const unsigned lt = (1u << lane) - 1u;
	int m = 0;
	for (int base = 0; base < n; base += 32) {          /* warp-uniform trip count */
		const int i = base + (int)lane;
		bool keep = false;
		uint32_t enc = 0, key = 0;
		if (i < n) {
			enc = __ldg(sl + 4 + i);
			key = __ldg(sl + 4 + RTX_SCAP + i);
			keep = true;
			if ((enc & 7u) == 0u) {
				const uint32_t tri = enc >> 3;
				const float4 lo = __ldg(sc.leafbox + 2 * RTX_IDX(tri, sc.num_tris)), hi = __ldg(sc.leafbox + 2 * RTX_IDX(tri, sc.num_tris) + 1);
				float e;
				keep = frustum_box(lo.x, lo.y, lo.z, hi.x, hi.y, hi.z, f, e);
			}
		}
		const unsigned b = __ballot_sync(0xffffffffu, keep);
		const int k = m + __popc(b & lt);
		if (keep && k < RTX_CCAP) { out[4 + k] = enc; out[4 + RTX_CCAP + k] = key; }
		m += __popc(b);
	}
	if (lane == 0) {
		out[0] = m <= RTX_CCAP ? (uint32_t)m : 0xffffffffu;
		if (m > RTX_CCAP) atomicAdd(w.overflow_tiles, 1u);
	}
}

/* A packet's share of its tile's list: lanes test the candidates' leaf boxes against the packet's own
 * (16x8-pixel) frustum in parallel and compact the survivors, order preserved, into shared memory. */
RTX_DEV int filter_candidates(const SceneDev &sc, const Frustum &f, const uint32_t *__restrict__ list, int n,
                              uint32_t *__restrict__ s_cand, float *__restrict__ s_key, uint32_t lane)
{
	const unsigned lt = (1u << lane) - 1u;
	int m = 0;
	for (int base = 0; base < n; base += 32) {
		const int i = base + (int)lane;
		bool keep = false;
		uint32_t enc = 0;
		float key = 0.f;
		if (i < n) {
			enc = __ldg(list + 4 + i);
			key = __uint_as_float(__ldg(list + 4 + RTX_CCAP + i));
			keep = true;
			if ((enc & 7u) == 0u) {                       /* single-triangle leaf: its box is the triangle's leaf box */
				const uint32_t tri = enc >> 3;
				const float4 lo = __ldg(sc.leafbox + 2 * RTX_IDX(tri, sc.num_tris)), hi = __ldg(sc.leafbox + 2 * RTX_IDX(tri, sc.num_tris) + 1);
				float e;
				keep = frustum_box(lo.x, lo.y, lo.z, hi.x, hi.y, hi.z, f, e);
			}
		}
		const unsigned b = __ballot_sync(0xffffffffu, keep);
		if (keep) { const int k = (int)RTX_IDX(m + __popc(b & lt), RTX_CCAP); s_cand[k] = enc; s_key[k] = key; }
		m += __popc(b);
	}
	__syncwarp();
	return m;
}

/* Level 2: every ray runs the reference's leaf-box test + triangle test over the packet's candidate leaves,
 * nearest first; a ray stops caring once the entry depth exceeds its culling bound. */
template <int NR, bool COUNT>
RTX_DEV void intersect_candidates(const SceneDev &sc, const uint32_t *__restrict__ s_cand, const float *__restrict__ s_key, int ncand,
                                  const f3 (&d)[NR], uint32_t active, HitRec (&best)[NR], Counters *cnt)
{
	const float max_distance = 100000.0f;
	const f3 o = make_f3(0.0f, 0.0f, 2.0f);
	f3 id[NR];
	float cull[NR];
#pragma unroll
	for (int r = 0; r < NR; ++r) {
		id[r] = make_f3(rn_div(1.0f, d[r].x), rn_div(1.0f, d[r].y), rn_div(1.0f, d[r].z));
		cull[r] = (active >> r) & 1u ? __int_as_float(0x7f800000) : __int_as_float(0xff800000);
	}
	const float abs_margin = 1e-5f * fmaxf(2.0f, sc.scene_scale);
	unsigned long long visits = 0, tests = 0;
	for (int k = 0; k < ncand; ++k) {
		const uint32_t enc = s_cand[k];
		const float key = s_key[k];                  /* entry depth <= entry parameter of every ray (|(u,v,-1)| >= 1) */
		uint32_t want = 0;
#pragma unroll
		for (int r = 0; r < NR; ++r)
			if (key < cull[r]) want |= 1u << r;
		if (!__any_sync(0xffffffffu, want != 0)) break;       /* sorted: nobody wants the rest either */
		if (want) {
			const uint32_t first = enc >> 3, last = first + (enc & 7u);
			for (uint32_t tri = first; tri <= last; ++tri) {
				/* (lo.x, lo.y, lo.z - 2, A), (hi.x, hi.y, hi.z - 2, 0): the subtraction of the camera's z (:45-49) and the
				 * plane term A of the triangle (:71-72) do not depend on the ray and were rounded once at upload */
				const float4 lo = __ldg(sc.pleafbox + 2 * RTX_IDX(tri, sc.num_tris)), hi = __ldg(sc.pleafbox + 2 * RTX_IDX(tri, sc.num_tris) + 1);
				uint32_t m = 0;
#pragma unroll
				for (int r = 0; r < NR; ++r) {
					if (!((want >> r) & 1u)) continue;
					if (COUNT) ++visits;
					const float ax = rn_mul(lo.x, id[r].x), bx = rn_mul(hi.x, id[r].x), ay = rn_mul(lo.y, id[r].y), by = rn_mul(hi.y, id[r].y);
					const float az = rn_mul(lo.z, id[r].z), bz = rn_mul(hi.z, id[r].z);
					const float tmin = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), bz);          /* d.z < 0: the far z plane is the entry */
					const float tmax = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), az);
					if (tmin <= tmax && tmin < max_distance && tmax > 0.0f) m |= 1u << r;      /* intersect_kernel.cl:41-60 */
				}
				if (!m) continue;
				const float4 *q = sc.tris + 4 * RTX_IDX(tri, sc.num_tris);
				const float4 t0 = __ldg(q), t1 = __ldg(q + 1), t2 = __ldg(q + 2), t3 = __ldg(q + 3);
#pragma unroll
				for (int r = 0; r < NR; ++r) {
					if (!((m >> r) & 1u)) continue;
					TriHit h;
					if (COUNT) ++tests;
					if (!triangle_test_primary(t0, t1, t2, t3, lo.w, d[r], cull[r], h)) continue;
					if (!(h.dist < best[r].dist || (h.dist == best[r].dist && tri < best[r].tri))) continue;
					best[r].dist = h.dist; best[r].tri = tri; best[r].s = h.s; best[r].t = h.t;
					cull[r] = h.dist * 1.0001f + abs_margin;
				}
			}
		}
	}
	if (COUNT) {
		atomicAdd(&cnt->leafbox_tests, visits);
		atomicAdd(&cnt->tri_tests, tests);
	}
}

/* Persistent packet kernel: a warp's unit of work is an (8 RX) x (4 RY) pixel block, lane (lx, ly) of the
 * 8 x 4 lane grid owns the RX x RY pixels at (lx RX, ly RY).  32x32 tiles hold 32 / (RX RY) units. */
/* MODE 0: per-ray/packet traversal for every tile.
 * MODE 1: candidate lists only (tiles whose list overflowed are skipped) -- no traversal code in the kernel.
 * MODE 2: traversal for the tiles MODE 1 skipped; exits at once when k_frustum_collect saw no overflow. */
/* STORE: the fused tile store of rtx_render_store (a separate instantiation: the plain kernel keeps its registers). */
template <int BLOCK, int MIN_BLOCKS, int SMEM_STACK, bool COUNT, bool RECORD, int RX, int RY, int MODE, bool STORE = false>
__global__ void __launch_bounds__(BLOCK, MIN_BLOCKS)
k_render_packet(const SceneDev sc, const Work w, Counters *cnt)
{
	if (MODE == 2 && *w.overflow_tiles == 0u) return;
	constexpr int NR = RX * RY;
	constexpr uint32_t UPT = 32u / NR;                 /* units per tile */
	constexpr uint32_t UX = RTX_TILE / (8 * RX);       /* units per tile row */
	extern __shared__ __align__(16) unsigned char smem_raw[];
	uint2 *s_stack = reinterpret_cast<uint2 *>(smem_raw) + threadIdx.x;
	const uint32_t lane = threadIdx.x & 31u;
	/* per-warp candidate list of the frustum front end, after the stacks */
	uint32_t *s_cand = reinterpret_cast<uint32_t *>(smem_raw + (size_t)SMEM_STACK * BLOCK * sizeof(uint2)) + (threadIdx.x >> 5) * (2 * RTX_CCAP);
	float *s_key = reinterpret_cast<float *>(s_cand + RTX_CCAP);
	const f3 o = make_f3(0.0f, 0.0f, 2.0f);
	for (;;) {
		uint32_t unit = 0;
		if (lane == 0) unit = atomicAdd(w.counter, 1u);
		unit = __shfl_sync(0xffffffffu, unit, 0);
		if (unit >= w.tile_count * UPT) break;
		const uint32_t ltile = w.tile_begin + unit / UPT, sub = unit % UPT;
		int nlist = -1;
		uint32_t tx, ty;
		if (MODE != 0) {
			const uint4 hdr = __ldg(reinterpret_cast<const uint4 *>(w.lists + (size_t)ltile * RTX_LIST_STRIDE));   /* count, tile x, tile y */
			nlist = (int)hdr.x;
			if ((MODE == 1) == (nlist < 0)) continue;       /* MODE 1 takes listed tiles, MODE 2 the overflowed ones */
			tx = hdr.y;
			ty = hdr.z;
		} else {
			const uint32_t tile = ltile * w.world + w.rank;
			tx = tile % w.tiles_x;
			ty = tile / w.tiles_x;
		}
		const uint32_t px0 = (sub % UX) * (8 * RX) + (lane & 7u) * RX, py0 = (sub / UX) * (4 * RY) + (lane >> 3) * RY;
		f3 d[NR];
		HitRec best[NR];
		size_t out[NR];
		uint32_t valid = 0, packet = 0;
		int oct = -1;
		bool same = true;
#pragma unroll
		for (int r = 0; r < NR; ++r) {
			const uint32_t px = px0 + (r % RX), py = py0 + (r / RX);
			const uint32_t x = tx * RTX_TILE + px, y = ty * RTX_TILE + py;
			out[r] = (w.world > 1 && !w.rowmajor) ? (size_t)ltile * (RTX_TILE * RTX_TILE) + py * RTX_TILE + px : (size_t)y * w.cam.W + x;
			best[r].dist = __int_as_float(0x7f800000); best[r].tri = 0xffffffffu; best[r].s = best[r].t = 0.f;
			d[r] = primary_dir(w.cam, x, y);
			if (ty < w.tiles_y && x < w.cam.W && y < w.cam.H) {
				valid |= 1u << r;
				if (ray_is_plain(o, d[r])) {
					packet |= 1u << r;
					const int oc = (d[r].x < 0.0f ? 1 : 0) | (d[r].y < 0.0f ? 2 : 0) | (d[r].z < 0.0f ? 0 : 4);
					if (oct < 0) oct = oc; else same = same && oc == oct;
				}
			}
		}
		if (!w.ordered_ok) packet = 0;
		const unsigned anyp = __ballot_sync(0xffffffffu, packet != 0);
		if (MODE == 1) {
			if (anyp) {
				const Frustum f = make_frustum(w.cam, (float)(tx * RTX_TILE + (sub % UX) * (8 * RX)), (float)(ty * RTX_TILE + (sub / UX) * (4 * RY)),
				                               (float)(8 * RX), (float)(4 * RY), sc.scene_scale);
				const int m = filter_candidates(sc, f, w.lists + (size_t)ltile * RTX_LIST_STRIDE, nlist, s_cand, s_key, lane);
				intersect_candidates<NR, COUNT>(sc, s_cand, s_key, m, d, packet, best, cnt);
			}
		} else if (anyp) {
			/* all packet rays of the warp in one octant (d.z < 0)?  then the pre-swapped node copy */
			if (MODE == 2 && COUNT && lane == 0) atomicAdd(&cnt->overflow_packets, 1ull);
			const int src = __ffs(anyp) - 1;
			const int oct0 = __shfl_sync(0xffffffffu, oct, src);
			const bool uniform = __all_sync(0xffffffffu, packet == 0 || (same && oct == oct0)) && oct0 < 4;
			if (packet) {
				if (uniform) {
					traverse_packet<NR, SMEM_STACK, COUNT, 0>(sc, sc.pairs + (size_t)oct0 * sc.num_pairs * 4, s_stack, d, packet, best, cnt);
				} else {
#pragma unroll 1
					for (int r = 0; r < NR; ++r)
						if ((packet >> r) & 1u)
							traverse_ordered<SMEM_STACK, false, COUNT, false, 4>(sc, sc.pairs, nullptr, s_stack, o, d[r], 100000.0f, best[r], cnt);
				}
			}
		}
#pragma unroll
		for (int r = 0; r < NR; ++r) {
			if (!((valid >> r) & 1u)) continue;
			if (!((packet >> r) & 1u)) {                 /* zero direction component or deep tree: literal walk */
				if (COUNT) atomicAdd(&cnt->exact_rays, 1ull);
				walk_reference<COUNT>(sc, o, d[r], 100000.0f, best[r], cnt);
			}
			float value = 0.0f;
			if (best[r].tri != 0xffffffffu) value = shade_hit(sc.tnormals, best[r].tri, best[r].s, best[r].t, d[r], w.cam.shading);
			w.image[out[r]] = value;
			if (RECORD) {
				w.face_id[out[r]] = best[r].tri != 0xffffffffu ? best[r].tri * 3u : 0xffffffffu;
				w.dist[out[r]] = best[r].dist;
				if (w.hit_st) w.hit_st[out[r]] = make_float2(best[r].s, best[r].t);
			}
		}
		__syncwarp();
		if (STORE) {
			/* fused store: whoever finishes the tile's last unit sends the whole tile (compact, 4 KB, still in L2) to its
			 * place in the caller's image -- full 128-byte rows, posted writes, while the other warps keep tracing */
			unsigned int done = 0;
			if (lane == 0) { __threadfence(); done = atomicAdd(w.tile_done + ltile, 1u); }
			done = __shfl_sync(0xffffffffu, done, 0);
			if (done == UPT - 1u) {
				__threadfence();
				const float *src = w.image + (size_t)ltile * (RTX_TILE * RTX_TILE);
				const uint32_t x0 = tx * RTX_TILE, y0 = ty * RTX_TILE;
				if ((w.cam.W & 3u) == 0 && (reinterpret_cast<uintptr_t>(w.store_image) & 15u) == 0 && (reinterpret_cast<uintptr_t>(w.image) & 15u) == 0) {
#pragma unroll 2
					for (uint32_t i = lane; i < RTX_TILE * RTX_TILE / 4; i += 32) {
						const uint32_t py = i >> 3, px = (i & 7u) << 2;
						if (y0 + py < w.cam.H && x0 + px < w.cam.W)
							__stcs(reinterpret_cast<float4 *>(w.store_image + (size_t)(y0 + py) * w.cam.W + x0 + px),
							       __ldcg(reinterpret_cast<const float4 *>(src + py * RTX_TILE + px)));
					}
				} else {
					for (uint32_t i = lane; i < RTX_TILE * RTX_TILE; i += 32) {
						const uint32_t px = i & 31u, py = i >> 5;
						if (x0 + px < w.cam.W && y0 + py < w.cam.H) w.store_image[(size_t)(y0 + py) * w.cam.W + x0 + px] = __ldcg(src + i);
					}
				}
			}
			__syncwarp();
		}
	}
}

/* The reference's launch shape: one thread per pixel, literal walk
 * (opencl_host.cc:145 uses 16x16 work groups; 8x4 warps of a 32x32 tile here). */
template <bool COUNT, bool RECORD>
__global__ void __launch_bounds__(256)
k_render_exhaustive(const SceneDev sc, const Work w, Counters *cnt)
{
	const uint32_t unit = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	if (unit >= w.num_units) return;
	uint32_t x, y;
	size_t out;
	if (!unit_pixel(w, unit + w.tile_begin * 32u, threadIdx.x & 31u, x, y, out)) return;
	const f3 o = make_f3(0.0f, 0.0f, 2.0f);
	const f3 d = primary_dir(w.cam, x, y);
	HitRec best;
	best.dist = __int_as_float(0x7f800000); best.tri = 0xffffffffu; best.s = best.t = 0.f;
	walk_reference<COUNT>(sc, o, d, 100000.0f, best, cnt);
	float value = 0.0f;
	if (best.tri != 0xffffffffu) value = shade_hit(sc.tnormals, best.tri, best.s, best.t, d, w.cam.shading);
	w.image[out] = value;
	if (RECORD) {
		w.face_id[out] = best.tri != 0xffffffffu ? best.tri * 3u : 0xffffffffu;
		w.dist[out] = best.dist;
		if (w.hit_st) w.hit_st[out] = make_float2(best.s, best.t);
	}
}

/* ---------------------------- arbitrary rays ----------------------------- */

struct RayWork {
	const float4 *origins, *dirs;   /* NULL: generate (seed, first + i) */
	uint32_t seed;
	unsigned long long first;
	unsigned long long nrays;
	f3 bbmin, bbmax;
	float max_distance;
	unsigned int *counter;
	uint32_t *face_id;              /* optional */
	float *dist;                    /* optional */
	unsigned long long *hit_count, *sum_face_id;
	int ordered_ok;
	int exhaustive;
};

template <int BLOCK, int MIN_BLOCKS, int SMEM_STACK, bool TOP_SMEM, bool COUNT>
__global__ void __launch_bounds__(BLOCK, MIN_BLOCKS)
k_trace_rays(const SceneDev sc, const RayWork w, Counters *cnt)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	uint2 *s_stack_all = reinterpret_cast<uint2 *>(smem_raw);
	float4 *s_top = reinterpret_cast<float4 *>(smem_raw + (size_t)SMEM_STACK * BLOCK * sizeof(uint2));
	if (TOP_SMEM) {
		for (uint32_t i = threadIdx.x; i < sc.top_pairs * 4u; i += BLOCK) s_top[i] = __ldg(sc.pairs + i);
		__syncthreads();
	}
	uint2 *s_stack = s_stack_all + threadIdx.x;
	const uint32_t lane = threadIdx.x & 31u;
	const unsigned long long nunits = (w.nrays + 31ull) >> 5;
	unsigned long long hits = 0, idsum = 0;
	for (;;) {
		uint32_t unit = 0;
		if (lane == 0) unit = atomicAdd(w.counter, 1u);
		unit = __shfl_sync(0xffffffffu, unit, 0);
		if (unit >= nunits) break;
		const unsigned long long i = ((unsigned long long)unit << 5) + lane;
		if (i < w.nrays) {
			f3 o, d;
			if (w.origins) {
				const float4 oo = __ldg(w.origins + i), dd = __ldg(w.dirs + i);
				o = make_f3(oo.x, oo.y, oo.z);
				d = make_f3(dd.x, dd.y, dd.z);
			} else {
				random_ray(w.seed, w.first + i, w.bbmin, w.bbmax, o, d);
			}
			HitRec best;
			closest_hit<SMEM_STACK, TOP_SMEM, COUNT, false>(sc, s_top, s_stack, w.ordered_ok != 0 && !w.exhaustive, o, d, w.max_distance, best, cnt);
			const uint32_t fid = best.tri != 0xffffffffu ? best.tri * 3u : 0xffffffffu;
			if (w.face_id) w.face_id[i] = fid;
			if (w.dist) w.dist[i] = best.dist;
			if (fid != 0xffffffffu) { ++hits; idsum += fid; }
		}
		__syncwarp();
	}
	if (w.hit_count) {
		for (int s = 16; s > 0; s >>= 1) {
			hits += __shfl_xor_sync(0xffffffffu, hits, s);
			idsum += __shfl_xor_sync(0xffffffffu, idsum, s);
		}
		if (lane == 0 && hits) { atomicAdd(w.hit_count, hits); atomicAdd(w.sum_face_id, idsum); }
	}
}

/* --------------------------------------------------------------------------
 * Persistent traversal for rays that are NOT coherent (arbitrary / random rays, primary rays over geometry
 * finer than the pixel grid).  ncu on the while-while kernel with random rays (profiles/r1_c5_*): 5.1 of 32
 * lanes active on average, 1.3 in the triangle test -- every lane reaches its leaves at a different time and
 * finished rays leave their lane idle.  Here a warp keeps its lanes busy in three warp-synchronous phases:
 *   fetch    when fewer than RTX_PT_REFILL lanes still own a ray, the idle lanes pull new rays from the
 *            global counter (one atomicAdd per warp, __ballot_sync/__popc to hand them out);
 *   descend  lanes with an interior node test its two children; a lane that reaches a leaf parks it as
 *            "pending" and keeps descending from its stack (speculative traversal) until it meets a second
 *            leaf; the phase ends when fewer than RTX_PT_MIN_DESCEND lanes still descend;
 *   leaves   all lanes with a pending leaf run the triangle tests together.
 * Per ray the tests, the acceptance rule and the culling bound are those of traverse_ordered; a parked
 * leaf is merely tested a little later, against a bound that can only have become tighter.
 * ------------------------------------------------------------------------ */
#ifndef RTX_PT_REFILL
#define RTX_PT_REFILL 20
#endif
#ifndef RTX_PT_MIN_DESCEND
#define RTX_PT_MIN_DESCEND 16     /* 8 / 12 / 16 measured on C5: 23.45 / 22.77 / 22.41 ms per 2^26 rays */
#endif
#define RTX_PT_DONE ((int)0x80000000)
#define RTX_PT_NONE ((int)0x80000001)

struct PtRay {
	f3 o, d, id;
	float max_distance, inv_len, abs_margin, cull, maxid;
	HitRec best;
	int cur, pend, sp;
	unsigned long long index;
};

template <int SMEM_STACK>
RTX_DEV int pt_pop(PtRay &r, const uint2 *__restrict__ s_stack, const uint2 *l_stack, int stride)
{
	while (r.sp > 0) {
		--r.sp;
		const uint2 e = (SMEM_STACK > 0 && r.sp < SMEM_STACK) ? s_stack[r.sp * stride] : l_stack[r.sp - SMEM_STACK];
		if (__uint_as_float(e.y) < r.cull) return (int)e.x;
	}
	return RTX_PT_DONE;
}

/* SOURCE 0: rays from arrays or from the counter-based generator (RayWork); SOURCE 1: primary rays (Work). */
template <int BLOCK, int MIN_BLOCKS, int SMEM_STACK, bool COUNT, bool RECORD, int SOURCE>
__global__ void __launch_bounds__(BLOCK, MIN_BLOCKS)
k_trace_persistent(const SceneDev sc, const RayWork rw, const Work pw, Counters *cnt)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	uint2 *s_stack = reinterpret_cast<uint2 *>(smem_raw) + threadIdx.x;
	const int stride = blockDim.x;
	uint2 l_stack[RTX_STACK_MAX - SMEM_STACK];
	const uint32_t lane = threadIdx.x & 31u;
	const unsigned full = 0xffffffffu, lt = (1u << lane) - 1u;
	const unsigned long long total = SOURCE == 0 ? rw.nrays : (unsigned long long)pw.num_units * 32ull;
	unsigned int *counter = SOURCE == 0 ? rw.counter : pw.counter;
	const bool ordered_ok = SOURCE == 0 ? (rw.ordered_ok != 0 && !rw.exhaustive) : pw.ordered_ok != 0;
	unsigned long long hits = 0, idsum = 0, visits = 0, tests = 0, lbtests = 0;
	PtRay r;
	bool has_ray = false, exhausted = false;
	r.cur = RTX_PT_DONE; r.pend = RTX_PT_NONE; r.sp = 0;

	for (;;) {
		/* ---------------- fetch ---------------- */
		const unsigned idle = __ballot_sync(full, !has_ray);
		if (!exhausted && __popc(idle) >= 32 - RTX_PT_REFILL + 1) {
			unsigned long long base = 0;
			const int want = __popc(idle);
			if (lane == 0) base = (unsigned long long)atomicAdd(counter, (unsigned int)want);
			base = __shfl_sync(full, base, 0);
			if (base + want >= total) exhausted = true;
			if (!has_ray) {
				const unsigned long long idx = base + __popc(idle & lt);
				if (idx < total) {
					bool valid = true;
					size_t out = 0;
					if (SOURCE == 0) {
						if (rw.origins) {
							const float4 oo = __ldg(rw.origins + idx), dd = __ldg(rw.dirs + idx);
							r.o = make_f3(oo.x, oo.y, oo.z);
							r.d = make_f3(dd.x, dd.y, dd.z);
						} else {
							random_ray(rw.seed, rw.first + idx, rw.bbmin, rw.bbmax, r.o, r.d);
						}
						r.max_distance = rw.max_distance;
						r.index = idx;
					} else {
						uint32_t x, y;
						valid = unit_pixel(pw, (uint32_t)(idx >> 5) + pw.tile_begin * 32u, (uint32_t)(idx & 31ull), x, y, out);
						r.o = make_f3(0.0f, 0.0f, 2.0f);
						r.d = primary_dir(pw.cam, x, y);
						r.max_distance = 100000.0f;
						r.index = out;
					}
					if (valid) {
						r.best.dist = __int_as_float(0x7f800000); r.best.tri = 0xffffffffu; r.best.s = r.best.t = 0.f;
						const bool plain = ray_is_plain(r.o, r.d);
						if (ordered_ok && plain) {
							r.id = make_f3(rn_div(1.0f, r.d.x), rn_div(1.0f, r.d.y), rn_div(1.0f, r.d.z));
							r.inv_len = rsqrtf(fmaxf(r.d.x * r.d.x + r.d.y * r.d.y + r.d.z * r.d.z, 1e-30f));
							r.abs_margin = 1e-5f * fmaxf(fmaxf(fabsf(r.o.x), fabsf(r.o.y)), fmaxf(fabsf(r.o.z), sc.scene_scale)) * r.inv_len;
							r.cull = __int_as_float(0x7f800000);
							r.maxid = fminf(fmaxf(fmaxf(fabsf(r.id.x), fabsf(r.id.y)), fabsf(r.id.z)), 1e30f);   /* see box_slack */
							r.cur = 0; r.pend = RTX_PT_NONE; r.sp = 0;
						} else {       /* zero direction component / deep tree: the literal walk, right away */
							if (COUNT) atomicAdd(&cnt->exact_rays, 1ull);
							walk_reference<COUNT>(sc, r.o, r.d, r.max_distance, r.best, cnt);
							r.cur = RTX_PT_DONE; r.pend = RTX_PT_NONE; r.sp = 0;
						}
						has_ray = true;
					}
				}
			}
		}
		if (__ballot_sync(full, has_ray) == 0u) {
			if (exhausted) break;
			continue;
		}
		/* ---------------- descend ---------------- */
		for (;;) {
			const bool descending = has_ray && r.cur >= 0;
			if (__popc(__ballot_sync(full, descending)) < RTX_PT_MIN_DESCEND && __ballot_sync(full, has_ray && r.pend != RTX_PT_NONE) != 0u) break;
			if (__ballot_sync(full, descending) == 0u) break;
			if (descending) {
				const float4 *q = sc.pairs + 4 * RTX_IDX(r.cur, sc.pair_count);
				const f4x2 qa = ldg256(q), qb = ldg256(q + 2);
				const float4 q0 = qa.a, q1 = qa.b, q2 = qb.a, q3 = qb.b;
				if (COUNT) visits += 2;
				const Slab L = slab_interval<false, 4>(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, r.o, r.id);
				const Slab R = slab_interval<false, 4>(q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, r.o, r.id);
				const float cL = __fmaf_rn(-q1.w, r.maxid, L.tmin), cR = __fmaf_rn(-q3.w, r.maxid, R.tmin);
				const bool hitL = L.tmin <= L.tmax && L.tmin < r.max_distance && cL < r.cull && L.tmax > 0.0f;
				const bool hitR = R.tmin <= R.tmax && R.tmin < r.max_distance && cR < r.cull && R.tmax > 0.0f;
				int next;
				if (!(hitL || hitR)) {
					next = pt_pop<SMEM_STACK>(r, s_stack, l_stack, stride);
				} else {
					const int refL = __float_as_int(q1.z), refR = __float_as_int(q3.z);
					const bool r_first = hitR && (!hitL || R.tmin < L.tmin);
					if (hitL && hitR) {
						const uint2 e = make_uint2((uint32_t)(r_first ? refL : refR), __float_as_uint(r_first ? cL : cR));
						if (SMEM_STACK > 0 && r.sp < SMEM_STACK) s_stack[r.sp * stride] = e; else l_stack[RTX_IDX(r.sp - SMEM_STACK, RTX_STACK_MAX - SMEM_STACK)] = e;
						++r.sp;
					}
					next = r_first ? refR : refL;
				}
				/* a leaf is parked; with nothing parked yet the lane keeps descending from its stack */
				if (next < 0 && next != RTX_PT_DONE && r.pend == RTX_PT_NONE) {
					r.pend = next;
					next = pt_pop<SMEM_STACK>(r, s_stack, l_stack, stride);
				}
				r.cur = next;
			}
		}
		/* ---------------- leaves ---------------- */
		if (has_ray && r.pend != RTX_PT_NONE) {
			const uint32_t enc = ~(uint32_t)r.pend;
			const uint32_t first = enc >> 3, count = (enc & 7u) + 1u;
			for (uint32_t k = 0; k < count; ++k) {
				const uint32_t tri = first + k;
				const float4 *q = sc.tris + 4 * RTX_IDX(tri, sc.num_tris);
				TriHit h;
				if (COUNT) ++tests;
				const f4x2 ta = ldg256(q), tb = ldg256(q + 2);
				if (!triangle_test(ta.a, ta.b, tb.a, tb.b, r.o, r.d, r.cull, h)) continue;
				if (!(h.dist < r.best.dist || (h.dist == r.best.dist && tri < r.best.tri))) continue;
				if (sc.verify_leafbox) {
					const float4 lo = __ldg(sc.leafbox + 2 * RTX_IDX(tri, sc.num_tris)), hi = __ldg(sc.leafbox + 2 * RTX_IDX(tri, sc.num_tris) + 1);
					if (COUNT) ++lbtests;
					if (!aabb_exact(make_f3(lo.x, lo.y, lo.z), make_f3(hi.x, hi.y, hi.z), r.o, r.d, r.max_distance)) continue;
				}
				r.best.dist = h.dist; r.best.tri = tri; r.best.s = h.s; r.best.t = h.t;
				r.cull = (h.dist * r.inv_len) * 1.0001f + r.abs_margin;

			}
			r.pend = RTX_PT_NONE;
			if (r.cur < 0 && r.cur != RTX_PT_DONE) {        /* a second leaf was waiting: park it, move on */
				r.pend = r.cur;
				r.cur = pt_pop<SMEM_STACK>(r, s_stack, l_stack, stride);
			}
		}
		/* ---------------- retire ---------------- */
		if (has_ray && r.cur == RTX_PT_DONE && r.pend == RTX_PT_NONE) {
			const uint32_t fid = r.best.tri != 0xffffffffu ? r.best.tri * 3u : 0xffffffffu;
			if (SOURCE == 0) {
				if (rw.face_id) rw.face_id[r.index] = fid;
				if (rw.dist) rw.dist[r.index] = r.best.dist;
				if (fid != 0xffffffffu) { ++hits; idsum += fid; }
			} else {
				float value = 0.0f;
				if (r.best.tri != 0xffffffffu) value = shade_hit(sc.tnormals, r.best.tri, r.best.s, r.best.t, r.d, pw.cam.shading);
				pw.image[r.index] = value;
				if (RECORD) {
					pw.face_id[r.index] = fid;
					pw.dist[r.index] = r.best.dist;
					if (pw.hit_st) pw.hit_st[r.index] = make_float2(r.best.s, r.best.t);
				}
			}
			has_ray = false;
		}
	}
	if (SOURCE == 0 && rw.hit_count) {
		for (int s = 16; s > 0; s >>= 1) {
			hits += __shfl_xor_sync(full, hits, s);
			idsum += __shfl_xor_sync(full, idsum, s);
		}
		if (lane == 0 && hits) { atomicAdd(rw.hit_count, hits); atomicAdd(rw.sum_face_id, idsum); }
	}
	if (COUNT) {
		atomicAdd(&cnt->node_visits, visits);
		atomicAdd(&cnt->tri_tests, tests);
		atomicAdd(&cnt->leafbox_tests, lbtests);
	}
}

/* --------------------------------------------------------------------------
 * Ambient occlusion (intersect_kernel.cl:128-183, 214-277, 305-307): a second pass over the hit pixels.
 * The primary pass records (triangle, s, t); this kernel recomputes the hit point and the smooth normal with
 * the operations of :71-85 and :118-127 (so they are the values the reference held in registers), shoots
 * the occlusion rays and multiplies the pixel.  An occlusion ray only asks "does ANY candidate triangle
 * hit?" (:251,:264,:270 use the bool); candidates are, as everywhere, the triangles whose leaf box passes
 * the slab test with max_distance = AO_MAX_DISTANCE, so any traversal that finds one may stop.
 * Transcendentals follow the oracle's definition: evaluate in double, round once to float.
 * ------------------------------------------------------------------------ */
struct AoParams {
	int method;             /* 0 uniform rings, 1 random hemisphere */
	uint32_t samples;
	float max_distance;
	int alpha_min, alpha_max;
	const float4 *ring;     /* method 0: ring[0].x = bits(number of rays), ring[1..] = (xs, ys, zs, 0) per ray */
	uint32_t ring_cap;
};

RTX_DEV float t_sin(float x) { return (float)sin((double)x); }
RTX_DEV float t_cos(float x) { return (float)cos((double)x); }
RTX_DEV float t_acos(float x) { return (float)acos((double)x); }
RTX_DEV float t_cospi(float x) { return (float)cos(__dmul_rn(3.14159265358979323846, (double)x)); }
RTX_DEV float t_sinpi(float x) { return (float)sin(__dmul_rn(3.14159265358979323846, (double)x)); }

RTX_DEV f3 normalize3(f3 a)
{
	const float len = rn_sqrt(dot3(a, a));
	return make_f3(rn_div(a.x, len), rn_div(a.y, len), rn_div(a.z, len));
}
RTX_DEV f3 perturb_smallest(f3 h)       /* :156-164, :226-234 */
{
	if (fabsf(h.x) <= fabsf(h.y) && fabsf(h.x) <= fabsf(h.z)) h.x = 1.0f;
	else if (fabsf(h.y) <= fabsf(h.x) && fabsf(h.y) <= fabsf(h.z)) h.y = 1.0f;
	else if (fabsf(h.z) <= fabsf(h.x) && fabsf(h.z) <= fabsf(h.y)) h.z = 1.0f;
	return h;
}
/* basis_x * xs + basis_y * ys + basis_z * zs, per component ((x + y) + z) */
RTX_DEV f3 combine3(f3 bx, float xs, f3 by, float ys, f3 bz, float zs)
{
	return make_f3(rn_add(rn_add(rn_mul(bx.x, xs), rn_mul(by.x, ys)), rn_mul(bz.x, zs)),
	               rn_add(rn_add(rn_mul(bx.y, xs), rn_mul(by.y, ys)), rn_mul(bz.y, zs)),
	               rn_add(rn_add(rn_mul(bx.z, xs), rn_mul(by.z, ys)), rn_mul(bz.z, zs)));
}
RTX_DEV uint32_t ao_random_int(uint32_t (&v)[4])      /* :128-135 */
{
	const uint32_t t = v[0] ^ (v[0] << 11u);
	v[0] = v[1]; v[1] = v[2]; v[2] = v[3];
	return v[3] = v[3] ^ (v[3] >> 19u) ^ (t ^ (t >> 8u));
}
RTX_DEV float ao_random_float(uint32_t (&v)[4]) { return rn_mul(2.32830643653869629E-10f, (float)ao_random_int(v)); }

/* the literal walk, asking only for "any hit" */
__device__ __noinline__ bool walk_reference_any(const SceneDev &sc, f3 o, f3 d, float max_distance)
{
	const uint32_t n = __ldg(sc.ref_nodes);
	uint32_t tri = 0;
	for (uint32_t i = 0; i < n;) {
		const uint32_t node_count = __ldg(sc.ref_nodes + i);
		const float4 lo = __ldg(sc.ref_aabbs + 2 * (size_t)i), hi = __ldg(sc.ref_aabbs + 2 * (size_t)i + 1);
		if (!aabb_exact(make_f3(lo.x, lo.y, lo.z), make_f3(hi.x, hi.y, hi.z), o, d, max_distance)) {
			tri += (node_count + 1) >> 1;
			i += node_count;
		} else {
			if (node_count == 1) {
				const float4 *q = sc.tris + 4 * RTX_IDX(tri, sc.num_tris);
				TriHit h;
				if (triangle_test(__ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3), o, d, __int_as_float(0x7f800000), h)) return true;
				++tri;
			}
			++i;
		}
	}
	return false;
}

/* All occlusion rays of a pixel leave the same point p and a box is only entered with t_min < max_distance
 * (:60), i.e. within max_distance * |d| of p, |d| = 1 up to a few ulp (normalised bases, xs^2+ys^2+zs^2 = 1).
 * So only leaves whose box reaches into the cube p +- R, R = max_distance * 1.001 + 1e-5 * scale, can pass
 * (a box beyond the cube along axis k has t_k,min >= R / |d_k| > max_distance when d_k points towards it and
 * t_max < 0 when it points away; margins are 4 orders above the rounding of the slab products).  Walk down
 * from the root while exactly one child reaches into the cube: the rays of the pixel start at that pair
 * instead of the root.  -1: no leaf can pass at all. */
RTX_DEV int ao_entry_pair(const SceneDev &sc, f3 p, float max_distance)
{
	const float R = max_distance * 1.001f + 1e-5f * fmaxf(fmaxf(fabsf(p.x), fabsf(p.y)), fmaxf(fabsf(p.z), sc.scene_scale));
	if (!(R == R) || !(p.x == p.x) || !(p.y == p.y) || !(p.z == p.z)) return 0;
	const f3 lo = make_f3(p.x - R, p.y - R, p.z - R), hi = make_f3(p.x + R, p.y + R, p.z + R);
	int cur = 0;
	for (;;) {
		const float4 *q = sc.pairs + 4 * RTX_IDX(cur, sc.pair_count);
		const float4 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2), q3 = __ldg(q + 3);
		const bool inL = q0.x <= hi.x && q0.w >= lo.x && q0.y <= hi.y && q1.x >= lo.y && q0.z <= hi.z && q1.y >= lo.z;
		const bool inR = q2.x <= hi.x && q2.w >= lo.x && q2.y <= hi.y && q3.x >= lo.y && q2.z <= hi.z && q3.y >= lo.z;
		if (inL == inR) return inL ? cur : -1;
		const int ref = __float_as_int(inL ? q1.z : q3.z);
		if (ref < 0) return cur;                 /* the one child is a leaf: its parent pair is the entry */
		cur = ref;
	}
}

RTX_DEV bool any_hit(const SceneDev &sc, bool ordered_ok, int entry, f3 o, f3 d, float max_distance)
{
	const bool plain = ray_is_plain(o, d);
	if (!(ordered_ok && plain)) return walk_reference_any(sc, o, d, max_distance);
	if (entry < 0) return false;
	const f3 id = make_f3(rn_div(1.0f, d.x), rn_div(1.0f, d.y), rn_div(1.0f, d.z));
	int stack[RTX_STACK_MAX];
	int sp = 0, cur = entry;
	for (;;) {
		while (cur >= 0) {
			const float4 *q = sc.pairs + 4 * RTX_IDX(cur, sc.pair_count);
			const float4 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2), q3 = __ldg(q + 3);
			const Slab L = slab_interval<false, 4>(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, o, id);
			const Slab R = slab_interval<false, 4>(q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, o, id);
			const bool hitL = L.tmin <= L.tmax && L.tmin < max_distance && L.tmax > 0.0f;
			const bool hitR = R.tmin <= R.tmax && R.tmin < max_distance && R.tmax > 0.0f;
			if (!(hitL || hitR)) goto pop;
			const int refL = __float_as_int(q1.z), refR = __float_as_int(q3.z);
			const bool r_first = hitR && (!hitL || R.tmin < L.tmin);
			if (hitL && hitR) stack[RTX_IDX(sp++, RTX_STACK_MAX)] = r_first ? refL : refR;
			cur = r_first ? refR : refL;
		}
		{
			const uint32_t enc = ~(uint32_t)cur;
			const uint32_t first = enc >> 3, last = first + (enc & 7u);
			for (uint32_t tri = first; tri <= last; ++tri) {
				const float4 *q = sc.tris + 4 * RTX_IDX(tri, sc.num_tris);
				TriHit h;
				if (!triangle_test(__ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3), o, d, __int_as_float(0x7f800000), h)) continue;
				if (sc.verify_leafbox) {
					const float4 lo = __ldg(sc.leafbox + 2 * RTX_IDX(tri, sc.num_tris)), hi = __ldg(sc.leafbox + 2 * RTX_IDX(tri, sc.num_tris) + 1);
					if (!aabb_exact(make_f3(lo.x, lo.y, lo.z), make_f3(hi.x, hi.y, hi.z), o, d, max_distance)) continue;
				}
				return true;
			}
		}
pop:
		if (sp == 0) return false;
		cur = stack[--sp];
	}
}

/* The sample directions of the uniform method in the local basis (:236-246), once per frame: they depend on
 * (ring, ray) only.  One thread; the sequence is a few dozen to a few hundred entries. */
__global__ void k_ao_ring_table(AoParams ao, float4 *__restrict__ ring)
{
	if (blockIdx.x != 0 || threadIdx.x != 0) return;
	const double PI = 3.14159265358979323846, PI_2 = 1.57079632679489661923;
	const uint32_t circle_count = ao.samples;
	const float degrees = (float)(PI / 180);                                            /* :221 */
	const float alpha_min = rn_mul((float)ao.alpha_min, degrees), alpha_max = rn_mul((float)ao.alpha_max, degrees);
	uint32_t n = 0;
	for (uint32_t cc = 0; cc < circle_count; ++cc) {
		const float step = rn_div(alpha_max, (float)circle_count);                      /* :238 */
		const float angle = rn_add(rn_mul(step, (float)cc), alpha_min);                 /* :239 */
		const uint32_t ray_count = (uint32_t)__ddiv_rn(__dmul_rn(__dmul_rn(2.0, PI), (double)t_cos(angle)), (double)step); /* :240 */
		const float theta = (float)__dsub_rn(PI_2, (double)angle);                      /* :241 */
		for (uint32_t cr = 0; cr <= ray_count; ++cr) {                                  /* :242 (<=) */
			const float phi = (float)__ddiv_rn(__dmul_rn(__dmul_rn(2.0, PI), (double)cr), (double)ray_count); /* :243 */
			const float xs = rn_mul(t_sin(theta), t_cospi(phi));
			const float ys = t_cos(theta);
			const float zs = rn_mul(t_sin(theta), t_sinpi(phi));
			if (n < ao.ring_cap) ring[1 + n] = make_float4(xs, ys, zs, 0.0f);
			++n;
		}
	}
	ring[0] = make_float4(__uint_as_float(n < ao.ring_cap ? n : ao.ring_cap), 0.f, 0.f, 0.f);
}

/* :214-277 */
RTX_DEV float ambient_occlusion(const SceneDev &sc, bool ordered_ok, f3 point, f3 normal, uint32_t index, const AoParams &ao)
{
	const float k = rn_div(1.0f, 100000.0f);
	const f3 p = make_f3(rn_add(point.x, rn_mul(normal.x, k)), rn_add(point.y, rn_mul(normal.y, k)), rn_add(point.z, rn_mul(normal.z, k))); /* :215 */
	const int entry = ordered_ok ? ao_entry_pair(sc, p, ao.max_distance) : 0;
	uint32_t hits = 0;
	if (ao.method == 0) {                                                    /* :218-256 */
		const f3 basis_y = normal;
		const f3 h = perturb_smallest(basis_y);
		const f3 basis_x = normalize3(cross3(h, basis_y));
		const f3 basis_z = normalize3(cross3(basis_x, basis_y));
		/* (xs, ys, zs) of :244-246 depend on the ring and the ray only, not on the pixel: k_ao_ring_table */
		const uint32_t n = __float_as_uint(__ldg(ao.ring).x);
		for (uint32_t i = 0; i < n; ++i) {
			const float4 s = __ldg(ao.ring + 1 + i);
			const f3 dir = combine3(basis_x, s.x, basis_y, s.y, basis_z, s.z);   /* :248 */
			if (any_hit(sc, ordered_ok, entry, p, dir, ao.max_distance)) ++hits;
		}
		return rn_sub(1.0f, rn_div((float)hits, (float)n));                  /* :256 */
	}
	/* random hemisphere :257-275, sampler :153-183 */
	const f3 basis_y = normalize3(normal);
	const f3 h = perturb_smallest(basis_y);
	const f3 basis_x = normalize3(cross3(h, basis_y));
	const f3 basis_z = normalize3(cross3(basis_x, basis_y));
	uint32_t rng[4];
	const uint32_t seed = 536870923u * index;
	rng[0] = (123456789u ^ seed) * 88675123u;
	rng[1] = (362436069u ^ seed) * 123456789u;
	rng[2] = (521288629u ^ seed) * 362436069u;
	rng[3] = (88675123u ^ seed) * 521288629u;
	ao_random_int(rng);
	const uint32_t n = ao.samples + 1u;
	if (any_hit(sc, ordered_ok, entry, p, normal, ao.max_distance)) ++hits;
	for (uint32_t i = 0; i < n; ++i) {
		const float xi1 = ao_random_float(rng);
		const float xi2 = ao_random_float(rng);
		const float theta = t_acos(rn_sqrt(rn_sub(1.0f, xi1)));
		const float phi = (float)__dmul_rn(2.0, (double)xi2);
		const float xs = rn_mul(t_sin(theta), t_cospi(phi));
		const float ys = t_cos(theta);
		const float zs = rn_mul(t_sin(theta), t_sinpi(phi));
		const f3 dir = normalize3(combine3(basis_x, xs, basis_y, ys, basis_z, zs));
		if (any_hit(sc, ordered_ok, entry, p, dir, ao.max_distance)) ++hits;
	}
	return rn_sub(1.0f, rn_div((float)hits, (float)n));
}

/* One lane per pixel, 8x4-pixel blocks like the primary pass. */
__global__ void __launch_bounds__(128)
k_ambient_occlusion(const SceneDev sc, const Work w, const AoParams ao)
{
	const uint32_t unit = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	if (unit >= w.num_units) return;
	uint32_t x, y;
	size_t out;
	if (!unit_pixel(w, unit + w.tile_begin * 32u, threadIdx.x & 31u, x, y, out)) return;
	const uint32_t fid = w.face_id[out];
	if (fid == 0xffffffffu) return;
	const uint32_t tri = fid / 3u;
	const float2 st = w.hit_st[out];
	const f3 o = make_f3(0.0f, 0.0f, 2.0f);
	const f3 d = primary_dir(w.cam, x, y);
	/* the hit point as triangle_test computed it (:71-85) */
	const float4 *q = sc.tris + 4 * RTX_IDX(tri, sc.num_tris);
	const float4 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
	const f3 a = make_f3(q0.x, q0.y, q0.z), nrm = make_f3(q0.w, q1.w, q2.w);
	const f3 w0 = sub3(o, a);
	const float r = rn_div(-dot3(nrm, w0), dot3(nrm, d));
	const f3 point = make_f3(rn_add(o.x, rn_mul(r, d.x)), rn_add(o.y, rn_mul(r, d.y)), rn_add(o.z, rn_mul(r, d.z)));
	/* the smooth normal (:118-127) */
	const float4 n0 = __ldg(sc.tnormals + 3 * RTX_IDX(tri, sc.num_tris)), n1 = __ldg(sc.tnormals + 3 * RTX_IDX(tri, sc.num_tris) + 1), n2 = __ldg(sc.tnormals + 3 * RTX_IDX(tri, sc.num_tris) + 2);
	const float b0 = rn_sub(rn_sub(1.0f, st.x), st.y), b1 = st.x, b2 = st.y;
	const f3 normal = normalize3(make_f3(rn_add(rn_add(rn_mul(n0.x, b0), rn_mul(n1.x, b1)), rn_mul(n2.x, b2)),
	                                     rn_add(rn_add(rn_mul(n0.y, b0), rn_mul(n1.y, b1)), rn_mul(n2.y, b2)),
	                                     rn_add(rn_add(rn_mul(n0.z, b0), rn_mul(n1.z, b1)), rn_mul(n2.z, b2))));
	const float occlusion = ambient_occlusion(sc, w.ordered_ok != 0, point, normal, y * w.cam.W + x, ao);
	w.image[out] = rn_mul(w.image[out], occlusion);                         /* :306 */
}

/* ux[x] = (x + 0.5)/a - W/(2a), vy[y] = -((y + 0.5)/a - H/(2a)): intersect_kernel.cl:287-288, once per column / row */
__global__ void k_ray_tables(Camera cam, float *__restrict__ ux, float *__restrict__ vy)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < cam.W) ux[i] = rn_sub(rn_div(rn_add((float)i, 0.5f), cam.a), cam.w_over_2a);
	if (i < cam.H) vy[i] = -rn_sub(rn_div(rn_add((float)i, 0.5f), cam.a), cam.h_over_2a);
}

/* ------------------------------ image ops -------------------------------- */

/* RayTracer::resize (src/ray_tracer.cc:3-15): n x n box sum in (ssY, ssX)
 * order, (total / (n*n)) * 255, float -> unsigned char truncation. */
__global__ void k_resize_u8(const float *__restrict__ img, uint32_t total_width, uint32_t width, uint32_t height,
                            uint32_t n, unsigned char *__restrict__ out)
{
	const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
	if (x >= width || y >= height) return;
	float total = 0.0f;
	for (uint32_t sy = 0; sy < n; ++sy)
		for (uint32_t sx = 0; sx < n; ++sx)
			total = rn_add(total, img[(size_t)(y * n + sy) * total_width + (x * n + sx)]);
	const float v = rn_mul(rn_div(total, (float)(n * n)), 255.0f);
	out[(size_t)y * width + x] = (unsigned char)(int)v;
}

/* The same with the n samples of a row in one vector load (n = 2: 64-bit, n = 4: 128-bit; total_width = width * n
 * keeps every row segment aligned).  Same additions in the same order. */
template <int N>
__global__ void k_resize_u8_vec(const float *__restrict__ img, uint32_t total_width, uint32_t width, uint32_t height,
                                unsigned char *__restrict__ out)
{
	const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
	if (x >= width || y >= height) return;
	float total = 0.0f;
#pragma unroll
	for (uint32_t sy = 0; sy < N; ++sy) {
		const float *row = img + (size_t)(y * N + sy) * total_width + (size_t)x * N;
		if (N == 4) {
			const float4 v = __ldcs(reinterpret_cast<const float4 *>(row));      /* streamed: read once */
			total = rn_add(rn_add(rn_add(rn_add(total, v.x), v.y), v.z), v.w);
		} else {
			const float2 v = __ldcs(reinterpret_cast<const float2 *>(row));
			total = rn_add(rn_add(total, v.x), v.y);
		}
	}
	const float v = rn_mul(rn_div(total, (float)(N * N)), 255.0f);
	out[(size_t)y * width + x] = (unsigned char)(int)v;
}

/* RayTracer::resize on a rank's compact tiles: [ltile][32][32] floats -> [ltile][32/n][32/n] bytes (n divides 32,
 * so no output pixel straddles tiles or ranks).  Same summation order as k_resize_u8. */
__global__ void k_resize_tiles_u8(const float *__restrict__ tiles, uint32_t ntiles, uint32_t n, unsigned char *__restrict__ out)
{
	const uint32_t m = RTX_TILE / n;                 /* output pixels per tile side */
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= (size_t)ntiles * m * m) return;
	const uint32_t t = (uint32_t)(i / (m * m)), r = (uint32_t)(i % (m * m)), oy = r / m, ox = r % m;
	const float *src = tiles + (size_t)t * (RTX_TILE * RTX_TILE);
	float total = 0.0f;
	if (n == 4 && (reinterpret_cast<uintptr_t>(tiles) & 15u) == 0) {         /* one 128-bit load per sample row */
#pragma unroll
		for (uint32_t sy = 0; sy < 4; ++sy) {
			const float4 v = __ldcs(reinterpret_cast<const float4 *>(src + (oy * 4 + sy) * RTX_TILE + ox * 4));
			total = rn_add(rn_add(rn_add(rn_add(total, v.x), v.y), v.z), v.w);
		}
	} else {
		for (uint32_t sy = 0; sy < n; ++sy)
			for (uint32_t sx = 0; sx < n; ++sx)
				total = rn_add(total, src[(oy * n + sy) * RTX_TILE + (ox * n + sx)]);
	}
	out[i] = (unsigned char)(int)rn_mul(rn_div(total, (float)(n * n)), 255.0f);
}

/* RayTracer::resize of a rank's compact tiles written STRAIGHT into a row-major width x height byte image -- the
 * rank's own, or rank 0's mapped through NVLink peer memory (rtx_peer_open): resize + gather + de-interleave of the
 * byte path in one kernel, no collective.  One thread per output pixel, or per 4 adjacent ones when they can leave as
 * one aligned 32-bit store (peer stores of single bytes cost a 32-byte sector each).  Same summation order as
 * k_resize_u8 (ray_tracer.cc:3-15). */
RTX_DEV unsigned char resize_one(const float *__restrict__ src, uint32_t oy, uint32_t ox, uint32_t n)
{
	float total = 0.0f;
	if (n == 4) {
#pragma unroll
		for (uint32_t sy = 0; sy < 4; ++sy) {
			const float4 v = __ldcs(reinterpret_cast<const float4 *>(src + (oy * 4 + sy) * RTX_TILE + ox * 4));
			total = rn_add(rn_add(rn_add(rn_add(total, v.x), v.y), v.z), v.w);
		}
	} else {
		for (uint32_t sy = 0; sy < n; ++sy)
			for (uint32_t sx = 0; sx < n; ++sx)
				total = rn_add(total, src[(oy * n + sy) * RTX_TILE + (ox * n + sx)]);
	}
	return (unsigned char)(int)rn_mul(rn_div(total, (float)(n * n)), 255.0f);
}

__global__ void k_resize_tiles_u8_to(const float *__restrict__ tiles, uint32_t local_tiles, uint32_t n, uint32_t rank, uint32_t world,
                                     uint32_t tiles_x, uint32_t width, uint32_t height, unsigned char *__restrict__ image)
{
	const uint32_t m = RTX_TILE / n;                 /* output pixels per tile side */
	const bool quads = (m & 3u) == 0 && (width & 3u) == 0 && (reinterpret_cast<uintptr_t>(image) & 3u) == 0 &&
	                   (reinterpret_cast<uintptr_t>(tiles) & 15u) == 0;
	const uint32_t per_row = quads ? m / 4 : m, per_tile = per_row * m;
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= (size_t)local_tiles * per_tile) return;
	const uint32_t lt = (uint32_t)(i / per_tile), r = (uint32_t)(i % per_tile), oy = r / per_row, oq = r % per_row;
	const uint32_t tile = lt * world + rank, tx = tile % tiles_x, ty = tile / tiles_x;
	const float *src = tiles + (size_t)lt * (RTX_TILE * RTX_TILE);
	const uint32_t y = ty * m + oy;
	if (y >= height) return;
	if (quads) {
		const uint32_t x = tx * m + oq * 4;
		if (x >= width) return;                      /* width % 4 == 0: a quad is inside or outside as a whole */
		uchar4 v;
		v.x = resize_one(src, oy, oq * 4, n); v.y = resize_one(src, oy, oq * 4 + 1, n);
		v.z = resize_one(src, oy, oq * 4 + 2, n); v.w = resize_one(src, oy, oq * 4 + 3, n);
		*reinterpret_cast<uchar4 *>(image + (size_t)y * width + x) = v;
	} else {
		const uint32_t x = tx * m + oq;
		if (x < width) image[(size_t)y * width + x] = resize_one(src, oy, oq, n);
	}
}

/* A rank's compact float tiles written straight into a row-major W x H float image that the kernel can address: the
 * rank's own, a peer's (rtx_peer_open: the float gather without a collective), or page-locked HOST memory mapped into
 * the device (rtx_host_register: every rank's tiles leave over its own PCIe link).  One CTA per tile, 128-bit stores
 * (a tile row = 128 contiguous bytes). */
__global__ void __launch_bounds__(256)
k_store_tiles(const float *__restrict__ tiles, uint32_t tile_begin, uint32_t local_tiles, uint32_t rank, uint32_t world, uint32_t tiles_x,
              uint32_t W, uint32_t H, float *__restrict__ image)
{
	const uint32_t lt = tile_begin + blockIdx.x;           /* local tiles [tile_begin, local_tiles) */
	if (lt >= local_tiles) return;
	const uint32_t tile = lt * world + rank, tx = tile % tiles_x, ty = tile / tiles_x;
	const float *src = tiles + (size_t)lt * (RTX_TILE * RTX_TILE);
	const uint32_t x0 = tx * RTX_TILE, y0 = ty * RTX_TILE;
	const bool vec = (W & 3u) == 0 && (reinterpret_cast<uintptr_t>(image) & 15u) == 0 && (reinterpret_cast<uintptr_t>(tiles) & 15u) == 0;
	if (vec) {
		const uint32_t py = threadIdx.x >> 3, px = (threadIdx.x & 7u) << 2;
		if (y0 + py < H && x0 + px < W)              /* W % 4 == 0: four pixels are inside or outside together */
			__stcs(reinterpret_cast<float4 *>(image + (size_t)(y0 + py) * W + x0 + px),
			       __ldcs(reinterpret_cast<const float4 *>(src + py * RTX_TILE + px)));
	} else {
		for (uint32_t i = threadIdx.x; i < RTX_TILE * RTX_TILE; i += blockDim.x) {
			const uint32_t px = i & 31u, py = i >> 5;
			if (x0 + px < W && y0 + py < H) image[(size_t)(y0 + py) * W + x0 + px] = src[i];
		}
	}
}

/* --------------------------------------------------------------------------
 * Ordering the ranks of the peer-memory paths without a collective: a monotonic frame counter per rank in rank 0's
 * memory.  A rank signals after its store kernel (k_peer_signal, next in its stream: the stores of the finished kernel
 * have been performed, the fence orders the flag behind them); rank 0 waits for every counter before it reads the frame
 * (k_peer_wait); a rank waits for rank 0's "consumed" counter before it overwrites a buffer.  Every wait gives up after
 * max_cycles and records that in *timed_out instead of hanging the GPU.  Waiter and signaller run on DIFFERENT GPUs.
 * ------------------------------------------------------------------------ */
__global__ void k_peer_signal(unsigned int *flag, unsigned int value)
{
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	__threadfence_system();
	asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(flag), "r"(value) : "memory");
}

__global__ void k_peer_wait(const unsigned int *flags, unsigned int count, unsigned int value, unsigned int *timed_out, long long max_cycles)
{
	const unsigned int t = threadIdx.x;
	if (blockIdx.x != 0 || t >= count) return;
	const long long t0 = clock64();
	for (;;) {
		unsigned int v;
		asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + t) : "memory");
		if ((int)(v - value) >= 0) break;
		if (clock64() - t0 > max_cycles) { atomicExch(timed_out, 1u); break; }
		__nanosleep(200);
	}
	__threadfence_system();
}

/* rank-major gathered compact u8 tiles -> row-major width x height byte image (rank 0).  One thread per output row of
 * a tile (m = 32 / n bytes; one 64-bit load + store when m == 8 and everything is aligned). */
__global__ void k_deinterleave_u8(const unsigned char *__restrict__ gathered, uint32_t world, uint32_t tiles_per_rank,
                                  uint32_t tiles_x, uint32_t tiles_y, uint32_t n, uint32_t width, uint32_t height,
                                  unsigned char *__restrict__ image)
{
	const uint32_t m = RTX_TILE / n;
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= (size_t)tiles_x * tiles_y * m) return;
	const uint32_t tile = (uint32_t)(i / m), oy = (uint32_t)(i % m);
	const uint32_t rank = tile % world, ltile = tile / world;
	const unsigned char *src = gathered + ((size_t)rank * tiles_per_rank + ltile) * (m * m) + oy * m;
	const uint32_t tx = tile % tiles_x, ty = tile / tiles_x;
	const uint32_t x0 = tx * m, y = ty * m + oy;
	if (y >= height) return;
	unsigned char *dst = image + (size_t)y * width + x0;
	if (m == 8 && x0 + 8 <= width && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 7u) == 0) {
		*reinterpret_cast<uint2 *>(dst) = __ldcs(reinterpret_cast<const uint2 *>(src));
	} else {
		for (uint32_t k = 0; k < m; ++k)
			if (x0 + k < width) dst[k] = src[k];
	}
}

/* rank-major gathered compact tile buffers -> row-major image (rank 0) */
__global__ void k_deinterleave(const float *__restrict__ gathered, uint32_t world, uint32_t tiles_per_rank,
                               uint32_t tiles_x, uint32_t tiles_y, uint32_t W, uint32_t H, float *__restrict__ image)
{
	const uint32_t tile = blockIdx.x;
	if (tile >= tiles_x * tiles_y) return;
	const uint32_t rank = tile % world, ltile = tile / world;
	const float *src = gathered + ((size_t)rank * tiles_per_rank + ltile) * (RTX_TILE * RTX_TILE);
	const uint32_t tx = tile % tiles_x, ty = tile / tiles_x;
	for (uint32_t i = threadIdx.x; i < RTX_TILE * RTX_TILE; i += blockDim.x) {
		const uint32_t px = i & 31u, py = i >> 5;
		const uint32_t x = tx * RTX_TILE + px, y = ty * RTX_TILE + py;
		if (x < W && y < H) image[(size_t)y * W + x] = src[i];
	}
}

/* ------------------- device-side pieces of the flatten ------------------- */

/* Leaf slots of the node pairs get the slack of their triangles (tri_slack_rel), in all four octant copies.  A leaf
 * that needs more than the default raises every ancestor's slot to its slack: with the parent links of the device
 * flatten (parent[q] = (pair << 1) | slot that refers to q) right here, by atomicMax on the bits (slacks are >= 0, so
 * their bit patterns order like the values); without them *fat = 1 tells the host to run k_slack_relax. */
__global__ void k_slack_leaves(float4 *__restrict__ pairs, const float4 *__restrict__ tris, uint32_t num_pairs, uint32_t stride,
                               const uint32_t *__restrict__ parent, unsigned int *fat)
{
	const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= num_pairs) return;
	for (int k = 0; k < 2; ++k) {
		const float4 a = pairs[4 * (size_t)p + 2 * k], b = pairs[4 * (size_t)p + 2 * k + 1];
		const int ref = __float_as_int(b.z);
		if (ref >= 0) continue;
		const uint32_t enc = ~(uint32_t)ref, first = enc >> 3, count = (enc & 7u) + 1u;
		float rel = RTX_SLACK_REL;
		for (uint32_t t = first; t < first + count; ++t) rel = fmaxf(rel, tri_slack_rel(tris[4 * (size_t)t + 3]));
		if (rel == RTX_SLACK_REL) continue;
		const float slack = box_slack_rel(a.x, a.y, a.z, a.w, b.x, b.y, rel);      /* copy 0 is unswapped: (lo.xyz, hi.x), (hi.y, hi.z) */
		for (int v = 0; v < 4; ++v) pairs[((size_t)v * stride + p) * 4 + 2 * k + 1].w = slack;
		if (!parent) { *fat = 1u; continue; }
		const unsigned int bits = __float_as_uint(slack);
		for (uint32_t q = p; q != 0;) {
			const uint32_t e = parent[q], pp = e >> 1, slot = e & 1u;
			unsigned int *w0 = reinterpret_cast<unsigned int *>(&pairs[(size_t)pp * 4 + 2 * slot + 1].w);
			if (*reinterpret_cast<volatile unsigned int *>(w0) >= bits) break;      /* whoever raised it carries on upwards */
			for (int v = 0; v < 4; ++v)
				atomicMax(reinterpret_cast<unsigned int *>(&pairs[((size_t)v * stride + pp) * 4 + 2 * slot + 1].w), bits);
			q = pp;
		}
	}
}

/* One bottom-up relaxation step: the slack of a child slot is at least the slack of the child pair's own slots.
 * Run `depth` times (only when k_slack_leaves found a fat leaf). */
__global__ void k_slack_relax(float4 *__restrict__ pairs, uint32_t num_pairs, uint32_t stride)
{
	const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= num_pairs) return;
	for (int k = 0; k < 2; ++k) {
		const float4 b = pairs[4 * (size_t)p + 2 * k + 1];
		const int ref = __float_as_int(b.z);
		if (ref < 0) continue;
		const float child = fmaxf(pairs[4 * (size_t)ref + 1].w, pairs[4 * (size_t)ref + 3].w);
		if (child > b.w)
			for (int v = 0; v < 4; ++v) pairs[((size_t)v * stride + p) * 4 + 2 * k + 1].w = child;
	}
}

/* --------------------------------------------------------------------------
 * Upload, device side: the invariants of the reference's pre-order tree (bvh.cc:98-162) and the two
 * prefix counts the flatten kernel needs, computed from the raw `nodes` array on the device -- no host pass
 * over the tree (20 M nodes at C4).
 *   first_leaf[i]  leaves before node i in pre-order  = exclusive sum of [nodes[j] == 1]
 *   pair_idx[i]    flattened internal nodes before i  = exclusive sum of [node j becomes a pair]
 *   depth          1 + most "pair" ancestors of any pair: +1 at j+1 and -1 at j+nodes[j] for every pair j,
 *                  summed in pre-order (k_tree_check scatters the +-1, the scan integrates them)
 * Three kernels: per-block partial sums, one block over the partials, per-block scan + write.
 * ------------------------------------------------------------------------ */
struct TreeResult {
	unsigned int bad_node;      /* smallest index that breaks an invariant, or 0xffffffff */
	unsigned int bad_face;      /* 1 if a face index is out of range */
	unsigned int num_pairs;
	unsigned int depth;
	unsigned int loose;         /* 1 if some child box is not inside its parent's box (k_tree_check) */
};

#define RTX_SCAN_ITEMS 16       /* nodes per thread */
#define RTX_SCAN_BLOCK 256

RTX_DEV bool tree_is_pair(uint32_t size, uint32_t i, uint32_t leaf_size) { return size > 1 && (((size + 1) >> 1) > leaf_size || i == 0); }

/* Also checks what the re-ordered traversals rely on beyond the topology (point 1): every child box
 * lies inside its parent's box, so that a leaf box passing the slab test implies all its ancestors pass.  bvh.cc
 * builds such trees; a caller-supplied tree that does not is rendered with the literal walk (res->loose). */
__global__ void k_tree_check(const uint32_t *__restrict__ nodes, const float4 *__restrict__ aabbs, uint32_t n, uint32_t leaf_size,
                             int *__restrict__ delta, TreeResult *res)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t size = nodes[i];
	if (size == 1) return;
	bool ok = (size & 1u) != 0 && size >= 3 && (uint64_t)i + size <= n;
	uint32_t l = 0;
	if (ok) {
		l = nodes[i + 1];
		ok = (l & 1u) != 0 && l + 2 <= size && nodes[i + 1 + l] == size - 1 - l;
	}
	if (!ok) { atomicMin(&res->bad_node, i); return; }
	{
		const float4 plo = aabbs[2 * (size_t)i], phi = aabbs[2 * (size_t)i + 1];
		bool nested = true;
		for (int k = 0; k < 2; ++k) {
			const size_t c = k == 0 ? (size_t)i + 1 : (size_t)i + 1 + l;
			const float4 lo = aabbs[2 * c], hi = aabbs[2 * c + 1];
			nested = nested && lo.x >= plo.x && lo.y >= plo.y && lo.z >= plo.z && hi.x <= phi.x && hi.y <= phi.y && hi.z <= phi.z;   /* false for NaN */
		}
		if (!nested) res->loose = 1u;
	}
	if (tree_is_pair(size, i, leaf_size)) {
		atomicAdd(delta + i + 1, 1);
		if (i + size < n) atomicAdd(delta + i + size, -1);
	}
}

/* block sums of the three scanned quantities -> partials[3][nblocks] */
__global__ void __launch_bounds__(RTX_SCAN_BLOCK)
k_tree_partials(const uint32_t *__restrict__ nodes, const int *__restrict__ delta, uint32_t n, uint32_t leaf_size, int *__restrict__ partials)
{
	__shared__ int s[3][RTX_SCAN_BLOCK / 32];
	const uint32_t base = (blockIdx.x * RTX_SCAN_BLOCK + threadIdx.x) * RTX_SCAN_ITEMS;
	int a = 0, b = 0, c = 0;
	for (uint32_t k = 0; k < RTX_SCAN_ITEMS; ++k) {
		const uint32_t i = base + k;
		if (i < n) {
			const uint32_t size = nodes[i];
			a += size == 1;
			b += tree_is_pair(size, i, leaf_size);
			c += delta[i];
		}
	}
	for (int o = 16; o > 0; o >>= 1) {
		a += __shfl_xor_sync(0xffffffffu, a, o);
		b += __shfl_xor_sync(0xffffffffu, b, o);
		c += __shfl_xor_sync(0xffffffffu, c, o);
	}
	if ((threadIdx.x & 31) == 0) { s[0][threadIdx.x >> 5] = a; s[1][threadIdx.x >> 5] = b; s[2][threadIdx.x >> 5] = c; }
	__syncthreads();
	if (threadIdx.x < 3) {
		int t = 0;
		for (int w = 0; w < RTX_SCAN_BLOCK / 32; ++w) t += s[threadIdx.x][w];
		partials[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = t;
	}
}

/* exclusive scan of each row of partials, in place; one block, one warp per row */
__global__ void k_tree_spine(int *__restrict__ partials, uint32_t nblocks, TreeResult *res)
{
	const uint32_t row = threadIdx.x >> 5, lane = threadIdx.x & 31u;
	if (row >= 3) return;
	int *p = partials + (size_t)row * nblocks;
	int carry = 0;
	for (uint32_t base = 0; base < nblocks; base += 32) {
		const uint32_t i = base + lane;
		const int v = i < nblocks ? p[i] : 0;
		int x = v;
		for (int o = 1; o < 32; o <<= 1) {
			const int y = __shfl_up_sync(0xffffffffu, x, o);
			if ((int)lane >= o) x += y;
		}
		if (i < nblocks) p[i] = carry + x - v;
		carry += __shfl_sync(0xffffffffu, x, 31);
	}
	if (row == 1 && lane == 0) res->num_pairs = (unsigned int)carry;
}

__global__ void __launch_bounds__(RTX_SCAN_BLOCK)
k_tree_scan(const uint32_t *__restrict__ nodes, const int *__restrict__ delta, uint32_t n, uint32_t leaf_size,
            const int *__restrict__ partials, uint32_t *__restrict__ first_leaf, uint32_t *__restrict__ pair_idx, TreeResult *res)
{
	__shared__ int s[3][RTX_SCAN_BLOCK / 32];
	const uint32_t base = (blockIdx.x * RTX_SCAN_BLOCK + threadIdx.x) * RTX_SCAN_ITEMS;
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	int a = 0, b = 0, c = 0;
	for (uint32_t k = 0; k < RTX_SCAN_ITEMS; ++k) {
		const uint32_t i = base + k;
		if (i < n) {
			const uint32_t size = nodes[i];
			a += size == 1;
			b += tree_is_pair(size, i, leaf_size);
			c += delta[i];
		}
	}
	/* exclusive scan of the thread sums over the block */
	int xa = a, xb = b, xc = c;
	for (int o = 1; o < 32; o <<= 1) {
		const int ya = __shfl_up_sync(0xffffffffu, xa, o), yb = __shfl_up_sync(0xffffffffu, xb, o), yc = __shfl_up_sync(0xffffffffu, xc, o);
		if ((int)lane >= o) { xa += ya; xb += yb; xc += yc; }
	}
	if (lane == 31) { s[0][warp] = xa; s[1][warp] = xb; s[2][warp] = xc; }
	__syncthreads();
	int oa = partials[blockIdx.x], ob = partials[(size_t)gridDim.x + blockIdx.x], oc = partials[2 * (size_t)gridDim.x + blockIdx.x];
	for (uint32_t w = 0; w < warp; ++w) { oa += s[0][w]; ob += s[1][w]; oc += s[2][w]; }
	oa += xa - a; ob += xb - b; oc += xc - c;
	int deepest = 0;
	for (uint32_t k = 0; k < RTX_SCAN_ITEMS; ++k) {
		const uint32_t i = base + k;
		if (i < n) {
			const uint32_t size = nodes[i];
			first_leaf[i] = (uint32_t)oa;
			pair_idx[i] = (uint32_t)ob;
			oc += delta[i];                       /* pair ancestors of node i */
			const bool pair = tree_is_pair(size, i, leaf_size);
			if (pair && oc + 1 > deepest) deepest = oc + 1;
			oa += size == 1;
			ob += pair;
		}
	}
	for (int o = 16; o > 0; o >>= 1) deepest = max(deepest, __shfl_xor_sync(0xffffffffu, deepest, o));
	if (lane == 0 && deepest > 0) atomicMax(&res->depth, (unsigned int)deepest);
}

/* The whole flatten in one pass over the reference's pre-order nodes (rtx_upload's default path).
 * A node is an internal node of the flattened tree iff its subtree holds more than `leaf_size` triangles
 * (or it is the root); in pre-order its pair index is the number of such nodes before it, and the first
 * leaf of its subtree is the number of leaves before it -- two prefix counts the host computes while it
 * validates the arrays.  Internal nodes write their pair (both children) into the four octant copies;
 * leaves write their triangle record, leaf box and corner normals. */
__global__ void k_flatten_nodes(const uint32_t *__restrict__ nodes, const float4 *__restrict__ ref_aabbs,
                                const uint32_t *__restrict__ first_leaf, const uint32_t *__restrict__ pair_idx,
                                const uint32_t *__restrict__ faces, const float4 *__restrict__ verts,
                                const float4 *__restrict__ vnormals, uint32_t nnodes, uint32_t num_pairs, uint32_t leaf_size,
                                float4 *__restrict__ pairs, float4 *__restrict__ tris, float4 *__restrict__ leafbox,
                                float4 *__restrict__ tnormals, uint32_t nverts, TreeResult *res, uint32_t *__restrict__ parent,
                                float4 *__restrict__ pleafbox)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nnodes) return;
	if (res && res->bad_node != 0xffffffffu) return;          /* malformed tree (k_tree_check): nothing below is safe */
	const uint32_t size = nodes[i];
	if (size == 1) {
		const uint32_t t = first_leaf[i];
		uint32_t i0 = faces[3 * (size_t)t], i1 = faces[3 * (size_t)t + 1], i2 = faces[3 * (size_t)t + 2];
		if (i0 >= nverts || i1 >= nverts || i2 >= nverts) {    /* reported by rtx_upload; keep the reads in range */
			if (res) res->bad_face = 1u;
			i0 = i1 = i2 = 0;
		}
		const float4 A = verts[i0], B = verts[i1], C = verts[i2];
		const f3 a = make_f3(A.x, A.y, A.z);
		const f3 u = sub3(make_f3(B.x, B.y, B.z), a);                 /* :68 */
		const f3 v = sub3(make_f3(C.x, C.y, C.z), a);                 /* :69 */
		const f3 n = cross3(u, v);                                    /* :70 */
		const float uu = dot3(u, u), uv = dot3(u, v), vv = dot3(v, v);/* :87-89 */
		const float D = rn_sub(rn_mul(uv, uv), rn_mul(uu, vv));       /* :93 */
		float4 *q = tris + 4 * (size_t)t;
		q[0] = make_float4(a.x, a.y, a.z, n.x);
		q[1] = make_float4(u.x, u.y, u.z, n.y);
		q[2] = make_float4(v.x, v.y, v.z, n.z);
		q[3] = make_float4(uu, uv, vv, D);
		float4 lo = ref_aabbs[2 * (size_t)i], hi = ref_aabbs[2 * (size_t)i + 1];
		lo.w = 0.f; hi.w = 0.f;
		leafbox[2 * (size_t)t] = lo;
		leafbox[2 * (size_t)t + 1] = hi;
		if (pleafbox) {
			pleafbox[2 * (size_t)t] = make_float4(lo.x, lo.y, rn_sub(lo.z, 2.0f), triangle_primary_A(a, n));
			pleafbox[2 * (size_t)t + 1] = make_float4(hi.x, hi.y, rn_sub(hi.z, 2.0f), 0.f);
		}
		float4 n0 = vnormals[i0], n1 = vnormals[i1], n2 = vnormals[i2];
		n0.w = n1.w = n2.w = 0.f;
		tnormals[3 * (size_t)t] = n0;
		tnormals[3 * (size_t)t + 1] = n1;
		tnormals[3 * (size_t)t + 2] = n2;
	}
	if (nnodes == 1) {   /* a single triangle: pair 0 = {the leaf, an unreachable far box} */
		const float4 lo = ref_aabbs[0], hi = ref_aabbs[1];
		const int ref = (int)~0u;                                     /* leaf (first 0, count 1) */
		for (int v = 0; v < 4; ++v) {
			float4 *p = pairs + 4 * (size_t)v;
			p[0] = make_float4((v & 1) ? hi.x : lo.x, (v & 2) ? hi.y : lo.y, lo.z, (v & 1) ? lo.x : hi.x);
			p[1] = make_float4((v & 2) ? lo.y : hi.y, hi.z, __int_as_float(ref), box_slack(lo.x, lo.y, lo.z, hi.x, hi.y, hi.z));
			p[2] = make_float4(3e38f, 3e38f, 3e38f, 3e38f);
			p[3] = make_float4(3e38f, 3e38f, __int_as_float(ref), 0.f);
		}
		return;
	}
	const uint32_t lv = (size + 1) >> 1;
	if (size == 1 || !(lv > leaf_size || i == 0)) return;
	const uint32_t p = pair_idx[i];
	const uint32_t child[2] = { i + 1, i + 1 + nodes[i + 1] };
	float4 a[2], b[2];
#pragma unroll
	for (int k = 0; k < 2; ++k) {
		const uint32_t c = child[k];
		const uint32_t clv = (nodes[c] + 1) >> 1;
		const int ref = clv > leaf_size ? (int)pair_idx[c] : (int)~((first_leaf[c] << 3) | (clv - 1));
		if (ref >= 0 && parent) parent[ref] = (p << 1) | (uint32_t)k;           /* for k_slack_leaves */
		const float4 lo = ref_aabbs[2 * (size_t)c], hi = ref_aabbs[2 * (size_t)c + 1];
		a[k] = make_float4(lo.x, lo.y, lo.z, hi.x);
		b[k] = make_float4(hi.y, hi.z, __int_as_float(ref), box_slack(lo.x, lo.y, lo.z, hi.x, hi.y, hi.z));
	}
#pragma unroll
	for (int v = 0; v < 4; ++v) {       /* octant copies: swap lo/hi of x (bit 0) and of y (bit 1) */
		float4 *dst = pairs + ((size_t)v * num_pairs + p) * 4;
#pragma unroll
		for (int k = 0; k < 2; ++k) {
			float4 x = a[k], y = b[k];
			if (v & 1) { const float t = x.x; x.x = x.w; x.w = t; }
			if (v & 2) { const float t = x.y; x.y = y.x; y.x = t; }
			dst[2 * k] = x;
			dst[2 * k + 1] = y;
		}
	}
}


/* triangle records, leaf boxes and corner normals from the upload arrays */
__global__ void k_build_triangles(const uint32_t *__restrict__ faces, const float4 *__restrict__ verts,
                                  const float4 *__restrict__ vnormals, const uint32_t *__restrict__ leaf_node,
                                  const float4 *__restrict__ ref_aabbs, uint32_t ntris,
                                  float4 *__restrict__ tris, float4 *__restrict__ leafbox, float4 *__restrict__ tnormals,
                                  float4 *__restrict__ pleafbox)
{
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= ntris) return;
	const uint32_t i0 = faces[3 * (size_t)t], i1 = faces[3 * (size_t)t + 1], i2 = faces[3 * (size_t)t + 2];
	const float4 A = verts[i0], B = verts[i1], C = verts[i2];
	const f3 a = make_f3(A.x, A.y, A.z);
	const f3 u = sub3(make_f3(B.x, B.y, B.z), a);                 /* :68 */
	const f3 v = sub3(make_f3(C.x, C.y, C.z), a);                 /* :69 */
	const f3 n = cross3(u, v);                                    /* :70 */
	const float uu = dot3(u, u), uv = dot3(u, v), vv = dot3(v, v);/* :87-89 */
	const float D = rn_sub(rn_mul(uv, uv), rn_mul(uu, vv));       /* :93 */
	float4 *q = tris + 4 * (size_t)t;
	q[0] = make_float4(a.x, a.y, a.z, n.x);
	q[1] = make_float4(u.x, u.y, u.z, n.y);
	q[2] = make_float4(v.x, v.y, v.z, n.z);
	q[3] = make_float4(uu, uv, vv, D);
	const uint32_t node = leaf_node[t];
	float4 lo = ref_aabbs[2 * (size_t)node], hi = ref_aabbs[2 * (size_t)node + 1];
	lo.w = 0.f; hi.w = 0.f;
	leafbox[2 * (size_t)t] = lo;
	leafbox[2 * (size_t)t + 1] = hi;
	if (pleafbox) {
		pleafbox[2 * (size_t)t] = make_float4(lo.x, lo.y, rn_sub(lo.z, 2.0f), triangle_primary_A(a, n));
		pleafbox[2 * (size_t)t + 1] = make_float4(hi.x, hi.y, rn_sub(hi.z, 2.0f), 0.f);
	}
	float4 n0 = vnormals[i0], n1 = vnormals[i1], n2 = vnormals[i2];
	n0.w = n1.w = n2.w = 0.f;
	tnormals[3 * (size_t)t] = n0;
	tnormals[3 * (size_t)t + 1] = n1;
	tnormals[3 * (size_t)t + 2] = n2;
}

/* ------------------------- cache bandwidth probe ------------------------- */

/* Streams a buffer with 128-bit loads; the roofline denominators for scenes
 * that live in L2/L1 come from here (MEASURED_PEAKS.json only has HBM).
 * per_cta_vecs == 0: every CTA sweeps the whole buffer, L1 bypassed (ld.cg):
 * L2 bandwidth when the buffer fits L2, HBM when it does not.
 * per_cta_vecs  > 0: each CTA re-reads its own slice through L1 (ld.ca). */
__global__ void __launch_bounds__(512)
k_probe_bw(const float4 *__restrict__ buf, size_t n_vecs, int iters, size_t per_cta_vecs, float *__restrict__ sink)
{
	float acc = 0.f;
	if (per_cta_vecs == 0) {
		const size_t stride = (size_t)gridDim.x * blockDim.x;
		for (int it = 0; it < iters; ++it) {
			size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
			for (; i + 3 * stride < n_vecs; i += 4 * stride) {
				const float4 a = __ldcg(buf + i), b = __ldcg(buf + i + stride), c = __ldcg(buf + i + 2 * stride), d = __ldcg(buf + i + 3 * stride);
				acc += a.x + b.y + c.z + d.w;
			}
			for (; i < n_vecs; i += stride) acc += __ldcg(buf + i).x;
		}
	} else {
		const float4 *mine = buf + ((size_t)blockIdx.x * per_cta_vecs) % (n_vecs - per_cta_vecs + 1);
		for (int it = 0; it < iters; ++it) {
			size_t i = threadIdx.x;
			for (; i + 3 * blockDim.x < per_cta_vecs; i += 4 * blockDim.x) {
				const float4 a = __ldca(mine + i), b = __ldca(mine + i + blockDim.x), c = __ldca(mine + i + 2 * blockDim.x), d = __ldca(mine + i + 3 * blockDim.x);
				acc += a.x + b.y + c.z + d.w;
			}
			for (; i < per_cta_vecs; i += blockDim.x) acc += __ldca(mine + i).x;
		}
	}
	if (acc == 12345.678f) *sink = acc;
}
