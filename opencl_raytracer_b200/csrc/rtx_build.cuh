/*
 * rtx_build.cuh -- the reference's longest-axis BVH builder (src/bvh.cc:59-162, aabb.cc, triangle.cc) on the
 * device, emitting the SAME arrays bvh.cc does (pre-order `nodes`, (min,max) `aabbs`, leaf-order `triangles`, and
 * the leaf-ordered faces of render.cc:88-95), so it is parity-neutral.
 *
 * bvh.cc recurses: box of the triangles + box of their centroids, cut at the midpoint of the centroid box's longest
 * axis, STABLE partition into "centroid inside the lower half" / rest, fix-ups when a side is empty, left subtree
 * first.  Because the partition is stable and left precedes right, the recursion is a sequence of in-place stable
 * partitions of ONE id array whose final order is the leaf order, a subtree over n triangles occupies 2n-1
 * consecutive node slots, and every quantity of a node depends only on the set of ids in its segment -- so all
 * segments of one depth can be processed together:
 *
 *   k_bvh_accumulate   per id: min/max of its triangle box and centroid into its segment's accumulators (order-
 *                      preserving uint encoding of floats, atomicMin/Max; warp-level REDUX when a warp sits in one
 *                      segment).  min/max are exact and order-free, so the boxes are bit-identical to the
 *                      sequential std::min/std::max of aabb.cc:2-13 (up to the sign of a zero extreme).
 *   k_bvh_split        per segment: write nodes[] / aabbs[], longest axis (aabb.cc:14-23), cut = (max+min)/2 (:69)
 *   flag scan          goes-left flag = !(c > cut || c < lo) (aabb.cc:24-31 on the cut axis); global exclusive sum
 *   k_bvh_scatter      stable partition inside every segment, the empty-side fix-ups of bvh.cc:85-93, children's
 *                      segment records
 * until no segment holds more than one triangle; k_bvh_leaves then writes the leaves.  One level = 6 launches.
 */
#pragma once
#include "rtx_device.cuh"

RTX_DEV uint32_t f2ord(float f) { const uint32_t b = __float_as_uint(f); return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u); }
RTX_DEV float ord2f(uint32_t e) { return __uint_as_float(e ^ ((e >> 31) ? 0x80000000u : 0xffffffffu)); }

#define RTX_BVH_ACC 12        /* per segment: box lo[3] hi[3], centroid box lo[3] hi[3] (encoded); after the split: axis, cut, lo */

struct BvhSeg { uint32_t start, n, node; };     /* segment of the id array a position belongs to, and its pre-order node slot */

/* triangle.cc:4-22: centroid ((a+b)+c)/3 and vertex min/max, once per triangle (input order) */
__global__ void k_bvh_prepare(const uint32_t *__restrict__ faces, const float4 *__restrict__ verts, uint32_t ntris, uint32_t nverts,
                              float *__restrict__ tc, float *__restrict__ tlo, float *__restrict__ thi,
                              uint32_t *__restrict__ ids, BvhSeg *__restrict__ seg, uint32_t *__restrict__ acc, unsigned int *bad_face)
{
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= ntris) return;
	uint32_t i0 = faces[3 * (size_t)t], i1 = faces[3 * (size_t)t + 1], i2 = faces[3 * (size_t)t + 2];
	if (i0 >= nverts || i1 >= nverts || i2 >= nverts) { *bad_face = 1u; i0 = i1 = i2 = 0; }
	const float4 A = verts[i0], B = verts[i1], C = verts[i2];
	const float a[3] = { A.x, A.y, A.z }, b[3] = { B.x, B.y, B.z }, c[3] = { C.x, C.y, C.z };
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		tc[3 * (size_t)t + k] = rn_div(rn_add(rn_add(a[k], b[k]), c[k]), 3.0f);
		const float mbc = c[k] < b[k] ? c[k] : b[k], Mbc = b[k] < c[k] ? c[k] : b[k];     /* std::min / std::max keep the FIRST of equals */
		tlo[3 * (size_t)t + k] = mbc < a[k] ? mbc : a[k];                                 /* triangle.cc:7-22 */
		thi[3 * (size_t)t + k] = a[k] < Mbc ? Mbc : a[k];
	}
	ids[t] = t;
	seg[t] = BvhSeg{ 0u, ntris, 0u };
	if (t == 0) {
#pragma unroll
		for (int k = 0; k < 3; ++k) {
			acc[k] = acc[6 + k] = f2ord(3.402823466e+38f);              /* AABB(): min = +max, max = -max (aabb.h:22-24) */
			acc[3 + k] = acc[9 + k] = f2ord(-3.402823466e+38f);
		}
	}
}

#define RTX_BVH_ACC_ITEMS 8      /* consecutive ids per thread in k_bvh_accumulate */

RTX_DEV void bvh_acc_flush(uint32_t *__restrict__ acc, uint32_t start, const uint32_t (&v)[RTX_BVH_ACC])
{
	uint32_t *a = acc + (size_t)start * RTX_BVH_ACC;
#pragma unroll
	for (int k = 0; k < RTX_BVH_ACC; ++k) {
		if ((k % 6) < 3) atomicMin(a + k, v[k]); else atomicMax(a + k, v[k]);
	}
}

/* Every thread folds RTX_BVH_ACC_ITEMS consecutive ids: runs of one segment are reduced in registers and only a run
 * that ends inside the thread goes to memory at once; the thread's LAST run is then merged with the neighbouring
 * lanes' (segments are contiguous, so equal keys are adjacent lanes): one REDUX per quantity when the whole warp sits
 * in one segment, a segmented shuffle reduction otherwise, and only the first lane of each run issues the atomics.
 * The top levels (a few huge segments) thus cost N / 256 atomics per address instead of N. */
__global__ void k_bvh_accumulate(const uint32_t *__restrict__ ids, const BvhSeg *__restrict__ seg, uint32_t ntris,
                                 const float *__restrict__ tc, const float *__restrict__ tlo, const float *__restrict__ thi,
                                 uint32_t *__restrict__ acc)
{
	const uint32_t lane = threadIdx.x & 31u;
	const size_t base = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * RTX_BVH_ACC_ITEMS;
	uint32_t v[RTX_BVH_ACC];
#pragma unroll
	for (int k = 0; k < RTX_BVH_ACC; ++k) v[k] = (k % 6) < 3 ? 0xffffffffu : 0u;       /* neutral for min / max */
	uint32_t cur = 0xffffffffu;                                                   /* segment of the run being folded */
	for (uint32_t j0 = 0; j0 < RTX_BVH_ACC_ITEMS; j0 += 4) {
		/* four ids at a time: all their loads are issued before the first is folded (the per-triangle data is a
		 * random access, and a thread that waits for one element at a time is bound by that latency) */
		BvhSeg sg[4];
		uint32_t e[4][9];
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			const size_t i = base + j0 + j;
			sg[j] = i < ntris ? seg[i] : BvhSeg{ 0u, 0u, 0u };
		}
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			if (sg[j].n <= 1) continue;
			const size_t t = ids[base + j0 + j];
#pragma unroll
			for (int k = 0; k < 3; ++k) {
				e[j][k] = f2ord(tlo[3 * t + k]);
				e[j][3 + k] = f2ord(thi[3 * t + k]);
				e[j][6 + k] = f2ord(tc[3 * t + k]);
			}
		}
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			if (sg[j].n <= 1) continue;
			if (cur != 0xffffffffu && sg[j].start != cur) {                           /* a run ended inside this thread */
				bvh_acc_flush(acc, cur, v);
#pragma unroll
				for (int k = 0; k < RTX_BVH_ACC; ++k) v[k] = (k % 6) < 3 ? 0xffffffffu : 0u;
			}
			cur = sg[j].start;
#pragma unroll
			for (int k = 0; k < 3; ++k) {
				v[k] = min(v[k], e[j][k]);
				v[3 + k] = max(v[3 + k], e[j][3 + k]);
				v[6 + k] = min(v[6 + k], e[j][6 + k]);
				v[9 + k] = max(v[9 + k], e[j][6 + k]);
			}
		}
	}
	const bool act = cur != 0xffffffffu;
	const uint32_t s0 = __shfl_sync(0xffffffffu, cur, 0);
	if (__all_sync(0xffffffffu, act && cur == s0)) {                              /* the whole warp in one segment */
#pragma unroll
		for (int k = 0; k < RTX_BVH_ACC; ++k)
			v[k] = (k % 6) < 3 ? __reduce_min_sync(0xffffffffu, v[k]) : __reduce_max_sync(0xffffffffu, v[k]);
		if (lane != 0) return;
	} else {
		const uint32_t key = act ? cur : 0xfffffffeu - lane;                      /* idle lanes: runs of their own */
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t k2 = __shfl_down_sync(0xffffffffu, key, o);
			const bool same = lane + o < 32u && k2 == key;
#pragma unroll
			for (int k = 0; k < RTX_BVH_ACC; ++k) {
				const uint32_t u = __shfl_down_sync(0xffffffffu, v[k], o);
				if (same) v[k] = (k % 6) < 3 ? min(v[k], u) : max(v[k], u);
			}
		}
		const uint32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
		if (!act || (lane != 0 && prev == key)) return;                           /* only run heads continue */
	}
	bvh_acc_flush(acc, cur, v);
}

/* the first position of every live segment: node record + split plane; acc[0..2] <- (axis, cut, lo) */
__global__ void k_bvh_split(const BvhSeg *__restrict__ seg, uint32_t ntris, uint32_t *__restrict__ acc,
                            uint32_t *__restrict__ nodes, float4 *__restrict__ aabbs)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= ntris) return;
	const BvhSeg s = seg[i];
	if (s.n <= 1 || s.start != i) return;
	uint32_t *a = acc + (size_t)i * RTX_BVH_ACC;
	float lo[3], hi[3], clo[3], chi[3];
#pragma unroll
	for (int k = 0; k < 3; ++k) { lo[k] = ord2f(a[k]); hi[k] = ord2f(a[3 + k]); clo[k] = ord2f(a[6 + k]); chi[k] = ord2f(a[9 + k]); }
	nodes[s.node] = 2u * s.n - 1u;                                           /* bvh.cc:154-157: subtree size */
	aabbs[2 * (size_t)s.node] = make_float4(lo[0], lo[1], lo[2], 0.f);
	aabbs[2 * (size_t)s.node + 1] = make_float4(hi[0], hi[1], hi[2], 0.f);
	const float d0 = rn_sub(chi[0], clo[0]), d1 = rn_sub(chi[1], clo[1]), d2 = rn_sub(chi[2], clo[2]);
	int axis = 2;                                                            /* aabb.cc:14-23 */
	if (d0 >= d1 && d0 >= d2) axis = 0;
	else if (d1 >= d0 && d1 >= d2) axis = 1;
	const float cut = rn_div(rn_add(chi[axis], clo[axis]), 2.0f);            /* bvh.cc:69 */
	a[0] = (uint32_t)axis;
	a[1] = __float_as_uint(cut);
	a[2] = __float_as_uint(clo[axis]);
}

/* 1 if position i's triangle goes to the left child (aabb.cc:24-31 on the cut axis; the other axes always pass) */
RTX_DEV uint32_t bvh_goes_left(uint32_t i, const uint32_t *ids, const BvhSeg *seg, const uint32_t *acc, const float *tc)
{
	const BvhSeg s = seg[i];
	if (s.n <= 1) return 0u;
	const uint32_t *a = acc + (size_t)s.start * RTX_BVH_ACC;
	const float c = tc[3 * (size_t)ids[i] + a[0]];
	return (c > __uint_as_float(a[1]) || c < __uint_as_float(a[2])) ? 0u : 1u;
}

#define RTX_BVH_ITEMS 8
#define RTX_BVH_BLOCK 256

__global__ void __launch_bounds__(RTX_BVH_BLOCK)
k_bvh_flag_partials(const uint32_t *__restrict__ ids, const BvhSeg *__restrict__ seg, const uint32_t *__restrict__ acc,
                    const float *__restrict__ tc, uint32_t ntris, uint32_t *__restrict__ partials)
{
	__shared__ uint32_t s[RTX_BVH_BLOCK / 32];
	const uint32_t base = (blockIdx.x * RTX_BVH_BLOCK + threadIdx.x) * RTX_BVH_ITEMS;
	uint32_t a = 0;
	for (uint32_t k = 0; k < RTX_BVH_ITEMS; ++k)
		if (base + k < ntris) a += bvh_goes_left(base + k, ids, seg, acc, tc);
	a = __reduce_add_sync(0xffffffffu, a);
	if ((threadIdx.x & 31u) == 0) s[threadIdx.x >> 5] = a;
	__syncthreads();
	if (threadIdx.x == 0) {
		uint32_t t = 0;
		for (int w = 0; w < RTX_BVH_BLOCK / 32; ++w) t += s[w];
		partials[blockIdx.x] = t;
	}
}

__global__ void k_bvh_spine(uint32_t *__restrict__ partials, uint32_t nblocks)      /* one warp: exclusive scan in place */
{
	const uint32_t lane = threadIdx.x & 31u;
	uint32_t carry = 0;
	for (uint32_t base = 0; base < nblocks; base += 32) {
		const uint32_t i = base + lane;
		const uint32_t v = i < nblocks ? partials[i] : 0u;
		uint32_t x = v;
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
			if ((int)lane >= o) x += y;
		}
		if (i < nblocks) partials[i] = carry + x - v;
		carry += __shfl_sync(0xffffffffu, x, 31);
	}
}

/* scan[i] = number of goes-left flags before position i; scan[ntris] = total */
__global__ void __launch_bounds__(RTX_BVH_BLOCK)
k_bvh_flag_scan(const uint32_t *__restrict__ ids, const BvhSeg *__restrict__ seg, const uint32_t *__restrict__ acc,
                const float *__restrict__ tc, uint32_t ntris, const uint32_t *__restrict__ partials, uint32_t *__restrict__ scan)
{
	__shared__ uint32_t s[RTX_BVH_BLOCK / 32];
	const uint32_t base = (blockIdx.x * RTX_BVH_BLOCK + threadIdx.x) * RTX_BVH_ITEMS;
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	uint32_t f[RTX_BVH_ITEMS], a = 0;
	for (uint32_t k = 0; k < RTX_BVH_ITEMS; ++k) {
		f[k] = base + k < ntris ? bvh_goes_left(base + k, ids, seg, acc, tc) : 0u;
		a += f[k];
	}
	uint32_t x = a;
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
		if ((int)lane >= o) x += y;
	}
	if (lane == 31) s[warp] = x;
	__syncthreads();
	uint32_t off = partials[blockIdx.x];
	for (uint32_t w = 0; w < warp; ++w) off += s[w];
	off += x - a;
	for (uint32_t k = 0; k < RTX_BVH_ITEMS; ++k) {
		if (base + k < ntris) scan[base + k] = off;
		off += f[k];
		if (base + k + 1 == ntris) scan[ntris] = off;
	}
}

/* stable partition of every live segment (bvh.cc:73-83), the fix-ups for an empty side (:85-93: the LAST id of the
 * full side moves over), and the segment records of the two children (left node = node+1, right = node+2*nL) */
__global__ void k_bvh_scatter(const uint32_t *__restrict__ ids, const BvhSeg *__restrict__ seg, const uint32_t *__restrict__ acc_in,
                              const float *__restrict__ tc, const uint32_t *__restrict__ scan, uint32_t ntris,
                              uint32_t *__restrict__ ids_out, BvhSeg *__restrict__ seg_out, uint32_t *__restrict__ acc_out, unsigned int *live)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= ntris) return;
	const BvhSeg s = seg[i];
	if (s.n <= 1) { ids_out[i] = ids[i]; seg_out[i] = s; return; }
	const uint32_t nl0 = scan[s.start + s.n] - scan[s.start];
	const uint32_t pos = i - s.start, lrank = scan[i] - scan[s.start];
	const uint32_t left = bvh_goes_left(i, ids, seg, acc_in, tc);
	uint32_t nl, dest;
	if (nl0 == 0) {                      /* nothing went left: the last id becomes the left child */
		nl = 1;
		dest = pos == s.n - 1 ? 0u : pos + 1;
	} else if (nl0 == s.n) {             /* nothing went right: the last id becomes the right child */
		nl = s.n - 1;
		dest = pos;
	} else {
		nl = nl0;
		dest = left ? lrank : nl + (pos - lrank);
	}
	const bool in_left = dest < nl;
	const BvhSeg child = in_left ? BvhSeg{ s.start, nl, s.node + 1 } : BvhSeg{ s.start + nl, s.n - nl, s.node + 2 * nl };
	const uint32_t d = s.start + dest;
	ids_out[d] = ids[i];
	seg_out[d] = child;
	if (d == child.start && child.n > 1) {           /* fresh accumulators for a child that will be split again */
		uint32_t *a = acc_out + (size_t)d * RTX_BVH_ACC;
#pragma unroll
		for (int k = 0; k < 3; ++k) {
			a[k] = a[6 + k] = f2ord(3.402823466e+38f);
			a[3 + k] = a[9 + k] = f2ord(-3.402823466e+38f);
		}
		atomicAdd(live, 1u);
	}
}

/* bvh.cc:118-130 + render.cc:88-95: the leaves, the leaf-order triangle ids and the leaf-ordered faces */
__global__ void k_bvh_leaves(const uint32_t *__restrict__ ids, const BvhSeg *__restrict__ seg, uint32_t ntris,
                             const float *__restrict__ tlo, const float *__restrict__ thi, const uint32_t *__restrict__ faces,
                             uint32_t *__restrict__ nodes, float4 *__restrict__ aabbs, uint32_t *__restrict__ triangles,
                             uint32_t *__restrict__ sorted_faces)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= ntris) return;
	const size_t t = ids[i];
	const uint32_t node = seg[i].node;
	nodes[node] = 1u;
	aabbs[2 * (size_t)node] = make_float4(tlo[3 * t], tlo[3 * t + 1], tlo[3 * t + 2], 0.f);
	aabbs[2 * (size_t)node + 1] = make_float4(thi[3 * t], thi[3 * t + 1], thi[3 * t + 2], 0.f);
	triangles[i] = (uint32_t)t;
	sorted_faces[3 * (size_t)i] = faces[3 * t];
	sorted_faces[3 * (size_t)i + 1] = faces[3 * t + 1];
	sorted_faces[3 * (size_t)i + 2] = faces[3 * t + 2];
}

/* --------------------------------------------------------------------------
 * Vertex normals on the device: compute_vertex_normals (src/mesh.cc:95-139).  The reference adds each face's
 * un-normalised normal to its three vertices IN FACE ORDER, so a vertex's sum depends on that order; here every
 * vertex gets the list of its (face, corner) entries, sorts it and adds in the same order.
 *   k_vn_faces       per face: (b-a) x (c-a) with one rounding per operation, and its length (skip when 0)
 *   k_vn_count/fill  per corner: incidence lists by vertex (counting sort; the fill order is arbitrary)
 *   k_vn_vertices    per vertex: sort its list by entry index, add in order, divide by the length unless 0
 * ------------------------------------------------------------------------ */
__global__ void k_vn_faces(const uint32_t *__restrict__ faces, const float4 *__restrict__ verts, uint32_t nfaces, uint32_t nverts,
                           float4 *__restrict__ fnormal, uint32_t *__restrict__ count)
{
	const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= nfaces) return;
	const uint32_t i0 = faces[3 * (size_t)f], i1 = faces[3 * (size_t)f + 1], i2 = faces[3 * (size_t)f + 2];
	if (i0 >= nverts || i1 >= nverts || i2 >= nverts) { fnormal[f] = make_float4(0.f, 0.f, 0.f, 0.f); return; }   /* reported by the builder */
	const float4 A = verts[i0], B = verts[i1], C = verts[i2];
	const f3 u = make_f3(rn_sub(B.x, A.x), rn_sub(B.y, A.y), rn_sub(B.z, A.z));
	const f3 v = make_f3(rn_sub(C.x, A.x), rn_sub(C.y, A.y), rn_sub(C.z, A.z));
	const f3 n = cross3(u, v);
	const float len = rn_sqrt(dot3(n, n));
	fnormal[f] = make_float4(n.x, n.y, n.z, len == 0.0f ? 0.0f : 1.0f);
	atomicAdd(count + i0, 1u);
	atomicAdd(count + i1, 1u);
	atomicAdd(count + i2, 1u);
}

__global__ void k_vn_fill(const uint32_t *__restrict__ faces, uint32_t nfaces, uint32_t nverts, const uint32_t *__restrict__ offset,
                          uint32_t *__restrict__ cursor, uint32_t *__restrict__ list)
{
	const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= nfaces) return;
	const uint32_t i[3] = { faces[3 * (size_t)f], faces[3 * (size_t)f + 1], faces[3 * (size_t)f + 2] };
	if (i[0] >= nverts || i[1] >= nverts || i[2] >= nverts) return;
#pragma unroll
	for (uint32_t k = 0; k < 3; ++k) list[offset[i[k]] + atomicAdd(cursor + i[k], 1u)] = 3u * f + k;
}

__global__ void k_vn_vertices(const uint32_t *__restrict__ offset, uint32_t *__restrict__ list, const float4 *__restrict__ fnormal,
                              uint32_t nverts, float4 *__restrict__ vnormals)
{
	const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
	if (v >= nverts) return;
	const uint32_t b = offset[v], e = offset[v + 1];
	for (uint32_t i = b + 1; i < e; ++i) {                   /* insertion sort: valences are small */
		const uint32_t x = list[i];
		uint32_t j = i;
		while (j > b && list[j - 1] > x) { list[j] = list[j - 1]; --j; }
		list[j] = x;
	}
	float nx = 0.f, ny = 0.f, nz = 0.f;
	for (uint32_t i = b; i < e; ++i) {
		const float4 n = fnormal[list[i] / 3u];
		if (n.w == 0.0f) continue;                               /* mesh.cc:112-114: zero-length face normals are skipped */
		nx = rn_add(nx, n.x); ny = rn_add(ny, n.y); nz = rn_add(nz, n.z);
	}
	const float len = rn_sqrt(rn_add(rn_add(rn_mul(nx, nx), rn_mul(ny, ny)), rn_mul(nz, nz)));
	if (len > 0.0f) { nx = rn_div(nx, len); ny = rn_div(ny, len); nz = rn_div(nz, len); }
	vnormals[v] = make_float4(nx, ny, nz, 0.f);
}

/* generic exclusive sum of a u32 array (n + 1 outputs: out[n] = total), three launches like the flag scan */
__global__ void __launch_bounds__(RTX_BVH_BLOCK)
k_scan_partials(const uint32_t *__restrict__ in, uint32_t n, uint32_t *__restrict__ partials)
{
	__shared__ uint32_t s[RTX_BVH_BLOCK / 32];
	const uint32_t base = (blockIdx.x * RTX_BVH_BLOCK + threadIdx.x) * RTX_BVH_ITEMS;
	uint32_t a = 0;
	for (uint32_t k = 0; k < RTX_BVH_ITEMS; ++k)
		if (base + k < n) a += in[base + k];
	a = __reduce_add_sync(0xffffffffu, a);
	if ((threadIdx.x & 31u) == 0) s[threadIdx.x >> 5] = a;
	__syncthreads();
	if (threadIdx.x == 0) {
		uint32_t t = 0;
		for (int w = 0; w < RTX_BVH_BLOCK / 32; ++w) t += s[w];
		partials[blockIdx.x] = t;
	}
}

__global__ void __launch_bounds__(RTX_BVH_BLOCK)
k_scan_apply(const uint32_t *__restrict__ in, uint32_t n, const uint32_t *__restrict__ partials, uint32_t *__restrict__ out)
{
	__shared__ uint32_t s[RTX_BVH_BLOCK / 32];
	const uint32_t base = (blockIdx.x * RTX_BVH_BLOCK + threadIdx.x) * RTX_BVH_ITEMS;
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	uint32_t f[RTX_BVH_ITEMS], a = 0;
	for (uint32_t k = 0; k < RTX_BVH_ITEMS; ++k) {
		f[k] = base + k < n ? in[base + k] : 0u;
		a += f[k];
	}
	uint32_t x = a;
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
		if ((int)lane >= o) x += y;
	}
	if (lane == 31) s[warp] = x;
	__syncthreads();
	uint32_t off = partials[blockIdx.x];
	for (uint32_t w = 0; w < warp; ++w) off += s[w];
	off += x - a;
	for (uint32_t k = 0; k < RTX_BVH_ITEMS; ++k) {
		if (base + k < n) out[base + k] = off;
		off += f[k];
		if (base + k + 1 == n) out[n] = off;
	}
}
