/*
 * scene_prep.cc -- host scene preparation in the reference's upload format.
 * C ABI: include/rtx_scene.h.  Pure host C++ (no CUDA); built with
 * -ffp-contract=off so every float operation rounds once, like the
 * reference's own host build.
 *
 * The tree that comes out is, array for array, the one BVH::buildBVH produces
 * with CUT_LONGEST_AXIS (reference src/bvh.cc:59-162): pre-order `nodes`
 * holding subtree sizes, `aabbs` as (min,max) float4 pairs, one triangle per
 * leaf, `triangles` in leaf order.  The construction differs: per-triangle
 * centroids and boxes are computed once, ranges are partitioned in place, and
 * because a subtree over k triangles always has exactly 2k-1 nodes the slot of
 * every child is known before it is built, so disjoint subtrees are built by
 * worker threads with no ordering between them.
 */
#include "rtx_scene.h"

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_error;

struct V4 { float x, y, z, w; };

struct rtx_scene_impl {
	std::vector<V4> vertices, normals, aabbs;
	std::vector<uint32_t> orig_faces, faces, triangles, nodes;
};

int fail(int code, const std::string &msg)
{
	g_error = msg;
	return code;
}

/* ---------------------------------------------------------------- OFF --- */

/* Tokens are whitespace separated; numbers are parsed with strtof/strtoul,
 * which is what iostream extraction does underneath (mesh.cc:29-47). */
struct Tokens {
	const char *p, *end;
	bool next(const char *&b, const char *&e)
	{
		while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r' || *p == '\f' || *p == '\v')) ++p;
		if (p >= end) return false;
		b = p;
		while (p < end && !(*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r' || *p == '\f' || *p == '\v')) ++p;
		e = p;
		return true;
	}
	bool next_float(float &v)
	{
		const char *b, *e;
		if (!next(b, e)) return false;
		char *q = nullptr;
		v = std::strtof(b, &q);
		return q != b;
	}
	bool next_uint(unsigned long &v)
	{
		const char *b, *e;
		if (!next(b, e)) return false;
		char *q = nullptr;
		v = std::strtoul(b, &q, 10);
		return q != b;
	}
};

int read_off(const char *path, std::vector<float> &verts3, std::vector<uint32_t> &faces, size_t &nverts)
{
	if (!path || !*path) return fail(RTX_SCENE_ERR_ARG, "No filename given");
	FILE *f = std::fopen(path, "rb");
	if (!f) return fail(RTX_SCENE_ERR_IO, std::string("Cannot read file: ") + path);
	std::string text;
	char buf[1 << 16];
	size_t got;
	while ((got = std::fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, got);
	std::fclose(f);
	text.push_back('\0');
	Tokens tk{ text.data(), text.data() + text.size() - 1 };
	const char *b, *e;
	if (!tk.next(b, e) || e - b != 3 || std::memcmp(b, "OFF", 3) != 0)
		return fail(RTX_SCENE_ERR_FORMAT, "File not recognized as OFF model");
	unsigned long nv = 0, nf = 0, ne = 0;
	if (!tk.next_uint(nv) || !tk.next_uint(nf) || !tk.next_uint(ne))
		return fail(RTX_SCENE_ERR_FORMAT, "OFF header truncated");
	nverts = nv;
	verts3.resize(3 * nv);
	for (size_t i = 0; i < 3 * (size_t)nv; ++i)
		if (!tk.next_float(verts3[i])) return fail(RTX_SCENE_ERR_FORMAT, "OFF vertex list truncated");
	faces.clear();
	faces.reserve(3 * nf);
	for (size_t i = 0; i < nf; ++i) {
		unsigned long n = 0;
		if (!tk.next_uint(n)) return fail(RTX_SCENE_ERR_FORMAT, "OFF face list truncated");
		if (n != 3) return fail(RTX_SCENE_ERR_FORMAT, "Invalid face with != 3 vertices");
		unsigned long id[3] = { 0, 0, 0 };
		for (int j = 0; j < 3; ++j)
			if (!tk.next_uint(id[j])) return fail(RTX_SCENE_ERR_FORMAT, "OFF face list truncated");
		for (int j = 0; j < 3; ++j) faces.push_back((uint32_t)std::min<unsigned long>(id[j], 0xfffffffful));
	}
	return RTX_SCENE_OK;
}

/* ------------------------------------------------------------ normals --- */

/* Area-weighted vertex normals (mesh.cc:95-139): face normal = (b-a) x (c-a)
 * accumulated un-normalised on its three vertices unless its length is 0;
 * each sum divided by its length unless that is 0. */
void vertex_normals(rtx_scene_impl &s)
{
	const size_t nv = s.vertices.size();
	s.normals.assign(nv, V4{ 0, 0, 0, 0 });
	const std::vector<uint32_t> &F = s.orig_faces;
	for (size_t i = 0; i + 2 < F.size(); i += 3) {
		const V4 &a = s.vertices[F[i]], &b = s.vertices[F[i + 1]], &c = s.vertices[F[i + 2]];
		const float ux = b.x - a.x, uy = b.y - a.y, uz = b.z - a.z;
		const float vx = c.x - a.x, vy = c.y - a.y, vz = c.z - a.z;
		const float nx = uy * vz - vy * uz;
		const float ny = uz * vx - vz * ux;
		const float nz = ux * vy - vx * uy;
		const float len = std::sqrt(nx * nx + ny * ny + nz * nz);
		if (len == 0) continue;
		for (int k = 0; k < 3; ++k) {
			V4 &n = s.normals[F[i + k]];
			n.x += nx; n.y += ny; n.z += nz;
		}
	}
	for (size_t i = 0; i < nv; ++i) {
		V4 &n = s.normals[i];
		const float len = std::sqrt(n.x * n.x + n.y * n.y + n.z * n.z);
		if (len > 0) { n.x /= len; n.y /= len; n.z /= len; }
	}
}

/* ---------------------------------------------------------------- BVH --- */

struct Box {
	float lo[3], hi[3];
	void reset()
	{
		for (int k = 0; k < 3; ++k) { lo[k] = std::numeric_limits<float>::max(); hi[k] = -std::numeric_limits<float>::max(); }
	}
	void grow(const float *l, const float *h)
	{
		for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], l[k]); hi[k] = std::max(hi[k], h[k]); }
	}
};

struct Builder {
	rtx_scene_impl &s;
	std::vector<float> centroid;   /* 3 per triangle: ((v0+v1)+v2)/3  (triangle.cc:4-6) */
	std::vector<float> tlo, thi;   /* 3 per triangle: vertex min/max  (triangle.cc:7-22) */
	std::vector<uint32_t> ids;     /* working permutation of triangle ids */
	std::vector<uint32_t> scratch;

	/* parallel task pool: a task = (first id slot, count, first node slot) */
	struct Task { size_t b, n, node; };
	std::mutex mu;
	std::condition_variable cv;
	std::vector<Task> queue;
	size_t pending = 0;
	size_t spawn_threshold = 0;

	bool sah = false;              /* BVH::Method::SURFACE_AREA_HEURISTIC (bvh.cc:178-236) instead of CUT_LONGEST_AXIS */
	std::vector<float> suffix;     /* SAH sweep: boxes of ids[i..end), 6 floats each */

	explicit Builder(rtx_scene_impl &sc) : s(sc) {}

	void prepare()
	{
		const std::vector<uint32_t> &F = s.orig_faces;
		const size_t nt = F.size() / 3;
		centroid.resize(3 * nt); tlo.resize(3 * nt); thi.resize(3 * nt);
		for (size_t t = 0; t < nt; ++t) {
			const float *a = &s.vertices[F[3 * t]].x, *b = &s.vertices[F[3 * t + 1]].x, *c = &s.vertices[F[3 * t + 2]].x;
			for (int k = 0; k < 3; ++k) {
				centroid[3 * t + k] = ((a[k] + b[k]) + c[k]) / 3.0f;
				tlo[3 * t + k] = std::min(a[k], std::min(b[k], c[k]));
				thi[3 * t + k] = std::max(a[k], std::max(b[k], c[k]));
			}
		}
		ids.resize(nt); scratch.resize(nt);
		if (sah) suffix.resize(6 * nt);
		for (size_t t = 0; t < nt; ++t) ids[t] = (uint32_t)t;
		s.nodes.assign(2 * nt - 1, 0);
		s.aabbs.assign(2 * (2 * nt - 1), V4{ 0, 0, 0, 0 });
		s.triangles.assign(nt, 0);
	}

	void put_box(size_t node, const Box &bb)
	{
		s.aabbs[2 * node] = V4{ bb.lo[0], bb.lo[1], bb.lo[2], 0 };
		s.aabbs[2 * node + 1] = V4{ bb.hi[0], bb.hi[1], bb.hi[2], 0 };
	}

	/* Split ids[b, b+n) like cutFacesLongestAxis (bvh.cc:59-94); returns the
	 * size of the left part.  Order inside each part is the input order. */
	size_t split(size_t b, size_t n, Box &bb)
	{
		Box cb;
		bb.reset(); cb.reset();
		for (size_t i = b; i < b + n; ++i) {
			const size_t t = ids[i];
			cb.grow(&centroid[3 * t], &centroid[3 * t]);
			bb.grow(&tlo[3 * t], &thi[3 * t]);
		}
		const float d0 = cb.hi[0] - cb.lo[0], d1 = cb.hi[1] - cb.lo[1], d2 = cb.hi[2] - cb.lo[2];
		int axis = 2;                                   /* aabb.cc:14-23 */
		if (d0 >= d1 && d0 >= d2) axis = 0;
		else if (d1 >= d0 && d1 >= d2) axis = 1;
		const float cut = (cb.hi[axis] + cb.lo[axis]) / 2;   /* bvh.cc:69 */
		const float lo = cb.lo[axis];
		size_t nl = 0, nr = 0;
		for (size_t i = b; i < b + n; ++i) {
			const uint32_t t = ids[i];
			const float c = centroid[3 * (size_t)t + axis];
			if (c > cut || c < lo) scratch[b + nr++] = t;   /* aabb.cc:24-31, other axes always inside */
			else ids[b + nl++] = t;
		}
		if (nl == 0) {                                   /* bvh.cc:85-88 */
			ids[b] = scratch[b + nr - 1];
			nl = 1; --nr;
		} else if (nr == 0) {                            /* bvh.cc:90-93 */
			scratch[b] = ids[b + nl - 1];
			nr = 1; --nl;
		}
		std::copy(scratch.begin() + b, scratch.begin() + b + nr, ids.begin() + b + nl);
		return nl;
	}

	/* getSurfaceArea (bvh.cc:37-42): float products and sums, doubled in double, returned as float */
	static float surface_area(const Box &bb)
	{
		const float w = bb.hi[0] - bb.lo[0], h = bb.hi[1] - bb.lo[1], d = bb.hi[2] - bb.lo[2];
		return (float)(2.0 * (double)(w * h + h * d + d * w));
	}

	/* Split ids[b, b+n) like cutFacesSAH (bvh.cc:178-236); returns the size of the left part.  The reference sorts the
	 * node's ids by descending centroid along each axis with std::sort, sweeps every cut position (it recomputes the
	 * right-hand box for every position: O(n^2); min / max are exact, so a suffix scan gives the same boxes), keeps the
	 * first strictly cheaper (axis, position), sorts once more along the best axis unless that was the last one, and cuts.
	 * The same std::sort calls on the same sequences here, so ties between equal centroids fall the same way (same
	 * libstdc++); the cost expression keeps the reference's float / double mix. */
	size_t split_sah(size_t b, size_t n, Box &bb)
	{
		bb.reset();
		for (size_t i = b; i < b + n; ++i) bb.grow(&tlo[3 * (size_t)ids[i]], &thi[3 * (size_t)ids[i]]);
		const float cBV2 = 1.0, cObj = 1.0;
		const float sa_current = surface_area(bb);
		size_t best_axis = 0, best_pos = 1;
		float min_costs = std::numeric_limits<float>::max();
		const auto first = ids.begin() + (std::ptrdiff_t)b, last = first + (std::ptrdiff_t)n;
		for (size_t axis = 0; axis < 3; ++axis) {
			const float *cen = centroid.data();
			std::sort(first, last, [cen, axis](size_t i, size_t j) { return cen[3 * i + axis] > cen[3 * j + axis]; });
			if (n < 3) continue;                         /* no interior cut position to evaluate (the loop at :204 is empty) */
			/* suffix[i] = box of ids[b+i .. b+n) */
			Box acc;
			acc.reset();
			for (size_t i = n; i-- > 1;) {
				const size_t t = ids[b + i];
				acc.grow(&tlo[3 * t], &thi[3 * t]);
				float *o = &suffix[6 * (b + i)];
				o[0] = acc.lo[0]; o[1] = acc.lo[1]; o[2] = acc.lo[2]; o[3] = acc.hi[0]; o[4] = acc.hi[1]; o[5] = acc.hi[2];
			}
			Box left;
			left.reset();
			left.grow(&tlo[3 * (size_t)ids[b]], &thi[3 * (size_t)ids[b]]);
			double count_left = 1, count_right = (double)(n - 1);
			for (size_t i = 1; i < n - 1; ++i) {
				Box right;
				const float *o = &suffix[6 * (b + i)];
				right.lo[0] = o[0]; right.lo[1] = o[1]; right.lo[2] = o[2]; right.hi[0] = o[3]; right.hi[1] = o[4]; right.hi[2] = o[5];
				const float sa_left = surface_area(left), sa_right = surface_area(right);
				const float costs = (float)(cBV2 + (sa_left / sa_current) * count_left * cObj + (sa_right / sa_current) * count_right * cObj);
				if (costs < min_costs) { min_costs = costs; best_pos = i; best_axis = axis; }
				const size_t t = ids[b + i];
				left.grow(&tlo[3 * t], &thi[3 * t]);
				++count_left;
				--count_right;
			}
		}
		if (best_axis < 2) {
			const float *cen = centroid.data();
			const size_t axis = best_axis;
			std::sort(first, last, [cen, axis](size_t i, size_t j) { return cen[3 * i + axis] > cen[3 * j + axis]; });
		}
		return best_pos;
	}

	/* Build the subtree over ids[b, b+n) into node slots [node, node+2n-1);
	 * its leaves are leaf-order positions [b, b+n). */
	void build(size_t b, size_t n, size_t node, bool may_spawn)
	{
		/* explicit stack of right siblings still to do */
		std::vector<Task> todo;
		todo.push_back(Task{ b, n, node });
		while (!todo.empty()) {
			Task t = todo.back();
			todo.pop_back();
			for (;;) {
				if (t.n == 1) {                          /* bvh.cc:118-130 */
					const size_t tri = ids[t.b];
					Box bb;
					bb.reset();
					bb.grow(&tlo[3 * tri], &thi[3 * tri]);
					put_box(t.node, bb);
					s.nodes[t.node] = 1;
					s.triangles[t.b] = (uint32_t)tri;
					break;
				}
				Box bb;
				const size_t nl = sah ? split_sah(t.b, t.n, bb) : split(t.b, t.n, bb);
				put_box(t.node, bb);
				s.nodes[t.node] = (uint32_t)(2 * t.n - 1);
				const Task right{ t.b + nl, t.n - nl, t.node + 2 * nl };
				if (may_spawn && right.n >= spawn_threshold) spawn(right);
				else todo.push_back(right);
				t = Task{ t.b, nl, t.node + 1 };
			}
		}
	}

	void spawn(const Task &t)
	{
		std::lock_guard<std::mutex> lk(mu);
		queue.push_back(t);
		++pending;
		cv.notify_one();
	}

	void worker()
	{
		for (;;) {
			Task t;
			{
				std::unique_lock<std::mutex> lk(mu);
				cv.wait(lk, [&] { return !queue.empty() || pending == 0; });
				if (queue.empty()) return;
				t = queue.back();
				queue.pop_back();
			}
			build(t.b, t.n, t.node, true);
			{
				std::lock_guard<std::mutex> lk(mu);
				if (--pending == 0) cv.notify_all();
			}
		}
	}

	void run(int nthreads)
	{
		const size_t nt = ids.size();
		if (nthreads <= 0) nthreads = (int)std::max(1u, std::thread::hardware_concurrency());
		if (nthreads == 1 || nt < 65536) {
			build(0, nt, 0, false);
			return;
		}
		spawn_threshold = std::max<size_t>(16384, nt / (64 * (size_t)nthreads));
		spawn(Task{ 0, nt, 0 });
		std::vector<std::thread> th;
		for (int i = 0; i < nthreads; ++i) th.emplace_back([this] { worker(); });
		for (auto &t : th) t.join();
	}
};

int finish(rtx_scene_impl *s, const std::vector<float> &verts3, size_t nverts, const uint32_t *faces, size_t nfaces,
           int nthreads, rtx_scene **out, int method = RTX_SCENE_BVH_LONGEST_AXIS)
{
	s->vertices.resize(nverts);
	for (size_t i = 0; i < nverts; ++i) s->vertices[i] = V4{ verts3[3 * i], verts3[3 * i + 1], verts3[3 * i + 2], 0 };
	s->orig_faces.clear();
	s->orig_faces.reserve(3 * nfaces);
	size_t skipped = 0;
	for (size_t f = 0; f < nfaces; ++f) {                /* mesh.cc:44-59 */
		const uint32_t *v = faces + 3 * f;
		if (v[0] >= nverts || v[1] >= nverts || v[2] >= nverts) { ++skipped; continue; }
		s->orig_faces.insert(s->orig_faces.end(), v, v + 3);
	}
	if (skipped) std::fprintf(stderr, "rtx_scene: skipped %zu faces naming a vertex >= %zu\n", skipped, nverts);
	if (s->orig_faces.empty()) { delete s; return fail(RTX_SCENE_ERR_EMPTY, "mesh has no usable triangle"); }
	vertex_normals(*s);
	Builder bld(*s);
	bld.sah = method == RTX_SCENE_BVH_SAH;
	bld.prepare();
	bld.run(nthreads);
	const size_t nt = s->triangles.size();
	s->faces.resize(3 * nt);                             /* render.cc:88-95 */
	for (size_t i = 0; i < nt; ++i) {
		const size_t f = 3 * (size_t)s->triangles[i];
		s->faces[3 * i] = s->orig_faces[f];
		s->faces[3 * i + 1] = s->orig_faces[f + 1];
		s->faces[3 * i + 2] = s->orig_faces[f + 2];
	}
	*out = reinterpret_cast<rtx_scene *>(s);
	return RTX_SCENE_OK;
}

inline const rtx_scene_impl *impl(const rtx_scene *s) { return reinterpret_cast<const rtx_scene_impl *>(s); }

} /* namespace */

extern "C" {

int rtx_scene_from_off_method(const char *path, int method, int nthreads, rtx_scene **out)
{
	if (!out) return fail(RTX_SCENE_ERR_ARG, "null output pointer");
	*out = nullptr;
	if (method != RTX_SCENE_BVH_LONGEST_AXIS && method != RTX_SCENE_BVH_SAH) return fail(RTX_SCENE_ERR_ARG, "unknown BVH method");
	std::vector<float> verts3;
	std::vector<uint32_t> faces;
	size_t nverts = 0;
	const int rc = read_off(path, verts3, faces, nverts);
	if (rc != RTX_SCENE_OK) return rc;
	return finish(new rtx_scene_impl, verts3, nverts, faces.data(), faces.size() / 3, nthreads, out, method);
}

int rtx_scene_from_mesh_method(const float *verts3, size_t nverts, const uint32_t *faces, size_t nfaces, int method, int nthreads,
                               rtx_scene **out)
{
	if (!out) return fail(RTX_SCENE_ERR_ARG, "null output pointer");
	*out = nullptr;
	if (method != RTX_SCENE_BVH_LONGEST_AXIS && method != RTX_SCENE_BVH_SAH) return fail(RTX_SCENE_ERR_ARG, "unknown BVH method");
	if (!verts3 || !faces || nverts == 0 || nfaces == 0) return fail(RTX_SCENE_ERR_ARG, "empty mesh");
	std::vector<float> v(verts3, verts3 + 3 * nverts);
	return finish(new rtx_scene_impl, v, nverts, faces, nfaces, nthreads, out, method);
}

int rtx_scene_from_off(const char *path, int nthreads, rtx_scene **out)
{
	return rtx_scene_from_off_method(path, RTX_SCENE_BVH_LONGEST_AXIS, nthreads, out);
}

int rtx_scene_from_mesh(const float *verts3, size_t nverts, const uint32_t *faces, size_t nfaces, int nthreads, rtx_scene **out)
{
	return rtx_scene_from_mesh_method(verts3, nverts, faces, nfaces, RTX_SCENE_BVH_LONGEST_AXIS, nthreads, out);
}

void rtx_scene_free(rtx_scene *scene) { delete reinterpret_cast<rtx_scene_impl *>(scene); }

void rtx_scene_counts(const rtx_scene *scene, size_t counts[5])
{
	const rtx_scene_impl *s = impl(scene);
	counts[0] = s->faces.size();
	counts[1] = s->nodes.size();
	counts[2] = s->aabbs.size();
	counts[3] = s->vertices.size();
	counts[4] = s->normals.size();
}

const uint32_t *rtx_scene_faces(const rtx_scene *scene) { return impl(scene)->faces.data(); }
const uint32_t *rtx_scene_triangles(const rtx_scene *scene) { return impl(scene)->triangles.data(); }
const uint32_t *rtx_scene_orig_faces(const rtx_scene *scene) { return impl(scene)->orig_faces.data(); }
const uint32_t *rtx_scene_nodes(const rtx_scene *scene) { return impl(scene)->nodes.data(); }
const float *rtx_scene_aabbs(const rtx_scene *scene) { return &impl(scene)->aabbs.data()->x; }
const float *rtx_scene_vertices(const rtx_scene *scene) { return &impl(scene)->vertices.data()->x; }
const float *rtx_scene_normals(const rtx_scene *scene) { return &impl(scene)->normals.data()->x; }

const char *rtx_scene_last_error(void) { return g_error.c_str(); }

} /* extern "C" */
