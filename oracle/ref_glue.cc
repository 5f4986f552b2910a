/*
 * ref_glue.cc -- C entry points over the REFERENCE'S OWN code, compiled from
 * /root/reference where it lies (never copied into this repository):
 *   - src/intersect_kernel.cl, included as C++ through ref_shim/cl_compat.h
 *     after the two vector-literal rewrites of oracle/Makefile
 *     (`(float4) (` -> `mk4(`, `(int2) (` -> `mk2(`);
 *   - src/{mesh,bvh,aabb,triangle,ray_tracer}.cc and include/compiler_options.h,
 *     unmodified (g++ 13 needs -include cstdint/string/iterator).
 * Output: oracle/_ref/libref_oracle.so (git-ignored).  TEST INFRASTRUCTURE ONLY:
 * it validates oracle/rt_oracle.c and the product's scene preparation, and is
 * the "reference" CPU baseline of bench.py.
 *
 * Kernel build options mirrored from opencl_host.cc:42-53 for `render -a 0`:
 * WIDTH/HEIGHT/FOCAL_LENGTH (variables here, macros there), SHADING_ENABLE
 * (both variants are compiled), AO_ENABLE undefined.  AO_METHOD=1 because the
 * UNIFORM branch writes to a const float4 (intersect_kernel.cl:225-234) and is
 * rejected by a conforming compiler.
 */
#include <pthread.h>
#include <sched.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "ref_shim/cl_compat.h"
#include "rt_oracle.h" /* orc_scene: the five upload arrays as raw pointers */

thread_local uint cl_compat_gid[2];

static int g_width, g_height;
static float g_focal;
#define WIDTH g_width
#define HEIGHT g_height
#define FOCAL_LENGTH g_focal
/* AO parameters that the kernel only uses inside expressions can be variables; AO_NUM_SAMPLES and AO_METHOD
 * steer the preprocessor (intersect_kernel.cl:218,257,305), so every (method, samples) pair the tests use is
 * its own compiled copy of the kernel text. */
static float g_ao_max_distance = .2f;
static int g_ao_alpha_min = 4, g_ao_alpha_max = 90;
#define AO_MAX_DISTANCE g_ao_max_distance
#define AO_ALPHA_MIN g_ao_alpha_min
#define AO_ALPHA_MAX g_ao_alpha_max

#define inline static inline
#define AO_METHOD 1
#define AO_NUM_SAMPLES 0
namespace shaded {
#define SHADING_ENABLE
#include "kernel_as_cpp.inc"
#undef SHADING_ENABLE
}
namespace flat {
#include "kernel_as_cpp.inc"
}
#undef AO_METHOD
#undef AO_NUM_SAMPLES
#define AO_ENABLE
#define SHADING_ENABLE
#define AO_METHOD 0
#define AO_NUM_SAMPLES 1
namespace ao_m0_n1 {
#include "kernel_as_cpp.inc"
}
#undef AO_METHOD
#undef AO_NUM_SAMPLES
#define AO_METHOD 0
#define AO_NUM_SAMPLES 2
namespace ao_m0_n2 {
#include "kernel_as_cpp.inc"
}
#undef AO_METHOD
#undef AO_NUM_SAMPLES
#define AO_METHOD 0
#define AO_NUM_SAMPLES 3
namespace ao_m0_n3 {
#include "kernel_as_cpp.inc"
}
#undef AO_METHOD
#undef AO_NUM_SAMPLES
#define AO_METHOD 0
#define AO_NUM_SAMPLES 4
namespace ao_m0_n4 {
#include "kernel_as_cpp.inc"
}
#undef AO_METHOD
#undef AO_NUM_SAMPLES
#define AO_METHOD 1
#define AO_NUM_SAMPLES 1
namespace ao_m1_n1 {
#include "kernel_as_cpp.inc"
}
#undef AO_METHOD
#undef AO_NUM_SAMPLES
#define AO_METHOD 1
#define AO_NUM_SAMPLES 2
namespace ao_m1_n2 {
#include "kernel_as_cpp.inc"
}
#undef AO_METHOD
#undef AO_NUM_SAMPLES
#define AO_METHOD 1
#define AO_NUM_SAMPLES 3
namespace ao_m1_n3 {
#include "kernel_as_cpp.inc"
}
#undef AO_METHOD
#undef AO_NUM_SAMPLES
#define AO_METHOD 1
#define AO_NUM_SAMPLES 4
namespace ao_m1_n4 {
#include "kernel_as_cpp.inc"
}
#undef AO_METHOD
#undef AO_NUM_SAMPLES
#undef AO_ENABLE
#undef SHADING_ENABLE
#undef inline

typedef void (*ref_kernel_fn)(const uint *, const uint *, const float4 *, const float4 *, const float4 *, float *);
static ref_kernel_fn ao_kernel(int method, int samples)
{
	if (method == 0 && samples == 1) return ao_m0_n1::intersect;
	if (method == 0 && samples == 2) return ao_m0_n2::intersect;
	if (method == 0 && samples == 3) return ao_m0_n3::intersect;
	if (method == 0 && samples == 4) return ao_m0_n4::intersect;
	if (method == 1 && samples == 1) return ao_m1_n1::intersect;
	if (method == 1 && samples == 2) return ao_m1_n2::intersect;
	if (method == 1 && samples == 3) return ao_m1_n3::intersect;
	if (method == 1 && samples == 4) return ao_m1_n4::intersect;
	return nullptr;
}

#include "bvh.h"
#include "compiler_options.h"
#include "mesh.h"
#include "ray_tracer.h"

struct ref_scene {
	Mesh mesh;
	std::vector<uint32_t> orig_faces;   /* mesh.faces before the leaf-order sort */
	std::vector<uint32_t> sorted_faces; /* render.cc:88-95 */
	std::vector<uint32_t> triangles;    /* leaf index -> OFF face id (bvh.cc:125) */
	std::vector<uint32_t> nodes;
	std::vector<Vec3f> aabbs;
};

static void zero_pad_lanes(std::vector<Vec3f> &v)
{
	/* Vec3f::fourth is not initialised by the 3-argument ctor nor copied by the
	 * copy ctor (vec3.h:15-26); define it as 0. */
	static_assert(sizeof(Vec3f) == 16, "Vec3f must be float4-compatible");
	for (auto &e : v) reinterpret_cast<float *>(&e)[3] = 0.0f;
}

static ref_scene *finish_scene(ref_scene *s, int bvh_method)
{
	compute_vertex_normals(&s->mesh);                      /* render.cc:56 */
	BVH bvh(bvh_method == 1 ? BVH::Method::SURFACE_AREA_HEURISTIC : BVH::Method::CUT_LONGEST_AXIS);
	bvh.buildBVH(s->mesh);                                 /* render.cc:76-80 */
	s->orig_faces = s->mesh.faces;
	s->sorted_faces.reserve(s->mesh.faces.size());
	for (std::size_t i = 0; i < bvh.triangles.size(); ++i) { /* render.cc:88-95 */
		const uint32_t faceID = bvh.triangles[i] * 3;
		s->sorted_faces.push_back(s->mesh.faces[faceID]);
		s->sorted_faces.push_back(s->mesh.faces[faceID + 1]);
		s->sorted_faces.push_back(s->mesh.faces[faceID + 2]);
	}
	s->triangles.swap(bvh.triangles);
	s->nodes.swap(bvh.nodes);
	s->aabbs.swap(bvh.aabbs);
	zero_pad_lanes(s->mesh.vertices);
	zero_pad_lanes(s->mesh.vnormals);
	zero_pad_lanes(s->aabbs);
	return s;
}

static int affinity_cpus(std::vector<int> &cpus)
{
	cpu_set_t set;
	if (sched_getaffinity(0, sizeof set, &set) == 0)
		for (int i = 0; i < CPU_SETSIZE; ++i)
			if (CPU_ISSET(i, &set)) cpus.push_back(i);
	return (int)cpus.size();
}

template <typename F>
static void parallel_rows(unsigned nitems, int nthreads, F &&body)
{
	std::vector<int> cpus;
	const int ncpu = affinity_cpus(cpus);
	if (nthreads <= 0) nthreads = ncpu > 0 ? ncpu : 1;
	std::atomic<unsigned> next{0};
	auto work = [&](int t, bool pin) {
		if (pin && ncpu > 0) {                             /* SURVEY App. B.3 */
			cpu_set_t one;
			CPU_ZERO(&one);
			CPU_SET(cpus[t % ncpu], &one);
			pthread_setaffinity_np(pthread_self(), sizeof one, &one);
		}
		for (;;) {
			const unsigned item = next.fetch_add(1, std::memory_order_relaxed);
			if (item >= nitems) break;
			body(item);
		}
	};
	if (nthreads == 1) { work(0, false); return; }
	std::vector<std::thread> th;
	for (int t = 0; t < nthreads; ++t) th.emplace_back(work, t, true);
	for (auto &t : th) t.join();
}

extern "C" {

ref_scene *ref_scene_from_off(const char *path, int bvh_method)
{
	ref_scene *s = new ref_scene;
	try {
		load_off_mesh(path, &s->mesh);                     /* render.cc:55 */
		if (s->mesh.faces.empty()) { delete s; return nullptr; }
		return finish_scene(s, bvh_method);
	} catch (const std::exception &e) {
		std::fprintf(stderr, "ref_scene_from_off: %s\n", e.what());
		delete s;
		return nullptr;
	}
}

ref_scene *ref_scene_from_mesh(const float *verts3, size_t nverts, const uint32_t *faces, size_t nfaces, int bvh_method)
{
	if (!verts3 || !faces || nfaces == 0) return nullptr;
	ref_scene *s = new ref_scene;
	s->mesh.vertices.reserve(nverts);
	for (size_t i = 0; i < nverts; ++i)
		s->mesh.vertices.push_back(Vec3f(verts3[3 * i], verts3[3 * i + 1], verts3[3 * i + 2]));
	s->mesh.faces.assign(faces, faces + 3 * nfaces);
	return finish_scene(s, bvh_method);
}

void ref_scene_free(ref_scene *s) { delete s; }

/* counts: [0] face indices, [1] nodes, [2] aabb vectors, [3] vertices, [4] normals */
void ref_scene_counts(const ref_scene *s, size_t out[5])
{
	out[0] = s->sorted_faces.size();
	out[1] = s->nodes.size();
	out[2] = s->aabbs.size();
	out[3] = s->mesh.vertices.size();
	out[4] = s->mesh.vnormals.size();
}
const uint32_t *ref_scene_faces(const ref_scene *s) { return s->sorted_faces.data(); }
const uint32_t *ref_scene_orig_faces(const ref_scene *s) { return s->orig_faces.data(); }
const uint32_t *ref_scene_triangles(const ref_scene *s) { return s->triangles.data(); }
const uint32_t *ref_scene_nodes(const ref_scene *s) { return s->nodes.data(); }
const float *ref_scene_aabbs(const ref_scene *s) { return reinterpret_cast<const float *>(s->aabbs.data()); }
const float *ref_scene_vertices(const ref_scene *s) { return reinterpret_cast<const float *>(s->mesh.vertices.data()); }
const float *ref_scene_normals(const ref_scene *s) { return reinterpret_cast<const float *>(s->mesh.vnormals.data()); }

/* The reference kernel (`intersect`, intersect_kernel.cl:278-310) over the
 * NDRange width x height; rows [row_begin,row_end) step row_step. */
int ref_render(const orc_scene *sc, unsigned width, unsigned height, float focal_length, int shading,
               unsigned row_begin, unsigned row_end, unsigned row_step, float *image, int nthreads)
{
	if (!sc || !image || width == 0 || height == 0) return -1;
	if (row_step == 0) row_step = 1;
	if (row_end > height) row_end = height;
	if (row_begin >= row_end) return 0;
	g_width = (int)width;
	g_height = (int)height;
	g_focal = focal_length;
	const float4 *aabbs = reinterpret_cast<const float4 *>(sc->aabbs);
	const float4 *verts = reinterpret_cast<const float4 *>(sc->vertices);
	const float4 *norms = reinterpret_cast<const float4 *>(sc->normals);
	const unsigned nitems = (row_end - row_begin + row_step - 1) / row_step;
	parallel_rows(nitems, nthreads, [&](unsigned item) {
		const unsigned y = row_begin + item * row_step;
		for (unsigned x = 0; x < width; ++x) {
			cl_compat_gid[0] = x;
			cl_compat_gid[1] = y;
			if (shading) shaded::intersect(sc->faces, sc->nodes, aabbs, verts, norms, image);
			else flat::intersect(sc->faces, sc->nodes, aabbs, verts, norms, image);
		}
	});
	return 0;
}

/* The reference kernel with ambient occlusion (AO_ENABLE, intersect_kernel.cl:214-277, 305-307).
 * method 0 = uniform rings, 1 = random hemisphere; samples in 1..4 (compiled copies).  Returns -2 for a
 * (method, samples) pair that was not compiled. */
int ref_render_ao(const orc_scene *sc, unsigned width, unsigned height, float focal_length, int method, int samples,
                  float ao_max_distance, int alpha_min, int alpha_max,
                  unsigned row_begin, unsigned row_end, unsigned row_step, float *image, int nthreads)
{
	if (!sc || !image || width == 0 || height == 0) return -1;
	const ref_kernel_fn fn = ao_kernel(method, samples);
	if (!fn) return -2;
	if (row_step == 0) row_step = 1;
	if (row_end > height) row_end = height;
	if (row_begin >= row_end) return 0;
	g_width = (int)width;
	g_height = (int)height;
	g_focal = focal_length;
	g_ao_max_distance = ao_max_distance;
	g_ao_alpha_min = alpha_min;
	g_ao_alpha_max = alpha_max;
	const float4 *aabbs = reinterpret_cast<const float4 *>(sc->aabbs);
	const float4 *verts = reinterpret_cast<const float4 *>(sc->vertices);
	const float4 *norms = reinterpret_cast<const float4 *>(sc->normals);
	const unsigned nitems = (row_end - row_begin + row_step - 1) / row_step;
	parallel_rows(nitems, nthreads, [&](unsigned item) {
		const unsigned y = row_begin + item * row_step;
		for (unsigned x = 0; x < width; ++x) {
			cl_compat_gid[0] = x;
			cl_compat_gid[1] = y;
			fn(sc->faces, sc->nodes, aabbs, verts, norms, image);
		}
	});
	return 0;
}

/* The reference's scene_intersect (intersect_kernel.cl:184-213) per ray. */
int ref_trace_rays(const orc_scene *sc, const float *origins, const float *dirs, size_t nrays, float max_distance,
                   uint32_t *face_id, float *distance, int nthreads)
{
	if (!sc || (nrays && (!origins || !dirs))) return -1;
	const float4 *aabbs = reinterpret_cast<const float4 *>(sc->aabbs);
	const float4 *verts = reinterpret_cast<const float4 *>(sc->vertices);
	const float4 *norms = reinterpret_cast<const float4 *>(sc->normals);
	const size_t chunk = 4096;
	const unsigned nitems = (unsigned)((nrays + chunk - 1) / chunk);
	parallel_rows(nitems, nthreads, [&](unsigned item) {
		const size_t b = (size_t)item * chunk, e = b + chunk < nrays ? b + chunk : nrays;
		for (size_t k = b; k < e; ++k) {
			shaded::Intersection isect;
			isect.distance = INFINITY;
			const float4 o(origins[4 * k], origins[4 * k + 1], origins[4 * k + 2], 0.0f);
			const float4 d(dirs[4 * k], dirs[4 * k + 1], dirs[4 * k + 2], 0.0f);
			const bool hit = shaded::scene_intersect(sc->nodes, aabbs, sc->faces, verts, norms, o, d, &isect, max_distance);
			if (face_id) face_id[k] = hit ? isect.face_id : 0xffffffffu;
			if (distance) distance[k] = hit ? isect.distance : INFINITY;
		}
	});
	return 0;
}

/* Primary-ray hit ids/distances: the kernel's own ray set-up (:284-295)
 * cannot be called in isolation, so this repeats those four statements and
 * then calls the reference's scene_intersect. */
int ref_primary_hits(const orc_scene *sc, unsigned width, unsigned height, float focal_length,
                     uint32_t *face_id, float *distance, int nthreads)
{
	if (!sc || width == 0 || height == 0) return -1;
	const float4 *aabbs = reinterpret_cast<const float4 *>(sc->aabbs);
	const float4 *verts = reinterpret_cast<const float4 *>(sc->vertices);
	const float4 *norms = reinterpret_cast<const float4 *>(sc->normals);
	const int W = (int)width, H = (int)height;
	parallel_rows(height, nthreads, [&](unsigned y) {
		for (unsigned x = 0; x < width; ++x) {
			const float4 camera_position = mk4(0.0f, 0.0f, 2.0f, 0.0f);
			const float a = focal_length * max(W, H);
			const float4 ray_dir = normalize(mk4(((float)x + 0.5f) / a - W / (2.0f * a), -(((float)y + 0.5f) / a - H / (2.0f * a)), -1.0f, 0.0f));
			shaded::Intersection isect;
			isect.distance = INFINITY;
			const bool hit = shaded::scene_intersect(sc->nodes, aabbs, sc->faces, verts, norms, camera_position, ray_dir, &isect, 100000.0f);
			const size_t idx = (size_t)y * width + x;
			if (face_id) face_id[idx] = hit ? isect.face_id : 0xffffffffu;
			if (distance) distance[idx] = hit ? isect.distance : INFINITY;
		}
	});
	return 0;
}

/* RayTracer::resize (src/ray_tracer.cc:3-15) through the reference's own class. */
void ref_resize(unsigned width, unsigned height, unsigned n_super_samples, float *tmp, unsigned char *image)
{
	RayTracer::Options o{ width, height, 1.f, n_super_samples, true, false, .2f, 0,
	                      RayTracer::AmbientOcclusionMethod::RANDOM, 4, 90, BVH::Method::CUT_LONGEST_AXIS };
	RayTracer rt(o);
	rt.resize(tmp, image);
}

/* totalWidth/totalHeight (include/ray_tracer.h:33-34) */
void ref_total_dims(unsigned width, unsigned height, unsigned n_super_samples, unsigned *tw, unsigned *th)
{
	RayTracer::Options o{ width, height, 1.f, n_super_samples, true, false, .2f, 0,
	                      RayTracer::AmbientOcclusionMethod::RANDOM, 4, 90, BVH::Method::CUT_LONGEST_AXIS };
	RayTracer rt(o);
	*tw = rt.totalWidth;
	*th = rt.totalHeight;
}

/* FOCAL_LENGTH as the kernel sees it: CompilerOptions::add(float)
 * (include/compiler_options.h:13-19) then the literal parser. */
float ref_focal_roundtrip(float focal)
{
	CompilerOptions co;
	co.add("F", focal);
	const std::string s = co.str(); /* "-DF=<digits>[.]f " */
	return std::strtof(s.c_str() + 4, nullptr);
}

int ref_online_cpus(void)
{
	std::vector<int> cpus;
	const int n = affinity_cpus(cpus);
	return n > 0 ? n : 1;
}

} /* extern "C" */
