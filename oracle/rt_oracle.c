/*
 * rt_oracle.c -- plain-C restatement of the reference's closest-hit path.
 * TEST INFRASTRUCTURE ONLY (see rt_oracle.h).  Build: oracle/Makefile
 *   gcc -O2 -ffp-contract=off -fno-fast-math  (x86-64 SSE2 float, no x87)
 *
 * Every function cites the lines of /root/reference it follows.  Arithmetic
 * is written one rounding per operation, in the order the kernel text spells
 * it; float4 values keep their w lane (always 0 here) so the
 * 4-lane dot/length of OpenCL C is reproduced literally.
 */
#define _GNU_SOURCE
#include "rt_oracle.h"

#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float x, y, z, w; } f4;

static inline f4 f4_make(float x, float y, float z, float w) { f4 r = { x, y, z, w }; return r; }
static inline f4 f4_load(const float *p) { f4 r = { p[0], p[1], p[2], p[3] }; return r; }
static inline f4 f4_load_w0(const float *p) { f4 r = { p[0], p[1], p[2], 0.0f }; return r; }
static inline f4 f4_add(f4 a, f4 b) { return f4_make(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
static inline f4 f4_sub(f4 a, f4 b) { return f4_make(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
static inline f4 f4_mul(f4 a, f4 b) { return f4_make(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
static inline f4 f4_scale(float s, f4 a) { return f4_make(s * a.x, s * a.y, s * a.z, s * a.w); }
/* OpenCL C dot(float4,float4): the pinned order is ((x+y)+z)+w. */
static inline float f4_dot(f4 a, f4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
/* OpenCL C cross(float4,float4): w component is 0. */
static inline f4 f4_cross(f4 a, f4 b)
{
	return f4_make(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x, 0.0f);
}
static inline float f4_length(f4 a) { return sqrtf(f4_dot(a, a)); }
static inline f4 f4_normalize(f4 a)
{
	const float l = f4_length(a);
	return f4_make(a.x / l, a.y / l, a.z / l, a.w / l);
}
/* OpenCL C max/min on floats: max(x,y) = x < y ? y : x; min(x,y) = y < x ? y : x. */
static inline float cl_max(float a, float b) { return a < b ? b : a; }
static inline float cl_min(float a, float b) { return b < a ? b : a; }
/* OpenCL C clamp(x,lo,hi) = fmin(fmax(x,lo),hi). */
static inline float cl_clamp(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

/* intersect_kernel.cl:7-12 */
typedef struct {
	uint32_t face_id;
	f4 barycentric;
	f4 position;
	float distance;
} isect_t;

/* include/compiler_options.h:13-19 + opencl_host.cc:45: "-DFOCAL_LENGTH=" <<
 * v, then '.' if v is integral, then 'f'.  ostream's default float format is
 * %g with precision 6; the OpenCL front end then parses the literal. */
float orc_focal_roundtrip(float focal)
{
	char buf[64];
	snprintf(buf, sizeof buf, "%g", (double)focal);
	return strtof(buf, NULL);
}

/* intersect_kernel.cl:21-61.  Slab test; three separate divides; NaN never
 * rejects because every rejection is an `a > b` comparison. */
static inline int aabb_intersect(const float *bb, f4 ray_pos, f4 ray_dir, float max_distance)
{
	const float *lo = bb, *hi = bb + 4;
	float t_min, t_max, ty_min, ty_max, tz_min, tz_max;
	float div = 1.0f / ray_dir.x;                         /* :23 */
	if (div >= 0) {                                       /* :24-31 */
		t_min = (lo[0] - ray_pos.x) * div;
		t_max = (hi[0] - ray_pos.x) * div;
	} else {
		t_min = (hi[0] - ray_pos.x) * div;
		t_max = (lo[0] - ray_pos.x) * div;
	}
	div = 1 / ray_dir.y;                                  /* :32 */
	if (div >= 0) {                                       /* :33-40 */
		ty_min = (lo[1] - ray_pos.y) * div;
		ty_max = (hi[1] - ray_pos.y) * div;
	} else {
		ty_min = (hi[1] - ray_pos.y) * div;
		ty_max = (lo[1] - ray_pos.y) * div;
	}
	if (t_min > ty_max || ty_min > t_max)                 /* :41-43 */
		return 0;
	t_min = cl_max(t_min, ty_min);                        /* :44 */
	t_max = cl_min(t_max, ty_max);                        /* :45 */
	div = 1 / ray_dir.z;                                  /* :46 */
	if (div >= 0) {                                       /* :47-54 */
		tz_min = (lo[2] - ray_pos.z) * div;
		tz_max = (hi[2] - ray_pos.z) * div;
	} else {
		tz_min = (hi[2] - ray_pos.z) * div;
		tz_max = (lo[2] - ray_pos.z) * div;
	}
	if (t_min > tz_max || tz_min > t_max)                 /* :55-57 */
		return 0;
	t_min = cl_max(t_min, tz_min);                        /* :58 */
	t_max = cl_min(t_max, tz_max);                        /* :59 */
	return t_min < max_distance && t_max > 0;             /* :60 */
}

int orc_aabb_intersect(const float *bb, const float *pos4, const float *dir4, float max_distance)
{
	return aabb_intersect(bb, f4_load(pos4), f4_load(dir4), max_distance);
}

/* intersect_kernel.cl:65-114.  geomalgorithms plane + parametric test. */
static inline int triangle_intersect(f4 ta, f4 tb, f4 tc, uint32_t face_id, f4 ray_pos, f4 ray_dir, isect_t *isect)
{
	const float EPSILON2 = 0.000001f;                     /* :66 */
	const f4 u = f4_sub(tb, ta);                          /* :68 */
	const f4 v = f4_sub(tc, ta);                          /* :69 */
	const f4 n = f4_cross(u, v);                          /* :70 */
	const f4 w0 = f4_sub(ray_pos, ta);                    /* :71 */
	const float a = -f4_dot(n, w0);                       /* :72 */
	const float b = f4_dot(n, ray_dir);                   /* :73 */
	if (fabsf(b) < EPSILON2)                              /* :75-77 */
		return 0;
	const float r = a / b;                                /* :79 */
	if ((double)r < 0.0)                                  /* :80 (double literal) */
		return 0;
	const f4 ip = f4_add(ray_pos, f4_scale(r, ray_dir));  /* :85 */
	const float uu = f4_dot(u, u);                        /* :87 */
	const float uv = f4_dot(u, v);                        /* :88 */
	const float vv = f4_dot(v, v);                        /* :89 */
	const f4 w = f4_sub(ip, ta);                          /* :90 */
	const float wu = f4_dot(u, w);                        /* :91 */
	const float wv = f4_dot(w, v);                        /* :92 */
	const float D = uv * uv - uu * vv;                    /* :93 */
	const float s = (uv * wv - vv * wu) / D;              /* :95 */
	if (s < -0.00001f || (double)s > 1.00001)             /* :96 (1.00001 is a double literal) */
		return 0;
	const float t = (uv * wu - uu * wv) / D;              /* :100 */
	if (t < -0.00001f || (double)(s + t) > 1.00001)       /* :101 */
		return 0;
	const float distance = f4_length(f4_sub(ip, ray_pos));/* :106 */
	if (isect->distance > distance) {                     /* :107-112 strict: first in leaf order wins ties */
		isect->face_id = face_id;
		isect->barycentric = f4_make(1.0f - s - t, s, t, 0);
		isect->position = ip;
		isect->distance = distance;
	}
	return 1;
}

/* intersect_kernel.cl:184-213.  Stackless pre-order walk with subtree skip. */
static inline int scene_intersect(const orc_scene *sc, f4 ray_pos, f4 ray_dir, isect_t *isect,
                                  float max_distance, orc_counters *c)
{
	const uint32_t *nodes = sc->nodes, *faces = sc->faces;
	const float *aabbs = sc->aabbs, *vertices = sc->vertices;
	int is_intersecting = 0;
	uint32_t triangle_index = 0;
	uint64_t visits = 0;
	for (uint32_t i = 0; i < nodes[0];) {                 /* :187 */
		const uint32_t node_count = nodes[i];             /* :188 */
		++visits;
		if (!aabb_intersect(aabbs + ((size_t)i << 3), ray_pos, ray_dir, max_distance)) { /* :189 */
			triangle_index += (node_count + 1) >> 1;      /* :191 */
			i += node_count;                              /* :192 */
		} else {
			if (c) c->box_hits++;
			if (node_count == 1) {                        /* :195 */
				const uint32_t face_id = triangle_index * 3; /* :197 */
				const int hit = triangle_intersect(
					f4_load(vertices + 4 * (size_t)faces[face_id + 0]),
					f4_load(vertices + 4 * (size_t)faces[face_id + 1]),
					f4_load(vertices + 4 * (size_t)faces[face_id + 2]),
					face_id, ray_pos, ray_dir, isect);    /* :198-206 */
				is_intersecting |= hit;
				if (c) { c->tri_tests++; c->tri_hits += (uint64_t)hit; }
				++triangle_index;                         /* :207 */
			}
			++i;                                          /* :209 */
		}
	}
	if (c) {
		c->rays++;
		c->node_visits += visits;
		c->hit_rays += (uint64_t)is_intersecting;
		if (visits > c->max_visits) c->max_visits = visits;
	}
	return is_intersecting;
}

/* intersect_kernel.cl:115-117 */
static inline float shade(f4 ray_dir, f4 normal)
{
	return cl_clamp(-f4_dot(normal, ray_dir), 0.f, 1.f);
}

/* intersect_kernel.cl:118-127 */
static inline f4 get_smooth_normal(const orc_scene *sc, const isect_t *isect)
{
	const uint32_t v0 = sc->faces[isect->face_id + 0];
	const uint32_t v1 = sc->faces[isect->face_id + 1];
	const uint32_t v2 = sc->faces[isect->face_id + 2];
	const f4 bx = f4_make(isect->barycentric.x, isect->barycentric.x, isect->barycentric.x, isect->barycentric.x);
	const f4 by = f4_make(isect->barycentric.y, isect->barycentric.y, isect->barycentric.y, isect->barycentric.y);
	const f4 bz = f4_make(isect->barycentric.z, isect->barycentric.z, isect->barycentric.z, isect->barycentric.z);
	return f4_normalize(f4_add(f4_add(
		f4_mul(f4_load(sc->normals + 4 * (size_t)v0), bx),
		f4_mul(f4_load(sc->normals + 4 * (size_t)v1), by)),
		f4_mul(f4_load(sc->normals + 4 * (size_t)v2), bz)));
}

/* ---------------- ambient occlusion (intersect_kernel.cl:128-183, 214-277) ---------------- */

/* Transcendentals: OpenCL leaves their last bits to the device; the oracle pins them as "evaluate in double,
 * round once to float" (same definition in ref_shim/cl_compat.h and in the CUDA kernels). */
static inline float t_sin(float x) { return (float)sin((double)x); }
static inline float t_cos(float x) { return (float)cos((double)x); }
static inline float t_acos(float x) { return (float)acos((double)x); }
static inline float t_cospi(float x) { return (float)cos(3.14159265358979323846 * (double)x); }
static inline float t_sinpi(float x) { return (float)sin(3.14159265358979323846 * (double)x); }

static inline f4 f4_muls(f4 a, float s) { return f4_make(a.x * s, a.y * s, a.z * s, a.w * s); }

/* :128-135 xorshift128 */
static inline uint32_t random_int(uint32_t v[4])
{
	const uint32_t t = v[0] ^ (v[0] << 11u);
	v[0] = v[1]; v[1] = v[2]; v[2] = v[3];
	return v[3] = v[3] ^ (v[3] >> 19u) ^ (t ^ (t >> 8u));
}
/* :136-142 */
static inline void random_initialize_seed(uint32_t v[4], uint32_t seed)
{
	v[0] = (123456789u ^ seed) * 88675123u;
	v[1] = (362436069u ^ seed) * 123456789u;
	v[2] = (521288629u ^ seed) * 362436069u;
	v[3] = (88675123u ^ seed) * 521288629u;
	random_int(v);
}
/* :150-152 */
static inline float random_float(uint32_t v[4]) { return 2.32830643653869629E-10f * (float)random_int(v); }

/* the "replace the smallest component by 1" step shared by :156-164 and :226-234 */
static inline f4 perturb_smallest(f4 h)
{
	if (fabsf(h.x) <= fabsf(h.y) && fabsf(h.x) <= fabsf(h.z)) h.x = 1.0f;
	else if (fabsf(h.y) <= fabsf(h.x) && fabsf(h.y) <= fabsf(h.z)) h.y = 1.0f;
	else if (fabsf(h.z) <= fabsf(h.x) && fabsf(h.z) <= fabsf(h.y)) h.z = 1.0f;
	return h;
}

static inline int any_hit(const orc_scene *sc, f4 p, f4 dir, float max_distance, orc_counters *c)
{
	isect_t scratch;                       /* :249,:262,:269 pass an uninitialised Intersection; only the bool is used */
	memset(&scratch, 0, sizeof scratch);
	scratch.distance = INFINITY;
	return scene_intersect(sc, p, dir, &scratch, max_distance, c);
}

/* :214-277 */
static float ambient_occlusion(const orc_scene *sc, f4 point, f4 normal, uint32_t index, const orc_ao *ao, orc_counters *c)
{
	const f4 p = f4_add(point, f4_muls(normal, 1.0f / 100000.0f));                 /* :215 */
	uint32_t hits = 0;
	const float max_distance = ao->max_distance;                                    /* :217 */
	if (ao->method == 0) {                                                          /* AO_METHOD_UNIFORM :218-256 */
		uint32_t n = 0;
		const uint32_t circle_count = ao->samples;
		const float degrees = (float)(M_PI / 180);                                  /* :221 (M_PI is a double here) */
		const float alpha_min = (float)ao->alpha_min * degrees;
		const float alpha_max = (float)ao->alpha_max * degrees;
		const f4 basis_y = normal;
		const f4 h = perturb_smallest(basis_y);                                     /* :225-234 */
		const f4 basis_x = f4_normalize(f4_cross(h, basis_y));
		const f4 basis_z = f4_normalize(f4_cross(basis_x, basis_y));
		for (uint32_t cc = 0; cc < circle_count; ++cc) {
			const float step = alpha_max / (float)circle_count;                     /* :238 */
			const float angle = (step * (float)cc) + alpha_min;                     /* :239 */
			const uint32_t ray_count = (uint32_t)(((double)2.0f * M_PI * (double)t_cos(angle)) / (double)step); /* :240 */
			const float theta = (float)(M_PI_2 - (double)angle);                    /* :241 */
			for (uint32_t cr = 0; cr <= ray_count; ++cr) {                          /* :242 (<=) */
				const float phi = (float)(((double)2.0f * M_PI * (double)cr) / (double)ray_count); /* :243 */
				const float xs = t_sin(theta) * t_cospi(phi);
				const float ys = t_cos(theta);
				const float zs = t_sin(theta) * t_sinpi(phi);
				const f4 ray_dir = f4_add(f4_add(f4_muls(basis_x, xs), f4_muls(basis_y, ys)), f4_muls(basis_z, zs)); /* :248 */
				++n;
				if (any_hit(sc, p, ray_dir, max_distance, c)) ++hits;
			}
		}
		return 1.0f - ((float)hits / (float)n);                                     /* :256 */
	}
	/* AO_METHOD_RANDOM :257-275, sampler :153-183 */
	const f4 basis_y = f4_normalize(normal);
	const f4 h = perturb_smallest(basis_y);
	const f4 basis_x = f4_normalize(f4_cross(h, basis_y));
	const f4 basis_z = f4_normalize(f4_cross(basis_x, basis_y));
	uint32_t rng[4];
	random_initialize_seed(rng, 536870923u * index);                                /* :169 */
	uint32_t n = ao->samples;
	++n;                                                                            /* :263 */
	if (any_hit(sc, p, normal, max_distance, c)) ++hits;                            /* :264 */
	for (uint32_t i = 0; i < n; ++i) {                                              /* :267 */
		const float xi1 = random_float(rng);
		const float xi2 = random_float(rng);
		const float theta = t_acos(sqrtf(1.0f - xi1));                              /* :175 */
		const float phi = (float)(2.0 * (double)xi2);                               /* :177 */
		const float xs = t_sin(theta) * t_cospi(phi);
		const float ys = t_cos(theta);
		const float zs = t_sin(theta) * t_sinpi(phi);
		const f4 direction = f4_add(f4_add(f4_muls(basis_x, xs), f4_muls(basis_y, ys)), f4_muls(basis_z, zs));
		if (any_hit(sc, p, f4_normalize(direction), max_distance, c)) ++hits;
	}
	return 1.0f - ((float)hits / (float)n);                                         /* :275 */
}

/* ---- extensions used only by configs C3 (jitter) and C5 (random rays) ---- */

static inline uint32_t mix32(uint32_t x)
{
	x ^= x >> 16; x *= 0x7feb352du;
	x ^= x >> 15; x *= 0x846ca68bu;
	x ^= x >> 16;
	return x;
}
static inline float u01(uint32_t h) { return (float)(h >> 8) * 5.9604644775390625e-08f; /* 2^-24 */ }

void orc_jitter(uint32_t seed, uint32_t x, uint32_t y, float *jx, float *jy)
{
	const uint32_t h1 = mix32(mix32(x ^ seed) + y);
	const uint32_t h2 = mix32(h1 + 0x9e3779b9u);
	*jx = u01(h1);
	*jy = u01(h2);
}

void orc_gen_random_rays(uint32_t seed, uint64_t first, size_t n, const float *bbmin3, const float *bbmax3,
                         float *origins, float *dirs)
{
	for (size_t k = 0; k < n; ++k) {
		const uint64_t id = first + k;
		uint32_t h = mix32((uint32_t)id ^ seed);
		h = mix32(h + (uint32_t)(id >> 32) + 0x9e3779b9u);
		float *o = origins + 4 * k, *d = dirs + 4 * k;
		for (int a = 0; a < 3; ++a) {
			h = mix32(h + 0x9e3779b9u);
			/* root box shrunk by 1 % about its centre */
			const float ext = bbmax3[a] - bbmin3[a];
			const float lo = bbmin3[a] + 0.005f * ext;
			o[a] = lo + u01(h) * (0.99f * ext);
		}
		o[3] = 0.0f;
		/* Marsaglia (1972): uniform direction without trigonometry. */
		for (;;) {
			h = mix32(h + 0x9e3779b9u);
			const float p = 2.0f * u01(h) - 1.0f;
			h = mix32(h + 0x9e3779b9u);
			const float q = 2.0f * u01(h) - 1.0f;
			const float s = p * p + q * q;
			if (s >= 1.0f) continue;
			const float f = 2.0f * sqrtf(1.0f - s);
			d[0] = p * f; d[1] = q * f; d[2] = 1.0f - 2.0f * s; d[3] = 0.0f;
			break;
		}
	}
}

/* intersect_kernel.cl:278-310 for one pixel. */
static inline void pixel(const orc_scene *sc, int W, int H, float focal, int shading, uint32_t jitter_seed, const orc_ao *ao,
                         uint32_t x, uint32_t y, float *value, uint32_t *face_id, float *distance, orc_counters *c)
{
	const f4 camera_position = f4_make(0.0f, 0.0f, 2.0f, 0.0f);       /* :284 */
	const float a = focal * (float)(W < H ? H : W);                   /* :285 int max, then int->float */
	float jx = 0.5f, jy = 0.5f;
	if (jitter_seed) orc_jitter(jitter_seed, x, y, &jx, &jy);
	const f4 ray_dir = f4_normalize(f4_make(
		((float)x + jx) / a - (float)W / (2.0f * a),                  /* :287 */
		-(((float)y + jy) / a - (float)H / (2.0f * a)),               /* :288 */
		-1.0f, 0.0f));
	const float max_distance = 100000.0f;                             /* :292 */
	isect_t isect;
	memset(&isect, 0, sizeof isect);
	isect.distance = INFINITY;                                        /* :294 */
	const int hit = scene_intersect(sc, camera_position, ray_dir, &isect, max_distance, c); /* :295 */
	float v = 1.0f;                                                   /* :296 */
	if (!hit) {
		v = 0.0f;                                                     /* :297-299 */
	} else {
		const f4 normal = get_smooth_normal(sc, &isect);              /* :301 */
		if (shading) v = shade(ray_dir, normal);                      /* :302-304 */
		if (ao && ao->enable && ao->samples > 0)                      /* :305-307; AO rays are not counted */
			v *= ambient_occlusion(sc, isect.position, normal, y * (uint32_t)W + x, ao, NULL);
	}
	*value = v;                                                       /* :309 */
	if (face_id) *face_id = hit ? isect.face_id : ORC_NO_HIT;
	if (distance) *distance = hit ? isect.distance : INFINITY;
}

/* ------------------------- pinned worker pool --------------------------- */

int orc_online_cpus(void)
{
	cpu_set_t set;
	if (sched_getaffinity(0, sizeof set, &set) == 0) {
		const int n = CPU_COUNT(&set);
		if (n > 0) return n;
	}
	return 1;
}

typedef struct job {
	/* shared */
	const orc_scene *sc;
	volatile uint64_t next;      /* dynamic work counter */
	uint64_t nitems;             /* rows or ray chunks   */
	int mode;                    /* 0 render, 1 rays     */
	/* render */
	int W, H; float focal; int shading; uint32_t jitter_seed;
	const orc_ao *ao;
	unsigned row_begin, row_step;
	float *image; uint32_t *face_id; float *distance;
	/* rays */
	const float *origins, *dirs; size_t nrays; float max_distance;
	int want_counters;
} job;

typedef struct worker {
	job *jb;
	int cpu;
	orc_counters c;
	pthread_t th;
} worker;

#define RAY_CHUNK 4096u

static void *worker_main(void *arg)
{
	worker *w = (worker *)arg;
	job *jb = w->jb;
	if (w->cpu >= 0) {            /* un-pinned threads pile up on one CPU */
		cpu_set_t one;
		CPU_ZERO(&one);
		CPU_SET(w->cpu, &one);
		pthread_setaffinity_np(pthread_self(), sizeof one, &one);
	}
	orc_counters *c = jb->want_counters ? &w->c : NULL;
	for (;;) {
		const uint64_t item = __atomic_fetch_add(&jb->next, 1, __ATOMIC_RELAXED);
		if (item >= jb->nitems) break;
		if (jb->mode == 0) {
			const uint32_t y = jb->row_begin + (uint32_t)item * jb->row_step;
			for (uint32_t x = 0; x < (uint32_t)jb->W; ++x) {
				const size_t idx = (size_t)y * (size_t)jb->W + x;    /* :281 */
				pixel(jb->sc, jb->W, jb->H, jb->focal, jb->shading, jb->jitter_seed, jb->ao, x, y,
				      jb->image + idx,
				      jb->face_id ? jb->face_id + idx : NULL,
				      jb->distance ? jb->distance + idx : NULL, c);
			}
		} else {
			const size_t b = (size_t)item * RAY_CHUNK;
			const size_t e = b + RAY_CHUNK < jb->nrays ? b + RAY_CHUNK : jb->nrays;
			for (size_t k = b; k < e; ++k) {
				isect_t isect;
				memset(&isect, 0, sizeof isect);
				isect.distance = INFINITY;
				const int hit = scene_intersect(jb->sc, f4_load_w0(jb->origins + 4 * k), f4_load_w0(jb->dirs + 4 * k),
				                                &isect, jb->max_distance, c);
				if (jb->face_id) jb->face_id[k] = hit ? isect.face_id : ORC_NO_HIT;
				if (jb->distance) jb->distance[k] = hit ? isect.distance : INFINITY;
			}
		}
	}
	return NULL;
}

static int run_job(job *jb, int nthreads, orc_counters *out)
{
	cpu_set_t set;
	int cpus[CPU_SETSIZE], ncpu = 0;
	if (sched_getaffinity(0, sizeof set, &set) == 0)
		for (int i = 0; i < CPU_SETSIZE; ++i)
			if (CPU_ISSET(i, &set)) cpus[ncpu++] = i;
	if (nthreads <= 0) nthreads = ncpu > 0 ? ncpu : 1;
	worker *ws = (worker *)calloc((size_t)nthreads, sizeof *ws);
	if (!ws) return -1;
	for (int t = 0; t < nthreads; ++t) {
		ws[t].jb = jb;
		ws[t].cpu = ncpu > 0 ? cpus[t % ncpu] : -1;
	}
	if (nthreads == 1) {
		ws[0].cpu = -1;           /* caller's thread, caller's placement */
		worker_main(&ws[0]);
	} else {
		for (int t = 0; t < nthreads; ++t)
			if (pthread_create(&ws[t].th, NULL, worker_main, &ws[t]) != 0) { free(ws); return -1; }
		for (int t = 0; t < nthreads; ++t) pthread_join(ws[t].th, NULL);
	}
	if (out) {
		memset(out, 0, sizeof *out);
		for (int t = 0; t < nthreads; ++t) {
			out->rays += ws[t].c.rays;
			out->node_visits += ws[t].c.node_visits;
			out->box_hits += ws[t].c.box_hits;
			out->tri_tests += ws[t].c.tri_tests;
			out->tri_hits += ws[t].c.tri_hits;
			out->hit_rays += ws[t].c.hit_rays;
			if (ws[t].c.max_visits > out->max_visits) out->max_visits = ws[t].c.max_visits;
		}
	}
	free(ws);
	return 0;
}

int orc_render(const orc_scene *scene, unsigned width, unsigned height, float focal_length, int shading,
               uint32_t jitter_seed, unsigned row_begin, unsigned row_end, unsigned row_step,
               float *image, uint32_t *face_id, float *distance, orc_counters *counters, int nthreads)
{
	return orc_render_ao(scene, width, height, focal_length, shading, jitter_seed, NULL, row_begin, row_end, row_step,
	                     image, face_id, distance, counters, nthreads);
}

int orc_render_ao(const orc_scene *scene, unsigned width, unsigned height, float focal_length, int shading,
                  uint32_t jitter_seed, const orc_ao *ao, unsigned row_begin, unsigned row_end, unsigned row_step,
                  float *image, uint32_t *face_id, float *distance, orc_counters *counters, int nthreads)
{
	if (!scene || !image || !scene->nodes || scene->nnodes == 0 || width == 0 || height == 0) return -1;
	if (row_step == 0) row_step = 1;
	if (row_end > height) row_end = height;
	if (row_begin >= row_end) { if (counters) memset(counters, 0, sizeof *counters); return 0; }
	job jb;
	memset(&jb, 0, sizeof jb);
	jb.sc = scene; jb.mode = 0;
	jb.W = (int)width; jb.H = (int)height; jb.focal = focal_length; jb.shading = shading; jb.jitter_seed = jitter_seed;
	jb.ao = ao;
	jb.row_begin = row_begin; jb.row_step = row_step;
	jb.nitems = (row_end - row_begin + row_step - 1) / row_step;
	jb.image = image; jb.face_id = face_id; jb.distance = distance;
	jb.want_counters = counters != NULL;
	return run_job(&jb, nthreads, counters);
}

int orc_trace_rays(const orc_scene *scene, const float *origins, const float *dirs, size_t nrays, float max_distance,
                   uint32_t *face_id, float *distance, orc_counters *counters, int nthreads)
{
	if (!scene || !scene->nodes || scene->nnodes == 0) return -1;
	if (nrays == 0) { if (counters) memset(counters, 0, sizeof *counters); return 0; }
	if (!origins || !dirs) return -1;
	job jb;
	memset(&jb, 0, sizeof jb);
	jb.sc = scene; jb.mode = 1;
	jb.origins = origins; jb.dirs = dirs; jb.nrays = nrays; jb.max_distance = max_distance;
	jb.nitems = (nrays + RAY_CHUNK - 1) / RAY_CHUNK;
	jb.face_id = face_id; jb.distance = distance;
	jb.want_counters = counters != NULL;
	return run_job(&jb, nthreads, counters);
}

/* src/ray_tracer.cc:3-15: n x n box sum in (ssY, ssX) order, then
 * (total / (n*n)) * 255 with the implicit float -> unsigned char conversion.
 * n*n is an unsigned int converted to float for the division. */
void orc_resize(const float *tmp, unsigned total_width, unsigned width, unsigned height, unsigned n, unsigned char *image)
{
	for (unsigned y = 0; y < height; ++y) {
		for (unsigned x = 0; x < width; ++x) {
			float total = 0;
			for (unsigned ssY = 0; ssY < n; ++ssY)
				for (unsigned ssX = 0; ssX < n; ++ssX)
					total += tmp[(size_t)(y * n + ssY) * total_width + (x * n + ssX)];
			image[(size_t)y * width + x] = (unsigned char)((total / (float)(n * n)) * 255);
		}
	}
}
