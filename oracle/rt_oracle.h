/*
 * rt_oracle.h -- CPU oracle for the closest-hit path of magcks/opencl_raytracer.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library, and only as the checker / the timed
 * CPU baseline.  The product path (opencl_raytracer_b200/) never links it.
 *
 * What it is: a plain-C restatement of the reference's OpenCL kernel
 * (src/intersect_kernel.cl) with the arithmetic pinned to one legal OpenCL
 * behaviour: IEEE-754 binary32, round-to-nearest-even, no FMA contraction,
 * correctly rounded '/' and sqrt, float4 dot summed x,y,z,w left to right.
 * Parity pin: the reference ships no tests or golden vectors for this path
 * (SURVEY.md section 4), so the restatement is pinned against the reference's
 * own kernel text compiled for the CPU (oracle/_ref/libref_oracle.so, built by
 * oracle/Makefile from /root/reference) -- bit-exact on every pixel -- and
 * against the committed golden vectors in tests/golden/ generated from that
 * library.
 *
 * All vector arrays use the reference's 16-byte Vec3f/float4 stride
 * (include/vec3.h:93-94: three floats + one pad lane).
 */
#ifndef RT_ORACLE_H
#define RT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* The five arrays render.cc:98 hands to OpenCLHost::upload. */
typedef struct orc_scene {
	const uint32_t *faces;    /* leaf-ordered vertex ids, 3 per triangle (render.cc:88-95) */
	size_t          nfaceidx; /* = 3 * triangles */
	const uint32_t *nodes;    /* pre-order subtree sizes (bvh.cc:115-162) */
	size_t          nnodes;
	const float    *aabbs;    /* 2 float4 (min,max) per node */
	const float    *vertices; /* float4 per vertex */
	size_t          nverts;
	const float    *normals;  /* float4 per vertex */
} orc_scene;

/* Per-call traversal statistics under the reference's exhaustive walk. */
typedef struct orc_counters {
	uint64_t rays;
	uint64_t node_visits;   /* V: boxes slab-tested          */
	uint64_t box_hits;      /*    boxes that passed          */
	uint64_t tri_tests;     /* T: triangle_intersect calls   */
	uint64_t tri_hits;      /*    calls that returned true   */
	uint64_t hit_rays;      /* h: rays with any hit          */
	uint64_t max_visits;    /* max node visits of a ray      */
} orc_counters;

#define ORC_NO_HIT 0xffffffffu

/* compiler_options.h:13-19: FOCAL_LENGTH reaches the kernel through
 * `ostream << float` (6 significant digits) and the OpenCL compiler's
 * literal parser. */
float orc_focal_roundtrip(float focal);

/* intersect_kernel.cl:21-61 */
int orc_aabb_intersect(const float *bb /* 8 floats */, const float *pos4,
                       const float *dir4, float max_distance);

/* intersect_kernel.cl:278-310 for every pixel of a width x height image
 * (the super-sampled dimensions).  Any of face_id / distance / counters may
 * be NULL.  face_id[i] = 3 * leaf index, or ORC_NO_HIT; distance[i] = +inf on
 * miss.  Rows [row_begin,row_end) step row_step are rendered (others are
 * left untouched) so a bounded sample can be timed.  nthreads <= 0 uses one
 * pinned worker per CPU in the affinity mask.  jitter_seed == 0 reproduces
 * the reference's regular grid (+0.5f); otherwise the +0.5f offsets are
 * replaced by the hash documented in DESIGN.md (extension, config C3). */
int orc_render(const orc_scene *scene, unsigned width, unsigned height,
               float focal_length, int shading, uint32_t jitter_seed,
               unsigned row_begin, unsigned row_end, unsigned row_step,
               float *image, uint32_t *face_id, float *distance,
               orc_counters *counters, int nthreads);

/* Ambient occlusion options of the kernel (opencl_host.cc:47-53: AO_ENABLE, AO_MAX_DISTANCE, AO_NUM_SAMPLES,
 * AO_METHOD, AO_ALPHA_MIN, AO_ALPHA_MAX). */
typedef struct orc_ao {
	int      enable;
	int      method;        /* 0 uniform rings (:218-256), 1 random hemisphere (:257-275) */
	uint32_t samples;
	float    max_distance;
	int      alpha_min, alpha_max;
} orc_ao;

/* orc_render with the kernel's ambient-occlusion stage (intersect_kernel.cl:214-277, :305-307); ao may be NULL. */
int orc_render_ao(const orc_scene *scene, unsigned width, unsigned height,
                  float focal_length, int shading, uint32_t jitter_seed, const orc_ao *ao,
                  unsigned row_begin, unsigned row_end, unsigned row_step,
                  float *image, uint32_t *face_id, float *distance,
                  orc_counters *counters, int nthreads);

/* intersect_kernel.cl:184-213 for arbitrary rays (config C5).
 * origins/dirs: 4 floats per ray (w ignored = 0). */
int orc_trace_rays(const orc_scene *scene, const float *origins,
                   const float *dirs, size_t nrays, float max_distance,
                   uint32_t *face_id, float *distance,
                   orc_counters *counters, int nthreads);

/* The counter-based ray generator of config C5: fills
 * origins/dirs (4 floats per ray) for ray ids [first, first+n). */
void orc_gen_random_rays(uint32_t seed, uint64_t first, size_t n,
                         const float *bbmin3, const float *bbmax3,
                         float *origins, float *dirs);

/* src/ray_tracer.cc:3-15 */
void orc_resize(const float *tmp, unsigned total_width, unsigned width,
                unsigned height, unsigned n, unsigned char *image);

/* the +0.5f replacement of config C3's jittered variant */
void orc_jitter(uint32_t seed, uint32_t x, uint32_t y, float *jx, float *jy);

int orc_online_cpus(void);

#ifdef __cplusplus
}
#endif
#endif
