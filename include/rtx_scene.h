/*
 * rtx_scene.h -- host-side scene preparation in the reference's own data
 * format (C ABI, no CUDA).  Library: opencl_raytracer_b200/lib/librtx_scene.so
 *
 * The render CLI keeps using the reference's mesh.cc / bvh.cc (compiled
 * unmodified).  This library exists for callers that do
 * not have the reference tree (the Python host layer, bench.py, the GPU box):
 * it produces, bit for bit, the five arrays that src/render.cc:86-98 hands to
 * OpenCLHost::upload:
 *
 *   reference interface replaced                      entry point here
 *   ------------------------------------------------  -----------------------
 *   load_off_mesh            include/mesh.h:18        rtx_scene_from_off
 *   compute_vertex_normals   include/mesh.h:22        (inside both builders)
 *   BVH::buildBVH            include/bvh.h:13         (inside both builders)
 *   leaf-order face sort     src/render.cc:88-95      rtx_scene_faces
 *
 * Both of BVH::Method's builders: CUT_LONGEST_AXIS (bvh.cc:59-94, the CLI's
 * default) and SURFACE_AREA_HEURISTIC (bvh.cc:178-236, `-r sah`).  The SAH
 * split makes the reference's std::sort calls on the same sequences and keeps
 * its float / double cost expression, so it emits the reference's tree (golden
 * arrays from the reference builder in tests/golden/soup_sah.npz); the
 * right-hand boxes of the sweep come from a suffix scan instead of being
 * recomputed per cut position (min / max are exact), O(n log n) per node
 * instead of O(n^2).
 */
#ifndef RTX_SCENE_H
#define RTX_SCENE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rtx_scene rtx_scene;

#define RTX_SCENE_OK            0
#define RTX_SCENE_ERR_ARG       1  /* null/empty argument ("No filename given", mesh.cc:8-10) */
#define RTX_SCENE_ERR_IO        2  /* "Cannot read file", mesh.cc:13-15 */
#define RTX_SCENE_ERR_FORMAT    3  /* "File not recognized as OFF model" / "Invalid face with != 3 vertices" */
#define RTX_SCENE_ERR_EMPTY     4  /* no usable triangle (the reference would index nodes.at(0) of an empty tree) */

#define RTX_SCENE_BVH_LONGEST_AXIS 0  /* BVH::Method::CUT_LONGEST_AXIS */
#define RTX_SCENE_BVH_SAH          1  /* BVH::Method::SURFACE_AREA_HEURISTIC */

/* Load an OFF file, compute area-weighted vertex normals, build the BVH.
 * nthreads <= 0: one builder thread per online CPU (the result does not
 * depend on the thread count). */
int rtx_scene_from_off(const char *path, int nthreads, rtx_scene **out);

/* Same from in-memory arrays: verts3 = 3 floats per vertex, faces = 3 vertex
 * ids per triangle (faces naming a vertex >= nverts are skipped like
 * mesh.cc:48-53). */
int rtx_scene_from_mesh(const float *verts3, size_t nverts, const uint32_t *faces, size_t nfaces,
                        int nthreads, rtx_scene **out);

/* The same with the builder named (RTX_SCENE_BVH_*), bvh.h:13 `BVH(Method)`. */
int rtx_scene_from_off_method(const char *path, int method, int nthreads, rtx_scene **out);
int rtx_scene_from_mesh_method(const float *verts3, size_t nverts, const uint32_t *faces, size_t nfaces, int method,
                               int nthreads, rtx_scene **out);

void rtx_scene_free(rtx_scene *scene);

/* counts[0] face indices (3 per triangle), [1] nodes, [2] aabb vectors
 * (2 per node), [3] vertices, [4] vertex normals */
void rtx_scene_counts(const rtx_scene *scene, size_t counts[5]);

const uint32_t *rtx_scene_faces(const rtx_scene *scene);      /* leaf-ordered vertex ids */
const uint32_t *rtx_scene_triangles(const rtx_scene *scene);  /* leaf index -> input face id (BVH::triangles) */
const uint32_t *rtx_scene_orig_faces(const rtx_scene *scene); /* accepted input faces, input order */
const uint32_t *rtx_scene_nodes(const rtx_scene *scene);      /* pre-order subtree sizes (BVH::nodes) */
const float    *rtx_scene_aabbs(const rtx_scene *scene);      /* float4 (min,max) per node (BVH::aabbs) */
const float    *rtx_scene_vertices(const rtx_scene *scene);   /* float4 per vertex, w = 0 */
const float    *rtx_scene_normals(const rtx_scene *scene);    /* float4 per vertex, w = 0 */

/* message of the last failure on the calling thread */
const char *rtx_scene_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
