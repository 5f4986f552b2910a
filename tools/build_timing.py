"""Device BVH build (rtx_upload_mesh) against the host builder of rtx_scene.h (development tool).
usage: python tools/build_timing.py [soup] [sibenik] [bunny] [c4]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opencl_raytracer_b200 import host, scene as scn, scenes
from oracle import pyoracle as po

def mesh(name):
    if name == "sibenik":
        return scenes.sibenik_standin()
    if name == "soup":
        return scenes.random_soup(100000, seed=5, size=0.05)
    v, f = po.read_mesh_bin(po.staged_bunny_path())
    if name == "c4":
        v, f = scenes.subdivided(v, f)
    return v, f

for name in sys.argv[1:] or ["sibenik", "bunny", "soup", "c4"]:
    v, f = mesh(name)
    t = time.perf_counter(); sc = scn.scene_from_mesh(v, f, name=name); t_host = time.perf_counter() - t
    rt = host.RayTracer(host.Options(width=64, height=64, nSuperSamples=1))
    with host.CudaHost(rt) as h:
        best, wall = 1e9, 1e9
        for _ in range(3):
            t = time.perf_counter(); h.upload_mesh(sc.vertices, sc.orig_faces, sc.normals); w = time.perf_counter() - t
            ms, levels = h.build_stats()
            best, wall = min(best, ms), min(wall, w)
        nodes, aabbs, tri, faces = h.download_tree()
        same = (np.array_equal(nodes, sc.nodes) and np.array_equal(tri, sc.triangles) and np.array_equal(faces, sc.faces)
                and np.array_equal(aabbs.view(np.uint32), np.ascontiguousarray(sc.aabbs, np.float32).reshape(-1, 4).view(np.uint32)))
        t = time.perf_counter(); h.upload_scene(sc); t_up = time.perf_counter() - t
    print("%-8s %9d triangles: host build (load+normals+BVH) %.1f ms | device build %.2f ms in %d levels, rtx_upload_mesh %.1f ms wall "
          "(rtx_upload of the host-built arrays %.1f ms) | identical arrays: %s" % (name, sc.num_triangles, t_host * 1e3, best, levels, wall * 1e3, t_up * 1e3, same), flush=True)
