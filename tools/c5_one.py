import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opencl_raytracer_b200 import host, scene as scn, scenes
from oracle import pyoracle as po
v, f = scenes.sibenik_standin()
sib = scn.scene_from_mesh(v, f, name="sibenik_standin")
lo, hi = sib.root_box()
rt = host.RayTracer(host.Options(width=32, height=32, nSuperSamples=1))
bad = 3664471
base = bad & ~31
o, d = po.gen_random_rays(1234, base, 32, lo, hi)
k = bad - base
ref = po.trace_rays(sib, o, d, 100000.0, want_counters=True)
print("oracle", ref.face_id[k], ref.distance[k])
# the two triangles
for t in (60132, 60133):
    fv = sib.faces[3*t:3*t+3]
    print("tri", t, sib.vertices[fv, :3].tolist())
nodes = sib.nodes; leaf_idx = np.nonzero(nodes == 1)[0]
for t in (60132, 60133):
    print("leafbox", t, sib.aabbs.reshape(-1, 2, 4)[leaf_idx[t]].tolist())
for kern, name in ((1, "refill"), (0, "plain")):
    for leaf in (1, 2, 4):
        with host.CudaHost(rt) as h:
            h.set_tunable(host.TUNE_LEAF_SIZE, leaf)
            h.set_tunable(host.TUNE_INCOHERENT_KERNEL, kern)
            h.set_tunable(host.TUNE_COUNTERS, 1)
            h.upload_scene(sib)
            fid, dist = h.trace_rays(o, d)
            st = h.stats()
            one = h.trace_rays(o[k:k+1], d[k:k+1])
            st1 = h.stats()
            rep = h.trace_rays(np.repeat(o[k:k+1], 32, 0), np.repeat(d[k:k+1], 32, 0))
            print(name, "leaf", leaf, "warp ctx:", fid[k], dist[k], "mismatches in warp", int((fid != ref.face_id).sum()),
                  "| alone:", one[0][0], one[1][0], "visits", st1["node_visits"], "tris", st1["tri_tests"], "| x32:", set(rep[0].tolist()))
with host.CudaHost(rt) as h:
    h.set_tunable(host.TUNE_KERNEL, host.KERNEL_EXHAUSTIVE)
    h.upload_scene(sib)
    one = h.trace_rays(o[k:k+1], d[k:k+1])
    print("exhaustive alone:", one[0][0], one[1][0])
