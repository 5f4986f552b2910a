"""One launch of the arbitrary-ray kernel on 2^26 generated rays (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opencl_raytracer_b200 import host, scene as scn, scenes
v, f = scenes.sibenik_standin(); sc = scn.scene_from_mesh(v, f)
rt = host.RayTracer(host.Options(width=32, height=32, nSuperSamples=1))
with host.CudaHost(rt) as h:
    h.upload_scene(sc)
    h.trace_random_rays(1234, 0, 1 << 24)
    h.trace_random_rays(1234, 0, 1 << 26)
    print("2^26 rays: %.3f ms" % h.stats()["kernel_ms"])
