"""C3 frame rendered as `world` emulated ranks on one GPU (two-level frustum pass per rank) == the single-context frame."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from opencl_raytracer_b200 import host, scene as scn, scenes
v, f = scenes.sibenik_standin(); sc = scn.scene_from_mesh(v, f)
rt = host.RayTracer(host.Options(width=3840, height=2160, nSuperSamples=16))
with host.CudaHost(rt) as h:
    h.upload_scene(sc); h(); single = h.download()
for world in [int(x) for x in sys.argv[1:]] or [8, 3]:
    tx, ty, tpr = host.tile_layout(rt.totalWidth, rt.totalHeight, world)
    gathered = torch.zeros(world * tpr * 1024, dtype=torch.float32, device="cuda")
    for r in range(world):
        with host.CudaHost(rt, tile_rank=r, tile_world=world) as c:
            c.upload_scene(sc)
            c.bind_output(gathered[r * tpr * 1024:(r + 1) * tpr * 1024].data_ptr(), tpr * 1024)
            c()
    torch.cuda.synchronize()
    with host.CudaHost(rt, tile_rank=0, tile_world=world) as c:
        c.upload_scene(sc)
        c.deinterleave_async(gathered.data_ptr(), world); c.synchronize()
        img = c.download()
    print("world %d: %d of %d pixels differ from the single-context frame" % (world, int((img != single).sum()), img.size), flush=True)
