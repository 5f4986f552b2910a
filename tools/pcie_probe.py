"""Aggregate device->host bandwidth of N rank processes (diagnostic for the multi-GPU end-to-end path).
   torchrun --nproc-per-node N tools/pcie_probe.py
 a) every rank DMA-copies 256 MiB into its OWN pinned buffer (cudaMemcpyAsync), all ranks at once
 b) every rank's store kernel writes its tiles of the C3 frame into its OWN page-locked full-size image (zero copy)
 c) ... into ONE image shared by all ranks (what bench.py's e2e does)"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opencl_raytracer_b200 import host, multigpu, scene as scn, scenes  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr)


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
        torch.cuda.synchronize()
        dist.barrier()
    return (time.perf_counter() - t0) / n


nbytes = 256 << 20
src = torch.empty(nbytes, dtype=torch.uint8, device=dev)
dst = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
t = timed(lambda: dst.copy_(src, non_blocking=True))
if rank == 0:
    print("a) DMA, private pinned buffers: %.2f ms -> %.1f GB/s per rank, %.1f GB/s aggregate" % (t * 1e3, nbytes / t / 1e9, world * nbytes / t / 1e9), flush=True)

v, f = scenes.sibenik_standin()
sc = scn.scene_from_mesh(v, f)
rt = host.RayTracer(host.Options(width=3840, height=2160, nSuperSamples=16))
r = multigpu.TiledRenderer(rt, sc, rank, world, lr, gather="float")
r.render_frame()
torch.cuda.synchronize()
frame = rt.totalWidth * rt.totalHeight * 4
st = torch.cuda.current_stream().cuda_stream
own = torch.empty((rt.totalHeight, rt.totalWidth), dtype=torch.float32).pin_memory()
t = timed(lambda: r.host.store_tiles_async(own.data_ptr(), st))
if rank == 0:
    print("b) store kernel, private page-locked images: %.2f ms -> %.1f GB/s per rank, %.1f GB/s aggregate" % (t * 1e3, frame / world / t / 1e9, frame / t / 1e9), flush=True)
shared = multigpu.SharedHostImage(rt, rank, world)
t = timed(lambda: r.host.store_tiles_async(shared.device_ptr, st))
if rank == 0:
    print("c) store kernel, ONE shared image: %.2f ms -> %.1f GB/s per rank, %.1f GB/s aggregate" % (t * 1e3, frame / world / t / 1e9, frame / t / 1e9), flush=True)
# d) same shared image, but each rank writes a contiguous slab of rows instead of interleaved 128-byte pieces (DMA)
rows = rt.totalHeight // world
img = torch.empty((rt.totalHeight, rt.totalWidth), dtype=torch.float32, device=dev)
dst_view = torch.from_numpy(np.asarray(shared.array))[rank * rows:(rank + 1) * rows]
lib = torch.cuda.cudart()
t = timed(lambda: r.host.copy_to_host(shared.array[rank * rows:(rank + 1) * rows], img.data_ptr() + rank * rows * rt.totalWidth * 4))
if rank == 0:
    print("d) DMA of a contiguous slab of rows per rank into the shared image: %.2f ms -> %.1f GB/s aggregate" % (t * 1e3, frame / t / 1e9), flush=True)
shared.close()
r.close()
dist.barrier()
dist.destroy_process_group()
