"""Device time of ONE rank's share of the C3 frame (development tool): what a rank of an N-GPU run executes between
the barrier and the NCCL gather, emulated on one GPU (tile_rank 0 of tile_world N).

usage: python tools/rank_timing.py [world ...]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from opencl_raytracer_b200 import host, scene as scn, scenes  # noqa: E402


def main():
    worlds = [int(x) for x in sys.argv[1:]] or [1, 2, 4, 8]
    v, f = scenes.sibenik_standin()
    sc = scn.scene_from_mesh(v, f, name="sibenik_standin")
    rt = host.RayTracer(host.Options(width=3840, height=2160, nSuperSamples=16))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for world in worlds:
        with host.CudaHost(rt, tile_rank=0, tile_world=world) as h:
            h.upload_scene(sc)
            _, n = h.device_image()
            local = torch.zeros(n, dtype=torch.float32, device="cuda")
            h.bind_output(local.data_ptr(), n)
            m = 32 // rt.n
            tpr = host.tile_layout(rt.totalWidth, rt.totalHeight, world)[2]
            u8 = torch.zeros(tpr * m * m, dtype=torch.uint8, device="cuda")
            stream = torch.cuda.current_stream().cuda_stream
            times, ktimes = [], []
            for it in range(8):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                h.render_async(stream)
                if world > 1:
                    h.resize_u8_async(u8.data_ptr(), u8.numel(), stream)
                else:
                    h.resize_u8_async(0, 0, stream)
                e1.record()
                torch.cuda.synchronize()
                if it >= 3:
                    times.append(e0.elapsed_time(e1))
                    ktimes.append(h.stats()["kernel_ms"])
            t, k = min(times), min(ktimes)
            print("world %d rank 0: trace+resize %.3f ms (trace %.3f ms), launches %d; ideal = %.3f ms" % (
                world, t, k, h.stats()["kernel_launches"] + 1, 0.0))


if __name__ == "__main__":
    main()
