"""Two ranks on two GPUs (NCCL): every way a tile-partitioned frame can reach rank 0 gives the single-GPU image.

Needs a box with >= 2 GPUs (`gpurun --gpus 2`); skipped on the one-GPU box.  The one-GPU emulation of the partition
(tests/test_parity_gpu.py::test_tile_partition_equals_single_image) and the gloo test (tests/test_multigpu_cpu.py)
cover the host logic everywhere else."""
import os
import socket

import numpy as np
import pytest

from conftest import require_gpu

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    from opencl_raytracer_b200 import host, multigpu, scene, scenes
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    import datetime
    import traceback
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank),
                            timeout=datetime.timedelta(seconds=90))           # a rank that died must not cost ten minutes
    try:
        v, f = scenes.sibenik_standin(detail=0.35)
        sc = scene.scene_from_mesh(v, f)
        rt = host.RayTracer(host.Options(width=416, height=232, nSuperSamples=16))     # 1664 x 928 rays: 52 x 29 tiles
        for mode, sync in (("float", ""), ("u8", ""), ("p2p_u8", "flags"), ("p2p_float", "flags"), ("p2p_u8", "allreduce"), ("p2p_float", "allreduce")):
            r = multigpu.TiledRenderer(rt, sc, rank, world, rank, gather=mode, sync=sync or "flags")
            for _ in range(5):                      # both of the alternating p2p images get used, ranks run ahead of each other
                r.render_frame()
            if rank == 0:
                img = r.download_u8() if mode.endswith("u8") else r.download()
                np.save(os.path.join(out_dir, mode + ("_" + sync if sync == "allreduce" else "") + ".npy"), img)
            torch.cuda.synchronize()
            dist.barrier()
            r.close()
        # the caller's host image shared by the ranks: every rank stores its own tiles into it
        shared = multigpu.SharedHostImage(rt, rank, world)
        r = multigpu.TiledRenderer(rt, sc, rank, world, rank, gather="float")
        r.host.render_async(torch.cuda.current_stream().cuda_stream)
        r.host.store_tiles_async(shared.device_ptr, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            np.save(os.path.join(out_dir, "shared_host.npy"), np.array(shared.array))
        dist.barrier()
        r.close()
        shared.close()
    except Exception:
        with open(os.path.join(out_dir, "error_rank%d.txt" % rank), "w") as fh:
            fh.write(traceback.format_exc())
        raise
    finally:
        dist.destroy_process_group()


def test_two_gpus_every_gather_mode(tmp_path, po):
    host = require_gpu()
    if host.device_count() < 2:
        pytest.skip("one GPU on this box")
    import torch.multiprocessing as mp
    from opencl_raytracer_b200 import scene, scenes
    try:
        mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    except Exception as e:
        notes = "".join(open(str(p)).read() for p in sorted(tmp_path.glob("error_rank*.txt")))
        pytest.fail("a rank failed: %s\n%s" % (e, notes))
    v, f = scenes.sibenik_standin(detail=0.35)
    sc = scene.scene_from_mesh(v, f)
    rt = host.RayTracer(host.Options(width=416, height=232, nSuperSamples=16))
    with host.CudaHost(rt) as h:
        h.upload_scene(sc)
        h()
        want_f, want_b = h.download(), h.download_u8()
    ref = po.render(sc, rt.totalWidth, rt.totalHeight, 1.0, True, want_ids=False).image
    assert np.array_equal(want_f, ref)
    for mode in ("float", "p2p_float", "p2p_float_allreduce", "shared_host"):
        assert np.array_equal(np.load(str(tmp_path / (mode + ".npy"))), want_f), mode
    for mode in ("u8", "p2p_u8", "p2p_u8_allreduce"):
        assert np.array_equal(np.load(str(tmp_path / (mode + ".npy"))), want_b), mode
