"""NCCL gather latency on the box (development tool): torchrun --nproc-per-node N tools/gather_latency.py"""
import os
import torch
import torch.distributed as dist

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr)
for nbytes in (8294400 // world, 530841600 // world):
    local = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    out = torch.empty(world * nbytes, dtype=torch.uint8, device=dev) if rank == 0 else None
    outs = list(out.chunk(world)) if rank == 0 else None
    full = torch.empty(world * nbytes, dtype=torch.uint8, device=dev)
    tok = torch.zeros(1, dtype=torch.int32, device=dev)

    def t_gather():
        dist.gather(local, outs, dst=0)

    def t_allgather():
        dist.all_gather_into_tensor(full, local)

    def t_sendrecv():
        if rank == 0:
            ops = [dist.P2POp(dist.irecv, outs[r], r) for r in range(1, world)]
        else:
            ops = [dist.P2POp(dist.isend, local, 0)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()

    for name, fn in (("gather", t_gather), ("all_gather_into_tensor", t_allgather), ("batch_isend_irecv", t_sendrecv)):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(20):
            dist.all_reduce(tok)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = torch.tensor([sorted(ts)[len(ts) // 2]], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print("world %d, %9d B per rank: %-24s median %.1f us (max over ranks)" % (world, nbytes, name, t.item() * 1e3), flush=True)
dist.destroy_process_group()
