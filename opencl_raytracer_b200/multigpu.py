"""Tile-partitioned rendering over the GPUs of one node: one process per GPU.

The scene is replicated; the super-sampled image is cut into 32x32-pixel tiles
and tile ``t`` is rendered by rank ``t % world`` (interleaved, so sky and
geometry are spread evenly).  Each rank's kernel writes its tiles into a compact
``[local_tile][32][32]`` float buffer owned by torch.  How the frame reaches rank 0
is the ``gather`` mode of ``TiledRenderer``:

* ``"float"`` / ``"u8"``: ONE ``torch.distributed.gather`` (NCCL over NVLink on the GPU
  box, gloo on CPU in the tests) of the float tiles -- or of the bytes after every rank
  applied ``RayTracer::resize`` to its own tiles --, then a de-interleave kernel on rank 0;
* ``"p2p_u8"`` / ``"p2p_float"``: no collective on the data path.  Rank 0's final image
  is mapped into every rank process (CUDA IPC) and each rank's resize / store kernel
  writes its share straight into it over NVLink; a one-element all-reduce orders the ranks.

``SharedHostImage`` is the same idea for the caller's HOST image: one page-locked buffer
mapped by every rank, so each rank's tiles leave over its own PCIe link.
``world == 1`` skips all communication.

``torch`` is plumbing here (device memory, streams, the process group); tracing,
resizing, storing and de-interleaving go through the C ABI (include/rtx_b200.h).
"""
from __future__ import annotations

import numpy as np

TILE = 32


def tile_counts(total_width: int, total_height: int, world: int):
    """(tiles_x, tiles_y, tiles_per_rank) -- same arithmetic as rtx_tile_layout."""
    tx = (total_width + TILE - 1) // TILE
    ty = (total_height + TILE - 1) // TILE
    return tx, ty, (tx * ty + world - 1) // world


def local_tiles(total_width: int, total_height: int, rank: int, world: int) -> int:
    tx, ty, _ = tile_counts(total_width, total_height, world)
    return (tx * ty + world - 1 - rank) // world


def pack_tiles(image: np.ndarray, rank: int, world: int) -> np.ndarray:
    """Host model of what rank `rank`'s kernel writes: its tiles of a row-major image, compact and padded."""
    H, W = image.shape
    tx, ty, tpr = tile_counts(W, H, world)
    out = np.zeros((tpr, TILE, TILE), image.dtype)
    for lt in range(local_tiles(W, H, rank, world)):
        t = lt * world + rank
        y0, x0 = (t // tx) * TILE, (t % tx) * TILE
        blk = image[y0:y0 + TILE, x0:x0 + TILE]
        out[lt, :blk.shape[0], :blk.shape[1]] = blk
    return out.reshape(-1)


def unpack_tiles(gathered: np.ndarray, total_width: int, total_height: int, world: int) -> np.ndarray:
    """Host model of the de-interleave kernel: rank-major compact buffers -> row-major image."""
    tx, ty, tpr = tile_counts(total_width, total_height, world)
    g = np.asarray(gathered).reshape(world, tpr, TILE, TILE)
    img = np.zeros((total_height, total_width), g.dtype)
    for t in range(tx * ty):
        y0, x0 = (t // tx) * TILE, (t % tx) * TILE
        h, w = min(TILE, total_height - y0), min(TILE, total_width - x0)
        img[y0:y0 + h, x0:x0 + w] = g[t % world, t // world, :h, :w]
    return img


def gather_to_rank0(local, world: int, rank: int, group=None):
    """The frame's single collective.  `local`: this rank's compact tile tensor.  Returns the
    rank-major concatenation on rank 0 (None elsewhere).  world == 1: no communication."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local
    if rank == 0:
        out = torch.empty(world * local.numel(), dtype=local.dtype, device=local.device)
        dist.gather(local, list(out.chunk(world)), dst=0, group=group)
        return out
    dist.gather(local, None, dst=0, group=group)
    return None


class TiledRenderer:
    """One rank of a tile-partitioned render (CUDA).  Usage, on every rank::

        r = TiledRenderer(rt, scene, rank, world, device)
        r.render_frame()                # kernel -> (gather -> de-interleave on rank 0), all on torch's stream
        img = r.download()              # rank 0 only

    gather (what a frame ends with on rank 0, and how it gets there):
      "float"      ONE NCCL gather of the float tiles + de-interleave kernel: the image ``download(float*)`` returns
      "u8"         every rank applies RayTracer::resize to its own tiles, ONE NCCL gather of the bytes + de-interleave
                   ((n*n*4)x less traffic into rank 0; needs sqrt(nSuperSamples) to divide 32)
      "p2p_u8"     no collective on the data path: every rank's resize kernel stores its bytes straight into its slot of
                   a buffer in rank 0's memory through NVLink peer memory (CUDA IPC mapping, set up once); the ranks are
                   ordered by a one-element all-reduce or by frame counters in that memory (``sync``); rank 0 de-interleaves.  Two buffers
                   alternate, so rank 0 may read frame k while k+1 is written.
      "p2p_float"  the float image: in every rank's traversal kernel the warp that finishes a tile sends it to its place in
                   rank 0's row-major image (rtx_render_store_async on the peer mapping), so the transfer overlaps the
                   tracing; same ordering.  (Writing each pixel remotely as it is shaded, rtx_bind_output_image, moves
                   8-byte pieces over NVLink: 1.68 ms against 1.03 ms of tracing at 4 GPUs.)
    """

    def __init__(self, rt, scene, rank: int, world: int, device: int, jitter_seed: int = 0, gather: str = "float", sync: str = "allreduce"):
        """sync (p2p modes): how the ranks are ordered once their stores are queued --
        "allreduce"  a one-element NCCL all-reduce on the same stream: it completes on rank 0 only when every rank's store
                     kernel has finished (default);
        "flags"      frame counters in rank 0's memory (rtx_peer_signal_async / rtx_peer_wait_async): a rank raises its counter
                     behind its store kernel, rank 0 waits for all of them, and a rank waits for rank 0's "consumed" counter
                     before it overwrites a buffer.  No collective at all in the frame; waits time out instead of hanging.
        Measured equal within noise (8 GPUs, C3: 0.607 ms with the counters, 0.594 ms with the all-reduce): what rank 0
        waits for is the slowest rank's tracing, not the mechanism."""
        import torch
        from . import host
        self.torch = torch
        self.rank, self.world, self.rt = rank, world, rt
        if world == 1 and gather.startswith("p2p_"):
            gather = gather[4:]
        self.gather = gather
        if gather not in ("float", "u8", "p2p_u8", "p2p_float"):
            raise ValueError("unknown gather mode %r" % gather)
        if gather.endswith("u8") and world > 1 and TILE % rt.n != 0:
            raise ValueError("u8 gather needs sqrt(nSuperSamples) to divide %d" % TILE)
        self.dev = torch.device("cuda", device)
        self.host = host.CudaHost(rt, device=device, jitter_seed=jitter_seed, tile_rank=rank, tile_world=world)
        self.host.upload_scene(scene)
        _, n = self.host.device_image()
        self.local = torch.zeros(n, dtype=torch.float32, device=self.dev)
        self.host.bind_output(self.local.data_ptr(), n)
        self.gathered = None
        self.kernel_launches = 0
        self.frame = 0
        self.timing = False          # record torch events around the multi-GPU phases (phase_ms)
        self._ev = None
        self.peer = None             # p2p modes: the two final images in rank 0's memory, as seen from this rank
        self.sync = sync
        self.flags = None            # p2p modes, sync="flags": 128 counters in rank 0's memory ([r] arrived, [64] consumed, [65] timed out)
        if gather == "u8" and world > 1:
            m = TILE // rt.n
            self.local_u8 = torch.zeros(tile_counts(rt.totalWidth, rt.totalHeight, world)[2] * m * m, dtype=torch.uint8, device=self.dev)
        if gather.startswith("p2p_"):
            import torch.distributed as dist
            m = TILE // rt.n if gather == "p2p_u8" else 0
            self.u8_per_rank = tile_counts(rt.totalWidth, rt.totalHeight, world)[2] * m * m
            nbytes = world * self.u8_per_rank if gather == "p2p_u8" else rt.totalWidth * rt.totalHeight * 4
            # collective set-up that cannot leave a rank waiting: every step ends with all ranks knowing whether it worked
            box, err = [None], None
            if rank == 0:
                try:
                    owned = [self.host.peer_alloc(nbytes) for _ in range(2)] + [self.host.peer_alloc(512)]
                    box = [[h for _, h in owned]]
                    self.peer = [p for p, _ in owned[:2]]
                    self.flags = owned[2][0]
                except host.RtxError as e:
                    err = str(e)
            dist.broadcast_object_list(box, src=0)
            if rank != 0 and box[0] is not None:
                try:
                    opened = [self.host.peer_open(h) for h in box[0]]
                    self.peer, self.flags = opened[:2], opened[2]
                except host.RtxError as e:
                    err = str(e)
            ok = torch.tensor([0 if (err or box[0] is None) else 1], dtype=torch.int32, device=self.dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                if self.peer and rank == 0:
                    for p in self.peer + [self.flags]:
                        self.host.peer_free(p)
                self.peer = None
                self.host.close()
                raise RuntimeError("peer-memory gather unavailable on this box (rank %d: %s)" % (rank, err or "another rank failed"))
            self.flag = torch.zeros(1, dtype=torch.int32, device=self.dev)

    def _mark(self, name):
        if self.timing:
            e = self.torch.cuda.Event(enable_timing=True)
            e.record()
            self._ev.append((name, e))

    def render_frame(self):
        torch = self.torch
        stream = torch.cuda.current_stream(self.dev).cuda_stream
        if self.timing:
            self._ev = []
        flags_sync = self.peer is not None and self.sync == "flags"
        if flags_sync and self.rank != 0 and self.frame >= 2:
            # this frame's buffer was last used by frame - 2: rank 0 must have consumed that one
            self.host.peer_wait_async(self.flags + 4 * 64, 1, self.frame - 1, self.flags + 4 * 65, stream)
        if self.gather == "p2p_float":
            # the traversal kernel sends every finished tile to its place in rank 0's image (128-byte rows over NVLink)
            self.host.render_store_async(self.peer[self.frame & 1], stream)
        else:
            self.host.render_async(stream)
        self.kernel_launches += self.host.last_launches()
        self._mark("resize")
        if self.gather == "u8":
            if self.world == 1:
                self.host.resize_u8_async(0, 0, stream)
                self.kernel_launches += 1
            else:
                self.host.resize_u8_async(self.local_u8.data_ptr(), self.local_u8.numel(), stream)
                self.kernel_launches += 1
                self._mark("gather")
                self.gathered = gather_to_rank0(self.local_u8, self.world, self.rank)
                self._mark("deinterleave")
                if self.rank == 0:
                    self.host.deinterleave_u8_async(self.gathered.data_ptr(), self.world, stream)
                    self.kernel_launches += 1
        elif self.gather == "float" and self.world > 1:
            self._mark("gather")
            self.gathered = gather_to_rank0(self.local, self.world, self.rank)
            self._mark("deinterleave")
            if self.rank == 0:
                self.host.deinterleave_async(self.gathered.data_ptr(), self.world, stream)
                self.kernel_launches += 1
        elif self.gather.startswith("p2p_"):
            import torch.distributed as dist
            target = self.peer[self.frame & 1]
            if self.gather == "p2p_u8":
                # resize this rank's tiles and store the bytes, compact and contiguous, into this rank's slot of the buffer
                # in rank 0's memory: one kernel, coalesced remote writes (storing every 8-byte tile row at its final
                # row-major place instead cost 0.07 ms against 0.02 ms at 4 GPUs)
                self.host.resize_u8_async(target + self.rank * self.u8_per_rank, self.u8_per_rank, stream)
                self.kernel_launches += 1
            self._mark("gather")
            if flags_sync:
                self.host.peer_signal_async(self.flags + 4 * self.rank, self.frame + 1, stream)     # behind this rank's stores
                if self.rank == 0:
                    self.host.peer_wait_async(self.flags, self.world, self.frame + 1, self.flags + 4 * 65, stream)
                self.kernel_launches += 1 if self.rank else 2
            else:
                dist.all_reduce(self.flag)                        # rendezvous: every rank's stores have landed
            if self.rank == 0 and self.gather == "p2p_u8":
                self._mark("deinterleave")
                self.host.deinterleave_u8_async(target, self.world, stream)
                self.kernel_launches += 1
            if flags_sync and self.rank == 0:
                self.host.peer_signal_async(self.flags + 4 * 64, self.frame + 1, stream)            # frame consumed
                self.kernel_launches += 1
            self.last_target = target
        self._mark("end")
        self.frame += 1

    def phase_ms(self):
        """Device time per phase of the last frame rendered with ``timing`` (and TUNE_PHASE_TIMING) set: the launch groups
        of rtx_render_async (host.PHASES) + resize / gather / deinterleave of this module.  Call after a synchronize."""
        out = dict(self.host.phase_ms())
        ev = self._ev or []
        for (name, a), (_, b) in zip(ev, ev[1:]):
            out[name] = out.get(name, 0.0) + a.elapsed_time(b)
        return out

    def _check_flags(self):
        if self.peer is not None and self.sync == "flags":
            words = self.host.copy_to_host(np.zeros(128, np.uint32), self.flags)
            if words[65]:
                raise RuntimeError("a rank did not arrive within the time-out of the peer-memory frame counters: %s" % words[:self.world])

    def download(self):
        assert self.rank == 0
        self.torch.cuda.synchronize(self.dev)
        self._check_flags()
        if self.gather == "p2p_float":
            return self.host.copy_to_host(np.empty((self.rt.totalHeight, self.rt.totalWidth), np.float32), self.last_target)
        return self.host.download()

    def download_u8(self):
        assert self.rank == 0
        self.torch.cuda.synchronize(self.dev)
        self._check_flags()
        return self.host.download_u8()

    def close(self):
        if self.peer:
            self.torch.cuda.synchronize(self.dev)
            if self.world > 1:
                import torch.distributed as dist
                dist.barrier()                        # nobody unmaps or frees while a peer may still store
            if self.rank == 0:
                for p in self.peer + [self.flags]:
                    self.host.peer_free(p)
            else:
                for p in self.peer + [self.flags]:
                    self.host.peer_close(p)
            self.peer = None
        self.host.close()


class SharedHostImage:
    """The caller's float image as ONE page-locked host buffer that every rank process maps (a file in /dev/shm) and
    registers with CUDA: each rank's tiles then leave the device over that rank's own PCIe link
    (rtx_store_tiles_async with the mapped pointer) instead of funnelling through rank 0's.  Collective constructor."""

    def __init__(self, rt, rank: int, world: int, directory: str = "/dev/shm"):
        import os
        import torch.distributed as dist
        from . import host
        self.host_mod, self.rank = host, rank
        shape = (rt.totalHeight, rt.totalWidth)
        box = [None]
        self.array = None
        if rank == 0:
            self.path = os.path.join(directory, "rtx_b200_image_%d_%d" % (os.getpid(), id(self) & 0xffff))
            try:
                fd = os.open(self.path, os.O_CREAT | os.O_RDWR, 0o600)
                try:
                    os.posix_fallocate(fd, 0, shape[0] * shape[1] * 4)     # ENOSPC here, not SIGBUS at the first store
                finally:
                    os.close(fd)
                self.array = np.memmap(self.path, dtype=np.float32, mode="r+", shape=shape)
                box = [self.path]
            except OSError as e:
                try:
                    os.unlink(self.path)
                except OSError:
                    pass
                box = [None]
                self.error = str(e)
        if world > 1:
            dist.broadcast_object_list(box, src=0)
        if box[0] is None:
            raise RuntimeError("no room for the shared host image in %s: %s" % (directory, getattr(self, "error", "rank 0 failed")))
        if rank != 0:
            self.path = box[0]
            self.array = np.memmap(self.path, dtype=np.float32, mode="r+", shape=shape)
        self.array[::1024] = 0            # touch: the pages exist before they are pinned
        self.device_ptr = host.host_register(self.array)
        if world > 1:
            dist.barrier()
        if rank == 0:
            os.unlink(self.path)          # the mappings keep it alive

    def close(self):
        if self.array is not None:
            self.host_mod.host_unregister(self.array)
            self.array = None
