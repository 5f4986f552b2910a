"""Tile-partitioned rendering over the GPUs of one node: one process per GPU.

The scene is replicated; the super-sampled image is cut into 32x32-pixel tiles
and tile ``t`` is rendered by rank ``t % world`` (interleaved, so sky and
geometry are spread evenly).  Each rank's kernel writes its tiles into a compact
``[local_tile][32][32]`` float buffer owned by torch; the only communication of a
frame is ONE ``torch.distributed.gather`` of those buffers to rank 0 (NCCL over
NVLink on the GPU box, gloo on CPU in the tests), followed on rank 0 by a
de-interleave kernel into the row-major image that ``download`` returns.
``world == 1`` skips the collective entirely.

``torch`` is plumbing here (device memory, streams, the process group); tracing
and de-interleaving go through the C ABI (include/rtx_b200.h).
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
    """

    def __init__(self, rt, scene, rank: int, world: int, device: int, jitter_seed: int = 0, gather: str = "float"):
        """gather = "float": the frame's collective moves the float tiles (what ``download(float*)`` needs);
        gather = "u8": every rank applies RayTracer::resize to its own tiles first and the collective moves
        bytes -- (n*n*4)x less traffic into rank 0; needs sqrt(nSuperSamples) to divide 32."""
        import torch
        from . import host
        self.torch = torch
        self.rank, self.world, self.rt = rank, world, rt
        self.gather = gather
        if gather == "u8" and world > 1 and TILE % rt.n != 0:
            raise ValueError("u8 gather needs sqrt(nSuperSamples) to divide %d" % TILE)
        self.dev = torch.device("cuda", device)
        self.host = host.CudaHost(rt, device=device, jitter_seed=jitter_seed, tile_rank=rank, tile_world=world)
        self.host.upload_scene(scene)
        _, n = self.host.device_image()
        self.local = torch.zeros(n, dtype=torch.float32, device=self.dev)
        self.host.bind_output(self.local.data_ptr(), n)
        self.gathered = None
        self.kernel_launches = 0
        if gather == "u8" and world > 1:
            m = TILE // rt.n
            self.local_u8 = torch.zeros(tile_counts(rt.totalWidth, rt.totalHeight, world)[2] * m * m, dtype=torch.uint8, device=self.dev)

    def render_frame(self):
        torch = self.torch
        stream = torch.cuda.current_stream(self.dev).cuda_stream
        self.host.render_async(stream)
        self.kernel_launches += self.host.last_launches()
        if self.gather == "u8":
            if self.world == 1:
                self.host.resize_u8_async(0, 0, stream)
                self.kernel_launches += 1
            else:
                self.host.resize_u8_async(self.local_u8.data_ptr(), self.local_u8.numel(), stream)
                self.kernel_launches += 1
                self.gathered = gather_to_rank0(self.local_u8, self.world, self.rank)
                if self.rank == 0:
                    self.host.deinterleave_u8_async(self.gathered.data_ptr(), self.world, stream)
                    self.kernel_launches += 1
        elif self.world > 1:
            self.gathered = gather_to_rank0(self.local, self.world, self.rank)
            if self.rank == 0:
                self.host.deinterleave_async(self.gathered.data_ptr(), self.world, stream)
                self.kernel_launches += 1

    def download(self):
        assert self.rank == 0
        self.torch.cuda.synchronize(self.dev)
        return self.host.download()

    def download_u8(self):
        assert self.rank == 0
        self.torch.cuda.synchronize(self.dev)
        return self.host.download_u8()

    def close(self):
        self.host.close()
