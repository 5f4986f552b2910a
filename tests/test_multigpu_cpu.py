"""The N>1 host logic on CPU: world_size-2/3 gloo process groups run the same partition -> single gather ->
de-interleave sequence the GPU path uses (opencl_raytracer_b200/multigpu.py), with the oracle's image standing in
for what each rank's kernel would have written into its compact tile buffer."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from opencl_raytracer_b200 import host, multigpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, image, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        H, W = image.shape
        local = torch.from_numpy(multigpu.pack_tiles(image, rank, world))
        _, _, tpr = multigpu.tile_counts(W, H, world)
        assert local.numel() == tpr * 1024
        gathered = multigpu.gather_to_rank0(local, world, rank)
        if rank == 0:
            np.save(out_path, multigpu.unpack_tiles(gathered.numpy(), W, H, world))
        else:
            assert gathered is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partition_gather_deinterleave(tmp_path, po, soup_scene, world):
    W, H = 150, 70                                    # not multiples of the 32-pixel tile
    image = po.render(soup_scene, W, H, 1.0, True, want_ids=False).image
    out = str(tmp_path / "img.npy")
    mp.spawn(_worker, args=(world, _free_port(), image, out), nprocs=world, join=True)
    assert np.array_equal(np.load(out), image)


def test_layout_matches_the_c_abi():
    for W, H, world in ((3840, 2160, 8), (15360, 8640, 4), (33, 17, 2), (150, 70, 3), (32, 32, 1)):
        assert multigpu.tile_counts(W, H, world) == host.tile_layout(W, H, world)
        tx, ty, tpr = multigpu.tile_counts(W, H, world)
        counts = [multigpu.local_tiles(W, H, r, world) for r in range(world)]
        assert sum(counts) == tx * ty and max(counts) == tpr


def test_pack_unpack_roundtrip():
    rng = np.random.default_rng(0)
    img = rng.random((70, 150), dtype=np.float32)
    for world in (1, 2, 5, 8):
        g = np.concatenate([multigpu.pack_tiles(img, r, world) for r in range(world)])
        assert np.array_equal(multigpu.unpack_tiles(g, 150, 70, world), img)
