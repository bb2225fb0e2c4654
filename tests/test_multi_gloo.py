"""N > 1 host-side logic on CPU: two gloo ranks each render their interleaved row strips (the oracle
stands in for the GPU renderer — this test is about the sharding / gather / reassembly plumbing in
aicb200.multi), rank 0 reassembles, and the frame must equal the 1-rank frame byte for byte."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, width, height, result_path):
    sys.path.insert(0, os.path.join(ROOT, "all-is-cubes_b200"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import aicb200
    import orc
    from aicb200 import multi, scenes

    space = scenes.small_mixed_scene(n=10, seed=11)
    opts = aicb200.GraphicsOptions(view_distance=50.0)
    cam = scenes.standard_camera(space, opts, width, height)
    o = orc.OracleScene(space)
    part = o.render(cam, opts, shard=(multi.STRIP_ROWS, rank, world), n_threads=2)["srgb8"]
    rows = multi.shard_rows(height, rank, world)
    assert part.shape[0] == len(rows) * width
    n_max = multi.max_shard_rows(height, world) * width
    local = torch.zeros((n_max, 4), dtype=torch.uint8)
    local[: part.shape[0]] = torch.from_numpy(part)
    frame = multi.gather_frame(local, height, width, rank, world)
    if rank == 0:
        full = o.render(cam, opts, n_threads=2)["srgb8"].reshape(height, width, 4)
        np.save(result_path, np.array([int(np.array_equal(frame.numpy(), full))]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,height", [(2, 50), (2, 33), (3, 70)])
def test_gloo_row_strip_gather_reassembles_the_frame(tmp_path, world, height):
    port = 29600 + (os.getpid() + world * 7 + height) % 300
    result = str(tmp_path / "ok.npy")
    mp.spawn(_worker, args=(world, port, 40, height, result), nprocs=world, join=True)
    assert np.load(result)[0] == 1


def test_shard_rows_partition():
    sys.path.insert(0, os.path.join(ROOT, "all-is-cubes_b200"))
    from aicb200 import multi
    for h in (1, 15, 16, 17, 1080, 2160):
        for world in (1, 2, 4, 8):
            allrows = np.concatenate([multi.shard_rows(h, r, world) for r in range(world)])
            assert sorted(allrows.tolist()) == list(range(h))
