"""CPU, world_size 2 over gloo: the N>1 plumbing of the path — static file sharding (no data-path
collective) and the single parameter-block broadcast at init."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist

    from basic_pitch_b200 import ICASSP_2022_MODEL_PATH, engine, weights

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. the parameter block: rank 0 holds the real tensors, the others start from garbage
        w = weights.load(ICASSP_2022_MODEL_PATH)
        flat = np.concatenate([w[k].ravel() for k in weights.EXPECTED_SHAPES])
        t = torch.from_numpy(flat.copy()) if rank == 0 else torch.full((flat.size,), float("nan"))
        dist.broadcast(t, src=0)
        assert np.array_equal(t.numpy(), flat)
        # 2. static sharding of 10 files of uneven length: every file exactly once, order preserved
        lengths = [220500, 44100, 1, 36164, 36165, 500000, 22050, 3840, 80000, 220500]
        lo, hi = engine.shard_range(len(lengths), rank, world)
        mine = list(range(lo, hi))
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        assert sum(gathered, []) == list(range(len(lengths)))
        # 3. per-rank totals reduce to the global total (what bench.py reports as whole-job units)
        tot = torch.tensor([float(sum(lengths[i] for i in mine))])
        dist.all_reduce(tot)
        assert tot.item() == float(sum(lengths))
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


@pytest.mark.parametrize("n,world", [(0, 1), (1, 4), (7, 2), (10, 8), (10000, 8), (13, 5)])
def test_shard_range_partitions(n, world):
    from basic_pitch_b200.engine import shard_range

    parts = [shard_range(n, r, world) for r in range(world)]
    assert parts[0][0] == 0 and parts[-1][1] == n
    for (a, b), (c, d) in zip(parts, parts[1:]):
        assert b == c and b >= a and d >= c
    sizes = [b - a for a, b in parts]
    assert max(sizes) - min(sizes) <= 1
