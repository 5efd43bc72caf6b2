"""The N > 1 path on CPU: 2 gloo ranks, clips sharded contiguously, one all-gather, shard-invariant result."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from use_b200.distributed import sample_sharded, shard_range


def test_shard_range_covers_everything_once():
    for n in (0, 1, 7, 8, 2048):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _fake_sample(y_local, clip0):
    """Stand-in for ScoreModel.sample: depends on the data and on the GLOBAL clip index (like the Philox streams)."""
    idx = torch.arange(clip0, clip0 + y_local.shape[0], dtype=y_local.dtype)[:, None]
    return y_local * 2.0 + idx


def _worker(rank, world, port, n_clips, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    y = torch.arange(n_clips * 5, dtype=torch.float32).reshape(n_clips, 5)
    out = sample_sharded(_fake_sample, y)
    ref = _fake_sample(y, 0)
    q.put((rank, bool(torch.equal(out, ref))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [8, 7, 1])  # even split, ragged, and B < world (rank 1 owns nothing: ADVICE r1)
def test_two_rank_gloo_shard_and_gather(n_clips):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_clips, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, True), (1, True)]


def test_single_process_passthrough():
    y = torch.ones(3, 4)
    assert torch.equal(sample_sharded(_fake_sample, y), _fake_sample(y, 0))
