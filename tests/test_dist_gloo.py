"""world_size-2 gloo test (CPU) of the instance sharding + final gather used by bench.py --gpus N."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from trajoptmpcreference_b200 import dist as bdist


def test_shard_range_covers_everything():
    for total in (1, 7, 8192, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [bdist.shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_pack_roundtrip():
    x = torch.arange(2 * 4 * 5, dtype=torch.float64).reshape(2, 4, 5)
    u = torch.arange(2 * 2 * 4, dtype=torch.float64).reshape(2, 2, 4) + 100
    st = torch.tensor([[1, 1, 1, 6], [2, 3, 9, 4]], dtype=torch.int32)
    p = bdist.pack_results(x, u, st)
    x2, u2, st2 = bdist.unpack_results(p, 4, 2, 5, 4)
    assert torch.equal(x, x2) and torch.equal(u, u2) and torch.equal(st.to(torch.int64), st2)


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nx, nu, N = 4, 2, 6
    lo, hi = bdist.shard_range(total, world, rank)
    ids = torch.arange(lo, hi, dtype=torch.float64)
    # stand-in for the per-shard solve: every output row carries its global instance id
    x = ids[:, None, None] + torch.zeros((hi - lo, nx, N), dtype=torch.float64)
    u = -ids[:, None, None] + torch.zeros((hi - lo, nu, N - 1), dtype=torch.float64)
    st = torch.stack([torch.arange(lo, hi), torch.full((hi - lo,), rank)], dim=1)
    full = bdist.all_gather_results(bdist.pack_results(x, u, st), total)
    xg, ug, sg = bdist.unpack_results(full, nx, nu, N, 2)
    ok = (full.shape[0] == total and torch.equal(xg[:, 0, 0], torch.arange(total, dtype=torch.float64))
          and torch.equal(ug[:, 1, 2], -torch.arange(total, dtype=torch.float64)) and torch.equal(sg[:, 0], torch.arange(total)))
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [10, 11])
def test_gather_world2_gloo(total):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_bench_workload_shards():
    """bench.workload_goals: C4 at one GPU (default_rng(1)); C5 shards at N GPUs tile ONE default_rng(2) stream (contiguous, disjoint);
    `c5=True` at one GPU is shard 0 of that stream (scaling series compare like with like); strong scaling (--total) partitions the same
    stream into larger shards."""
    import sys
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, ROOT)
    import bench
    B = 64
    c4 = bench.workload_goals(1, 0, B)
    assert np.array_equal(c4, bench.goals(B, 1)) and np.all(c4[:, 6:] == 0) and np.all(np.abs(c4[:, :6]) <= 0.5)
    for world in (2, 4, 8):
        shards = [bench.workload_goals(world, r, B) for r in range(world)]
        assert np.array_equal(np.concatenate(shards), bench.goals(B * world, 2))
    # shard 0 of the stream is a prefix of every larger partition of it (numpy Generator.uniform fills row-major)
    assert np.array_equal(bench.workload_goals(1, 0, B, c5=True), bench.workload_goals(8, 0, B))
    assert np.array_equal(bench.workload_goals(2, 0, 4 * B)[:B], bench.workload_goals(8, 0, B))
    assert not np.array_equal(c4, bench.workload_goals(1, 0, B, c5=True))
