"""world_size-2 test (gloo, CPU) of the multi-GPU plumbing of bench.py: interleaved sharding by QP index, the
max-over-ranks timing reduction and the sum reductions — everything of the N>1 path that is not the CUDA kernel.
The per-rank "solve" is the CPU oracle on a tiny shard (test infrastructure standing in for the device)."""
import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, total, out_dir):
    import torch
    import torch.distributed as dist
    import bench
    import ssqp_b200 as S
    from oracle import ssqp_oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = bench.shard_indices(rank, world, total)
    c = S.workloads.config2(nb=total, N=30)
    r = O.solve_batch(c["V"], c["A"], c["G"], c["q"][idx], c["b"][idx], c["g"][idx], c["d"][idx], c["u"][idx], nthreads=1)
    my_ms = 10.0 * (rank + 1)
    ms, e2e, sums = bench.reduce_over_ranks(my_ms, 2 * my_ms, [float((r["status"] > 0).sum()), float(r["status"].sum())],
                                            world, torch.device("cpu"))
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), idx=idx, x=r["x"], status=r["status"], ms=ms, e2e=e2e, sums=sums)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_reductions(tmp_path):
    import torch.multiprocessing as mp
    import bench
    import ssqp_b200 as S
    from oracle import ssqp_oracle as O
    total, world = 10, 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, total, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / ("rank%d.npz" % r)) for r in range(world)]
    # shards partition the batch, interleaved
    allidx = np.concatenate([p["idx"] for p in parts])
    assert sorted(allidx.tolist()) == list(range(total))
    assert parts[0]["idx"].tolist() == list(range(0, total, 2)) and parts[1]["idx"].tolist() == list(range(1, total, 2))
    # the gathered result equals the single-process result
    c = S.workloads.config2(nb=total, N=30)
    ref = O.solve_batch(c["V"], c["A"], c["G"], c["q"], c["b"], c["g"], c["d"], c["u"], nthreads=1)
    X = np.empty_like(ref["x"]); st = np.empty_like(ref["status"])
    for p in parts:
        X[p["idx"]] = p["x"]; st[p["idx"]] = p["status"]
    assert np.array_equal(st, ref["status"]) and np.array_equal(X, ref["x"])
    # timing = max over ranks, counters = sum over ranks, identical on every rank
    for p in parts:
        assert float(p["ms"]) == 20.0 and float(p["e2e"]) == 40.0
        assert p["sums"].tolist() == [float((ref["status"] > 0).sum()), float(ref["status"].sum())]


def test_shard_indices_cover_uneven_batches():
    import bench
    for total, world in ((7, 2), (65536, 8), (5, 8)):
        parts = [bench.shard_indices(r, world, total) for r in range(world)]
        assert sorted(np.concatenate(parts).tolist()) == list(range(total))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
