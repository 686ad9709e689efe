"""CPU: the N > 1 host logic on the gloo backend with world_size 2 (no GPU, rendezvous on 127.0.0.1)."""
import importlib
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = importlib.import_module("3dhumanposeestimation_b200.dist")
    assert d.world() == (rank, world)
    lo, hi = d.shard_range(257, rank, world)
    covered = torch.zeros(257)
    covered[lo:hi] = 1
    dist.all_reduce(covered)
    ms = 10.0 + 5.0 * rank                      # rank 1 is slower
    out = dict(shard=(lo, hi), covered_once=bool((covered == 1).all()), max_ms=d.max_over_ranks(ms),
               total=d.sum_over_ranks(hi - lo), thr=d.job_throughput(256.0, ms), seed=d.rank_seed(42, rank))
    # the data-parallel gradient exchange: bucketed all-reduce of a range of the flat gradient buffer, tail first
    grad = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    d.allreduce_range(grad, 100, 900, 256)
    want = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    want[100:900] = torch.arange(100, 900, dtype=torch.float32) * 3          # 1x + 2x over the two ranks
    out["allreduce_ok"] = bool(torch.equal(grad, want))
    # the bf16 wire format of the same exchange (Trainer(grad_wire="bf16")): small integers are exact in bf16
    g16 = (torch.arange(1000, dtype=torch.float32) % 64 * (rank + 1)).bfloat16()
    d.allreduce_range(g16, 0, 1000, 300)
    out["allreduce_bf16_ok"] = bool(torch.equal(g16.float(), torch.arange(1000, dtype=torch.float32) % 64 * 3))
    ret[rank] = out
    dist.destroy_process_group()


def test_sharding_and_timing_reductions_world2():
    world, port = 2, _free_port()
    ret = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    r0, r1 = ret[0], ret[1]
    assert r0["shard"] == (0, 129) and r1["shard"] == (129, 257)
    assert r0["covered_once"] and r1["covered_once"]
    assert r0["max_ms"] == r1["max_ms"] == 15.0           # max over ranks, never the local time
    assert r0["total"] == r1["total"] == 257
    assert abs(r0["thr"] - 512.0 / 15e-3) < 1e-6          # all units / slowest rank
    assert r0["seed"] != r1["seed"]
    assert r0["allreduce_ok"] and r1["allreduce_ok"]
    assert r0["allreduce_bf16_ok"] and r1["allreduce_bf16_ok"]


def test_single_process_defaults():
    d = importlib.import_module("3dhumanposeestimation_b200.dist")
    assert d.world() == (0, 1)
    assert d.shard_range(10, 0, 1) == (0, 10)
    assert [d.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert d.max_over_ranks(3.5) == 3.5 and d.job_throughput(100, 50.0) == 2000.0
    assert d.bucket_ranges(10, 100, 40) == [(60, 100), (20, 60), (10, 20)]      # model tail first
    assert d.bucket_ranges(5, 5, 8) == []


def test_gradient_sections_are_coalesced_without_gaps_or_repeats():
    """dist.SectionCoalescer (the Trainer's exchange schedule): announced sections arrive tail first; exchanges cover every
    element exactly once, none is shorter than the threshold except the one that reaches the head."""
    d = importlib.import_module("3dhumanposeestimation_b200.dist")
    n = 1000
    cuts = [1000, 950, 700, 690, 400, 390, 120, 0]                 # backward announces [950,1000), [700,950), ...
    for min_elems in (0, 100, 300, 10_000):
        c = d.SectionCoalescer(min_elems)
        sent = []
        for hi, lo in zip(cuts[:-1], cuts[1:]):
            sent += c.add(lo, hi)
        assert c.flush() == []                                      # the head was announced: nothing is left pending
        cover = torch.zeros(n)
        for a, b in sent:
            cover[a:b] += 1
        assert bool((cover == 1).all()), (min_elems, sent)
        assert all(b - a >= min_elems or a == 0 for a, b in sent), (min_elems, sent)
        if min_elems == 0:
            assert sent == list(zip(cuts[1:], cuts[:-1]))
        if min_elems == 10_000:
            assert sent == [(0, 1000)]
    c = d.SectionCoalescer(500)
    assert c.add(900, 1000) == [] and c.add(100, 200) == [(900, 1000)]     # not adjacent: the pending range goes first
    assert c.flush() == [(100, 200)] and c.add(5, 5) == []
