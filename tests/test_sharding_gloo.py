"""The N > 1 path on CPU: two gloo ranks shard a batch by track id ranges,
transport their shards (with the CPU build of the core standing in for the
GPU engine), all-reduce the tallies and must reproduce the single process
run: identical per-track states and counts."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import harness
from pumas_b200.sharding import allreduce_tallies, shard_range, step_offset
from workloads import WORKLOADS

N_TOTAL = 96
SEED = 5


def tallies(states, events):
    sel = events == harness.EVENT_LIMIT_DISTANCE
    w = states["weight"][sel]
    hist, _ = np.histogram(np.log(states["energy"][sel]), bins=8,
                           range=(np.log(1e-3), np.log(1e3)))
    return np.concatenate([[w.sum(), (w * w).sum(), len(states), sel.sum()],
                           hist]).astype(np.float64)


def run_rank(rank, world, port, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dump, builder, kwargs = WORKLOADS["straight_mixed"]
    engine = harness.Port(dump)
    begin, end = shard_range(N_TOTAL, rank, world)
    w = builder(lambda name: {"StandardRock": 0}[name], N_TOTAL, **kwargs)
    states = w["states"][begin:end]
    out = engine.transport(w["settings"], w["geometry"], states, seed=SEED,
                           offset=begin)
    t = torch.from_numpy(tallies(out[0], out[1]))
    allreduce_tallies(t)
    np.save(os.path.join(result_dir, "states_%d.npy" % rank), out[0])
    if rank == 0:
        np.save(os.path.join(result_dir, "tally.npy"), t.numpy())
    dist.destroy_process_group()


def test_two_ranks_match_one(tmp_path):
    if not harness.have_port():
        pytest.skip("oracle/_port not built")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(run_rank, args=(2, port, str(tmp_path)), nprocs=2, join=True)

    dump, builder, kwargs = WORKLOADS["straight_mixed"]
    engine = harness.Port(dump)
    w = builder(lambda name: {"StandardRock": 0}[name], N_TOTAL, **kwargs)
    full = engine.transport(w["settings"], w["geometry"], w["states"],
                            seed=SEED, offset=0)
    shards = np.concatenate([np.load(tmp_path / ("states_%d.npy" % r))
                             for r in range(2)])
    for f in harness.STATE_FIELDS:
        assert np.array_equal(shards[f], full[0][f]), f
    expected = tallies(full[0], full[1])
    got = np.load(tmp_path / "tally.npy")
    assert np.array_equal(got[2:], expected[2:])          # counts: exact
    assert np.allclose(got[:2], expected[:2], rtol=1e-13)  # fp64 sums


def test_shard_ranges_partition():
    for n, world in ((10, 3), (96, 2), (7, 8), (1 << 21, 8)):
        ranges = [shard_range(n, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    ids = {step_offset(s, r, 4, 100) for s in range(3) for r in range(4)}
    assert len(ids) == 12 and min(ids) == 0 and max(ids) == 1100
