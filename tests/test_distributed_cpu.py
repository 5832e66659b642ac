"""The N>1 path on CPU: world_size-2 gloo.  Each rank computes its shard of histories (with the oracle standing
in for the GPU, this being a CPU test), the integer tallies are all-reduced through minimc_b200.distributed, and
the result must equal the single-process run bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import util
from minimc_b200 import distributed


def test_shard_partitions_exactly():
    for n in (0, 1, 7, 100000, 2 ** 31 + 5):
        for p in (1, 2, 3, 8):
            ranges = [distributed.shard(11, n, r, p) for r in range(p)]
            assert ranges[0][0] == 11 and sum(c for _, c in ranges) == n
            for (a, ca), (b, _) in zip(ranges, ranges[1:]):
                assert a + ca == b


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world_size, port, n_histories, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    flat = util.flat_from_xml(util.deck_text("fissile_slab", "surface", histories=n_histories))
    first, count = distributed.shard(0, n_histories, rank, world_size)
    scores, squares, counters, status = util.oracle_problem(flat).run(count, first=first)
    t_scores = torch.from_numpy(scores.astype(np.int64))
    t_squares = torch.from_numpy(squares.astype(np.int64))
    t_counters = torch.tensor(list(counters.values()), dtype=torch.int64)
    distributed.allreduce_sum_(t_scores, t_squares, t_counters)
    if rank == 0:
        np.save(os.path.join(out_dir, "scores.npy"), t_scores.numpy())
        np.save(os.path.join(out_dir, "squares.npy"), t_squares.numpy())
        np.save(os.path.join(out_dir, "counters.npy"), t_counters.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_reduction_equals_single_run(tmp_path):
    n = 20001
    mp.spawn(_worker, args=(2, _free_port(), n, os.fspath(tmp_path)), nprocs=2, join=True)
    flat = util.flat_from_xml(util.deck_text("fissile_slab", "surface", histories=n))
    scores, squares, counters, _ = util.oracle_problem(flat).run(n)
    assert np.array_equal(np.load(tmp_path / "scores.npy"), scores.astype(np.int64))
    assert np.array_equal(np.load(tmp_path / "squares.npy"), squares.astype(np.int64))
    assert np.load(tmp_path / "counters.npy").tolist() == list(counters.values())
