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


# ---------------------------------------------------------------- k-eigenvalue bank exchange
def _site_ids(first, count):
    """`count` fake 64-byte sites whose first 8 bytes hold their global index."""
    a = np.zeros((count, 8), np.int64)
    a[:, 0] = np.arange(first, first + count)
    return a


def test_exchange_plan_covers_every_needed_site():
    """Every rank's needed range is tiled exactly by what the ranks send it, for balanced, skewed and empty banks."""
    for n_total in (1, 7, 1000, 12345):
        for counts in ([5], [3, 4], [0, 9, 1], [100, 0, 0, 7], [400, 300, 200, 100, 50, 25, 12, 6], [1000, 1, 1]):
            P = len(counts)
            offsets = np.concatenate([[0], np.cumsum(counts)])
            for rank in range(P):
                plan = distributed.exchange_plan(counts, n_total, rank)
                first, count = plan["need"]
                i_lo, n = distributed.shard(0, n_total, rank, P)
                if n:
                    js = [i * sum(counts) // n_total for i in range(i_lo, i_lo + n)]
                    assert first == js[0] and first + count - 1 == js[-1]
                covered = 0
                for peer, dst, c in plan["recv"]:
                    assert dst == covered  # in rank order, contiguous
                    assert offsets[peer] <= first + dst and first + dst + c <= offsets[peer + 1]
                    covered += c
                assert covered == count
                # symmetric: what I send to q is what q expects from me
                for peer, start, c in plan["send"]:
                    theirs = [r for r in distributed.exchange_plan(counts, n_total, peer)["recv"] if r[0] == rank]
                    assert len(theirs) == 1 and theirs[0][2] == c


def _exchange_worker(rank, world_size, port, counts, n_total, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    offsets = np.concatenate([[0], np.cumsum(counts)])
    local = torch.from_numpy(_site_ids(int(offsets[rank]), counts[rank]).view(np.uint8).reshape(-1).copy())
    # what a rank does after mmc_generation_run: all-gather the counts, plan, exchange
    mine = torch.tensor([counts[rank]], dtype=torch.int64)
    gathered = [torch.zeros_like(mine) for _ in range(world_size)]
    dist.all_gather(gathered, mine)
    assert [int(t) for t in gathered] == list(counts)
    plan = distributed.exchange_plan([int(t) for t in gathered], n_total, rank)
    piece = distributed.exchange_bank(local, plan, rank)
    first, count = plan["need"]
    got = piece.numpy()[:count * 64].view(np.int64).reshape(count, 8)[:, 0]
    np.save(os.path.join(out_dir, f"piece{rank}.npy"), np.concatenate([[first], got]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("counts,n_total", [((700, 650), 1500), ((10, 900, 3), 1000)])
def test_gloo_bank_exchange_delivers_the_global_order(tmp_path, counts, n_total):
    """world_size 2 and 3 over gloo: after the exchange every rank holds exactly the global site range its next
    sources are drawn from, in global order -- the precondition of mmc_bank_resample."""
    P = len(counts)
    mp.spawn(_exchange_worker, args=(P, _free_port(), counts, n_total, os.fspath(tmp_path)), nprocs=P, join=True)
    M = sum(counts)
    for rank in range(P):
        data = np.load(tmp_path / f"piece{rank}.npy")
        first, got = int(data[0]), data[1:]
        i_lo, n = distributed.shard(0, n_total, rank, P)
        js = np.array([i * M // n_total for i in range(i_lo, i_lo + n)])
        assert first == js[0] and np.array_equal(got, np.arange(js[0], js[-1] + 1))


def test_c_abi_exchange_plan_equals_the_python_plan():
    """mmc_exchange_plan (the host arithmetic of mmc_bank_exchange; no device, no NCCL) against
    distributed.exchange_plan on random bank sizes, including empty ranks, one rank and more ranks than sources."""
    import random
    from minimc_b200 import capi
    rng = random.Random(20261018)
    for P in (1, 2, 3, 8):
        for _ in range(40):
            n_total = rng.choice([1, 5, 1000, 12345, 10**9 + 7])
            counts = [rng.choice([0, rng.randrange(0, 3 * max(n_total // P, 1) + 1)]) for _ in range(P)]
            if sum(counts) == 0:
                counts[rng.randrange(P)] = 1
            for rank in range(P):
                a = capi.exchange_plan(counts, n_total, rank)
                b = distributed.exchange_plan(counts, n_total, rank)
                assert a["need"] == tuple(b["need"]) and a["m_total"] == b["m_total"]
                assert a["send"] == [tuple(x) for x in b["send"]] and a["recv"] == [tuple(x) for x in b["recv"]]
