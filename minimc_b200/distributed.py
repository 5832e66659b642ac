"""Multi-GPU plumbing: one process per GPU, torch.distributed over NCCL (gloo in CPU tests).

Fixed source.  Histories are independent (history i is fully determined by seed + i, FixedSource.cpp:61), so the path
shards with NO data-path collective: rank r of P owns the contiguous range [r*N/P, (r+1)*N/P) and the only exchange is
one all-reduce of the integer tallies and counters at the end (the multi-GPU form of
`solver_estimator_set += worker_estimator_set.get()`, FixedSource.cpp:31-33).  The sums are exact 64-bit integers, so
the result is bit-identical for any P.

K-eigenvalue.  One exchange step per generation (DESIGN.md "k-eigenvalue"): rank r transports source indices
[r*N/P, (r+1)*N/P); its fission bank comes back ordered by (source index, ordinal), so the global bank is the
concatenation of the ranks' banks.  An all-gather of the P bank counts gives every rank the global offsets and
k = M / N; the comb resampling (source i <- site floor(i*M/N)) then tells each rank which contiguous global range of
sites its next sources need, and only the parts of that range that live on other ranks move -- grouped point-to-point
sends/receives (ncclSend/ncclRecv under NCCL), typically a few sites to each neighbour because the per-rank counts
differ from M/P only by the statistical imbalance.  Tallies and counters are all-reduced once at the end.  Every step
is order-based, so k of every cycle, the banks and the tallies are bit-identical for any P.
"""
from __future__ import annotations

SITE_BYTES = 64


def shard(first: int, n: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous history range of `rank`: [first + rank*n//P, first + (rank+1)*n//P)."""
    lo = rank * n // world_size
    hi = (rank + 1) * n // world_size
    return first + lo, hi - lo


def allreduce_sum_(*tensors, group=None) -> None:
    """In-place SUM all-reduce of int64 tally / counter tensors (device tensors under NCCL)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)


# ------------------------------------------------------------------------------------- fission-bank exchange
def needed_sites(n_total: int, m_total: int, rank: int, world_size: int) -> tuple[int, int]:
    """Global fission-site range [first, first + count) that the next sources of `rank` are drawn from: source i uses
    site floor(i * M / N) for i in the rank's source range."""
    i_lo, n = shard(0, n_total, rank, world_size)
    if n == 0 or m_total == 0:
        return 0, 0
    j_lo = i_lo * m_total // n_total
    j_hi = (i_lo + n - 1) * m_total // n_total
    return j_lo, j_hi - j_lo + 1


def exchange_plan(counts: list[int], n_total: int, rank: int) -> dict:
    """Who sends which part of its ordered local bank to whom.  counts[r] = sites banked by rank r this generation.

    Returns {"need": (first, count), "send": [(peer, local_start, count)...], "recv": [(peer, dst_start, count)...]}
    with peers in rank order; the rank's own contribution appears in both lists with peer == rank (a local copy)."""
    world_size = len(counts)
    m_total = sum(counts)
    offsets = [0]
    for c in counts:
        offsets.append(offsets[-1] + c)
    need_first, need_count = needed_sites(n_total, m_total, rank, world_size)
    send, recv = [], []
    for peer in range(world_size):
        # what `peer` needs from me
        p_first, p_count = needed_sites(n_total, m_total, peer, world_size)
        lo, hi = max(p_first, offsets[rank]), min(p_first + p_count, offsets[rank + 1])
        if hi > lo:
            send.append((peer, lo - offsets[rank], hi - lo))
        # what I need from `peer`
        lo, hi = max(need_first, offsets[peer]), min(need_first + need_count, offsets[peer + 1])
        if hi > lo:
            recv.append((peer, lo - need_first, hi - lo))
    return {"need": (need_first, need_count), "send": send, "recv": recv, "m_total": m_total}


def exchange_bank(local_bank, plan: dict, rank: int, group=None):
    """Moves the planned site ranges between ranks.  `local_bank` is a uint8 tensor [>= counts[rank] * 64] holding this
    rank's ordered fission bank; returns a uint8 tensor holding global sites plan["need"] in order."""
    import torch
    import torch.distributed as dist
    need_first, need_count = plan["need"]
    out = torch.empty(max(need_count, 1) * SITE_BYTES, dtype=torch.uint8, device=local_bank.device)
    ops, keep = [], []
    for peer, start, count in plan["send"]:
        piece = local_bank[start * SITE_BYTES:(start + count) * SITE_BYTES]
        if peer == rank:
            continue
        piece = piece.contiguous()
        keep.append(piece)
        ops.append(dist.P2POp(dist.isend, piece, peer, group=group))
    for peer, dst, count in plan["recv"]:
        view = out[dst * SITE_BYTES:(dst + count) * SITE_BYTES]
        if peer == rank:
            src = next(s for p, s, c in plan["send"] if p == rank)
            view.copy_(local_bank[src * SITE_BYTES:(src + count) * SITE_BYTES])
        else:
            ops.append(dist.P2POp(dist.irecv, view, peer, group=group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return out


class KEigenvalue:
    """K-eigenvalue power iteration over P GPUs (one process each).  `world` is a capi.World on this rank's device,
    `source` / `estimators` the capi descriptors; every rank must call solve() with the same arguments."""

    def __init__(self, world, source, estimators, batchsize: int, inactive: int, active: int, *, tracking=0,
                 bank_capacity_factor: float = 4.0, group=None):
        self.world, self.source, self.estimators = world, source, estimators
        self.batchsize, self.inactive, self.active = int(batchsize), int(inactive), int(active)
        self.tracking, self.group = tracking, group
        self.bank_capacity_factor = bank_capacity_factor

    def solve(self, device=None, stream=None) -> dict:
        import numpy as np
        import torch
        import torch.distributed as dist
        from . import capi

        distributed = dist.is_available() and dist.is_initialized()
        rank = dist.get_rank(self.group) if distributed else 0
        P = dist.get_world_size(self.group) if distributed else 1
        device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        N = self.batchsize
        first, n_local = shard(0, N, rank, P)
        capacity = int(self.bank_capacity_factor * max(n_local, 1)) + 1024
        # a real (non-default) stream: handle 0 would mean "the library's own stream" to the C ABI
        s = stream if stream is not None and stream.cuda_stream != 0 else torch.cuda.Stream(device=device)
        s.wait_stream(torch.cuda.current_stream(device))
        handle = s.cuda_stream
        with torch.cuda.stream(s):
            bank_source = torch.zeros(max(n_local, 1) * SITE_BYTES, dtype=torch.uint8, device=device)
            bank_fission = torch.zeros(capacity * SITE_BYTES, dtype=torch.uint8, device=device)
            n_out = torch.zeros(1, dtype=torch.int64, device=device)
            errors = torch.zeros(1, dtype=torch.int64, device=device)
            bins = max(self.estimators.total_bins, 1)
            scores = torch.zeros(bins, dtype=torch.int64, device=device)
            squares = torch.zeros(bins, dtype=torch.int64, device=device)
            counters = torch.zeros(len(capi.Counters._fields_), dtype=torch.int64, device=device)
            # KEigenvalue.cpp:29-33: source.Sample(s), s = 1 .. batchsize; this rank's part
            self.world.source_bank_sample(self.source, 1, first, n_local, bank_source.data_ptr(), stream=handle)
        k_cycle, bank_sizes = [], []
        with torch.cuda.stream(s):  # the C ABI launches, torch copies and the NCCL calls all order on `s`
            for cycle in range(self.inactive + self.active):
                self.world.generation_run(
                    bank_source.data_ptr(), n_local, self.estimators, cycle >= self.inactive, bank_fission.data_ptr(),
                    capacity, n_out.data_ptr(), scores.data_ptr(), squares.data_ptr(), counters.data_ptr(),
                    tracking=self.tracking, stream=handle)
                if distributed and P > 1:
                    gathered = [torch.zeros_like(n_out) for _ in range(P)]
                    dist.all_gather(gathered, n_out, group=self.group)
                    counts = [int(v) for v in torch.stack(gathered).cpu().flatten().tolist()]
                else:
                    counts = [int(n_out.item())]
                # decided alike on every rank (each holds all counts): a rank-local check would leave the other
                # ranks waiting in the exchange for a peer that has raised
                for r in range(P):
                    if counts[r] > int(self.bank_capacity_factor * max(shard(0, N, r, P)[1], 1)) + 1024:
                        raise capi.MinimcError(capi.ERR_CAPACITY,
                                               f"fission bank overflow on rank {r}: raise bank_capacity_factor")
                M = sum(counts)
                k_cycle.append(M / N)
                bank_sizes.append(M)
                if M == 0:
                    raise capi.MinimcError(capi.ERR_PHYSICS, "the fission chain died out (empty fission bank)")
                plan = exchange_plan(counts, N, rank)
                piece = exchange_bank(bank_fission, plan, rank, group=self.group) if P > 1 else bank_fission
                need_first, need_count = plan["need"]
                self.world.bank_resample(piece.data_ptr(), need_first, need_count, M, N, first, n_local,
                                         bank_source.data_ptr(), errors.data_ptr(), stream=handle)
                if P > 1:
                    piece.record_stream(s)
        with torch.cuda.stream(s):
            allreduce_sum_(scores, squares, counters, errors, group=self.group)
        s.synchronize()
        if int(errors.item()):
            raise capi.MinimcError(capi.ERR_INVALID, "bank resampling read outside its slice")
        c = dict(zip([n for n, _ in capi.Counters._fields_], counters.tolist()))
        if c["n_lost"]:
            raise capi.MinimcError(capi.ERR_LOST_PARTICLE, f"{c['n_lost']} particle(s) outside every cell")
        if c["n_physics_errors"]:
            raise capi.MinimcError(capi.ERR_PHYSICS, f"{c['n_physics_errors']} unreachable-branch event(s)")
        if c["n_capacity_overflow"]:
            raise capi.MinimcError(capi.ERR_CAPACITY, f"{c['n_capacity_overflow']} capacity overflow(s)")
        k = np.array(k_cycle)
        act = k[self.inactive:]
        k_mean = float(act.mean()) if len(act) else 0.0
        k_std = float(act.std(ddof=1) / np.sqrt(len(act))) if len(act) > 1 else 0.0
        return {"k_cycle": k, "k_mean": k_mean, "k_std": k_std, "bank_sizes": bank_sizes,
                "scores": scores.cpu().numpy().astype(np.float64)[:self.estimators.total_bins],
                "square_scores": squares.cpu().numpy().astype(np.float64)[:self.estimators.total_bins], "counters": c}
