#!/usr/bin/env python
"""Benchmark of the history transport loop: particle histories per second.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload NAME]

Workloads (config.workload) -- BASELINE.json `configs`:
  single_zone (default)   configs[1] benchmarks/single_zone.xml as shipped: continuous-energy + S(a,b) thermal
                          scattering, 5 cm slab of H-in-H2O at 450 K, surface tracking, one current estimator with
                          103 x 101 = 10403 bins.  The deck's *.hdf5 are git-lfs pointers, so the
                          tables are the synthetic full-shape tables of minimc_b200/ce_decks.py ("data": "synthetic"),
                          the same files for the GPU and for the reference binary.
  multigroup_critical     configs[0] test/multigroup_critical.xml: 1-group infinite medium, c = 0.25, plus a
                          `current` estimator so the tally path is live; north_star's ">= 1e9 multigroup histories/s
                          per B200" is quoted on this one.  It is also measured (3 steps) inside the default run and
                          reported under "multigroup" in the same JSON line.
  continuous_temperature  configs[4]: cell delta tracking, linear T(x), 202 energy bins.

  keigenvalue_mg / keigenvalue_ce   the k-eigenvalue power iteration (BASELINE.json's metric is quoted on "active
                          cycles"): a step is ONE ACTIVE GENERATION of --histories-per-gpu source particles per GPU, the
                          warm-up steps are the inactive generations; the bank exchange between the ranks
                          (mmc_bank_exchange) and the final all-reduce are inside the timed region.  Both are also
                          measured briefly inside the default run and reported under "keigenvalue".

A step = one pass of the hot path over --histories-per-gpu histories per GPU (weak scaling, the default): rank r
transports histories [r*n, (r+1)*n) of the N*n total, then the integer tallies and counters are summed with ONE
all-reduce of the packed words (mmc_tally_allreduce: the library's own NCCL call).  `--scaling strong
--total-histories T` splits a fixed batch of T histories over the N GPUs instead (BASELINE configs[4] "scaled to 1e8").
`value`  : device-timed (CUDA events on the launch stream), tables resident in HBM.
`e2e`    : the same step through the reference-facing call with HOST buffers: the C++ host flattens its World again
           and uploads the tables from pinned host memory (mmc_world_update), then Driver::Solve() ->
           mmc_fixed_source_run (H2D of the bin boundaries, D2H of tallies + counters into a pinned mirror).
`--impl reference` times the reference's own C++ (oracle/_ref/ref_harness = /root/reference/src behind shims) on the
host cores on a bounded sample of the same deck; without a prebuilt oracle/_ref it says so.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, os.fspath(ROOT))

METRIC = "particle histories/sec"
WORKLOADS = {
    "single_zone": {
        "name": "single_zone (BASELINE configs[1], benchmarks/single_zone.xml as shipped): CE + S(a,b) H-in-H2O slab "
                "5 cm at 450 K, surface tracking, 10403-bin current estimator, synthetic full-shape tables (rank 10, "
                "partitions up to 97x18x294)",
        "histories_per_gpu": 1 << 25, "cpu_rate_guess": 4.0e4},
    "continuous_temperature": {
        "name": "continuous_temperature (BASELINE configs[4]): CE + S(a,b) slab, cell delta tracking, linear T(x) "
                "300..600 K, synthetic full-shape tables",
        "histories_per_gpu": 1 << 25, "cpu_rate_guess": 4.0e4},
    "multi_zone": {
        "name": "multi_zone (BASELINE configs[2], benchmarks/multi_zone.xml): CE + S(a,b), 13 slab segments at "
                "300..600 K between 14 planes, surface tracking, 202-bin current estimator, synthetic full-shape tables",
        "histories_per_gpu": 1 << 25, "cpu_rate_guess": 4.0e4},
    "broomstick": {
        "name": "broomstick (BASELINE configs[3], benchmarks/broomstick.xml): CE + S(a,b), cylinder r = 1e-6 along x, "
                "at most one collision per history, 184 x 239 cosine x energy bins, synthetic full-shape tables",
        "histories_per_gpu": 1 << 25, "cpu_rate_guess": 4.0e5},
    "multigroup_critical": {
        "name": "multigroup_critical (BASELINE configs[0], SURVEY M1): 1-group infinite medium c=0.25, surface tracking",
        "histories_per_gpu": 1 << 30, "cpu_rate_guess": 4.0e6},
    "keigenvalue_mg": {
        "name": "k-eigenvalue power iteration, SURVEY M1k: test/multigroup_critical.xml as <keigenvalue> with capture 0.5 / "
                "scatter 0.25 / fission 0.25 / nubar 2.43 (k_inf = 0.81 analytic), surface tracking",
        "histories_per_gpu": 1 << 25, "cpu_rate_guess": 2.0e6},
    "keigenvalue_ce": {
        "name": "k-eigenvalue power iteration, continuous energy: fuel sphere r = 12 cm (fissile heavy nuclide + oxygen, "
                "free-gas scattering) in a 6 cm water shell, synthetic tables, surface tracking",
        "histories_per_gpu": 1 << 22, "cpu_rate_guess": 0.0},
}
K_WORKLOADS = ("keigenvalue_mg", "keigenvalue_ce")


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def dram_traffic_per_step(workload, schedule, counters_per_rank):
    """dram__bytes_read+write of one step's kernels from the committed ncu captures (profiles/traffic.json), if any:
    a per-launch figure for the fused kernel (tables and tallies only: it does not scale with histories), a
    per-event figure for the flight + S(a,b) kernel pair of the event-split schedule."""
    p = ROOT / "profiles" / "traffic.json"
    if not p.exists():
        return None
    t = json.loads(p.read_text())
    if schedule == "event":
        per_event = t.get(workload + "/event_bytes_per_event")
        return None if per_event is None else int(per_event * counters_per_rank["n_events"])
    return t.get(workload)


class ClockSampler:
    QUERY = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.index)], stdout=fd, stderr=subprocess.DEVNULL)
            os.close(fd)
        except OSError:
            self.proc = None

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def stop(self) -> dict:
        """Median SM clock and throttle reasons of the samples taken between mark_begin() and mark_end() (the timed
        region; nvidia-smi runs since before the warm-up, sampling every 50 ms).  A region shorter than the sampling
        period is represented by the samples nearest to it."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)  # let a sample land after the region
        self.proc.terminate()
        self.proc.wait()
        import datetime
        rows = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in Path(self.path).read_text().splitlines():
            f = [v.strip() for v in line.split(",")]
            if len(f) < 9:
                continue
            try:
                stamp = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((stamp, float(f[1]), float(f[2]), [n for n, v in zip(names, f[5:9]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        os.unlink(self.path)
        begin, end = getattr(self, "t_begin", 0.0), getattr(self, "t_end", float("inf"))
        inside = [r for r in rows if begin <= r[0] <= end]
        if not inside and rows:  # shorter than the sampling period: the two samples that bracket the region
            before = [r for r in rows if r[0] < begin][-1:]
            after = [r for r in rows if r[0] > end][:1]
            inside = before + after
        sm = sorted(r[1] for r in inside)
        reasons = sorted({n for r in inside for n in r[3]})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max((r[2] for r in inside), default=None),
                "samples": len(sm), "reasons": reasons}


ROOFLINE_NOTE = {
    "fused": "bytes = 72*births + 144*events + 16*scores + 144*banked (BASELINE.md s4) per launch of the fused kernel, "
             "which keeps particle state in registers and the tables in L2/SMEM: real DRAM traffic is far below this "
             "model and the kernel is fp64-issue bound, see DESIGN.md",
    "event": "bytes = 72*births + 144*events + 16*scores + 144*banked (BASELINE.md s4) per step; a 'launch' is the "
             "whole pass loop of one step (flight, boundary and S(a,b) kernel per event, kernel_split has their CUDA-event "
             "times and counts); particle state really streams through HBM here (traffic = ncu DRAM bytes). The "
             "dominant S(a,b) kernel is not HBM-bound: every lane gathers from its own table row, so the kernel pays "
             "for the number of gather instructions (up to 32 L1 wavefronts each). With evaluated tables per constant "
             "cell temperature a reconstruction is one load; the kernel kind that holds only the direct samplers runs "
             "32 warps per SM at 64 registers and 54 % of issue slots (ncu, profiles/r02s_*), stalls led by "
             "long_scoreboard (L1/L2 latency of the gathers) and fixed-latency fp64 chains: see DESIGN.md s3, s4.1, "
             "s7 and profiles/",
}


def algorithmic_bytes(c: dict) -> int:
    # BASELINE.md section 4 / SURVEY.md 8(d)
    return 72 * c["n_births"] + 144 * c["n_events"] + 16 * c["n_scores"] + 144 * c["n_banked"]


def deck_text(workload: str, table_dir, histories: int, threads: int) -> str:
    from minimc_b200 import ce_decks, decks
    if workload == "multigroup_critical":
        return decks.critical(histories=histories, threads=threads, estimators=[{"name": "leakage", "surface": "sphere"}])
    if workload == "single_zone":
        return ce_decks.single_zone_benchmark_deck(table_dir, histories=histories, threads=threads)
    if workload == "continuous_temperature":
        return ce_decks.continuous_temperature_deck(table_dir, histories=histories, threads=threads, n_energy_bins=201)
    if workload == "multi_zone":
        return ce_decks.multi_zone_deck(table_dir, histories=histories, threads=threads, n_energy_bins=201)
    if workload == "broomstick":
        return ce_decks.broomstick_deck(table_dir, histories=histories, threads=threads, n_energy_bins=238,
                                        n_cosine_bins=182)
    raise SystemExit(f"unknown workload {workload}")


def k_deck_text(workload: str, table_dir, histories: int, inactive: int, active: int) -> str:
    from minimc_b200 import ce_decks, decks
    if workload == "keigenvalue_mg":
        return decks.k_infinite(histories=histories, threads=1, inactive=inactive, active=active)
    return ce_decks.fissile_sphere_keigenvalue_deck(table_dir, histories=histories, threads=1, inactive=inactive, active=active)


def make_tables(workload: str):
    """Synthetic full-shape tables in a temp dir (the generator is deterministic, exact arithmetic only)."""
    if workload in ("multigroup_critical", "keigenvalue_mg"):
        return None
    from minimc_b200 import ce_decks
    d = tempfile.mkdtemp(prefix="mmc_tables_")
    ce_decks.generate_tables(d, "full")
    return d


# --------------------------------------------------------------------- CPU arm
REF_HARNESS = ROOT / "oracle" / "_ref" / "ref_harness"


def time_reference_cpu(workload: str, table_dir, histories: int, threads: int):
    """Wall time of Driver::Solve() of the reference's own code on `threads` host threads (oracle/_ref).  This is the
    one place bench.py executes anything under oracle/ (the cpu_baseline / reference arm)."""
    if not (REF_HARNESS.exists() and os.access(REF_HARNESS, os.X_OK)):
        return None, "reference binary oracle/_ref/ref_harness was not built"
    with tempfile.TemporaryDirectory() as d:
        path = Path(d) / "deck.xml"
        path.write_text(deck_text(workload, table_dir, histories, threads))
        # stdout (the .out text and the reference's progress line) to /dev/null, BASELINE.md s3
        p = subprocess.run([os.fspath(REF_HARNESS), "run", os.fspath(path)], stdout=subprocess.DEVNULL,
                           stderr=subprocess.PIPE, text=True)
        if p.returncode != 0:
            return None, f"reference failed: {p.stderr.strip()[-200:]}"
        return float(p.stderr.strip().split("solve_seconds=")[-1]), "reference"


def cpu_baseline_block(workload: str, table_dir, seconds_per_run: float = 5.0) -> dict:
    """the reference's own Driver::Solve() on the host cores with 1 thread and with all of them, best
    of 3 each, on a bounded sample of the same deck and tables (~seconds_per_run of CPU work per run)."""
    cores = os.cpu_count() or 1
    w = WORKLOADS[workload]
    probe = max(int(w["cpu_rate_guess"] * 0.25), 1000)
    seconds, kind = time_reference_cpu(workload, table_dir, probe, cores)
    if seconds is None:
        return {"value": None, "unit": "histories/s", "cores": cores, "kind": "reference", "sample": kind}
    rate_all = probe / max(seconds, 1e-3)
    out = {}
    for label, threads, rate in (("all", cores, rate_all), ("one", 1, rate_all / cores * 1.5)):
        sample = max(int(rate * seconds_per_run), 1000)
        best = None
        for _ in range(3):
            seconds, kind = time_reference_cpu(workload, table_dir, sample, threads)
            if seconds is not None:
                best = seconds if best is None else min(best, seconds)
        out[label] = (sample / best if best else None, sample, threads)
    return {"value": out["all"][0], "unit": "histories/s", "cores": out["all"][2], "kind": kind,
            "sample": f"{out['all'][1]} histories of the same deck and tables, Driver::Solve() wall time, best of 3, "
                      f"stdout to /dev/null, {out['all'][2]} threads",
            "one_thread": {"value": out["one"][0], "cores": 1,
                           "sample": f"{out['one'][1]} histories, best of 3, 1 thread"}}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    w = WORKLOADS[args.workload]
    if args.workload in K_WORKLOADS:
        return run_reference_arm_keigenvalue(args, cores)
    table_dir = make_tables(args.workload)
    # a step = a bounded sample of the workload, sized (from a short probe) to ~10 s so K + W steps end within minutes
    probe = max(int(w["cpu_rate_guess"] * 0.5), 1000)
    seconds, kind = time_reference_cpu(args.workload, table_dir, probe, cores)
    if seconds is None:
        print(json.dumps({"impl": "reference", "unavailable": kind}), flush=True)
        return
    sample = max(int(probe / max(seconds, 1e-3) * 10.0), probe)
    for _ in range(args.warmup):
        time_reference_cpu(args.workload, table_dir, max(sample // 10, 1000), cores)
    total = 0.0
    for _ in range(args.steps):
        seconds, kind = time_reference_cpu(args.workload, table_dir, sample, cores)
        total += seconds
    value = sample * args.steps / total
    line = {
        "metric": METRIC, "value": value, "unit": "histories/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
        "config": {"workload": w["name"], "histories_per_step": sample, "rng": "std::minstd_rand (reference)"},
        "cpu_baseline": {"value": value, "unit": "histories/s", "cores": cores, "kind": kind,
                         "sample": f"{sample} histories per step, Driver::Solve() wall time, {cores} threads"},
        "e2e": {"value": value, "unit": "histories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_reference_arm_keigenvalue(args, cores):
    """The reference's KEigenvalue::Solve is a stub (KEigenvalue.cpp:36-62): there is no reference k-eigenvalue to time.
    The multigroup power iteration is timed on the oracle's CPU restatement of the same definition (oracle/port.cpp,
    sequential by definition: one thread); the continuous-energy one has no CPU implementation at all."""
    if args.workload != "keigenvalue_mg":
        print(json.dumps({"impl": "reference", "unavailable": "the reference has no k-eigenvalue (KEigenvalue::Solve is a "
                          "stub) and the oracle restates the multigroup power iteration only"}), flush=True)
        return
    sys.path.insert(0, os.fspath(ROOT / "tests"))
    import util
    n = 400_000
    flat = util.flat_from_xml(k_deck_text(args.workload, None, n, args.warmup, args.steps))
    t0 = time.perf_counter()
    util.oracle_problem(flat).keigenvalue()
    seconds = time.perf_counter() - t0
    value = n * (args.warmup + args.steps) / seconds
    print(json.dumps({
        "metric": METRIC + " (active cycles)", "value": value, "unit": "histories/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * seconds / (args.warmup + args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
        "config": {"workload": WORKLOADS[args.workload]["name"], "histories_per_step": n, "rng": "std::minstd_rand"},
        "cpu_baseline": {"value": value, "unit": "histories/s", "cores": 1, "kind": "port",
                         "sample": f"{args.warmup + args.steps} generations of {n} source particles, oracle/port.cpp, 1 thread"},
        "e2e": {"value": value, "unit": "histories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), flush=True)


# --------------------------------------------------------------------- GPU arm
def measure(args, workload, n_per_gpu, steps, warmup, ctx, cpu_baseline: bool, total_histories=None, rng_mode=0):
    """One fixed-source workload on this rank's GPU; returns the result dict on rank 0 (None elsewhere).
    total_histories: strong scaling (a fixed batch split over the ranks) instead of n_per_gpu per rank."""
    import torch
    import torch.distributed as dist

    from minimc_b200 import capi, distributed

    rank, local_rank, world_size, dev, stream, comm = ctx
    w = WORKLOADS[workload]
    table_dir = make_tables(workload)
    strong = total_histories is not None
    if not strong:
        total_histories = n_per_gpu * world_size
    drv = capi.Driver(text=deck_text(workload, table_dir, total_histories, 1))
    drv.set_options(device=local_rank, rng_mode=rng_mode)
    first, count = distributed.shard(0, total_histories, rank, world_size)
    bins = max(drv.total_bins, 1)
    n_counters = len(capi.Counters._fields_)
    # [scores | squares | counters]: one packed buffer, one all-reduce per step
    tally = torch.zeros(2 * bins + n_counters, dtype=torch.int64, device=dev)
    scores, squares, counters = tally[:bins], tally[bins:2 * bins], tally[2 * bins:]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step():
        tally.zero_()
        drv.run_device(first, count, scores.data_ptr(), squares.data_ptr(), counters.data_ptr(), stream.cuda_stream)
        comm.allreduce(tally.data_ptr(), tally.numel(), stream.cuda_stream)

    def sync_all():
        torch.cuda.synchronize()
        if world_size > 1:
            dist.barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # nvidia-smi needs a few hundred ms before its first sample: it runs through the warm-up
    for _ in range(warmup):
        step()
    sync_all()

    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(steps)] for _ in range(4)]
    starts, ends, kernel_starts, kernel_ends = ev
    sync_all()
    sampler.mark_begin()
    for k in range(steps):
        flush.fill_(k)  # evict L2 between timed iterations (untimed)
        starts[k].record(stream)
        tally.zero_()
        kernel_starts[k].record(stream)
        drv.run_device(first, count, scores.data_ptr(), squares.data_ptr(), counters.data_ptr(), stream.cuda_stream)
        kernel_ends[k].record(stream)
        comm.allreduce(tally.data_ptr(), tally.numel(), stream.cuda_stream)
        ends[k].record(stream)
    sync_all()
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    step_ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
    kernel_ms = sum(s.elapsed_time(e) for s, e in zip(kernel_starts, kernel_ends))
    t = torch.tensor([step_ms, kernel_ms], dtype=torch.float64, device=dev)
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms, kernel_ms = t.tolist()
    launches_per_step = drv.last_launches
    schedule = "fused" if launches_per_step == 1 else "event"
    kernel_split = None
    if schedule == "event":
        # one more (untimed) step with CUDA events around every kernel: the flight / S(a,b) split of the device time
        drv.set_options(device=local_rank, profile=1, rng_mode=rng_mode)
        step()
        sync_all()
        flight_ms, tsl_ms, boundary_ms = drv.last_kernel_ms()
        drv.set_options(device=local_rank, rng_mode=rng_mode)
        kernel_split = {"event_flight_kernel_ms": flight_ms, "event_tsl_kernel_ms": tsl_ms,
                        "event_boundary_kernel_ms": boundary_ms,
                        "dominant": "event_tsl_kernel" if tsl_ms >= flight_ms else "event_flight_kernel",
                        "dominant_share": max(flight_ms, tsl_ms) / max(flight_ms + tsl_ms + boundary_ms, 1e-9),
                        "launches": launches_per_step}
    c = dict(zip([n for n, _ in capi.Counters._fields_], counters.tolist()))
    assert c["n_histories"] == total_histories, (c["n_histories"], total_histories)
    assert c["n_lost"] == 0 and c["n_physics_errors"] == 0 and c["n_capacity_overflow"] == 0, c
    value = total_histories * steps / (step_ms * 1e-3)

    # ---- e2e: Driver::Solve() with HOST buffers, tables uploaded again every step, over all `steps`
    e2e_steps = steps
    drv.set_comm(comm)  # this rank's share of the batch, the packed all-reduce inside Solve()
    h2d = drv.table_bytes + 8 * bins  # tables + (upper bound of) the bin-boundary array
    d2h = 2 * 8 * bins + 8 * n_counters
    # every step: the World is flattened again and its tables uploaded from pinned host memory (refresh_device ->
    # mmc_world_update), Driver::Solve() runs with HOST tally buffers (device -> host read of the result inside)
    for _ in range(2):  # untimed warm-up of the host-buffer path
        drv.refresh_device()
        drv.solve()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        drv.refresh_device()
        drv.solve()
    sync_all()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = total_histories * e2e_steps / t.item()
    drv.close()
    if rank != 0:
        return None

    peak, peak_src = measured_peak_gbs()
    per_rank = {k: v / world_size for k, v in c.items()}
    bytes_per_launch = algorithmic_bytes(per_rank)
    launch_s = kernel_ms * 1e-3 / steps
    achieved = bytes_per_launch / launch_s / 1e9
    result = {
        "value": value, "ms_per_step": step_ms / steps, "clocks": clocks, "launches_per_step": launches_per_step,
        "e2e": {"value": e2e_value, "unit": "histories/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": dram_traffic_per_step(workload, schedule, per_rank), "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": bytes_per_launch, "kernel_ms_per_launch": 1e3 * launch_s,
                     "schedule": schedule, "kernel_split": kernel_split,
                     "note": ROOFLINE_NOTE[schedule]},
        "counters_per_step": c,
        "config": {"workload": w["name"],
                   "histories_per_gpu_per_step": total_histories / world_size if strong else n_per_gpu,
                   "histories_per_step": total_histories,
                   "rng": "counter (Philox-2x32-10 per particle)" if rng_mode else "minstd_compat (bit-exact)",
                   "estimator_bins": int(bins), "l2": "flushed between timed steps (256 MiB write, untimed)",
                   "parallelism": f"histories sharded over {world_size} GPU(s), one all-reduce of the packed integer "
                                  f"tallies per step (mmc_tally_allreduce)"},
    }
    if cpu_baseline:
        result["cpu_baseline"] = cpu_baseline_block(workload, table_dir)
    return result


def measure_keigenvalue(workload, n_per_gpu, steps, warmup, ctx, cpu_baseline: bool):
    """K-eigenvalue power iteration through the reference-facing call (the C++ host's KEigenvalue::Solve with HOST
    result buffers): `warmup` inactive + `steps` active generations of n_per_gpu source particles per GPU.  The timed
    region is the active cycles as the host sees them (every cycle ends with a device synchronisation; bank exchange,
    final all-reduce and read-back of the tallies inside), maximum over the ranks."""
    import torch
    import torch.distributed as dist

    from minimc_b200 import capi

    rank, local_rank, world_size, dev, stream, comm = ctx
    w = WORKLOADS[workload]
    table_dir = make_tables(workload)
    N = n_per_gpu * world_size
    drv = capi.Driver(text=k_deck_text(workload, table_dir, N, warmup, steps))
    drv.set_options(device=local_rank)
    drv.set_comm(comm)
    torch.cuda.synchronize()
    if world_size > 1:
        dist.barrier()
    drv.solve()
    inactive_s, active_s = drv.cycle_seconds()
    t = torch.tensor([active_s, inactive_s], dtype=torch.float64, device=dev)
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    active_s, inactive_s = t.tolist()
    k_mean, k_std, k_cycle = drv.keff()
    kc_mean, kc_std, _, exchange_ms = drv.k_collision()
    c = drv.counters()
    launches = drv.last_launches
    drv.close()
    if rank != 0:
        return None
    cycles = warmup + steps
    peak, _ = measured_peak_gbs()
    # counters cover all cycles and all ranks; the algorithmic bytes of one generation on one GPU
    bytes_per_generation = algorithmic_bytes({k: v / cycles / world_size for k, v in c.items()})
    result = {
        "workload": w["name"], "value": N * steps / active_s, "unit": "histories/s (active cycles)",
        "ms_per_step": 1e3 * active_s / steps, "steps": steps, "warmup": warmup, "histories_per_gpu_per_step": n_per_gpu,
        "k_eff": k_mean, "k_eff_std": k_std, "k_collision": kc_mean, "k_collision_std": kc_std,
        "bank_exchange_ms_per_cycle": exchange_ms / cycles, "exchange": "ncclAllGather of {bank size, status} + grouped "
        "ncclSend/ncclRecv of the sites a rank needs from its neighbours (mmc_bank_exchange)" if world_size > 1 else "none (one rank)",
        "roofline_frac": bytes_per_generation / (active_s / steps) / 1e9 / peak,
        "events_per_history": c["n_events"] / max(c["n_histories"], 1), "kernels_per_generation": 4,
        "e2e": {"value": N * steps / active_s, "unit": "histories/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 8 + 8 * len(capi.Counters._fields_),
                "note": "the timed call IS the host-buffer call (Driver::Solve()); the banks never leave the device"},
    }
    if workload == "keigenvalue_mg":
        result["k_analytic"] = 0.81
        result["k_sigma_from_analytic"] = abs(k_mean - 0.81) / k_std if k_std > 0 else None
    if cpu_baseline and workload == "keigenvalue_mg":
        # the reference has no k-eigenvalue (stub): the CPU figure is the oracle's restatement of the same definition,
        # one thread, on a bounded sample (this is the cpu_baseline leg: the one place bench.py executes oracle/)
        sys.path.insert(0, os.fspath(ROOT / "tests"))
        import util
        flat = util.flat_from_xml(k_deck_text(workload, None, 400_000, 1, 3))
        t0 = time.perf_counter()
        util.oracle_problem(flat).keigenvalue()
        seconds = time.perf_counter() - t0
        result["cpu_baseline"] = {"value": 400_000 * 4 / seconds, "unit": "histories/s", "cores": 1, "kind": "port",
                                  "sample": "4 generations of 400000 source particles, oracle/port.cpp orc_keigenvalue_run, 1 thread"}
    return result


def run_gpu_arm(args):
    # NCCL and the CUDA runtime may print banners on the C-level stdout; the contract is ONE JSON line there.
    # Everything but the final line goes to stderr.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist

    from minimc_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available() or capi.load().mmc_device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: minimc_b200 has no CPU transport path")
    torch.cuda.set_device(local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    # a real (non-default) stream: its handle is what the C ABI launches on and what the CUDA events time
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    # this rank's communicator: the library's own NCCL calls (mmc_comm_create / mmc_tally_allreduce /
    # mmc_bank_exchange); torch.distributed only carries the unique id, the barrier and the max of the timings
    comm = capi.Comm.from_torch(local_rank) if world_size > 1 else capi.Comm(1, 0, None, local_rank)
    ctx = (rank, local_rank, world_size, dev, stream, comm)

    n_per_gpu = args.histories_per_gpu or WORKLOADS[args.workload]["histories_per_gpu"]
    want_cpu = world_size == 1 and not args.no_cpu_baseline
    if args.workload in K_WORKLOADS:
        k = measure_keigenvalue(args.workload, n_per_gpu, args.steps, args.warmup, ctx, cpu_baseline=want_cpu)
        if rank == 0:
            line = {
                "metric": METRIC + " (active cycles)", "value": k["value"], "unit": "histories/s", "n_gpus": world_size,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": k["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": k["workload"], "histories_per_gpu_per_step": n_per_gpu,
                           "rng": "minstd_compat (bit-exact)", "l2": "inputs (banks of 64-byte sites) larger than L2",
                           "parallelism": f"every generation sharded over {world_size} GPU(s), bank exchange per generation"},
                "e2e": k["e2e"], "gpu_launches": (args.steps + args.warmup) * 5, "keigenvalue": k,
                "roofline": {"bound": "hbm", "achieved": k["roofline_frac"] * measured_peak_gbs()[0], "peak": measured_peak_gbs()[0],
                             "unit": "GB/s", "frac": k["roofline_frac"], "traffic": None,
                             "note": "bytes = 72*births + 144*events + 16*scores + 144*banked per generation; fused kernel, "
                                     "particle state in registers: a model figure (DESIGN.md s4.2)"},
            }
            if "cpu_baseline" in k:
                line["cpu_baseline"] = k["cpu_baseline"]
            sys.stdout.flush()
            os.write(real_stdout, (json.dumps(line) + "\n").encode())
        comm.close()
        if world_size > 1:
            dist.destroy_process_group()
        return
    strong_total = args.total_histories if args.scaling == "strong" else None
    main = measure(args, args.workload, n_per_gpu, args.steps, args.warmup, ctx, cpu_baseline=want_cpu,
                   total_histories=strong_total, rng_mode=1 if args.rng == "counter" else 0)
    extra = None
    if args.workload != "multigroup_critical" and not args.no_multigroup:
        extra = measure(args, "multigroup_critical", WORKLOADS["multigroup_critical"]["histories_per_gpu"], 3, 3, ctx,
                        cpu_baseline=False)
    small = keig = counter = None
    if args.workload == "single_zone" and not args.no_extras and strong_total is None and not args.histories_per_gpu:
        # the same workload with MMC_RNG_COUNTER (a Philox-2x32-10 stream per particle; statistically equivalent)
        counter = measure(args, "single_zone", n_per_gpu, 3, 3, ctx, cpu_baseline=False, rng_mode=1)
        # the deck as shipped (benchmarks/single_zone.xml: 10^6 histories) -- a batch that ends before the GPU is full
        small = measure(args, "single_zone", 1_000_000, 5, 3, ctx, cpu_baseline=False)
        # BASELINE.json's metric names active k-eigenvalue cycles: both power iterations, briefly
        keig = {name: measure_keigenvalue(name, WORKLOADS[name]["histories_per_gpu"], 5, 2, ctx, cpu_baseline=want_cpu)
                for name in K_WORKLOADS}
    if rank == 0:
        line = {
            "metric": METRIC, "value": main["value"], "unit": "histories/s", "n_gpus": world_size, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True,
            "scaling": "strong" if strong_total is not None else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": main["config"], "clocks": main["clocks"],
            "e2e": main["e2e"], "gpu_launches": args.steps * main["launches_per_step"], "roofline": main["roofline"],
            "counters_per_step": main["counters_per_step"],
        }
        if "cpu_baseline" in main:
            line["cpu_baseline"] = main["cpu_baseline"]
        if extra is not None:
            line["multigroup"] = {"workload": extra["config"]["workload"], "value": extra["value"], "unit": "histories/s",
                                  "steps": 3, "ms_per_step": extra["ms_per_step"], "e2e": extra["e2e"],
                                  "roofline_frac": extra["roofline"]["frac"],
                                  "histories_per_gpu_per_step": extra["config"]["histories_per_gpu_per_step"]}
        if small is not None:
            line["deck_as_shipped"] = {"workload": "single_zone at the deck's own 10^6 histories per GPU and step",
                                       "value": small["value"], "unit": "histories/s", "steps": 5,
                                       "ms_per_step": small["ms_per_step"], "e2e": small["e2e"],
                                       "fraction_of_large_batch_rate": small["value"] / main["value"],
                                       "launches_per_step": small["launches_per_step"]}
        if counter is not None:
            line["counter_rng"] = {"rng": counter["config"]["rng"], "value": counter["value"], "unit": "histories/s",
                                   "steps": 3, "ms_per_step": counter["ms_per_step"], "e2e": counter["e2e"],
                                   "fraction_of_minstd_rate": counter["value"] / main["value"]}
        if keig is not None:
            line["keigenvalue"] = keig
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    comm.close()
    if world_size > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="single_zone", choices=sorted(WORKLOADS))
    ap.add_argument("--histories-per-gpu", type=int, default=0)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--total-histories", type=int, default=100_000_000,
                    help="--scaling strong: the batch split over the GPUs (BASELINE configs[4]: 1e8 histories)")
    ap.add_argument("--rng", default="minstd", choices=["minstd", "counter"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-multigroup", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the 10^6-history and k-eigenvalue lines of the default run")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
