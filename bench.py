#!/usr/bin/env python
"""Benchmark of the history transport loop: particle histories per second.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (config.workload): BASELINE.json configs[0] / SURVEY.md M1 = test/multigroup_critical.xml scaled --
one-group infinite medium, c = 0.25, surface tracking, minstd_compat RNG (bit-exact with the reference), plus a
`current` estimator on the sphere so the tally path is live.  configs[1..4] (continuous-energy + S(a,b) decks)
cannot be run by anyone here: every *.hdf5 they need is a git-lfs pointer; north_star's numeric
target (>= 1e9 multigroup histories/s per B200) is quoted on this multigroup workload.

A step = one pass of the hot path over HISTORIES_PER_GPU histories per GPU (weak scaling): rank r transports
histories [r*n, (r+1)*n) of the N*n total, then the integer tallies and counters are all-reduced (NCCL).
`value`  : device-timed (CUDA events on the launch stream), tables resident in HBM.
`e2e`    : the same step through the reference-facing C ABI with HOST buffers: flatten -> mmc_world_create (H2D
           of the tables) -> mmc_fixed_source_run (H2D of bin boundaries, D2H of tallies + counters) -> destroy.
`--impl reference` times the reference's own C++ (oracle/_ref/ref_harness = /root/reference/src behind shims) on
the host cores; without a prebuilt oracle/_ref it times the oracle port instead and says so.
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
sys.path.insert(0, os.fspath(ROOT / "tests"))

HISTORIES_PER_GPU = 1 << 30          # per step
CPU_SAMPLE_HISTORIES = 20_000_000    # bounded sample of the same workload for the CPU baseline
METRIC = "particle histories/sec"
WORKLOAD = "multigroup_critical (BASELINE configs[0], SURVEY M1): 1-group infinite medium c=0.25, surface tracking"


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def dram_traffic_per_launch():
    """dram__bytes_read+write per launch of the fused kernel from the committed ncu capture, if any."""
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        return json.loads(p.read_text()).get("fixed_source_kernel_dram_bytes_per_launch")
    return None


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
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
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=fd, stderr=subprocess.DEVNULL)
            os.close(fd)
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.proc.wait()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in Path(self.path).read_text().splitlines():
            f = [v.strip() for v in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def algorithmic_bytes(c: dict) -> int:
    # BASELINE.md section 4 / SURVEY.md 8(d)
    return 72 * c["n_births"] + 144 * c["n_events"] + 16 * c["n_scores"] + 144 * c["n_banked"]


def critical_deck(histories: int, threads: int) -> str:
    from minimc_b200 import decks
    return decks.critical(histories=histories, threads=threads, estimators=[{"name": "leakage", "surface": "sphere"}])


# --------------------------------------------------------------------- CPU arm
def time_reference_cpu(histories: int, threads: int):
    """Wall time of Driver::Solve() of the reference's own code on `threads` host threads."""
    from oracle import port_py
    deck = critical_deck(histories, threads)
    if port_py.ref_available():
        with tempfile.TemporaryDirectory() as d:
            path = Path(d) / "critical.xml"
            path.write_text(deck)
            _, seconds = port_py.ref_run(path)
        return seconds, "reference"
    import util
    prob = util.oracle_problem(util.flat_from_xml(deck))
    t0 = time.perf_counter()
    prob.run(threads=threads)
    return time.perf_counter() - t0, "port"


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # a step = a bounded sample of the workload, sized so K + W steps end within a few minutes
    sample = 5_000_000
    for _ in range(args.warmup):
        time_reference_cpu(sample // 10, cores)
    total = 0.0
    kind = "reference"
    for _ in range(args.steps):
        seconds, kind = time_reference_cpu(sample, cores)
        total += seconds
    value = sample * args.steps / total
    line = {
        "metric": METRIC, "value": value, "unit": "histories/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
        "config": {"workload": WORKLOAD, "histories_per_step": sample, "rng": "std::minstd_rand (reference)"},
        "cpu_baseline": {"value": value, "unit": "histories/s", "cores": cores, "kind": kind,
                         "sample": f"{sample} histories per step, Driver::Solve() wall time, {cores} threads"},
        "e2e": {"value": value, "unit": "histories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------- GPU arm
def run_gpu_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import util
    from minimc_b200 import capi, distributed

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available() or capi.load().mmc_device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: minimc_b200 has no CPU transport path")
    torch.cuda.set_device(local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    n_per_gpu = args.histories_per_gpu
    total_histories = n_per_gpu * world_size
    flat = util.flat_from_xml(critical_deck(min(total_histories, 2 ** 31 - 1), 1))
    fw = capi.FlatWorld(**flat["world"])
    world = capi.World(fw, device=local_rank)
    src, est = util.product_source(flat), util.product_estimators(flat)
    first, count = distributed.shard(0, total_histories, rank, world_size)

    scores = torch.zeros(max(est.total_bins, 1), dtype=torch.int64, device=dev)
    squares = torch.zeros_like(scores)
    counters = torch.zeros(len(capi.Counters._fields_), dtype=torch.int64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    # a real (non-default) stream: its handle is what the C ABI launches on and what the CUDA events time
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    def step():
        scores.zero_(), squares.zero_(), counters.zero_()
        world.fixed_source_run_device(src, est, 1, first, count, scores.data_ptr(), squares.data_ptr(),
                                      counters.data_ptr(), stream=stream.cuda_stream)
        distributed.allreduce_sum_(scores, squares, counters)

    def sync_all():
        torch.cuda.synchronize()
        if world_size > 1:
            dist.barrier()

    for _ in range(args.warmup):
        step()
    sync_all()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    kernel_starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    kernel_ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    sync_all()
    for k in range(args.steps):
        flush.fill_(k)  # evict L2 between timed iterations (untimed)
        starts[k].record(stream)
        scores.zero_(), squares.zero_(), counters.zero_()
        kernel_starts[k].record(stream)
        world.fixed_source_run_device(src, est, 1, first, count, scores.data_ptr(), squares.data_ptr(),
                                      counters.data_ptr(), stream=stream.cuda_stream)
        kernel_ends[k].record(stream)
        distributed.allreduce_sum_(scores, squares, counters)
        ends[k].record(stream)
    sync_all()
    clocks = sampler.stop() if rank == 0 else None
    step_ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
    kernel_ms = sum(s.elapsed_time(e) for s, e in zip(kernel_starts, kernel_ends))
    t = torch.tensor([step_ms, kernel_ms], dtype=torch.float64, device=dev)
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms, kernel_ms = t.tolist()
    c = dict(zip([n for n, _ in capi.Counters._fields_], counters.tolist()))
    assert c["n_histories"] == total_histories, (c["n_histories"], total_histories)
    assert c["n_lost"] == 0 and c["n_physics_errors"] == 0 and c["n_capacity_overflow"] == 0
    value = total_histories * args.steps / (step_ms * 1e-3)

    # ---- e2e: the plugin call with HOST buffers (world upload + run + tallies back), every step
    e2e_steps = max(1, min(args.steps, 3))
    wd_bytes = sum(getattr(fw, name).nbytes for name in capi.FlatWorld.FIELDS)
    h2d = wd_bytes + 8 * 0  # tables (this deck has no bin-boundary arrays)
    d2h = 2 * 8 * max(est.total_bins, 1) + 8 * len(capi.Counters._fields_)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        w2 = capi.World(fw, device=local_rank)
        h_scores, h_squares, h_counters = w2.fixed_source_run(src, est, 1, first, count)
        w2.close()
    sync_all()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = total_histories * e2e_steps / t.item()

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        per_rank = {k: v / world_size for k, v in c.items()}
        bytes_per_launch = algorithmic_bytes(per_rank)
        launch_s = kernel_ms * 1e-3 / args.steps
        achieved = bytes_per_launch / launch_s / 1e9
        cpu_value, cpu = None, None
        if world_size == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            seconds, kind = time_reference_cpu(CPU_SAMPLE_HISTORIES, cores)
            cpu_value = CPU_SAMPLE_HISTORIES / seconds
            cpu = {"value": cpu_value, "unit": "histories/s", "cores": cores, "kind": kind,
                   "sample": f"{CPU_SAMPLE_HISTORIES} histories of the same deck, Driver::Solve() wall time, "
                             f"{cores} threads"}
        line = {
            "metric": METRIC, "value": value, "unit": "histories/s", "n_gpus": world_size, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "histories_per_gpu_per_step": n_per_gpu, "rng": "minstd_compat (bit-exact)",
                       "tracking": "surface", "estimators": 1, "l2": "flushed between timed steps (256 MiB write, untimed)",
                       "parallelism": f"histories sharded over {world_size} GPU(s), final all-reduce of integer tallies"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "histories/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps},
            "gpu_launches": args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": dram_traffic_per_launch(), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bytes_per_launch, "kernel_ms_per_launch": 1e3 * launch_s,
                         "note": "bytes = 72*births + 144*events + 16*scores + 144*banked (BASELINE.md s4); the fused "
                                 "kernel keeps particle state in registers, so real DRAM traffic is far below this"},
            "counters_per_step": c,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world_size > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--histories-per-gpu", type=int, default=HISTORIES_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
