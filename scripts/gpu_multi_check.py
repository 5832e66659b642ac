"""Multi-GPU check, launched with torchrun (one rank per GPU, NCCL): the P-rank k-eigenvalue run (bank exchange with
grouped send/recv, all-gather of bank counts, final all-reduce) and the P-rank fixed-source run must equal the
single-process oracle bit for bit."""
import json, os, sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import torch
import torch.distributed as dist
import util
from minimc_b200 import capi, decks, distributed

rank, local_rank, P = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
dev = torch.device("cuda", local_rank)
ok = True
for name in decks.KDECKS:
    text = decks.KDECKS[name]()
    flat = util.flat_from_xml(text)
    run = flat["run"]
    world = capi.World(capi.FlatWorld(**flat["world"]), device=local_rank)
    kd = distributed.KEigenvalue(world, util.product_source(flat), util.product_estimators(flat), run["histories"],
                                 run["inactive"], run["active"], tracking=run["tracking"])
    t0 = time.time()
    out = kd.solve(device=dev)
    dt = time.time() - t0
    if rank == 0:
        o_scores, o_squares, o_k, o_sizes, o_counters, status = util.oracle_problem(flat).keigenvalue()
        same = (np.array_equal(out["k_cycle"], o_k) and np.array_equal(out["scores"], o_scores)
                and np.array_equal(out["square_scores"], o_squares)
                and all(out["counters"][k] == v for k, v in o_counters.items()))
        ok = ok and same
        print(f"keig {name:12s} P={P} exact_vs_oracle={same} k={out['k_mean']:.5f}+-{out['k_std']:.5f} {dt*1e3:.0f} ms", flush=True)
# fixed source, sharded
flat = util.flat_from_xml(util.deck_text("fissile_slab", "delta", histories=200001))
world = capi.World(capi.FlatWorld(**flat["world"]), device=local_rank)
src, est = util.product_source(flat), util.product_estimators(flat)
first, count = distributed.shard(0, 200001, rank, P)
stream = torch.cuda.Stream(device=dev)
with torch.cuda.stream(stream):
    scores = torch.zeros(est.total_bins, dtype=torch.int64, device=dev)
    squares = torch.zeros_like(scores)
    counters = torch.zeros(len(capi.Counters._fields_), dtype=torch.int64, device=dev)
    world.fixed_source_run_device(src, est, 1, first, count, scores.data_ptr(), squares.data_ptr(), counters.data_ptr(),
                                  tracking=1, secondary_capacity=256, stream=stream.cuda_stream)
    distributed.allreduce_sum_(scores, squares, counters)
stream.synchronize()
if rank == 0:
    o_scores, o_squares, o_counters, _ = util.oracle_problem(flat).run(threads=8)
    same = np.array_equal(scores.cpu().numpy(), o_scores.astype(np.int64)) and np.array_equal(squares.cpu().numpy(), o_squares.astype(np.int64))
    ok = ok and same
    print(f"fixed-source fissile_slab P={P} exact_vs_oracle={same}", flush=True)
    print("MULTI-GPU PARITY", "OK" if ok else "FAILED", flush=True)
dist.destroy_process_group()
