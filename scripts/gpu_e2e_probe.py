"""Where the host-buffer (e2e) path spends its time: Driver::Solve() with and without dropping the device world."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from minimc_b200 import capi  # noqa: E402

for workload in ("single_zone", "continuous_temperature"):
    table_dir = bench.make_tables(workload)
    n = 1 << 23
    drv = capi.Driver(text=bench.deck_text(workload, table_dir, n, 1))
    drv.set_options(device=0)
    for label, release in (("keep world", False), ("release world", True), ("keep world", False), ("release world", True)):
        times = []
        for _ in range(3):
            t0 = time.perf_counter()
            if release:
                drv.release_device()
            t1 = time.perf_counter()
            drv.solve()
            t2 = time.perf_counter()
            times.append((round(1e3 * (t1 - t0), 1), round(1e3 * (t2 - t1), 1)))
        print(workload, label, "(release ms, solve ms):", times, flush=True)
    drv.close()
