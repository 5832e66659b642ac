"""GPU debugging aid: CUDA continuous-energy path vs the reference binary (oracle/_ref/ref_harness) on every CE deck:
event traces record by record, then tallies."""
import os, sys, tempfile, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from minimc_b200 import capi, ce_decks
from oracle import port_py
import util

size = sys.argv[1] if len(sys.argv) > 1 else "small"
n_trace = int(sys.argv[2]) if len(sys.argv) > 2 else 300
n_hist = int(sys.argv[3]) if len(sys.argv) > 3 else 20000
d = tempfile.mkdtemp()
ce_decks.generate_tables(d, size)
ok = True
for name, fn in ce_decks.CE_DECKS.items():
    for tracking in (None, "cell delta"):
        kw = {"histories": n_hist, "threads": os.cpu_count()}
        if name in ("single_zone", "free_gas_sphere"):
            kw["tracking"] = tracking
        elif name == "multi_zone":
            kw["tracking"] = tracking or "surface"
        elif tracking is not None:
            continue
        text = fn(d, **kw)
        path = os.path.join(d, f"{name}.xml")
        open(path, "w").write(text)
        ref = port_py.ref_trace(path, 0, n_trace)
        drv = capi.Driver(path)
        drv.set_options(secondary_capacity=256)
        try:
            mine = drv.trace(0, n_trace, cap=1 << 20)
        except capi.MinimcError as e:
            print(name, tracking, "TRACE ERROR", e); ok = False; mine = []
        bad = 0; first_bad = None; max_ulp = 0
        for i, (a, b) in enumerate(zip(mine, ref)):
            ta = (int(a.history), int(a.particle), int(a.event), int(a.cell), int(a.surface), int(a.rng_state))
            tb = (b["history"], b["particle"], b["event"], b["cell"], b["surface"], b["rng_state"])
            if a.event == 0: ta = ta[:3] + (None,) + ta[4:]; tb = tb[:3] + (None,) + tb[4:]
            pa = np.array(list(a.position) + list(a.direction) + [a.energy]); pb = np.array(list(b["position"]) + list(b["direction"]) + [b["energy"]])
            ulp = int(util.ulp_distance(pa, pb).max())
            if ta != tb or ulp:
                bad += 1
                if first_bad is None:
                    first_bad = (i, ta, tb, ulp, pa.tolist(), pb.tolist())
            if ta == tb: max_ulp = max(max_ulp, ulp)
        print(f"{name:24s} {str(tracking):10s} records {len(mine)}/{len(ref)} bad {bad} max_ulp(on matching records) {max_ulp}")
        if first_bad:
            print("   first mismatch:", first_bad[:4]); print("     mine", first_bad[4]); print("     ref ", first_bad[5])
            if first_bad[0] > 0:
                a = mine[first_bad[0] - 1]; print("     previous record (mine): event", a.event, "E", a.energy, "pos", list(a.position))
        ok = ok and bad == 0 and len(mine) == len(ref)
        # tallies
        t0 = time.time()
        try:
            sc, sq = drv.solve()
        except capi.MinimcError as e:
            print("   SOLVE ERROR", e); ok = False; continue
        dt = time.time() - t0
        out, secs = port_py.ref_run(path)
        _, ref_out = port_py.parse_out(out)
        c = drv.counters()
        mine_out = port_py.parse_out(drv.output())[1]
        diff = sum(1 for est in ref_out for k in ("mean", "std dev") for x, y in zip(ref_out[est][k], mine_out[est][k]) if x != y)
        total = sum(len(ref_out[est]["mean"]) * 2 for est in ref_out)
        print(f"   tallies: {diff}/{total} printed values differ; events/hist {c['n_events']/n_hist:.2f}; gpu {dt*1e3:.0f} ms, reference {secs*1e3:.0f} ms ({os.cpu_count()} threads)")
print("CE PARITY", "OK" if ok else "FAILED")
