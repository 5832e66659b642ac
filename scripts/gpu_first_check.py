"""First GPU check: CUDA path vs oracle port on every deck (traces + tallies), then a timing of M1."""
import sys, time, json
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from minimc_b200 import capi, decks
from oracle import port_py, flatten
import util

ok = True
# device math vs the box's libm
import ctypes
libm = ctypes.CDLL("libm.so.6")
libm.log.restype = ctypes.c_double; libm.log.argtypes = [ctypes.c_double]
libm.sin.restype = ctypes.c_double; libm.sin.argtypes = [ctypes.c_double]
libm.cos.restype = ctypes.c_double; libm.cos.argtypes = [ctypes.c_double]
rs = np.random.default_rng(7)
n = 200000
xl = np.concatenate([rs.random(n), 1 - rs.random(n) * 0.1, rs.random(n) * 1e3 + 1e-9])
got = capi.device_math(0, xl)[0]
want = np.array([libm.log(float(v)) for v in xl])
print("device log mismatches", int((got.view(np.uint64) != want.view(np.uint64)).sum()), "of", xl.size)
ok = ok and bool((got.view(np.uint64) == want.view(np.uint64)).all())
xs = np.concatenate([rs.random(n) * 2 * np.pi, rs.random(n) * 14 - 7, rs.random(n) * 1e4])
s_, c_ = capi.device_math(1, xs)
ws = np.array([libm.sin(float(v)) for v in xs]); wc = np.array([libm.cos(float(v)) for v in xs])
print("device sincos mismatches", int((s_.view(np.uint64) != ws.view(np.uint64)).sum()), int((c_.view(np.uint64) != wc.view(np.uint64)).sum()))
g2 = capi.device_math(2, xs)[0]; g3 = capi.device_math(3, xs)[0]
print("device sin/cos mismatches", int((g2.view(np.uint64) != ws.view(np.uint64)).sum()), int((g3.view(np.uint64) != wc.view(np.uint64)).sum()))
ok = ok and bool((g2.view(np.uint64) == ws.view(np.uint64)).all() and (g3.view(np.uint64) == wc.view(np.uint64)).all())
for name in decks.DECKS:
    for tracking in (None, "cell delta"):
        kw = {"estimators": decks.THREE_SHELL_ESTIMATORS} if name == "three_shells" else {}
        flat = flatten.flatten(decks.DECKS[name](tracking=tracking, **kw), is_text=True)
        prob = port_py.Problem(flat)
        world = util.product_world(flat)
        src = util.product_source(flat)
        est = util.product_estimators(flat)
        trk = flat["run"]["tracking"]
        # traces
        mine = world.trace(src, flat["run"]["seed"], 0, 500, tracking=trk, cap=1 << 18)
        ref = prob.trace(0, 500, cap=1 << 18)
        bad = 0; maxulp = 0
        if len(mine) != len(ref):
            print(name, tracking, "TRACE LENGTH", len(mine), len(ref)); ok = False
        for a, b in zip(mine, ref):
            if util.record_tuple(a) != util.record_tuple(b):
                bad += 1
                if bad < 4: print("   ", util.record_tuple(a), util.record_tuple(b))
            pa, da = util.record_vectors(a); pb, db = util.record_vectors(b)
            maxulp = max(maxulp, int(util.ulp_distance(da, db).max()), int(util.ulp_distance(pa, pb).max()))
        # tallies
        N = flat["run"]["histories"]
        t0 = time.time()
        sc, sq, cnt = world.fixed_source_run(src, est, flat["run"]["seed"], 0, N, tracking=trk, secondary_capacity=256)
        dt = time.time() - t0
        osc, osq, ocnt, st = prob.run()
        tb = int((sc != osc).sum() + (sq != osq).sum())
        cb = [k for k in ocnt if k in cnt and cnt[k] != ocnt[k]]
        print(f"{name:18s} {str(tracking):10s} records {len(ref)} bad {bad} max_ulp {maxulp} tally_bad {tb} counter_bad {cb} events {cnt['n_events']} {dt*1e3:.1f} ms")
        ok = ok and bad == 0 and tb == 0 and not cb and maxulp == 0
        world.close()
print("PARITY", "OK" if ok else "FAILED")

# timing M1
flat = flatten.flatten(decks.critical(), is_text=True)
world = util.product_world(flat); src = util.product_source(flat)
est = capi.Estimators([{"surface": 0}])
for N in (10**6, 10**7, 10**8, 10**9):
    for bps in (0, 1, 2):
        t0 = time.time()
        sc, sq, cnt = world.fixed_source_run(src, est, 1, 0, N, blocks_per_sm=bps)
        dt = time.time() - t0
        print(f"M1 N={N:.0e} blocks_per_sm={bps} {dt*1e3:.1f} ms  {N/dt:.3e} hist/s events/hist {cnt['n_events']/N:.4f}")
