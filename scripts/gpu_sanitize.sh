#!/bin/bash
# compute-sanitizer over small runs of every kernel family (memcheck, then racecheck and initcheck on the event-split path).
set -u
OUT=gpurun_out/${1:-san}
mkdir -p $OUT
cat > /tmp/san_run.py <<'PY'
import sys, tempfile
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import util
from minimc_b200 import capi, ce_decks, decks
d = tempfile.mkdtemp(); ce_decks.generate_tables(d, "small")
cases = [
    ("single_zone event", ce_decks.slab_deck(d, histories=3000, threads=1), dict(schedule=capi.SCHEDULE_EVENT, event_slots=1001)),
    ("single_zone event-only", ce_decks.slab_deck(d, histories=1500, threads=1), dict(schedule=capi.SCHEDULE_EVENT_ONLY, event_slots=301)),
    ("continuous_temperature event", ce_decks.continuous_temperature_deck(d, histories=2000, threads=1), dict()),
    ("free_gas_sphere event (fission deque)", ce_decks.free_gas_sphere_deck(d, histories=1500, threads=1), dict(secondary_capacity=256)),
    ("single_zone fused", ce_decks.slab_deck(d, histories=2000, threads=1), dict(schedule=capi.SCHEDULE_FUSED)),
    ("three_shells fused", decks.three_shells(histories=5000, estimators=decks.THREE_SHELL_ESTIMATORS), dict(secondary_capacity=256)),
    ("sensitivity_shells", decks.sensitivity_shells(histories=3000), dict(secondary_capacity=256)),
]
for name, text, opts in cases:
    drv = capi.Driver(text=text)
    drv.set_options(**opts)
    drv.solve()
    c = drv.counters()
    print(name, c["n_histories"], c["n_events"], "launches", drv.last_launches, flush=True)
PY
for tool in memcheck racecheck initcheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 7 python /tmp/san_run.py > $OUT/$tool.txt 2>&1
  echo "exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|launches" $OUT/$tool.txt | tail -12
  grep -m3 -A12 "=========.*\(Invalid\|Uninitialized\|Race\|hazard\)" $OUT/$tool.txt | head -40
done
