"""The reference-side adapter, compiled (INTEGRATION.md way B): oracle/_ref/gpu_adapter is the
REFERENCE'S OWN Driver / World / Source / EstimatorSet classes (/root/reference/src, unmodified, behind the shims) with
one new Driver subclass, GpuFixedSource (oracle/adapter/GpuFixedSource.cpp), whose Solve() flattens the reference's
World and calls mmc_world_create + mmc_fixed_source_run instead of FixedSource::Solve's worker pool
(FixedSource.cpp:22-36, Driver.cpp:19-35).  What it prints -- `batchsize` + the reference's EstimatorSet::to_string(),
minimc.cpp:20-21 -- must be the golden .out the reference binary wrote, byte for byte."""
import os
import subprocess
from pathlib import Path

import pytest

import util
from minimc_b200 import ce_decks

pytestmark = pytest.mark.gpu

ADAPTER = Path(__file__).resolve().parent.parent / "oracle" / "_ref" / "gpu_adapter"


def _run(deck_path):
    if not (ADAPTER.exists() and os.access(ADAPTER, os.X_OK)):
        pytest.skip("oracle/_ref/gpu_adapter was not built (needs /root/reference at build time)")
    p = subprocess.run([os.fspath(ADAPTER), os.fspath(deck_path)], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    assert "histories" in p.stderr  # the device counters it reports
    return p.stdout


@pytest.mark.parametrize("name,tracking", util.DECK_CASES)
def test_reference_driver_through_the_adapter_writes_the_golden_out(tmp_path, name, tracking):
    path = tmp_path / "deck.xml"
    path.write_text(util.deck_text(name, tracking))
    assert _run(path) == (util.GOLDEN / f"{name}__{tracking}.out").read_text()


@pytest.mark.parametrize("name,tag", [c for c in util.CE_CASE_IDS if c in util.CE_EXACT])
def test_reference_driver_through_the_adapter_continuous_energy(tmp_path, name, tag):
    tables = tmp_path / "tables"
    ce_decks.generate_tables(tables, "small")
    text = dict(((n, t), x) for n, t, x in util.ce_cases(tables))[(name, tag)]
    path = tmp_path / "deck.xml"
    path.write_text(text)
    assert _run(path) == (util.GOLDEN / "ce" / f"{name}__{tag}.out").read_text()
