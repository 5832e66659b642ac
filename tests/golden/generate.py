"""Regenerates the golden fixtures in this directory from the REFERENCE'S OWN CODE
(oracle/_ref/ref_harness = /root/reference/src/*.cpp behind the shims in oracle/shim).

Run in the build container (needs /root/reference to have built oracle/_ref):
    python tests/golden/generate.py
Fixtures (all produced by the reference binary, none by this repo's code):
    <deck>__<tracking>.out      the .out text of Driver::Create(deck)->Solve() (minimc.cpp:17-21)
    <deck>__<tracking>.trace    per-event records of the first TRACE_HISTORIES histories
    <deck>.world.json           the World flattened in the reference containers' iteration order
    rng.json                    std::minstd_rand + generate_canonical streams (libstdc++ 13)
"""
import json
import os
import subprocess
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
from minimc_b200 import decks  # noqa: E402
from oracle import port_py  # noqa: E402
sys.path.insert(0, str(ROOT / "tests"))
import util  # noqa: E402

TRACE_HISTORIES = 48
TRACKING = {"surface": None, "delta": "cell delta"}


def deck_text(name, tracking):
    kw = {"estimators": decks.THREE_SHELL_ESTIMATORS} if name == "three_shells" else {}
    return decks.DECKS[name](tracking=TRACKING[tracking], **kw)


def sensitivity_fixtures(tmp, tables):
    """sens/<deck>__<tracking>.out: decks with <perturbations> and <sensitivities> (Perturbation.cpp, Sensitivity.cpp)."""
    (HERE / "sens").mkdir(exist_ok=True)
    for name, tracking, text in util.sensitivity_cases(tables):
        path = tmp / f"sens_{name}.xml"
        path.write_text(text)
        out, _ = port_py.ref_run(path)
        (HERE / "sens" / f"{name}__{tracking}.out").write_text(out)


def main():
    if not port_py.ref_available():
        raise SystemExit("oracle/_ref/ref_harness missing: run `make -C oracle ref` where /root/reference exists")
    tmp = Path(tempfile.mkdtemp())
    if sys.argv[1:] == ["sens"]:  # only the sensitivity fixtures
        from minimc_b200 import ce_decks
        ce_decks.generate_tables(tmp / "tables", "small")
        sensitivity_fixtures(tmp, tmp / "tables")
        return
    for name in decks.DECKS:
        for tracking in TRACKING:
            path = tmp / f"{name}.xml"
            path.write_text(deck_text(name, tracking))
            out, _ = port_py.ref_run(path)
            (HERE / f"{name}__{tracking}.out").write_text(out)
            trace = subprocess.run([os.fspath(port_py.REF_HARNESS), "trace", os.fspath(path), "0", str(TRACE_HISTORIES)],
                                   capture_output=True, text=True, check=True).stdout
            (HERE / f"{name}__{tracking}.trace").write_text(trace)
        path.write_text(deck_text(name, "surface"))
        (HERE / f"{name}.world.json").write_text(port_py.ref_dump(path))
    # continuous-energy / thermal-scattering decks on the synthetic "small" tables (minimc_b200/ce_decks.py); the
    # table files are regenerated bit-identically wherever the tests run, so absolute paths are rewritten to @TABLES@
    from minimc_b200 import ce_decks
    tables = tmp / "tables"
    ce_decks.generate_tables(tables, "small")
    (HERE / "ce").mkdir(exist_ok=True)
    for name, tracking, text in util.ce_cases(tables):
        path = tmp / f"ce_{name}.xml"
        path.write_text(text)
        out, _ = port_py.ref_run(path)
        (HERE / "ce" / f"{name}__{tracking}.out").write_text(out)
        trace = subprocess.run([os.fspath(port_py.REF_HARNESS), "trace", os.fspath(path), "0", str(util.CE_TRACE_HISTORIES)],
                               capture_output=True, text=True, check=True).stdout
        (HERE / "ce" / f"{name}__{tracking}.trace").write_text(trace)
    sensitivity_fixtures(tmp, tables)
    rng = {}
    for seed in (1, 0, 2147483647, 2147483648, 12345, 4294967297):
        lines = subprocess.run([os.fspath(port_py.REF_HARNESS), "rng", str(seed), "8"], capture_output=True, text=True,
                               check=True).stdout.split()
        rng[str(seed)] = {"canonical": lines[0::2], "state": [int(s) for s in lines[1::2]]}
    (HERE / "rng.json").write_text(json.dumps(rng, indent=1))
    print("golden fixtures regenerated in", HERE)


if __name__ == "__main__":
    main()
