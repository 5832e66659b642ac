"""More than one GPU, one process per GPU, NO Python in the ranks: the C++ host's CLI (minimc_b200/runminimc_b200 =
minimc.cpp:10-25 on the GPU path) started once per GPU with MMC_WORLD_SIZE / MMC_RANK / MMC_DEVICE /
MMC_COMM_ID_FILE.  Rank 0 writes the NCCL unique id to the file, every rank builds its communicator
(mmc_comm_create), Driver::Solve() shards the batch -- fixed source: histories [r*N/P, (r+1)*N/P) and ONE packed
all-reduce of the integer tallies; k-eigenvalue: every generation sharded, mmc_bank_exchange (all-gather of the bank
sizes, grouped ncclSend / ncclRecv of the few sites a rank needs from its neighbours) and mmc_bank_resample.

Contract: every step is order-based or an exact integer sum, so the .out file, k of every cycle, every bank size and
the collision estimator of k are IDENTICAL for P = 1, 2, 4, 8.  Skipped on a box with one GPU."""
import os
import subprocess
from pathlib import Path

import pytest

from minimc_b200 import capi, ce_decks, decks

pytestmark = pytest.mark.gpu

CLI = Path(capi.LIB_PATH).parent / "runminimc_b200"


def _gpu_count():
    try:
        return capi.load().mmc_device_count()
    except OSError:
        return 0


def _rank_counts():
    n = _gpu_count()
    return [p for p in (2, 3, 4, 8) if p <= n]


def _run_cli(deck: Path, nranks: int, extra_env=None):
    """Runs the CLI on `nranks` GPUs; returns (rank 0's .out text, the cycle lines of every rank's stdout)."""
    procs = []
    id_file = deck.parent / f"nccl_id_{nranks}"
    for rank in range(nranks):
        env = dict(os.environ)
        env.update(extra_env or {})
        if nranks > 1:
            env.update(MMC_WORLD_SIZE=str(nranks), MMC_RANK=str(rank), MMC_DEVICE=str(rank), MMC_COMM_ID_FILE=os.fspath(id_file))
        procs.append(subprocess.Popen([os.fspath(CLI), os.fspath(deck)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = []
    for p in procs:
        try:
            out, err = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        assert p.returncode == 0, err[-2000:]
        outs.append([line for line in out.splitlines() if line.startswith(("cycle ", "k-effective"))])
    return deck.with_suffix(".out").read_text(), outs


@pytest.fixture(scope="module")
def tables(tmp_path_factory):
    d = tmp_path_factory.mktemp("tables_multi")
    ce_decks.generate_tables(d, "small")
    return d


def _cases(tables):
    return {
        "mg_k_slab": decks.k_slab(histories=50_000, inactive=2, active=5),
        "mg_k_infinite_delta": decks.k_infinite(histories=30_001, inactive=1, active=4, tracking="cell delta"),
        "ce_fissile_sphere_k": ce_decks.fissile_sphere_keigenvalue_deck(tables, histories=20_000, inactive=2, active=4),
        "mg_three_shells_fixed": decks.DECKS["three_shells"](histories=100_003, estimators=decks.THREE_SHELL_ESTIMATORS),
        "ce_single_zone_fixed": ce_decks.slab_deck(tables, histories=30_000, threads=1),
    }


@pytest.mark.skipif(_gpu_count() < 2, reason="needs at least two GPUs")
@pytest.mark.parametrize("name", ["mg_k_slab", "mg_k_infinite_delta", "ce_fissile_sphere_k", "mg_three_shells_fixed",
                                  "ce_single_zone_fixed", "mg_k_slab/counter", "ce_single_zone_fixed/counter"])
def test_cli_ranks_reproduce_the_single_process_run(tables, tmp_path, name):
    # "/counter": MMC_RNG_MODE=counter, the Philox build of every kernel (a particle's stream does not depend on the rank
    # that runs it, so the result is the same for any number of GPUs in this mode too)
    name, _, rng = name.partition("/")
    extra_env = {"MMC_RNG_MODE": "counter"} if rng else None
    text = _cases(tables)[name]
    single_dir = tmp_path / "p1"
    single_dir.mkdir()
    (single_dir / "deck.xml").write_text(text)
    ref_out, ref_lines = _run_cli(single_dir / "deck.xml", 1, extra_env)
    if rng:  # other random numbers than the minstd build
        plain_dir = tmp_path / "p1_minstd"
        plain_dir.mkdir()
        (plain_dir / "deck.xml").write_text(text)
        assert _run_cli(plain_dir / "deck.xml", 1)[0] != ref_out
    if "_k" in name:
        assert any(line.startswith("cycle ") for line in ref_lines[0])
    for nranks in _rank_counts():
        d = tmp_path / f"p{nranks}"
        d.mkdir()
        (d / "deck.xml").write_text(text)
        out, lines = _run_cli(d / "deck.xml", nranks, extra_env)
        assert out == ref_out, f"{name}: .out differs at {nranks} ranks"
        for rank_lines in lines:  # every rank prints the same k of every cycle, bank sizes and collision estimator
            assert rank_lines == ref_lines[0], f"{name}: cycle lines differ at {nranks} ranks"
