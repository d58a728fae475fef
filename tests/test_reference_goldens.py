"""Replays the reference's own hot-path golden vectors (tests/golden/reference_goldens.json,
lifted from test/units/simulation_cpu/*.cpp by tests/golden/extract_goldens.py) against
 (1) the CPU oracle and (2) the CPU emulation of the CUDA kernel's code."""
import json
import os

import numpy as np
import pytest

import emu_lib
from oracle import pyoracle

U64_MAX = (1 << 64) - 1
HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_goldens.json")) as fh:
    GOLDENS = json.load(fh)["cases"]
STEP_CASES = [c for c in GOLDENS if c["steps"]]
RANK_CASES = [c for c in GOLDENS if c["rank_calls"] and "rev_ranks_expected1" in c]


def _inputs(case):
    lefs = [list(l) for l in case["lef_sets"]["lefs"]]
    rev = [l[0] for l in lefs]
    fwd = [l[1] for l in lefs]
    ep = [l[2] for l in lefs]
    for i in case["released"]:
        rev[i] = fwd[i] = ep[i] = U64_MAX
    bars = case["barriers"]
    return dict(
        start=case.get("start", 0), end=case.get("end", max(f for f in fwd if f != U64_MAX) + 1000),
        rev=rev, fwd=fwd, ep=ep, rr=case["rev_ranks"], fr=case["fwd_ranks"],
        rm=case["rev_moves"], fm=case["fwd_moves"],
        bar_pos=[b[0] for b in bars], bar_dir=[b[1] for b in bars],
        bar_active=[case["barriers_active"]] * len(bars),
        prob_bypass=case["prob_bypass"], pblock_major=case["pblock_major"],
        pblock_minor=case["pblock_minor"], rng_seed=case["rng_seed"])


def _check(case, out):
    for key, got in (("rev_moves_expected", "rm"), ("fwd_moves_expected", "fm"),
                     ("rev_moves_adjusted", "rm"), ("fwd_moves_adjusted", "fm"),
                     ("rev_collisions_expected", "rc"), ("fwd_collisions_expected", "fc")):
        if key in case:
            assert list(map(int, out[got])) == case[key], (case["name"], key)
    for key, got in (("rev_ranks_final", "rr"), ("fwd_ranks_final", "fr")):
        for slot, v in case.get(key, {}).items():
            assert int(out[got][int(slot)]) == v, (case["name"], key)


def test_golden_file_is_complete():
    names = " ".join(c["name"] for c in GOLDENS)
    for i in range(1, 13):
        assert f"Simulation {i:03d}" in names
    assert len(STEP_CASES) >= 21 and len(RANK_CASES) == 2


@pytest.mark.parametrize("case", STEP_CASES, ids=[c["name"][:40] for c in STEP_CASES])
def test_oracle_reproduces_reference_goldens(case):
    kw = _inputs(case)
    out = pyoracle.collision_steps(case["steps"], kw.pop("start"), kw.pop("end"), **kw)
    _check(case, out)


@pytest.mark.parametrize("case", [c for c in STEP_CASES if not c["released"]],
                         ids=[c["name"][:40] for c in STEP_CASES if not c["released"]])
@pytest.mark.parametrize("vthreads", [1, 3, 64])
def test_kernel_emulation_reproduces_reference_goldens(case, vthreads):
    # the kernel's epoch body never sees unbound LEFs (bind_lefs runs first), so the one golden
    # with released LEFs ("Simulation 006") is replayed against the oracle only
    kw = _inputs(case)
    steps = list(case["steps"])
    # the emulation fuses adjust+clamp and the two move corrections
    if "adjust" in steps and "clamp" not in steps:
        pytest.skip("adjust without clamp is not a kernel phase")
    if ("correct_lef_bar" in steps) != ("correct_primary" in steps):
        # a lone correction step is equivalent when the other one has nothing to do
        pass
    out = emu_lib.collision_steps(steps, kw.pop("start"), kw.pop("end"),
                                  virtual_threads=vthreads, **kw)
    assert out["fault"] == 0
    _check(case, out)


@pytest.mark.parametrize("case", RANK_CASES, ids=[c["name"][:30] for c in RANK_CASES])
def test_rank_lefs_goldens(case):
    for which in ("1", "2"):
        lefs = case["lef_sets"]["lefs" + which]
        rev = [l[0] for l in lefs]
        fwd = [l[1] for l in lefs]
        ep = [l[2] for l in lefs]
        n = len(lefs)
        rr, fr = pyoracle.rank_lefs(rev, fwd, ep, range(n), range(n), init_buffers=True)
        assert list(map(int, rr)) == case["rev_ranks_expected" + which]
        assert list(map(int, fr)) == case["fwd_ranks_expected" + which]
        rr, fr = emu_lib.rank_lefs(rev, fwd, ep, np.arange(n), np.arange(n))
        assert list(map(int, rr)) == case["rev_ranks_expected" + which]
        assert list(map(int, fr)) == case["fwd_ranks_expected" + which]
