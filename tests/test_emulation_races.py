"""Race detection for the kernel source without a GPU: the SPMD emulation (tests/emu/
emu_mt_main.cpp -- every virtual thread of the CTA is an OS thread running the whole per-cell
program of sim_core.hpp, CTA barriers are pthread barriers) under ThreadSanitizer. A shared word
written by one virtual thread and touched by another with no CTA barrier in between is reported
whatever the interleaving was: the class of bug (`S.rng_pos` read by every thread while the leader
advances it, rank slots read while a neighbour swaps them) that showed up on the B200 as a
1-in-500-cells flake and that compute-sanitizer's racecheck traced. Accesses that
are shared by design go through MB_SHARED_* (relaxed atomics here, plain accesses on the device).
Each run must also reproduce the serial emulation bit for bit. Both RNG modes."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import emu_lib
from common import make_case, results_equal
from modle_b200 import buildutil, host

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu")
CXXFLAGS = ["-O1", "-g", "-fsanitize=thread", "-march=x86-64-v3", "-ffp-contract=off",
            "-fno-fast-math", "-std=c++17", "-pthread", "-Wno-unknown-pragmas"]

CASES = {
    "defaults": dict(size=1_500_000, ncells=2, nbar=25, target_contact_density=0.05),
    "frac_pblock_bypass": dict(size=1_500_000, ncells=1, nbar=40, target_contact_density=0.05,
                               lef_bar_major_collision_pblock=0.7,
                               lef_bar_minor_collision_pblock=0.2,
                               probability_of_extrusion_unit_bypass=0.3),
    "no_bypass": dict(size=1_500_000, ncells=1, nbar=30, target_contact_density=0.05,
                      probability_of_extrusion_unit_bypass=0.0),
    "always_bypass": dict(size=1_500_000, ncells=1, nbar=30, target_contact_density=0.05,
                          probability_of_extrusion_unit_bypass=1.0),
    "high_collision": dict(size=1_200_000, ncells=1, nbar=90, target_contact_density=0.03,
                           number_of_lefs_per_mbp=80.0, probability_of_extrusion_unit_bypass=0.01),
    "skip_burnin_epochs": dict(size=2_000_000, ncells=1, nbar=30, skip_burnin=1,
                               stopping_criterion=1, target_simulation_epochs=120),
    "tad_only_no_1d": dict(size=1_500_000, ncells=1, nbar=30, target_contact_density=0.05,
                           contact_sampling_strategy=3, track_1d_lef_position=0),
    "sub_interval_narrow_band": dict(size=4_000_000, start=1_000_000, end=2_600_000, ncells=1,
                                     nbar=30, diagonal_width=20_000, target_contact_density=0.3),
    # many LEFs per thread for a few epochs: merge ranking, wide scans, several refills
    "large_few_epochs": dict(size=60_000_000, ncells=1, nbar=700, debug_max_epochs=25),
    "large_skip_burnin": dict(size=40_000_000, ncells=1, nbar=500, skip_burnin=1,
                              debug_max_epochs=12),
}


def _tsan_compiler(tmp):
    """First C++ compiler that can link a ThreadSanitizer binary (the image's default CXX is a
    gcc build without libtsan; the distribution's g++ has it)."""
    src = os.path.join(tmp, "probe.cpp")
    with open(src, "w") as f:
        f.write("int main() { return 0; }\n")
    for cxx in (os.environ.get("CXX"), "/usr/bin/g++", "g++", "clang++"):
        if not cxx:
            continue
        try:
            r = subprocess.run([cxx, "-fsanitize=thread", "-o", os.path.join(tmp, "probe"), src],
                               capture_output=True)
        except OSError:
            continue
        if r.returncode == 0:
            return cxx
    return None


@pytest.fixture(scope="module")
def tsan_cxx(tmp_path_factory):
    cxx = _tsan_compiler(str(tmp_path_factory.mktemp("tsan_probe")))
    if cxx is None:
        pytest.skip("no C++ compiler with ThreadSanitizer support here")
    return cxx


@pytest.fixture(scope="module")
def tsan_binary(tsan_cxx):
    csrc = os.path.join(HERE, "..", "modle_b200", "csrc")
    deps = [os.path.join(EMU, f) for f in ("emu_mt_main.cpp", "emu_capi.cpp")] + [
        os.path.join(csrc, f) for f in ("sim_core.hpp", "sim_types.hpp", "cta.hpp", "launch_prep.hpp",
                                        "host_rng.hpp", "host.cpp", "status.hpp",
                                        "ziggurat_tables.inc")]
    cxx = tsan_cxx
    # MODLE_B200_EMU_DEFINES: the same experiment switches as tests/emu_lib.py, in a binary of its own
    defines = [d for d in os.environ.get("MODLE_B200_EMU_DEFINES", "").split(",") if d]
    flags = CXXFLAGS + ["-D" + d for d in defines]
    out = os.path.join(EMU, "emu_mt_main" if not defines else "emu_mt_main_variant")
    try:
        buildutil.ensure_built(
            out, deps, lambda tmp: [cxx] + flags + ["-o", tmp, os.path.join(EMU, "emu_mt_main.cpp"),
                                                   os.path.join(csrc, "host.cpp")],
            extra=" ".join(flags))
    except Exception as e:  # no libtsan in this toolchain
        pytest.skip(f"cannot build the ThreadSanitizer harness: {e}")
    return out


def _write_case(path, p, iv, bars, tasks):
    with open(path, "wb") as f:
        f.write(np.array([len(bars), len(tasks)], dtype=np.uint64).tobytes())
        f.write(bytes(p))
        f.write(bytes(iv))
        f.write(np.ascontiguousarray(bars).tobytes())
        f.write(np.ascontiguousarray(tasks).tobytes())


def _run(binary, tmp, name, mode, nthreads):
    p, iv, bars, tasks = make_case(seed=3, **CASES[name])
    case = os.path.join(tmp, f"{name}_{mode}.case")
    out = os.path.join(tmp, f"{name}_{mode}.out")
    _write_case(case, p, iv, bars, tasks)
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 exitcode=66")
    r = subprocess.run([binary, case, str(nthreads), str(mode), out], capture_output=True,
                       text=True, env=env, timeout=900)
    return (p, iv, bars, tasks), r, out


@pytest.mark.parametrize("mode", [0, 1], ids=["deterministic", "throughput"])
def test_spmd_emulation_is_race_free_and_matches_the_serial_emulation(tsan_binary, tmp_path, mode):
    names = sorted(CASES)
    with ThreadPoolExecutor(4) as ex:
        runs = list(ex.map(lambda n: _run(tsan_binary, str(tmp_path), n, mode, 8), names))
    emu_lib.set_rng_mode(mode)
    try:
        for name, ((p, iv, bars, tasks), r, out) in zip(names, runs):
            if "FATAL: ThreadSanitizer" in r.stderr:
                pytest.skip("ThreadSanitizer cannot run here: " + r.stderr.strip().splitlines()[0])
            assert "WARNING: ThreadSanitizer" not in r.stderr, (name, r.stderr[:3000])
            assert r.returncode == 0, (name, r.returncode, r.stderr[:2000])
            ref = emu_lib.simulate_interval(p, iv, bars, tasks, virtual_threads=8)
            nrows, ncols = host.band_shape(p, int(iv.end - iv.start))
            raw = np.fromfile(out, dtype=np.uint8)
            band = raw[:4 * (nrows * ncols + 1)].view(np.uint32)
            occ = raw[4 * (nrows * ncols + 1):].view(np.uint64)
            assert np.array_equal(band, ref[0]), name
            if p.track_1d_lef_position:
                assert np.array_equal(occ, ref[1]), name
            lines = [ln.split() for ln in r.stdout.splitlines()]
            for c, ln in enumerate(x for x in lines if x[0] == "cell"):
                st = ref[2][c]
                assert [int(ln[k]) for k in (3, 5, 7, 9, 11, 13)] == [
                    int(st[f]) for f in ("num_contacts", "num_epochs", "num_burnin_epochs",
                                         "num_lef_updates", "num_rng_draws", "device_fault")], name
            assert int(lines[-1][-1]) == ref[3], name
    finally:
        emu_lib.set_rng_mode(0)


def test_the_detector_sees_a_missing_barrier(tsan_binary, tsan_cxx, tmp_path):
    """Negative control: the same harness built from a copy of the kernel source in which the
    barrier between the threads' reads of `S.rng_pos` and the leader's update is removed (one of
    the two races behind the flake of DESIGN.md 3) must report a data race."""
    import shutil

    root = tmp_path / "tree"
    shutil.copytree(os.path.join(HERE, "..", "modle_b200", "csrc"), root / "modle_b200" / "csrc")
    shutil.copytree(os.path.join(HERE, "..", "include"), root / "include")
    os.makedirs(root / "tests" / "emu")
    for f in ("emu_mt_main.cpp", "emu_capi.cpp"):
        shutil.copy(os.path.join(EMU, f), root / "tests" / "emu" / f)
    src = root / "modle_b200" / "csrc" / "sim_core.hpp"
    text = src.read_text()
    good = ("    cta.sync();  // every thread has read `base` before the leader moves the stream "
            "position\n    if constexpr (!kCtr) {\n      MB_REGION(cta, tid) {\n"
            "        if (cta.leader(tid)) S.rng_pos = base + P.n_bar;")
    assert text.count(good) == 1
    src.write_text(text.replace(good, good.split("\n", 1)[1]))
    binary = str(root / "tests" / "emu" / "emu_mt_bug")
    subprocess.run([tsan_cxx] + CXXFLAGS +
                   ["-o", binary, str(root / "tests" / "emu" / "emu_mt_main.cpp"),
                    str(root / "modle_b200" / "csrc" / "host.cpp")], check=True)
    CASES["_control"] = CASES["defaults"]
    try:
        _, r, _ = _run(binary, str(tmp_path), "_control", 0, 8)
    finally:
        del CASES["_control"]
    if "FATAL: ThreadSanitizer" in r.stderr:
        pytest.skip("ThreadSanitizer cannot run here")
    assert "WARNING: ThreadSanitizer: data race" in r.stderr
