"""Small known-answer tests lifted from the reference's own unit tests (data only), run against
the oracle, the kernel's helpers (through the emulation library) and the host layer:

  collision word encoding      test/units/simulation_cpu/collision_encoding_test.cpp:27-176
  barrier occupancy <-> stp    test/units/simulation_internal/extrusion_barriers_test.cpp:35-96
  ContactMatrixDense updates   test/units/contact_matrix/contact_matrix_dense_test.cpp:35-44,90-121
  index encoding               test/units/contact_matrix/contact_matrix_internal_test.cpp:15-49
"""
import ctypes as C

import numpy as np
import pytest

import emu_lib
from modle_b200 import host
from oracle import pyoracle

# Collision<>::* (src/libmodle/cpu/include/modle/collision_encoding.hpp:91-96)
COLLISION, CHROM_BOUNDARY, LEF_BAR, PRIMARY, SECONDARY = 0x10, 0x08, 0x04, 0x02, 0x01
INDEX_MASK = (1 << 24) - 1  # the port keeps 24 index bits (the reference 59); DESIGN.md 3


def _impls():
    o, e = pyoracle.lib(), emu_lib.lib()
    for L, pre in ((o, "oracle"), (e, "emu")):
        w = getattr(L, pre + "_collision_word")
        q = getattr(L, pre + "_collision_query")
        w.restype = C.c_uint32
        w.argtypes = [C.c_uint64, C.c_uint32]
        q.restype = None
        q.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64)]
        yield pre, w, q


def query(w, q, idx, event, kind=0):
    out = (C.c_uint64 * 6)()
    q(w(idx, event), kind, out)
    keys = ("index", "event", "occurred", "avoided", "occurred_kind", "avoided_kind")
    return dict(zip(keys, [int(x) for x in out]))


@pytest.mark.parametrize("impl", ["oracle", "emu"])
def test_collision_encoding(impl):
    pre, w, q = next(t for t in _impls() if t[0] == impl)
    events = [LEF_BAR, PRIMARY, SECONDARY, CHROM_BOUNDARY, COLLISION | LEF_BAR,
              COLLISION | PRIMARY, COLLISION | SECONDARY, COLLISION | CHROM_BOUNDARY]
    for ev in events:  # "Collision encoding": round trip of (idx, event)
        r = query(w, q, 5, ev)
        assert (r["index"], r["event"]) == (5, ev)
    for idx in (123, INDEX_MASK):  # largest index next to every event
        for kind in (SECONDARY, LEF_BAR):
            r = query(w, q, idx, COLLISION | kind, kind)
            assert r["occurred_kind"] and r["index"] == idx
    r = query(w, q, 123, PRIMARY, PRIMARY)  # set_event(LEF_LEF_PRIMARY); set_idx(123)
    assert r["avoided_kind"] and r["index"] == 123
    # "... - Chromosomal boundaries / LEF-BAR / LEF-LEF primary / LEF-LEF secondary"
    for idx, kind in ((3, CHROM_BOUNDARY), (5, CHROM_BOUNDARY), (123, LEF_BAR), (123, PRIMARY),
                      (123, SECONDARY)):
        r = query(w, q, idx, COLLISION | kind, kind)
        assert r["occurred"] and r["occurred_kind"]
        assert not r["avoided"] and not r["avoided_kind"]
        r = query(w, q, idx, kind, kind)
        assert not r["occurred"] and not r["occurred_kind"]
        assert r["avoided"] and r["avoided_kind"]
        for other in (CHROM_BOUNDARY, LEF_BAR, PRIMARY, SECONDARY):
            if other != kind:  # decode_event() == (c | COLLISION) is an exact match
                assert not query(w, q, idx, COLLISION | kind, other)["occurred_kind"]


def test_barrier_occupancy_identities(product_lib):
    L = product_lib
    # "Extrusion barriers - occupancy": stp_active(occupancy) inverts occupancy(stp)
    for occ, stp_i in ((0.85, 0.7), (0.93, 0.65)):
        stp_a = L.modle_b200_stp_active_from_occupancy(stp_i, occ)
        assert L.modle_b200_occupancy_from_stp(stp_a, stp_i) == pytest.approx(occ, rel=1.2e-5)
    # "Extrusion barriers - compute_occupancy"
    f = L.modle_b200_occupancy_from_stp
    assert f(1.0, 0.0) == 1.0
    assert f(0.0, 1.0) == 0.0
    assert f(0.7, 0.7) == 0.5
    assert f(0.7, 0.5) == pytest.approx(0.625, rel=1.2e-5)
    assert f(1.0, 0.5) == 1.0
    assert f(0.7, 1.0) == 0.0


def _increment(band, nrows, ncols, b1, b2, missed):
    L = pyoracle.lib()
    L.oracle_band_increment.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64,
                                        C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]
    return L.oracle_band_increment(band.ctypes.data, nrows, ncols, 1, b1, b2, C.byref(missed))


def test_contact_matrix_increment_semantics():
    # "ContactMatrixDense simple" (10 rows x 100 columns)
    band = np.zeros(10 * 100 + 1, dtype=np.uint32)
    missed = C.c_uint64(0)
    _increment(band, 10, 100, 0, 0, missed)
    _increment(band, 10, 100, 0, 0, missed)
    assert band[0] == 2 and band.sum() == 2
    # "ContactMatrixDense in/decrement" (10 x 20): (11, 0) is 11 bins off the diagonal -> missed,
    # and a missed update does not count towards the total
    band = np.zeros(10 * 20 + 1, dtype=np.uint32)
    _increment(band, 10, 20, 0, 0, missed)
    _increment(band, 10, 20, 15, 15, missed)
    assert band.sum() == 2 and band[0] == 1 and band[15 * 10] == 1 and missed.value == 0
    assert _increment(band, 10, 20, 11, 0, missed) == 1
    assert band[0] == 1 and missed.value == 1 and band.sum() == 2
    # symmetric: (row, col) and (col, row) are the same pixel (transpose_coords)
    _increment(band, 10, 20, 3, 7, missed)
    _increment(band, 10, 20, 7, 3, missed)
    assert band[7 * 10 + 4] == 2


def test_index_encoding_roundtrip():
    # "ContactMatrix internal: encode_idx / decode_idx / roundtrip": pixel (row = |b1 - b2|,
    # col = max) <-> col * nrows + row; checked through the oracle's pixel loop
    nrows, ncols = 4, 9
    band = np.zeros(nrows * ncols + 1, dtype=np.uint32)
    want = []
    for col in range(ncols):
        for row in range(min(nrows, col + 1)):
            i = col * nrows + row
            band[i] = i + 1
            want.append((col - row, col, i + 1))
    px = pyoracle.band_to_pixels(band, nrows, ncols)
    got = sorted((int(a), int(b), int(c)) for a, b, c, _ in px.tolist())
    assert got == sorted(want)
    assert band[9] == 10 and (1, 2, 10) in got  # encode_idx(1, 2, nrows = 4) == 9


def test_burnin_statistics_reference_kats():
    """test/units/stats/descriptive_test.cpp "Mean" / "Sum of squared deviations" / "Variance" /
    "Standard Deviation" on {0..10}: mean 5, SSD 110, population variance 10, standard deviation
    3.1622776601683795 -- the statistics compute_loop_size_stats (simulation.cpp:795-819) feeds the
    burn-in history with, here through the oracle's and the kernel source's burn-in step."""
    v = np.arange(11, dtype=np.uint64)
    out = (C.c_double * 2)()
    L = pyoracle.lib()
    L.oracle_loop_size_stats.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_double)]
    L.oracle_loop_size_stats.restype = None
    L.oracle_loop_size_stats(v.ctypes.data, len(v), out)
    assert out[0] == 5.0 and out[1] == 3.1622776601683795
    assert out[1] ** 2 * len(v) == pytest.approx(110.0, rel=1e-15)
    E = emu_lib.lib()
    E.emu_loop_size_stats.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_double), C.c_int]
    for threads in (1, 4, 64):
        assert E.emu_loop_size_stats(v.ctypes.data, len(v), out, threads) == 0
        assert out[0] == 5.0
        assert out[1] == pytest.approx(3.1622776601683795 / 5.0, rel=1e-15)  # cv = sd / mean
    # a loop-size vector of the size of a real chromosome: kernel (tree order) vs oracle
    # (left to right, as the reference) agree to a few ulp ("FP fidelity")
    rng = np.random.default_rng(9)
    big = rng.integers(0, 400_000, 4979).astype(np.uint64)
    L.oracle_loop_size_stats(big.ctypes.data, len(big), out)
    mean_o, sd_o = out[0], out[1]
    assert E.emu_loop_size_stats(big.ctypes.data, len(big), out, 1024) == 0
    assert out[0] == mean_o
    assert abs(out[1] - sd_o / mean_o) <= 8 * np.spacing(sd_o / mean_o)


@pytest.mark.parametrize("impl", ["oracle", "kernel source"])
def test_bind_lefs_reference_properties(impl):
    """test/units/simulation_cpu/simulation_simple_unit_test.cpp "Bind LEFs 001 / 003": every LEF
    selected for binding ends up bound at one position inside [start, end) and both rank arrays
    are sorted. Here after the first epoch of a --skip-burnin run (all LEFs are selected at epoch
    0; one epoch of extrusion later a LEF is either released or holds rev <= fwd inside the
    interval, bound at epoch 0), and with extrusion speed 1 bp / no release the bound position
    itself is visible: fwd - rev <= 2."""
    from common import make_case

    for kw in (dict(), dict(rev_extrusion_speed=1, fwd_extrusion_speed=1,
                            rev_extrusion_speed_std=0.0, fwd_extrusion_speed_std=0.0,
                            avg_lef_processivity=10**12)):
        p, iv, bars, tasks = make_case(size=4_000_000, start=500_000, end=3_500_000, ncells=1,
                                       nbar=40, seed=2, skip_burnin=1, debug_max_epochs=1, **kw)
        snap = (pyoracle.snapshot_cell(p, iv, bars, tasks[0]) if impl == "oracle"
                else emu_lib.snapshot_cell(p, iv, bars, tasks[0], virtual_threads=32))
        n = int(iv.num_lefs)
        unbound = np.uint64(2**64 - 1)
        rev, fwd, ep = snap["rev_pos"], snap["fwd_pos"], snap["binding_epoch"]
        bound = ep != unbound
        assert snap["num_active_lefs"] == n and bound.sum() > 0.9 * n
        assert np.all(rev[bound] >= int(iv.start)) and np.all(fwd[bound] < int(iv.end))
        assert np.all(rev[bound] <= fwd[bound]) and np.all(ep[bound] == 0)
        assert np.all(rev[~bound] == unbound) and np.all(fwd[~bound] == unbound)
        if kw:
            assert bound.all() and np.all(fwd - rev <= 2)
        for ranks, pos in ((snap["rev_ranks"], rev), (snap["fwd_ranks"], fwd)):
            assert sorted(ranks.tolist()) == list(range(n))  # check_that_lefs_are_sorted_by_idx
