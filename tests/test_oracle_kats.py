"""Known-answer tests that pin the oracle's third-party arithmetic as far as this container
allows: xoshiro256++/SplitMix64, jump(), XXH3-64, and sanity of the restated
Boost.Random distributions (moments; exact draw counts)."""
import numpy as np
import pytest

from oracle import pyoracle


def test_prng_first_outputs():
    # PRNG(752741483): first output must exceed 0.75 * 2^64 for the reference's
    # "Simulation 011/012" goldens to hold (simulation_complex_unit_test.cpp:637-756)
    st = pyoracle.rng_seed(752741483)
    assert st == [0x2a3bc28b8fc13c5a, 0xdb997ec403e6d05d, 0x3b9841261cc6feca, 0x3943d8b92b198bdb]
    outs = []
    for _ in range(3):
        v, st = pyoracle.rng_next(st)
        outs.append(v)
    assert outs == [0xcc992fefaa72fc27, 0x71df6bfd251890fc, 0xe07ee158acf122d2]
    assert outs[0] / 2.0**64 > 0.75
    v, _ = pyoracle.rng_next(pyoracle.rng_seed(10556020843759504871))  # DEFAULT_PRNG
    assert v == 0x251cef6953ce03a9


def _matmul_cols(a, b):
    """columns of a∘b for GF(2) matrices stored as lists of 256-bit column ints"""
    out = []
    for col in b:
        acc = 0
        j = 0
        while col:
            if col & 1:
                acc ^= a[j]
            col >>= 1
            j += 1
        out.append(acc)
    return out


def _state_to_int(st):
    return st[0] | (st[1] << 64) | (st[2] << 128) | (st[3] << 192)


def test_jump_equals_transition_matrix_power():
    """jump() must equal 2^128 plain steps: check the polynomial against T^(2^128) obtained by
    128 squarings of the one-step transition matrix (independent of the jump constants)."""
    m64 = (1 << 64) - 1
    cols = []
    for j in range(256):
        st = [0, 0, 0, 0]
        st[j // 64] = 1 << (j % 64)
        cols.append(_state_to_int(pyoracle.rng_discard(st, 1)))
    # sanity of the one-step matrix itself
    st = pyoracle.rng_seed(42)
    x = _state_to_int(st)
    y = 0
    for j in range(256):
        if (x >> j) & 1:
            y ^= cols[j]
    assert y == _state_to_int(pyoracle.rng_discard(st, 1))
    m = cols
    for _ in range(128):
        m = _matmul_cols(m, m)
    for seed in (0, 1, 12345):
        st = pyoracle.rng_seed(seed)
        x = _state_to_int(st)
        y = 0
        for j in range(256):
            if (x >> j) & 1:
                y ^= m[j]
        jumped = pyoracle.rng_jump(st)
        assert y == _state_to_int(jumped)
        assert all(0 <= w <= m64 for w in jumped)


def test_xxh3_against_python_xxhash():
    xxhash = pytest.importorskip("xxhash")
    rng = np.random.default_rng(7)
    for n in list(range(17, 241)):
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        for seed in (0, 1, 0xDEADBEEFCAFEF00D):
            assert pyoracle.xxh3_64(data, seed) == xxhash.xxh3_64_intdigest(data, seed=seed), n


def test_interval_hash_probes():
    # values probed with xxHash 0.8.3: seed 0, whole chromosome
    assert pyoracle.interval_hash("chr20", 64444167, 0, 64444167, 0) == 0x50dbf797ede17321
    assert pyoracle.interval_hash("chr1", 248956422, 0, 248956422, 0) == 0xa53d8e35875b84b9
    assert pyoracle.interval_hash("chr2", 242193529, 0, 242193529, 0) == 0xd4233268d6f85130


def test_ziggurat_tables_shape():
    nx, ny, ex, ey = pyoracle.zig_tables()
    # leading / trailing literals of boost/random/normal_distribution.hpp (normal_table<>) and
    # exponential_distribution.hpp (exponential_table<>), Boost 1.88: bit-exact after parsing
    assert list(nx[:8]) == [3.7130862467403632609, 3.4426198558966521214, 3.2230849845786185446,
                            3.0832288582142137009, 2.9786962526450169606, 2.8943440070186706210,
                            2.8231253505459664379, 2.7611693723841538514]
    assert list(ny[:4]) == [0, 0.0026696290839025035092, 0.0055489952208164705392,
                            0.0086244844129304709682]
    assert list(ex[:4]) == [8.6971174701310497140, 7.6971174701310497140, 6.9410336293772123602,
                            6.4783784938325698538]
    assert nx[128] == 0.0 and ny[0] == 0.0 and ny[128] == 1.0
    assert ex[256] == 0.0 and ey[0] == 0.0 and ey[256] == 1.0
    assert np.all(np.diff(nx) < 0) and np.all(np.diff(ny) > 0)
    assert np.all(np.diff(ex) < 0) and np.all(np.diff(ey) > 0)
    # the tables solve the ziggurat equations: equal-area layers, the last one closing at f = 1
    v = nx[1] * ny[1] + np.sqrt(np.pi / 2) * __import__("math").erfc(nx[1] / np.sqrt(2))
    area = nx[1:128] * (ny[2:129] - ny[1:128])
    assert np.allclose(area, v, rtol=1e-13) and abs(nx[0] * ny[1] - v) < 1e-15
    ve = (ex[1] + 1) * ey[1]
    assert np.allclose(ex[1:256] * (ey[2:257] - ey[1:256]), ve, rtol=1e-13)
    # the kernel's copy (modle_b200/csrc) is the same generated file
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    a = open(os.path.join(here, "..", "oracle", "ziggurat_tables.inc")).read()
    b = open(os.path.join(here, "..", "modle_b200", "csrc", "ziggurat_tables.inc")).read()
    assert a == b


@pytest.mark.parametrize("kind,args,mean,var", [
    ("unit_normal", (), 0.0, 1.0),
    ("normal", (4000.0, 200.0), 4000.0, 200.0**2),
    ("unit_exponential", (), 1.0, 1.0),
    ("poisson", (6.9,), 6.9, 6.9),        # inversion branch (chr20: 1289/187)
    ("poisson", (26.6,), 26.6, 26.6),     # PTRD branch (chr1: 4979/187)
    ("binomial", (206.0, 1.0 / 6), 206 / 6, 206 * (1 / 6) * (5 / 6)),   # BTRD
    ("binomial", (20.0, 0.3), 6.0, 20 * 0.3 * 0.7),                     # inversion
    ("binomial", (500.0, 0.9), 450.0, 500 * 0.9 * 0.1),                 # mirrored BTRD
    ("canonical", (), 0.5, 1 / 12),
    ("uniform_int", (10.0, 29.0), 19.5, (20**2 - 1) / 12),
])
def test_distribution_moments(kind, args, mean, var):
    n = 400_000
    x, _, draws = pyoracle.sample(kind, n, pyoracle.rng_seed(2024), *args)
    se = np.sqrt(var / n)
    assert abs(x.mean() - mean) < 6 * se
    assert abs(x.var() - var) < 0.02 * var + 1e-9
    assert draws >= n


def test_normal_ziggurat_tail_and_shape():
    n = 2_000_000
    x, _, draws = pyoracle.sample("unit_normal", n, pyoracle.rng_seed(5))
    # fast path takes exactly one draw; the slow paths are rare
    assert 1.0 < draws / n < 1.06
    for q, expect in ((1.0, 0.158655), (2.0, 0.0227501), (3.0, 0.0013499), (3.6, 0.000159109)):
        frac = np.mean(x > q)
        assert abs(frac - expect) < 6 * np.sqrt(expect / n) + 1e-6
    assert abs(np.mean(x < -3.6) - 0.000159109) < 6 * np.sqrt(0.000159109 / n) + 1e-6


def test_draw_counts_of_elementary_distributions():
    st = pyoracle.rng_seed(1)
    assert pyoracle.sample("bernoulli", 100, st, 0.0)[2] == 0      # p == 0: no draw
    assert pyoracle.sample("bernoulli", 100, st, 0.3)[2] == 100
    assert pyoracle.sample("uniform_int", 100, st, 5.0, 5.0)[2] == 0  # empty range: no draw
    assert pyoracle.sample("canonical", 100, st)[2] == 100
    x, _, _ = pyoracle.sample("bernoulli", 200_000, st, 0.75)
    assert abs(x.mean() - 0.75) < 0.005


def test_gev_noise_matches_scipy_quantiles():
    stats = pytest.importorskip("scipy.stats")
    x, _, _ = pyoracle.sample("gev", 300_000, pyoracle.rng_seed(3), 0.0, 5000.0, 0.001)
    # modle's parameterisation: mu + sigma (1 - (-ln u)^xi) / xi  == scipy genextreme(c=xi)
    # evaluated at 1-u; compare a few quantiles
    qs = [0.1, 0.5, 0.9]
    ref = -stats.genextreme.ppf([1 - q for q in qs], c=0.001, loc=0.0, scale=5000.0)
    got = np.quantile(x, qs)
    assert np.allclose(np.sort(np.abs(got)), np.sort(np.abs(ref)), rtol=0.03)


def test_oracle_internal_state_log_invariants():
    """dump_stats restatement (simulation.cpp:995-1056): one record per completed epoch, taken
    after extrude; counts are consistent with each other and with the cell's stats."""
    from common import make_case

    p, iv, bars, tasks = make_case(size=4_000_000, ncells=2, nbar=60, target_contact_density=0.01)
    plain = pyoracle.simulate_interval(p, iv, bars, tasks, nthreads=2)
    band, occ, stats, missed, log = pyoracle.simulate_interval(p, iv, bars, tasks, nthreads=2,
                                                               log_capacity_per_cell=2000)
    assert np.array_equal(band, plain[0]) and np.array_equal(stats, plain[2])  # logging is passive
    for c in range(2):
        ne = int(stats["num_epochs"][c])
        r = log[c][:ne]
        assert np.array_equal(r["epoch"], np.arange(ne))
        assert not log[c][ne:].view(np.uint8).any()
        assert int(r["num_lefs"].astype(np.uint64).sum()) == int(stats["num_lef_updates"][c])
        assert (r["lefs_stalled_both"] <= np.minimum(r["lefs_stalled_rev"],
                                                     r["lefs_stalled_fwd"])).all()
        # every stall is exactly one of: interval boundary, LEF-BAR, primary or secondary LEF-LEF
        typed = r["lef_bar_collisions"] + r["lef_lef_primary_collisions"] + \
            r["lef_lef_secondary_collisions"]
        assert (typed <= r["lefs_stalled_rev"] + r["lefs_stalled_fwd"]).all()
        assert (r["barriers_occupied"] <= len(bars)).all()
        assert r["burnin"][0] == 1 and r["burnin"][-1] == 0
        assert (np.diff(r["burnin"].astype(np.int64)) <= 0).all()  # burn-in never resumes
        # epoch 0: the first LEF bound at epoch 0 has extruded once in each direction
        assert 0 < r["loop_size_sum"][0] <= 2 * 17000
