"""Test helpers: modle_tools eval-style comparisons of two band matrices (reference layout).

Pearson per stratum follows the definition the reference's `modle_tools eval` uses for diagonal
strata (src/modle_tools/eval.cpp:438-443 computes correlations over matching stripes of the two
matrices); the stratum-adjusted correlation is the HiCRep combination of the per-diagonal Pearson
coefficients, weighted by N_k * sqrt(var1_k * var2_k)."""
import numpy as np


def diagonals(band, nrows, ncols):
    """[d][j] view: counts of pixels (j - d, j), d = 0 .. nrows-1 (entries with j < d are 0)."""
    return np.ascontiguousarray(band[:nrows * ncols].reshape(ncols, nrows).T)


def per_diagonal_mean_var(band, nrows, ncols):
    D = diagonals(band, nrows, ncols).astype(np.float64)
    mean = np.zeros(nrows)
    var = np.zeros(nrows)
    for d in range(nrows):
        x = D[d, d:]
        if len(x):
            mean[d] = x.mean()
            var[d] = x.var()
    return mean, var, D.sum(axis=1)


def stratum_adjusted_correlation(b1, b2, nrows, ncols, max_d=None):
    D1 = diagonals(b1, nrows, ncols).astype(np.float64)
    D2 = diagonals(b2, nrows, ncols).astype(np.float64)
    num = den = 0.0
    for d in range(1, max_d or nrows):
        x, y = D1[d, d:], D2[d, d:]
        if len(x) < 3:
            continue
        vx, vy = x.var(), y.var()
        if vx == 0 or vy == 0:
            continue
        rho = ((x - x.mean()) * (y - y.mean())).mean() / np.sqrt(vx * vy)
        w = len(x) * np.sqrt(vx * vy)
        num += w * rho
        den += w
    return num / den if den else float("nan")


def stripe_pearson(b1, b2, nrows, ncols, direction="vertical"):
    """Per-bin Pearson correlation of matching stripes of two band matrices, the `modle_tools
    eval --metric pearson` definition (src/modle_tools/eval.cpp:460-470 fills one buffer of nrows
    pixels per bin with unsafe_get_column / unsafe_get_row and correlates the two buffers):
    vertical stripe of bin i = pixels (i - d, i), horizontal = pixels (i, i + d), d < nrows.
    Stripes without variance in either matrix give NaN (ignored by the callers' nanmedian)."""
    A = np.asarray(b1[:nrows * ncols], dtype=np.float64).reshape(ncols, nrows)
    B = np.asarray(b2[:nrows * ncols], dtype=np.float64).reshape(ncols, nrows)
    if direction == "horizontal":  # pixel (i, i + d) is stored in column i + d at row d
        def rows(M):
            H = np.zeros_like(M)
            for d in range(nrows):
                H[:ncols - d, d] = M[d:, d]
            return H
        A, B = rows(A), rows(B)
    a = A - A.mean(axis=1, keepdims=True)
    b = B - B.mean(axis=1, keepdims=True)
    den = np.sqrt((a * a).sum(axis=1) * (b * b).sum(axis=1))
    with np.errstate(invalid="ignore", divide="ignore"):
        return (a * b).sum(axis=1) / den


def c1_inputs(seeds=(1, 2, 3, 4)):
    """BASELINE C1 (chr20 shape, 512 cells, density 1) under several Config::seed values:
    {seed: (params, interval, barriers, tasks)}, (nrows, ncols)."""
    from modle_b200 import abi, host, workloads

    cfg, genome = workloads.config_c1(512)
    name, size, start, end, recs = genome[0]
    runs = {}
    for seed in seeds:
        p = cfg.params.copy()
        p.seed = seed
        bars = host.barriers_from_records(recs, p)
        iv = abi.Interval(size, start, end, host.compute_num_lefs(p, end - start))
        runs[seed] = (p, iv, bars, host.make_cell_tasks(p, name, iv))
    return runs, host.band_shape(cfg.params, size)


def c1_scale_gate(test, ora, nrows, ncols):
    """Gate (ii) of SURVEY 8c at the size of BASELINE C1 (7.73 M contacts over 7.73 M pixels):
    `test` = (band, occ1d, stats) of the run under test, `ora` = {seed: oracle result} for three
    other seeds; compared the way `modle_tools eval` compares matrices
    (src/modle_tools/eval.cpp:425-483). Raises AssertionError with the offending figures.

    Tolerances from five oracle seed pairs at this size (the oracle against itself gave: diagonals
    with >= 1e5 contacts -- the 20 nearest -- mean within 0.54 - 0.80 %, variance within 3.0 -
    4.3 %; all 63 diagonals with >= 1e4 contacts: mean within 2.0 - 3.1 %, variance within 6.7 -
    11.0 %, medians 0.27 - 0.47 % and 1.5 - 1.9 %; SCC 0.4790 - 0.4840; median per-stripe Pearson
    0.96499 - 0.96531; KS p of the burn-in lengths >= 0.24):
      * diagonals >= 1e5 contacts: mean within 1.5 %, variance within 8 %
      * diagonals >= 1e4 contacts: mean within 4.5 %, variance within 16 %; medians 0.9 % / 3.5 %
      * SCC against each oracle seed not more than 0.003 below the lowest SCC among the oracle
        runs themselves
      * median per-stripe Pearson (vertical and horizontal) not more than 0.0006 below the
        lowest of the oracle pairs
      * burn-in and total epochs: two-sample KS p > 0.001; 1D occupancy tracks correlate > 0.95
    """
    from scipy.stats import ks_2samp

    band, occ, stats = test
    seeds = sorted(ora)
    m_g, v_g, tot_g = per_diagonal_mean_var(band, nrows, ncols)
    for s in seeds[:2]:
        m_o, v_o, tot_o = per_diagonal_mean_var(ora[s][0], nrows, ncols)
        top = (tot_g >= 1e5) & (tot_o >= 1e5)
        big = (tot_g >= 1e4) & (tot_o >= 1e4)
        assert top.sum() >= 15 and big.sum() >= 55
        with np.errstate(invalid="ignore", divide="ignore"):
            rel_m, rel_v = np.abs(m_g / m_o - 1.0), np.abs(v_g / v_o - 1.0)
        assert rel_m[top].max() < 0.015 and rel_v[top].max() < 0.08, (rel_m[top].max(), rel_v[top].max())
        assert rel_m[big].max() < 0.045 and rel_v[big].max() < 0.16, (rel_m[big].max(), rel_v[big].max())
        assert np.median(rel_m[big]) < 0.009 and np.median(rel_v[big]) < 0.035
    pairs = [(a, b) for i, a in enumerate(seeds) for b in seeds[i + 1:]]
    scc_floor = min(stratum_adjusted_correlation(ora[a][0], ora[b][0], nrows, ncols, max_d=200)
                    for a, b in pairs)
    for s in seeds:
        scc = stratum_adjusted_correlation(band, ora[s][0], nrows, ncols, max_d=200)
        assert scc > scc_floor - 0.003, (s, scc, scc_floor)
    for direction in ("vertical", "horizontal"):
        floor = min(np.nanmedian(stripe_pearson(ora[a][0], ora[b][0], nrows, ncols, direction))
                    for a, b in pairs)
        for s in seeds:
            r = np.nanmedian(stripe_pearson(band, ora[s][0], nrows, ncols, direction))
            assert r > floor - 0.0006, (direction, s, r, floor)
    for f in ("num_burnin_epochs", "num_epochs"):
        assert ks_2samp(stats[f], ora[seeds[0]][2][f]).pvalue > 0.001, f
    assert np.corrcoef(occ.astype(np.float64), ora[seeds[0]][1].astype(np.float64))[0, 1] > 0.95
