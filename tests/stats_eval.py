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
