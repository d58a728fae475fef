"""Band -> sorted COO pixels (the .cool writer hand-off).

CPU part: the oracle's restatement of append_contact_matrix_to_cooler's pixel loop
(src/libmodle_io/contact_matrix_dense_io_impl.hpp:50-71) against an independent numpy formulation
of ContactMatrixDense::unsafe_get and against the reference's own index KAT
(contact_matrix_internal_test.cpp:15-49: encode_idx(1, 2, nrows=4) == 9).
GPU part (`-m gpu`): the CUDA kernels through the C ABI against the oracle, bit for bit."""
import ctypes as C

import numpy as np
import pytest

from modle_b200 import abi
from oracle import pyoracle

SHAPES = [(1, 1), (1, 7), (4, 4), (4, 3), (7, 50), (33, 31), (32, 64), (600, 97), (257, 300),
          (600, 1289), (3000, 41)]


def random_band(nrows, ncols, density, seed, big=False):
    """Band in the reference layout. Entries the matrix can never hold (i > j, only possible for
    j < nrows) stay zero, as ContactMatrixDense::increment leaves them."""
    rng = np.random.default_rng(seed)
    band = np.zeros(nrows * ncols + 1, dtype=np.uint32)
    body = band[:nrows * ncols].reshape(ncols, nrows)  # [j][i]
    mask = rng.random((ncols, nrows)) < density
    vals = rng.integers(1, 2**32 if big else 1000, size=(ncols, nrows), dtype=np.uint64)
    body[...] = np.where(mask, vals, 0).astype(np.uint32)
    i = np.arange(nrows)[None, :]
    j = np.arange(ncols)[:, None]
    body[i > j] = 0
    band[-1] = 12345  # trailing element (the "+1"), never a pixel
    return band


def numpy_pixels(band, nrows, ncols, bin_offset):
    """Dense symmetric matrix -> upper-triangle non-zeros in row-major order."""
    body = band[:nrows * ncols].reshape(ncols, nrows)
    j, i = np.nonzero(body)           # column (= max bin), distance
    r, c = j - i, j                   # bin1 <= bin2
    ok = r >= 0
    r, c, v = r[ok], c[ok], body[j[ok], i[ok]]
    order = np.lexsort((c, r))
    out = np.zeros(len(order), dtype=abi.pixel_dtype())
    out["bin1_id"] = r[order].astype(np.uint64) + np.uint64(bin_offset)
    out["bin2_id"] = c[order].astype(np.uint64) + np.uint64(bin_offset)
    out["count"] = v[order].view(np.int32)
    return out


def test_reference_index_kat():
    # encode_idx(1, 2, nrows=4) == 9: pixel (row 1, col 2) of the band is bins (1, 2)
    band = np.zeros(4 * 5 + 1, dtype=np.uint32)
    band[9] = 7
    px = pyoracle.band_to_pixels(band, 4, 5, bin_offset=100)
    assert px.tolist() == [(101, 102, 7, 0)]


@pytest.mark.parametrize("nrows,ncols", SHAPES)
@pytest.mark.parametrize("density", [0.0, 0.03, 0.6, 1.0])
def test_oracle_pixels_match_numpy(nrows, ncols, density):
    band = random_band(nrows, ncols, density, seed=nrows * 1000 + ncols)
    a = pyoracle.band_to_pixels(band, nrows, ncols, bin_offset=17)
    b = numpy_pixels(band, nrows, ncols, 17)
    assert np.array_equal(a, b)
    assert int(a["count"].astype(np.int64).sum()) == int(band[:-1].astype(np.int64).sum())


def test_oracle_pixels_int32_cast_and_order():
    band = random_band(5, 9, 0.8, seed=3, big=True)
    a = pyoracle.band_to_pixels(band, 5, 9)
    assert np.array_equal(a, numpy_pixels(band, 5, 9, 0))
    assert (a["count"] < 0).any()  # counts >= 2^31 wrap like conditional_static_cast<int32>
    keys = a["bin1_id"].astype(np.int64) * 1000 + a["bin2_id"].astype(np.int64)
    assert np.all(np.diff(keys) > 0)
    assert np.all(a["bin2_id"] - a["bin1_id"] < 5)


# ------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("nrows,ncols", SHAPES)
@pytest.mark.parametrize("density", [0.0, 0.03, 0.6, 1.0])
def test_cuda_pixels_match_oracle(gpu_ctx, nrows, ncols, density):
    band = random_band(nrows, ncols, density, seed=nrows * 77 + ncols, big=density == 0.6)
    a = gpu_ctx.band_to_pixels(band, nrows, ncols, bin_offset=123456)
    b = pyoracle.band_to_pixels(band, nrows, ncols, bin_offset=123456)
    assert np.array_equal(a, b)


@pytest.mark.gpu
def test_cuda_pixels_capacity_and_size_query(gpu_ctx):
    from modle_b200 import host

    band = random_band(6, 40, 0.5, seed=5)
    n = C.c_uint64(0)
    L = host.lib()
    assert L.modle_b200_band_to_pixels(gpu_ctx.handle, band.ctypes.data, 6, 40, 0, None, 0,
                                       C.byref(n)) == 0
    want = len(pyoracle.band_to_pixels(band, 6, 40))
    assert n.value == want
    out = np.zeros(want - 1, dtype=abi.pixel_dtype())
    rc = L.modle_b200_band_to_pixels(gpu_ctx.handle, band.ctypes.data, 6, 40, 0, out.ctypes.data,
                                     len(out), C.byref(n))
    assert rc == abi.ERR_INVALID_ARGUMENT and n.value == want
    assert not out["count"].any()


@pytest.mark.gpu
def test_cuda_pixels_of_a_simulated_band_on_device(gpu_ctx):
    """Device-resident flow at the C1 geometry: simulate -> count -> fill, nothing but the pixels
    leaves the GPU; compared with the oracle's pixels of the oracle's band."""
    import torch

    from common import make_case

    p, iv, bars, tasks = make_case(size=64_444_167, ncells=8, nbar=1132, name="chr20",
                                   target_contact_density=0.01)
    band, occ, stats, missed = gpu_ctx.simulate_interval(p, iv, bars, tasks)
    o_band = pyoracle.simulate_interval(p, iv, bars, tasks, nthreads=8)[0]
    assert np.array_equal(band, o_band)
    from modle_b200 import host

    nrows, ncols = host.band_shape(p, 64_444_167)
    dev = torch.device("cuda", 0)
    d_band = torch.from_numpy(band.view(np.int32)).to(dev)
    d_rows = torch.zeros(ncols + 1, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    gpu_ctx.count_pixels_device(d_band.data_ptr(), nrows, ncols, d_rows.data_ptr(), stream)
    nnz = int(d_rows[-1].item())
    d_px = torch.zeros(nnz * 3, dtype=torch.int64, device=dev)
    gpu_ctx.fill_pixels_device(d_band.data_ptr(), nrows, ncols, 1000, d_rows.data_ptr(),
                               d_px.data_ptr(), nnz, stream)
    torch.cuda.synchronize()
    px = d_px.cpu().numpy().view(abi.pixel_dtype())
    want = pyoracle.band_to_pixels(o_band, nrows, ncols, bin_offset=1000)
    assert np.array_equal(px, want)
    assert int(px["count"].sum()) == int(stats["num_contacts"].sum()) - missed
    rows = d_rows.cpu().numpy()
    assert np.array_equal(np.diff(rows), np.bincount((want["bin1_id"] - 1000).astype(np.int64),
                                                     minlength=ncols))


# ------------------------------------------------------------------------ 1D occupancy profile
@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 31, 12889, 242194])
def test_cuda_occupancy_profile_matches_oracle(gpu_ctx, n):
    rng = np.random.default_rng(n)
    occ = rng.integers(0, 2**40, n, dtype=np.uint64)
    occ[rng.integers(0, n)] = 2**53 + 12345  # beyond the exactly representable doubles
    a = gpu_ctx.lef_occupancy_profile(occ)
    b = pyoracle.lef_occupancy_profile(occ)
    assert a.dtype == np.float32 and np.array_equal(a.view(np.uint32), b.view(np.uint32))
    z = gpu_ctx.lef_occupancy_profile(np.zeros(17, dtype=np.uint64))
    assert np.isnan(z).all()  # 0 / 0, as in the reference


def test_oracle_occupancy_profile():
    p = pyoracle.lef_occupancy_profile(np.array([0, 5, 10, 2], dtype=np.uint64))
    assert p.tolist() == [0.0, 0.5, 1.0, np.float32(0.2)]
