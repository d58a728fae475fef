#!/usr/bin/env python3
"""Band -> sorted COO pixels in isolation: times modle_b200_count_pixels_device +
modle_b200_fill_pixels_device on device-resident bands and reports them against the HBM roofline.

    python scripts/bench_pixels.py [--out gpurun_out/pixels.json] [--reps 5]

Algorithmic bytes per call: the band is read once per pass (4 B/pixel, two passes), the row
offsets are written and read (8 B/row each), every non-zero pixel is written once (24 B).
Cases: chr1 at 5 kb bins (600 x 49,792 = 119.5 MB, the largest band of C2) filled by a loop-like
contact stream at density 1 contact/pixel, the same band fully dense, and the C5 band (3000 x
242,194 = 2.9 GB) at density 1. A CPU baseline (the oracle's pixel loop, 1 core) is timed on the
chr1 case.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import numpy as np
    import torch

    from modle_b200.simulation import Context

    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cases", default="chr1_loop,chr1_dense,c5_loop")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    ctx = Context(0)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    try:
        hbm = float(json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        hbm = 6650.0
    g = torch.Generator(device=dev)
    g.manual_seed(20260117)
    stream = torch.cuda.Stream(device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    results = []
    shapes = {"chr1_loop": (600, 49_792, "loop"), "chr1_dense": (600, 49_792, "dense"),
              "c5_loop": (3000, 242_194, "loop")}
    for case in args.cases.split(","):
        nrows, ncols, kind = shapes[case]
        npx = nrows * ncols
        band = torch.zeros(npx + 1, dtype=torch.int32, device=dev)
        if kind == "dense":
            band[:npx] = torch.randint(1, 1000, (npx,), device=dev, generator=g,
                                       dtype=torch.int32)
            # pixels the matrix cannot hold (distance > column) stay zero
            body = band[:npx].view(ncols, nrows)
            ii = torch.arange(nrows, device=dev)[None, :]
            jj = torch.arange(min(ncols, nrows), device=dev)[:, None]
            body[:min(ncols, nrows)][ii > jj] = 0
        else:
            missed = torch.zeros(1, dtype=torch.int64, device=dev)
            left = npx
            while left > 0:
                n = min(left, 1 << 27)
                b2 = torch.randint(0, ncols, (n,), device=dev, generator=g, dtype=torch.int64)
                d = torch.empty(n, device=dev, dtype=torch.float32).exponential_(1.0 / 100.0,
                                                                                generator=g)
                d = d.to(torch.int64).clamp_(0, nrows - 1)
                b1 = (b2 - d).clamp_(min=0).to(torch.int32).contiguous()
                b2 = b2.to(torch.int32).contiguous()
                torch.cuda.synchronize()  # the context's stream does not wait for torch's
                ctx.register_contacts_device(b1.data_ptr(), b2.data_ptr(), n, nrows, ncols,
                                             band.data_ptr(), missed.data_ptr(),
                                             torch.cuda.current_stream(dev).cuda_stream)
                torch.cuda.synchronize()
                left -= n
                del b1, b2, d
        rows = torch.zeros(ncols + 1, dtype=torch.int64, device=dev)
        torch.cuda.synchronize()
        ctx.count_pixels_device(band.data_ptr(), nrows, ncols, rows.data_ptr(), stream.cuda_stream)
        stream.synchronize()
        nnz = int(rows[-1].item())
        px = torch.empty(3 * nnz, dtype=torch.int64, device=dev)
        t_count, t_fill = [], []
        for r in range(args.reps + 2):
            flush.fill_(r & 0xFF)  # evict the band from the L2 between repetitions
            torch.cuda.synchronize()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record(stream)
            ctx.count_pixels_device(band.data_ptr(), nrows, ncols, rows.data_ptr(),
                                    stream.cuda_stream)
            e[1].record(stream)
            ctx.fill_pixels_device(band.data_ptr(), nrows, ncols, 0, rows.data_ptr(),
                                   px.data_ptr(), nnz, stream.cuda_stream)
            e[2].record(stream)
            stream.synchronize()
            if r >= 2:
                t_count.append(e[0].elapsed_time(e[1]))
                t_fill.append(e[1].elapsed_time(e[2]))
        # checks that hold at any size: pixels sorted, counts add up to the band's total
        v = px.view(-1, 3)
        assert int((v[:, 2] & 0xFFFFFFFF).sum().item()) == int(band[:npx].to(torch.int64).sum().item())
        key = v[:, 0] * (1 << 20) + v[:, 1]
        assert bool((key[1:] > key[:-1]).all().item()) if nnz > 1 else True
        assert bool(((v[:, 1] - v[:, 0]) < nrows).all().item())
        del key
        mc = sorted(t_count)[len(t_count) // 2]
        mf = sorted(t_fill)[len(t_fill) // 2]
        b_count = 4 * npx + 16 * ncols
        b_fill = 4 * npx + 8 * ncols + 24 * nnz
        res = {"case": case, "nrows": nrows, "ncols": ncols, "band_bytes": 4 * (npx + 1),
               "nonzero_pixels": nnz, "fill_fraction": nnz / npx,
               "count_ms": mc, "fill_ms": mf, "total_ms": mc + mf,
               "count_GBps": b_count / mc / 1e6, "fill_GBps": b_fill / mf / 1e6,
               "total_GBps": (b_count + b_fill) / (mc + mf) / 1e6,
               "algorithmic_bytes": b_count + b_fill, "hbm_peak_GBps": hbm,
               "frac_of_hbm_peak": (b_count + b_fill) / (mc + mf) / 1e6 / hbm,
               "pixels_per_s": nnz / ((mc + mf) * 1e-3),
               "l2_policy": "256 MB written between repetitions (band evicted from the L2)"}
        if case == "chr1_loop" and not args.no_cpu:
            from oracle import pyoracle

            h = band.cpu().numpy().view(np.uint32)
            t0 = time.perf_counter()
            ref = pyoracle.band_to_pixels(h, nrows, ncols, 0)
            dt = time.perf_counter() - t0
            assert np.array_equal(ref.view(np.int64).reshape(-1, 3), v.cpu().numpy())
            # the oracle wrapper runs the loop twice (size query + fill)
            res["cpu_baseline"] = {"ms": 1e3 * dt / 2, "cores": 1, "kind": "port",
                                   "pixels_per_s": nnz / (dt / 2)}
        results.append(res)
        print(json.dumps(res), flush=True)
        del band, rows, px, v
    ctx.close()
    if args.out:
        with open(args.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
