#!/usr/bin/env python3
"""Contact-register kernel in isolation (regime 2): replays a (bin1, bin2) stream
through modle_b200_register_contacts_device and reports contacts/s next to the two ceilings.

    python scripts/bench_register.py [--out gpurun_out/register.json]

Cases
  c5_hbm   BASELINE C5 geometry: chr2 at 1 kb bins, 3000 x 242,194 px = 2.9 GB (>> 126 MB L2):
           every contact is a random 4-byte read-modify-write in HBM. Hardware floor per contact:
           one 32 B sector in and one out = 64 B of DRAM traffic (+ 8 B to read the pair);
           algorithmic bytes: 8 B per contact (4 read + 4 written).
  c1_l2    C1 geometry: 600 x 12,889 px = 30.9 MB, L2 resident: bound by L2 atomic throughput.
Streams: "loop" (bin2 uniform, distance geometric with mean 100 bins, like loop contacts) and
"uniform" (uniform inside the band, the worst case for locality). The stream (2 x 4 B per contact)
is larger than L2, so nothing is cached between launches.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch

    from modle_b200.simulation import Context

    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--contacts", type=int, default=0)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    ctx = Context(0)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(
            os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    stream = torch.cuda.Stream(device=dev)
    results = []
    g = torch.Generator(device=dev)
    g.manual_seed(20260117)
    for case, nrows, ncols in (("c5_hbm", 3000, 242_194), ("c1_l2", 600, 12_889)):
        band = torch.zeros(nrows * ncols + 1, dtype=torch.int32, device=dev)
        missed = torch.zeros(1, dtype=torch.int64, device=dev)
        # C5 at its real density (1 contact per pixel) unless --contacts says otherwise
        n = args.contacts if (args.contacts or case != "c5_hbm") else nrows * ncols
        n = n or (1 << 27)
        for kind in ("loop", "uniform"):
            b2 = torch.randint(0, ncols, (n,), device=dev, generator=g, dtype=torch.int64)
            if kind == "loop":
                d = torch.empty(n, device=dev, dtype=torch.float32).exponential_(1.0 / 100.0,
                                                                                generator=g)
                d = d.to(torch.int64).clamp_(0, nrows - 1)
            else:
                d = torch.randint(0, nrows, (n,), device=dev, generator=g, dtype=torch.int64)
            b1 = (b2 - d).clamp_(min=0)
            b1 = b1.to(torch.int32).contiguous()
            b2 = b2.to(torch.int32).contiguous()
            del d
            times = []
            with torch.cuda.stream(stream):
                for r in range(max(1, args.reps) + 2):
                    band.zero_()
                    missed.zero_()
                    e0 = torch.cuda.Event(enable_timing=True)
                    e1 = torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    ctx.register_contacts_device(b1.data_ptr(), b2.data_ptr(), n, nrows, ncols,
                                                 band.data_ptr(), missed.data_ptr(),
                                                 stream.cuda_stream)
                    e1.record(stream)
                    stream.synchronize()
                    if r >= 2:
                        times.append(e0.elapsed_time(e1))
            total = int(band.to(torch.int64).sum().item()) + int(missed.item())
            assert total == n, (total, n)
            ms = sorted(times)[len(times) // 2]
            rate = n / (ms * 1e-3)
            results.append({
                "case": case, "stream": kind, "nrows": nrows, "ncols": ncols,
                "path": os.environ.get("MODLE_B200_REGISTER_PATH", "auto"),
                "band_bytes": 4 * (nrows * ncols + 1), "contacts": n, "ms": ms,
                "contacts_per_s": rate,
                "algorithmic_GBps": 8 * rate / 1e9,
                "sector_traffic_GBps": (64 + 8) * rate / 1e9,
                "hbm_peak_GBps": hbm,
                "frac_of_hbm_sector_ceiling": (64 + 8) * rate / 1e9 / hbm,
                "frac_of_hbm_algorithmic": 8 * rate / 1e9 / hbm,
            })
            print(json.dumps(results[-1]), flush=True)
            del b1, b2
        del band
    ctx.close()
    if args.out:
        with open(args.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
