#!/usr/bin/env python3
"""End to end through the C ABI, the way `modle simulate` strings the same steps together:

    chrom.sizes + barrier BED6 [+ intervals BED3]           modle_b200_genome_import
      -> per interval: cell tasks (seed, jump() per cell)   modle_b200_make_cell_tasks
      -> loop-extrusion simulation on the GPU               modle_b200_simulate_interval
      -> sorted COO pixels, ready for hictk append_pixels   modle_b200_band_to_pixels
      -> 1D LEF occupancy profile (bigWig values)           modle_b200_lef_occupancy_profile

    python examples/simulate_to_pixels.py hg38.chrom.sizes barriers.bed out_prefix \
        [--intervals regions.bed] [--ncells 64] [--target-contact-density 0.1] [--seed 0]
        [--throughput-mode]

Writes <out_prefix>.pixels.tsv (bin1_id, bin2_id, count: the cooler pixel table of the run) and
<out_prefix>.lef_occupancy.bedgraph. Needs a CUDA GPU (there is no CPU fallback). Plain-text
inputs; decompress .xz/.gz first. The .cool / bigWig containers themselves are written by
hictk / libBigWig in MoDLE and are out of scope here.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from modle_b200 import abi, host  # noqa: E402
from modle_b200.simulation import Context  # noqa: E402


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawTextHelpFormatter)
    ap.add_argument("chrom_sizes")
    ap.add_argument("extrusion_barriers")
    ap.add_argument("out_prefix")
    ap.add_argument("--intervals", default="")
    ap.add_argument("--ncells", type=int, default=64)
    ap.add_argument("--target-contact-density", type=float, default=0.1)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--throughput-mode", action="store_true",
                    help="counter-based draws (statistically equivalent, not bit-identical to the "
                         "reference's draw order; modle_b200_set_rng_mode)")
    args = ap.parse_args()

    p = host.default_params()
    p.num_cells = args.ncells
    p.target_contact_density = args.target_contact_density
    p.seed = args.seed
    host.transform_params(p)
    genome = host.import_genome(args.chrom_sizes, args.extrusion_barriers, p, args.intervals)
    ctx = Context(args.device, rng_mode=1 if args.throughput_mode else 0)
    with open(args.out_prefix + ".pixels.tsv", "w") as px_out, \
            open(args.out_prefix + ".lef_occupancy.bedgraph", "w") as occ_out:
        for g in genome:
            if len(g["barriers"]) == 0:  # scheduler_simulate.cpp:111-124: skipped
                continue
            size = g["end"] - g["start"]
            iv = abi.Interval(g["chrom_size"], g["start"], g["end"], host.compute_num_lefs(p, size))
            tasks = host.make_cell_tasks(p, g["chrom_name"], iv)
            band, occ, stats, missed = ctx.simulate_interval(p, iv, g["barriers"], tasks)
            nrows, ncols = host.band_shape(p, size)
            pixels = ctx.band_to_pixels(band, nrows, ncols, g["bin_offset"])
            np.savetxt(px_out, np.column_stack([pixels["bin1_id"], pixels["bin2_id"],
                                                pixels["count"]]), fmt="%d", delimiter="\t")
            profile = ctx.lef_occupancy_profile(occ)
            bs = int(p.bin_size)
            for i, v in enumerate(profile):
                lo = g["start"] + i * bs
                occ_out.write(f"{g['chrom_name']}\t{lo}\t{min(lo + bs, g['end'])}\t{v:.6g}\n")
            print(f"{g['chrom_name']}:{g['start']}-{g['end']}: {int(stats['num_contacts'].sum())} "
                  f"contacts ({missed} outside the band), {len(pixels)} non-zero pixels, "
                  f"{int(stats['num_epochs'].sum())} cell-epochs", file=sys.stderr)
    ctx.close()


if __name__ == "__main__":
    main()
