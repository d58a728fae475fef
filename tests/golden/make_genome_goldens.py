#!/usr/bin/env python3
"""Generates tests/golden/genome_goldens.json from the reference's example data set
(examples/data/hg38.chrom.sizes + hg38_extrusion_barriers.bed.xz) with the Python restatement of
the genome import (oracle/pygenome.py). Run in the build container (needs /root/reference):

    python tests/golden/make_genome_goldens.py

Stored per interval: number of barriers, number blocking REV-moving units ('+' motifs), sum of the
positions, XOR of the positions, first/last position, sum of stp_active rounded to 1e-9, and the
cooler bin offset -- enough to pin pos = (start + end + 1) / 2, the strand mapping, the
score -> stp arithmetic and the interval order without copying the data set."""
import json
import lzma
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pygenome

REF = "/root/reference/examples/data"


def summarise(genome):
    out = []
    for iv in genome:
        pos = [b[0] for b in iv["barriers"]]
        x = 0
        for p in pos:
            x ^= p
        out.append(dict(chrom=iv["chrom_name"], size=iv["chrom_size"], start=iv["start"],
                        end=iv["end"], bin_offset=iv["bin_offset"], n=len(pos),
                        n_block_rev=sum(1 for b in iv["barriers"] if b[3] == 1),
                        pos_sum=sum(pos), pos_xor=x, first=pos[0] if pos else None,
                        last=pos[-1] if pos else None,
                        stp_active_sum=round(sum(b[1] for b in iv["barriers"]), 9)))
    return out


def main():
    with tempfile.TemporaryDirectory() as d:
        bed = os.path.join(d, "barriers.bed")
        with open(bed, "wb") as f:
            f.write(lzma.open(os.path.join(REF, "hg38_extrusion_barriers.bed.xz")).read())
        g = pygenome.import_genome(os.path.join(REF, "hg38.chrom.sizes"), bed, 5000, 0.0, 0.7)
        iv_bed = os.path.join(d, "intervals.bed")
        with open(iv_bed, "w") as f:
            f.write("chr20\t10000000\t30000000\nchr1\t5000000\t9000000\nchr1\t200000000\t248956422\n")
        g2 = pygenome.import_genome(os.path.join(REF, "hg38.chrom.sizes"), bed, 5000, 0.0, 0.7,
                                    path_intervals=iv_bed)
    out = dict(source="paulsengroup/modle examples/data (hg38.chrom.sizes, "
                      "hg38_extrusion_barriers.bed.xz); defaults pbb 0.0, puu 0.7, 5 kb bins",
               whole_genome=summarise(g), sub_intervals=summarise(g2),
               sub_intervals_bed="chr20\t10000000\t30000000\nchr1\t5000000\t9000000\n"
                                 "chr1\t200000000\t248956422\n")
    with open(os.path.join(ROOT, "tests", "golden", "genome_goldens.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(sum(s["n"] for s in out["whole_genome"]), "barriers in", len(out["whole_genome"]),
          "intervals")


if __name__ == "__main__":
    main()
