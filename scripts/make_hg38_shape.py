#!/usr/bin/env python3
"""Derives modle_b200/data/hg38_shape.json from the reference's example inputs
(/root/reference/examples/data/hg38.chrom.sizes and hg38_extrusion_barriers.bed.xz): chromosome
names/sizes and, per chromosome, only the NUMBER of barriers, the '+' strand fraction and the
score range. bench.py regenerates synthetic barrier positions of that shape (there is no network
and the reference tree does not travel to the GPU box). Run in the build container only."""
import json
import lzma
import os

REF = "/root/reference/examples/data"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "modle_b200",
                   "data", "hg38_shape.json")
sizes = []
for line in open(os.path.join(REF, "hg38.chrom.sizes")):
    name, size = line.split()[:2]
    sizes.append((name, int(size)))
per = {n: dict(count=0, plus=0, smin=1.0, smax=0.0, ssum=0.0) for n, _ in sizes}
with lzma.open(os.path.join(REF, "hg38_extrusion_barriers.bed.xz"), "rt") as fh:
    for line in fh:
        f = line.rstrip("\n").split("\t")
        if len(f) < 6 or f[0] not in per:
            continue
        d = per[f[0]]
        s = float(f[4])
        d["count"] += 1
        d["plus"] += f[5] == "+"
        d["smin"] = min(d["smin"], s)
        d["smax"] = max(d["smax"], s)
        d["ssum"] += s
chroms = []
for n, size in sizes:
    d = per[n]
    chroms.append(dict(name=n, size=size, num_barriers=d["count"],
                       plus_fraction=round(d["plus"] / max(d["count"], 1), 4),
                       score_min=round(d["smin"], 4), score_max=round(d["smax"], 4),
                       score_mean=round(d["ssum"] / max(d["count"], 1), 4)))
json.dump(dict(assembly="hg38", source="examples/data of paulsengroup/modle v1.1.0 (counts only)",
               chromosomes=chroms), open(OUT, "w"), indent=1)
print(OUT, sum(c["num_barriers"] for c in chroms), "barriers", sum(c["size"] for c in chroms), "bp")
