"""Synthetic inputs in the shape of the BASELINE configs (no network: positions are generated,
only chromosome sizes and per-chromosome barrier counts come from the reference's examples)."""
import json
import os

import numpy as np

from . import host
from .simulation import Config

_HERE = os.path.dirname(os.path.abspath(__file__))


def hg38_shape():
    with open(os.path.join(_HERE, "data", "hg38_shape.json")) as fh:
        return json.load(fh)["chromosomes"]


def synthetic_barrier_records(size, count, plus_fraction=0.5, score_min=0.6, score_max=1.0,
                              seed=20260117):
    """Sorted unique positions in [0, size), strand Bernoulli(plus_fraction), score uniform
    (SURVEY 8d generator)."""
    rng = np.random.default_rng(seed)
    pos = np.sort(rng.choice(size, count, replace=False)) if count else []
    return [(int(x), "+" if rng.random() < plus_fraction else "-",
             float(rng.uniform(score_min, score_max))) for x in pos]


def genome(names=None, seed=20260117):
    """[(name, size, start, end, barrier_records)] for hg38-shaped chromosomes."""
    out = []
    for i, c in enumerate(hg38_shape()):
        if names is not None and c["name"] not in names:
            continue
        recs = synthetic_barrier_records(c["size"], c["num_barriers"], c["plus_fraction"],
                                         c["score_min"], c["score_max"], seed + i)
        out.append((c["name"], c["size"], 0, c["size"], recs))
    return out


def config_c1(ncells=512):
    """BASELINE C1: chr20, defaults."""
    return Config(num_cells=ncells).transform(), genome({"chr20"})


def config_c2(ncells=512):
    """BASELINE C2: genome-wide GRCh38, defaults."""
    return Config(num_cells=ncells).transform(), genome()


def config_c3(ncells=8192):
    """BASELINE C3: chr1, 8192 cells."""
    return Config(num_cells=ncells).transform(), genome({"chr1"})


def config_c4(ncells=512):
    """BASELINE C4: high-collision regime -- chr20 size, LEF density x4, dense synthetic barriers
    (1 per 15 kb), bypass probability 0.01."""
    size = 64_444_167
    cfg = Config(num_cells=ncells, number_of_lefs_per_mbp=80.0,
                 probability_of_extrusion_unit_bypass=0.01).transform()
    recs = synthetic_barrier_records(size, size // 15_000, seed=20260117 + 100)
    return cfg, [("chr20", size, 0, size, recs)]


def estimate_lef_updates(params, genome_list, mean_epochs=630):
    """Rough work estimate used only to size bounded CPU samples."""
    return sum(host.compute_num_lefs(params, e - s) for _, _, s, e, _ in genome_list) * \
        int(params.num_cells) * mean_epochs
