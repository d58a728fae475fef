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


C4_SIZE = 64_444_167
C5_SIZE = 242_193_529


def spec(name, ncells=None):
    """(Config overrides, genome) of a BASELINE config, as plain data: the product builds its
    Config from it (config_*), the CPU oracle its own parameters (oracle/pyparams.py)."""
    if name == "c1":    # chr20, defaults
        return dict(num_cells=ncells or 512), genome({"chr20"})
    if name == "c2":    # genome-wide GRCh38, defaults
        return dict(num_cells=ncells or 512), genome()
    if name == "c3":    # chr1, 8192 cells
        return dict(num_cells=ncells or 8192), genome({"chr1"})
    if name == "c4":    # high collision: LEF density x4, barriers 1 / 15 kb, bypass 0.01
        recs = synthetic_barrier_records(C4_SIZE, C4_SIZE // 15_000, seed=20260117 + 100)
        return dict(num_cells=ncells or 512, number_of_lefs_per_mbp=80.0,
                    probability_of_extrusion_unit_bypass=0.01), [("chr20", C4_SIZE, 0, C4_SIZE, recs)]
    if name == "c5":    # chr2 at 1 kb bins, 3 Mbp diagonal width (2.9 GB band)
        g = genome({"chr2"})
        return dict(num_cells=ncells or 512, bin_size=1000, diagonal_width=3_000_000), g
    raise ValueError(name)


def _config(name, ncells, **more):
    overrides, g = spec(name, ncells)
    overrides.update(more)
    return Config(**overrides).transform(), g


def config_c1(ncells=512):
    """BASELINE C1: chr20, defaults."""
    return _config("c1", ncells)


def config_c2(ncells=512):
    """BASELINE C2: genome-wide GRCh38, defaults."""
    return _config("c2", ncells)


def config_c3(ncells=8192):
    """BASELINE C3: chr1, 8192 cells."""
    return _config("c3", ncells)


def config_c4(ncells=512, **more):
    """BASELINE C4: high-collision regime -- chr20 size, LEF density x4, dense synthetic barriers
    (1 per 15 kb), bypass probability 0.01."""
    return _config("c4", ncells, **more)


def config_c5(ncells=512, **more):
    """BASELINE C5: fine-resolution contact scatter -- chr2, 1 kb bins, 3 Mbp diagonal width
    (3000 x 242,194 pixels = 2.9 GB band); `target_contact_density` sets the contact volume."""
    return _config("c5", ncells, **more)


def estimate_lef_updates(params, genome_list, mean_epochs=630):
    """Rough work estimate used only to size bounded CPU samples."""
    return sum(host.compute_num_lefs(params, e - s) for _, _, s, e, _ in genome_list) * \
        int(params.num_cells) * mean_epochs
