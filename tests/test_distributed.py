"""Multi-GPU host logic on CPU: the shard planner, and the sharded runner over a world_size-2
gloo group with the kernel emulation as the compute engine (results must equal the oracle's
unsharded run bit for bit, whatever the split)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from modle_b200 import distributed

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("slice_all", [False, True])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("num_lefs,cells", [([4979], 8192), ([1289, 900, 0, 3000, 40], 512),
                                            ([100] * 24, 7), ([5000, 10], 1)])
def test_plan_covers_every_cell_once(world, num_lefs, cells, slice_all):
    shards = distributed.plan_shards(num_lefs, cells, world, slice_all=slice_all)
    assert shards == distributed.plan_shards(num_lefs, cells, world, slice_all=slice_all)
    for i, n in enumerate(num_lefs):
        mine = sorted((s.cell_lo, s.cell_hi) for s in shards if s.interval == i)
        if n == 0:
            assert mine == []
            continue
        assert mine[0][0] == 0 and mine[-1][1] == cells
        assert all(a[1] == b[0] for a, b in zip(mine, mine[1:]))
    assert all(0 <= s.rank < world for s in shards)


def test_sliced_plan_gives_every_rank_the_same_mix_and_rotates_the_roots():
    sh = distributed.plan_shards([4979, 3000, 1289, 900], 512, 8, slice_all=True)
    for r in range(8):
        assert sorted((s.interval, s.cell_hi - s.cell_lo) for s in sh if s.rank == r) == \
            [(i, 64) for i in range(4)]
    roots = distributed.interval_roots(sorted(sh, key=lambda s: (s.interval, s.cell_lo)))
    assert [roots[i][0] for i in range(4)] == [0, 1, 2, 3]
    assert all(len(roots[i][1]) == 8 for i in range(4))


def test_cost_weights_follow_the_launch_geometry(product_lib):
    from modle_b200 import host

    # chr20 / chr13 / chr1 shapes: three, two and one cell per SM
    assert host.launch_geometry(1289, 1132)[:2] == (256, 3)
    assert host.launch_geometry(2287, 943)[:2] == (512, 2)
    assert host.launch_geometry(4979, 3518)[:2] == (1024, 1)
    assert host.launch_geometry(4979, 3518)[2] < 227 * 1024
    c20, c13, c1 = (distributed.cell_cost(*x) for x in ((1289, 1132), (2287, 943), (4979, 3518)))
    assert c20 < c13 < c1
    # a chr1 cell owns its SM: it costs more per LEF than three co-resident chr20 cells do
    assert c1 / 4979 > 0.9 * c20 / 1289 and c1 > 3.0 * c20
    assert distributed.cell_cost(0, 10) == 0.0

    class Iv:
        def __init__(self, n, nb):
            self.num_lefs, self.barriers = n, [0] * nb
    w = distributed.interval_weights([Iv(1289, 1132), Iv(900, 0), Iv(4979, 3518)])
    assert w[1] == 0.0 and w[0] == c20 and w[2] == c1


def test_plan_balances_and_prefers_whole_intervals():
    # one big chromosome, 8 ranks: equal cell ranges, one per rank
    sh = distributed.plan_shards([4979], 8192, 8)
    assert sorted(s.rank for s in sh) == list(range(8))
    assert {s.cell_hi - s.cell_lo for s in sh} == {1024}
    # many similar intervals: nothing is split
    sh = distributed.plan_shards([1000 + 10 * i for i in range(24)], 512, 4)
    assert all(s.cell_lo == 0 and s.cell_hi == 512 for s in sh)
    load = [sum(s.weight for s in sh if s.rank == r) for r in range(4)]
    assert max(load) <= 1.10 * sum(load) / 4
    roots = distributed.interval_roots(distributed.plan_shards([4979], 8192, 8))
    assert roots[0][0] == 0 and len(roots[0][1]) == 8


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_run_sharded_world2_gloo_throughput_mode_is_split_invariant(tmp_path):
    """Throughput mode: a cell's result depends on its task only, so the sharded world-2 run
    (chrA's cells split 2 + 4 over the ranks, summed with one reduce) equals the unsharded one."""
    import emu_lib
    from dist_worker import make_genome
    from modle_b200 import host

    port = _free_port()
    env = dict(os.environ, MODLE_B200_RNG_MODE="1")
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "dist_worker.py"), str(r), "2",
                               str(port), str(tmp_path), "1"], env=env) for r in range(2)]
    for pr in procs:
        assert pr.wait(timeout=600) == 0
    res = [np.load(tmp_path / f"rank{r}.npz") for r in range(2)]
    p, genome = make_genome()
    emu_lib.set_rng_mode(1)
    try:
        for idx, (name, iv, bars) in enumerate(genome):
            if len(bars) == 0:
                continue
            tasks = host.make_cell_tasks(p, name, iv)
            band, occ, stats, missed = emu_lib.simulate_interval(p, iv, bars, tasks)
            root = int([r for r in res if f"band{idx}" in r][0][f"root{idx}"][0])
            assert np.array_equal(res[root][f"band{idx}"], band), name
            assert np.array_equal(res[root][f"occ{idx}"][:len(occ)], occ), name
    finally:
        emu_lib.set_rng_mode(0)


@pytest.mark.parametrize("force_split", [False, True])
def test_run_sharded_world2_gloo_matches_oracle(tmp_path, force_split):
    from dist_worker import make_genome
    from oracle import pyoracle

    port = _free_port()
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "dist_worker.py"), str(r), "2",
                               str(port), str(tmp_path), "1" if force_split else "0"])
             for r in range(2)]
    for pr in procs:
        assert pr.wait(timeout=600) == 0
    res = [np.load(tmp_path / f"rank{r}.npz") for r in range(2)]
    p, genome = make_genome()
    for idx, (name, iv, bars) in enumerate(genome):
        if len(bars) == 0:  # skipped like the reference does (scheduler_simulate.cpp:111-124)
            assert all(f"band{idx}" not in r for r in res)
            continue
        from modle_b200 import host

        tasks = host.make_cell_tasks(p, name, iv)
        band, occ, stats, missed = pyoracle.simulate_interval(p, iv, bars, tasks, nthreads=4)
        holders = [r for r in res if f"band{idx}" in r]
        assert holders
        root = int(holders[0][f"root{idx}"][0])
        assert all(int(r[f"root{idx}"][0]) == root for r in holders)
        rr = res[root]
        assert np.array_equal(rr[f"band{idx}"], band), name
        assert np.array_equal(rr[f"occ{idx}"][:len(occ)], occ), name
        assert int(rr[f"missed{idx}"][0]) == missed
        assert sum(int(r[f"ncells{idx}"][0]) for r in holders) == len(tasks)
        assert sum(int(r[f"contacts{idx}"][0]) for r in holders) == int(stats["num_contacts"].sum())
    if force_split:
        assert int(res[0]["calls"][0]) == 1 and int(res[1]["calls"][0]) == 2


@pytest.mark.parametrize("faulty_rank", [0, 1])
def test_a_device_fault_on_one_rank_fails_the_sharded_run_on_every_rank(tmp_path, faulty_rank):
    """A faulting cell leaves a truncated band: run_sharded must not return it as a result, and
    every rank has to learn about it (MAX all-reduce of the fault code)."""
    port = _free_port()
    env = dict(os.environ, MODLE_B200_TEST_INJECT_FAULT=str(faulty_rank))
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "dist_worker.py"), str(r), "2",
                               str(port), str(tmp_path), "1"], env=env) for r in range(2)]
    assert [pr.wait(timeout=600) for pr in procs] == [7, 7]
    assert not list(tmp_path.glob("rank*.npz"))
