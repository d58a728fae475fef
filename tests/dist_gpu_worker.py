"""Worker of the multi-GPU test (tests/test_gpu_multi.py), launched by torch.distributed.run:
Simulation.run_simulate over NCCL with a plan that splits one interval's cells over the ranks;
rank 0 checks every root's result against the oracle."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.dirname(HERE), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    import torch
    import torch.distributed as dist

    from modle_b200 import abi, distributed, host, workloads
    from modle_b200.simulation import Config, Simulation
    from oracle import pyoracle

    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = Config(num_cells=16, target_contact_density=0.01).transform()
    sizes = [("chrA", 40_000_000, 700), ("chrB", 4_000_000, 70), ("chrC", 2_500_000, 0)]
    genome = [(n, s, 0, s, workloads.synthetic_barrier_records(s, nb, seed=31 + k))
              for k, (n, s, nb) in enumerate(sizes)]
    sim = Simulation(cfg, genome, device=local, rank=rank, world_size=world)
    # the plan run_simulate() makes (every rank computes the same one)
    shards = distributed.plan_shards(distributed.interval_weights(sim.intervals), 16, world)
    roots = distributed.interval_roots(shards)
    assert len(roots[0][1]) > 1, "chrA should be split over several ranks"
    sim.run_simulate()
    ok = 1
    p = cfg.params
    for idx, iv in enumerate(sim.intervals):
        if idx not in roots or roots[idx][0] != rank:
            continue
        tasks = host.make_cell_tasks(p, iv.chrom_name, iv.abi_interval())
        band, occ, stats, missed = pyoracle.simulate_interval(p, iv.abi_interval(), iv.barriers,
                                                              tasks, nthreads=8)
        good = (np.array_equal(iv.contacts, band) and np.array_equal(iv.lef_1d_occupancy, occ)
                and iv.missed_updates == missed)
        print(f"[rank {rank}] {iv.chrom_name}: root here, {int(band.sum())} contacts, "
              f"{'OK' if good else 'MISMATCH'}", flush=True)
        ok &= int(good)
    t = torch.tensor([ok], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    sim.close()
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
