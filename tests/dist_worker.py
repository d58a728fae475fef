"""Worker of the world_size-2 gloo test (tests/test_distributed.py). TEST INFRASTRUCTURE: the
compute engine here is the CPU emulation of the kernel (tests/emu), so that the sharding /
reduce logic of modle_b200.distributed runs without a GPU."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.dirname(HERE), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


from modle_b200.distributed import HostEngineHooks  # noqa: E402


class EmuEngine(HostEngineHooks):
    """Same interface as modle_b200.distributed.DeviceEngine, CPU tensors, emulated kernel."""

    def __init__(self):
        import torch

        self.torch = torch
        self.calls = 0

    def alloc_outputs(self, nrows, ncols):
        t = self.torch
        return (t.zeros(nrows * ncols + 1, dtype=t.int32), t.zeros(max(ncols, 1), dtype=t.int64),
                t.zeros(1, dtype=t.int64))

    def run(self, params, abi_interval, barriers, tasks, band, occ, missed):
        import emu_lib

        t = self.torch
        emu_lib.set_rng_mode(int(os.environ.get("MODLE_B200_RNG_MODE", "0")))
        b, o, st, ms = emu_lib.simulate_interval(params, abi_interval, barriers, tasks)
        band += t.from_numpy(b.view(np.int32))
        occ[:len(o)] += t.from_numpy(o.view(np.int64))
        missed += ms
        self.calls += 1
        if os.environ.get("MODLE_B200_TEST_INJECT_FAULT") == os.environ.get("MODLE_B200_TEST_RANK"):
            st = st.copy()
            st["device_fault"][-1] = 4  # what a cell that ran out of serial draws would report
        return t.from_numpy(st.view(np.uint8).copy()), None

    def join(self):
        pass


def make_genome():
    from common import make_case

    out = []
    p = None
    for k, (name, size, nbar) in enumerate([("chrA", 3_000_000, 40), ("chrB", 1_200_000, 15),
                                            ("chrC", 2_000_000, 0)]):
        p, iv, bars, tasks = make_case(size=size, ncells=6, nbar=nbar, seed=5 + k, name=name,
                                       target_contact_density=0.02)
        out.append((name, iv, bars))
    return p, out


class Interval:
    def __init__(self, name, iv, bars, params):
        from modle_b200 import host

        self.chrom_name = name
        self._iv = iv
        self.barriers = bars
        self.num_lefs = int(iv.num_lefs)
        self.nrows, self.ncols = host.band_shape(params, int(iv.end - iv.start))

    def abi_interval(self):
        return self._iv


def main(rank, world, port, outdir, force_split):
    import torch
    import torch.distributed as dist

    from modle_b200 import distributed

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p, genome = make_genome()
    intervals = [Interval(n, iv, b, p) for n, iv, b in genome]
    shards = None
    if force_split:  # chrA's cells over both ranks, chrB whole on rank 1
        S = distributed.Shard
        shards = [S(0, 0, 2, 0, 2.0), S(0, 2, 6, 1, 4.0), S(1, 0, 6, 1, 1.0)]
    eng = EmuEngine()
    os.environ["MODLE_B200_TEST_RANK"] = str(rank)
    try:
        out = distributed.run_sharded(eng, p, intervals, rank, world, dist, shards=shards)
    except Exception as e:  # the fault-injection test expects this on EVERY rank
        from modle_b200 import abi, host

        ok = isinstance(e, host.ModleB200Error) and e.code == abi.ERR_DEVICE_FAULT
        if not ok:
            import traceback

            traceback.print_exc()
        dist.destroy_process_group()
        sys.exit(7 if ok else 1)
    res = {}
    for idx, o in out.items():
        res[f"band{idx}"] = o["band"].numpy().view(np.uint32)
        res[f"occ{idx}"] = o["occ1d"].numpy().view(np.uint64)
        res[f"missed{idx}"] = np.array([int(o["missed"].item())])
        res[f"root{idx}"] = np.array([o["root"]])
        res[f"ncells{idx}"] = np.array([sum(len(s) for s in o["stats"])])
        res[f"contacts{idx}"] = np.array([sum(int(s["num_contacts"].sum()) for s in o["stats"])])
    res["calls"] = np.array([eng.calls])
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), **res)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5] == "1")
