"""First GPU contact: CUDA path vs oracle on a few small intervals + timing of one chr20-sized run."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from modle_b200 import abi, host
from modle_b200.simulation import Context
from oracle import pyoracle


def setup(size=3_000_000, ncells=2, nbar=40, seed=1, **kw):
    p = host.default_params()
    p.num_cells = ncells
    for k, v in kw.items():
        setattr(p, k, v)
    host.transform_params(p)
    rng = np.random.default_rng(seed)
    pos = np.sort(rng.choice(size, nbar, replace=False))
    recs = [(int(x), '+' if rng.random() < 0.5 else '-', float(rng.uniform(0.6, 1.0))) for x in pos]
    bars = host.barriers_from_records(recs, p)
    iv = abi.Interval(size, 0, size, host.compute_num_lefs(p, size))
    tasks = host.make_cell_tasks(p, "chrT", iv)
    return p, iv, bars, tasks


ctx = Context(0)
ok_all = True
for kw in [dict(size=3_000_000, ncells=4, target_contact_density=0.01),
           dict(size=8_000_000, ncells=6, nbar=150, target_contact_density=0.02),
           dict(size=5_000_000, ncells=3, nbar=90, target_contact_density=0.02,
                lef_bar_major_collision_pblock=0.8, lef_bar_minor_collision_pblock=0.1),
           dict(size=5_000_000, ncells=3, nbar=90, target_contact_density=0.02,
                number_of_lefs_per_mbp=80),
           dict(size=64_444_167, ncells=8, nbar=1132, target_contact_density=0.002),
           dict(size=248_956_422, ncells=4, nbar=3518, target_contact_density=0.0005)]:
    p, iv, bars, tasks = setup(**kw)
    t0 = time.time()
    a = pyoracle.simulate_interval(p, iv, bars, tasks, nthreads=os.cpu_count())
    t1 = time.time()
    try:
        b = ctx.simulate_interval(p, iv, bars, tasks)
    except Exception as e:
        print(kw, "GPU FAILED:", e)
        ok_all = False
        continue
    t2 = time.time()
    ok = np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[3] == b[3]
    okstats = all(np.array_equal(a[2][f], b[2][f]) for f in a[2].dtype.names)
    ok_all &= ok and okstats
    print(kw, "band", ok, "stats", okstats, "contacts", int(a[0].sum()),
          "epochs", a[2]['num_epochs'][:4], b[2]['num_epochs'][:4],
          "lef_updates", int(a[2]['num_lef_updates'].sum()),
          f"oracle {t1-t0:.2f}s gpu {t2-t1:.2f}s", flush=True)
    if not okstats:
        for f in a[2].dtype.names:
            print("   ", f, a[2][f][:6], b[2][f][:6])
print("ALL OK" if ok_all else "MISMATCH")

# timing: chr20-sized, 512 cells, default density
p, iv, bars, tasks = setup(size=64_444_167, ncells=512, nbar=1132)
t0 = time.time()
b = ctx.simulate_interval(p, iv, bars, tasks)
t1 = time.time()
lu = int(b[2]['num_lef_updates'].sum())
print(f"chr20-like 512 cells: {t1-t0:.3f}s, lef_updates {lu}, {lu/(t1-t0)/1e6:.1f} M LEF-updates/s, "
      f"contacts {int(b[0].sum())}, epochs mean {b[2]['num_epochs'].mean():.1f}", flush=True)
p, iv, bars, tasks = setup(size=248_956_422, ncells=512, nbar=3518)
t0 = time.time()
b = ctx.simulate_interval(p, iv, bars, tasks)
t1 = time.time()
lu = int(b[2]['num_lef_updates'].sum())
print(f"chr1-like 512 cells: {t1-t0:.3f}s, lef_updates {lu}, {lu/(t1-t0)/1e6:.1f} M LEF-updates/s, "
      f"contacts {int(b[0].sum())}, epochs mean {b[2]['num_epochs'].mean():.1f}", flush=True)
