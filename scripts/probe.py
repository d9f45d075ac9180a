"""Dev probe: times the map + loops on synthetic data (not a benchmark)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from cassiopeia_b200 import engine

def synth(n, m, S, seed=0, missing=0.1):
    rng = np.random.default_rng(seed)
    cm = rng.integers(0, S + 1, size=(n, m)).astype(np.int32)
    cm[rng.random((n, m)) < 0.6] = 0
    cm[rng.random((n, m)) < missing] = -1
    w = -np.log(rng.dirichlet(np.ones(S), size=m))
    w = np.concatenate([np.zeros((m, 1)), w], axis=1)
    return cm, w

def timeit(fn):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
    return r, e0.elapsed_time(e1)

print(engine.device_info())
for (n, m, S, algo, mode, dt, iters) in [
    (2000, 40, 50, "nj", "strict", "f64", -1),
    (10000, 40, 50, "upgma", "strict", "f64", -1),
    (10000, 40, 50, "nj", "fast", "f32", -1),
    (10000, 40, 50, "nj", "strict", "f64", 200),
    (30000, 60, 100, "nj", "fast", "f32", 200),
    (50000, 60, 100, "nj", "fast", "f32", 100),
]:
    cm, w = synth(n, m, S)
    wt = w if algo == "nj" else None
    D, tmap = timeit(lambda: engine.whd_map(cm, -1, wt, S, dtype=dt, method="exact"))
    D, tmap = timeit(lambda: engine.whd_map(cm, -1, wt, S, dtype=dt, method="exact"))
    (joins, last2), tloop = timeit(lambda: engine.agglomerate(D, n, algo, mode, iters))
    esz = 8 if dt == "f64" else 4
    it = len(joins)
    ks = np.arange(n, n - it, -1, dtype=np.float64)
    bytes_alg = (esz * ks * (ks - 1) / 2).sum()
    print(f"n={n} m={m} {algo} {mode} {dt}: map {tmap:.1f} ms ({n*(n-1)/2/tmap/1e6:.1f} Gpairs/s) "
          f"loop {it} iters {tloop:.1f} ms -> {bytes_alg/tloop/1e6:.0f} GB/s algorithmic, launches {engine.last_launch_count()}", flush=True)
    del D
