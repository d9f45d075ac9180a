import sys, os
import numpy as np, torch
sys.path.insert(0, ".")
from cassiopeia_b200 import engine

def synth(n, m, S, seed=0, missing=0.1):
    rng = np.random.default_rng(seed)
    cm = rng.integers(0, S + 1, size=(n, m)).astype(np.int32)
    cm[rng.random((n, m)) < 0.6] = 0
    cm[rng.random((n, m)) < missing] = -1
    w = -np.log(rng.dirichlet(np.ones(S), size=m))
    return cm, np.concatenate([np.zeros((m, 1)), w], axis=1)

def timeit(fn):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
    return r, e0.elapsed_time(e1)

for n in (3000, 10000, 20000):
    for algo, mode, dt in (("nj", "fast", "f32"), ("upgma", "strict", "f64")):
        cm, w = synth(n, 40, 50)
        wt = w if algo == "nj" else None
        res = {}
        for flag in ("0", "1", "0", "1"):
            os.environ["CAS_PERSISTENT"] = flag
            D = engine.whd_map(cm, -1, wt, 50, dtype=dt)
            (_, _), t = timeit(lambda: engine.agglomerate(D, n, algo, mode))
            res[flag] = t
        print(f"n={n} {algo} {mode}: multi-kernel {res['0']:.1f} ms, persistent {res['1']:.1f} ms", flush=True)
