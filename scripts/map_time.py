"""Device time of the exact (CUDA-core) map kernel on the bench workloads (CUDA events, best of 3)."""
import argparse, sys
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from cassiopeia_b200 import engine
ap = argparse.ArgumentParser()
ap.add_argument("--rates", type=float, nargs="+", default=[0.1, 0.3, 0.7])
a = ap.parse_args()
for (cells, chars, states, missing) in ((50000, 60, 100, 0.2), (50000, 40, 50, 0.1)):
    for rate in a.rates:
        args = argparse.Namespace(cells=cells, chars=chars, states=states, missing=missing, mutation_rate=rate, seed=0,
                                  algo="nj", priors=True)
        codes, missing_code, table, n_states = bench.make_workload(args)
        cm = torch.from_numpy(codes).cuda()
        n = codes.shape[0]
        cutfrac = float(((codes != 0) & (codes != missing_code)).mean())
        for dt in ("f64", "f32"):
            D = torch.empty((n, engine.padded_ld(n)), dtype=torch.float64 if dt == "f64" else torch.float32, device="cuda")
            best = 1e9
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                engine.whd_map(cm, missing_code, table, n_states, dtype=dt, method="exact", out=D)
                e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            print(f"{cells} x {chars} x {states}, rate {rate} (cut fraction {cutfrac:.3f}), {dt}: {best:.2f} ms = "
                  f"{n * (n - 1) / 2 / best / 1e6:.1f} Gpairs/s", flush=True)
            del D
