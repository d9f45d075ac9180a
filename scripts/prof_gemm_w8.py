"""One weighted map on the fixed-point int8 tensor path at a bench-like shape (for ncu)."""
import argparse, sys
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from cassiopeia_b200 import engine
ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=50000)
ap.add_argument("--chars", type=int, default=40)
ap.add_argument("--states", type=int, default=50)
a = ap.parse_args()
args = argparse.Namespace(cells=a.cells, chars=a.chars, states=a.states, missing=0.1, mutation_rate=0.1, seed=0,
                          algo="nj", priors=True)
codes, missing_code, table, n_states = bench.make_workload(args)
cm = torch.from_numpy(codes).cuda()
n = codes.shape[0]
D = torch.empty((n, engine.padded_ld(n)), dtype=torch.float32, device="cuda")
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    engine.whd_map(cm, missing_code, table, n_states, dtype="f32", method="tensor", out=D)
    e1.record(); torch.cuda.synchronize()
print(f"{a.cells} x {a.chars} x {a.states} weighted, fixed-point int8 tensor path: {e0.elapsed_time(e1):.2f} ms")
