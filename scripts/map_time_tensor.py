"""Weighted map: exact CUDA-core kernel vs the two tensor-core paths (device time, best of 3; max relative error)."""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from cassiopeia_b200 import engine
ap = argparse.ArgumentParser()
ap.add_argument("--rate", type=float, default=0.1)
a = ap.parse_args()
def t(fn):
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out
print("| cells x chars x states | exact ms | tensor fixed-point int8 ms | Gpairs/s | alg TFLOP/s | tensor bf16x3 ms | max rel err i8 | max rel err bf16 |")
print("|---|---:|---:|---:|---:|---:|---:|---:|")
for (cells, chars, states, missing) in ((20000, 40, 50, 0.1), (50000, 40, 50, 0.1), (50000, 60, 100, 0.2), (30000, 200, 100, 0.1)):
    args = argparse.Namespace(cells=cells, chars=chars, states=states, missing=missing, mutation_rate=a.rate, seed=0,
                              algo="nj", priors=True)
    codes, missing_code, table, n_states = bench.make_workload(args)
    cm = torch.from_numpy(codes).cuda()
    n = codes.shape[0]
    D0 = torch.empty((n, engine.padded_ld(n)), dtype=torch.float32, device="cuda")
    D1 = torch.empty_like(D0)
    te, _ = t(lambda: engine.whd_map(cm, missing_code, table, n_states, dtype="f32", method="exact", out=D0))
    os.environ["CAS_GEMM_WEIGHTED"] = "i8"
    ti, _ = t(lambda: engine.whd_map(cm, missing_code, table, n_states, dtype="f32", method="tensor", out=D1))
    ei = float(((D0[:, :n].double() - D1[:, :n].double()).abs() / D0[:, :n].double().abs().clamp(min=1.0)).max())
    os.environ["CAS_GEMM_WEIGHTED"] = "bf16"
    tb, _ = t(lambda: engine.whd_map(cm, missing_code, table, n_states, dtype="f32", method="tensor", out=D1))
    eb = float(((D0[:, :n].double() - D1[:, :n].double()).abs() / D0[:, :n].double().abs().clamp(min=1.0)).max())
    os.environ.pop("CAS_GEMM_WEIGHTED")
    pairs = n * (n - 1) / 2
    print(f"| {cells} x {chars} x {states} | {te:.2f} | {ti:.2f} | {pairs / ti / 1e6:.1f} | "
          f"{pairs * 2 * chars * (states + 3) / ti / 1e9:.0f} | {tb:.2f} | {ei:.1e} | {eb:.1e} |", flush=True)
    del D0, D1
