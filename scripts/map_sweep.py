"""BASELINE.json config 5: dissimilarity-map-only sweep (n, characters, states; priors on/off) on one
B200, exact CUDA-core kernel vs tcgen05 tensor kernel, with the CPU oracle (reference algorithm
restated in C, all host threads) timed beside it on a bounded row block.  Prints a markdown table."""
import argparse, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from cassiopeia_b200 import engine
from cassiopeia_b200.synthetic import simulate_character_matrix
from oracle import pyoracle as po

ap = argparse.ArgumentParser()
ap.add_argument("--quick", action="store_true")
ap.add_argument("--ns", default="", help="comma-separated cell counts (overrides the default grid)")
ap.add_argument("--ms", default="", help="comma-separated character counts")
ap.add_argument("--Ss", default="", help="comma-separated state counts")
a = ap.parse_args()

def gpu_time(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): out = fn()
    e1.record(); torch.cuda.synchronize()
    return out, e0.elapsed_time(e1) / reps

import json, os
PEAK_BF16, PEAK_HBM = 1382.8e12, 6545.9e9
if os.path.exists("MEASURED_PEAKS.json"):
    pk = json.load(open("MEASURED_PEAKS.json"))
    PEAK_BF16 = float(pk.get("bf16_tflops", PEAK_BF16 / 1e12)) * 1e12
    PEAK_HBM = float(pk.get("hbm_gbs", PEAK_HBM / 1e9)) * 1e9

def cublaslt_int8(nn=16384, reps=10):
    """cuBLASLt's int8 GEMM as torch._int_mm reaches it (reported for reference: it is far from the pipe's peak)."""
    A = torch.randint(-8, 8, (nn, nn), dtype=torch.int8, device="cuda")
    B = torch.randint(-8, 8, (nn, nn), dtype=torch.int8, device="cuda").t().contiguous().t()
    _, ms_ = gpu_time(lambda: torch._int_mm(A, B), reps)
    return 2.0 * nn ** 3 / (ms_ * 1e-3)

# measured denominator of the tensor rows: issue rate of tcgen05.mma kind::i8 on shared-memory operands
PEAK_I8 = engine.int8_mma_peak_tops() * 1e12
print(f"measured peaks: int8 tensor pipe (tcgen05.mma kind::i8 128x256x32 back to back, all SMs) {PEAK_I8 / 1e12:.0f} TOP/s; "
      f"cuBLASLt int8 GEMM via torch._int_mm at 16384^3 {cublaslt_int8() / 1e12:.0f} TOP/s; bf16 {PEAK_BF16 / 1e12:.0f} "
      f"TFLOP/s; HBM copy {PEAK_HBM / 1e9:.0f} GB/s\n")
ns = (1000, 5000, 20000, 50000) if a.quick else (1000, 2000, 5000, 10000, 20000, 50000, 100000, 200000)
ms = (10, 40, 200) if a.quick else (10, 40, 100, 200)
Ss = (10, 100) if a.quick else (10, 50, 100)
if a.ns: ns = tuple(int(x) for x in a.ns.split(","))
if a.ms: ms = tuple(int(x) for x in a.ms.split(","))
if a.Ss: Ss = tuple(int(x) for x in a.Ss.split(","))
print("| n | m | S | priors | exact ms | exact Gpairs/s | tensor ms | tensor Gpairs/s | tensor alg TFLOP/s | roofline bound Gpairs/s | tensor frac | CPU oracle Mpairs/s (threads) | max rel err |")
print("|---:|---:|---:|:---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
threads = po.max_threads()
for n in ns:
    for m in ms:
        for S in Ss:
            if n * n * 4 > 60e9 and not (n == 200000 and m == 40 and S == 50): continue  # 200k: the C4 shape only
            cm, priors, table = simulate_character_matrix(n, m, S, seed=1, mutation_rate=0.7, missing_fraction=0.1)
            with np.errstate(invalid="ignore"): w = np.nan_to_num(-np.log(table))
            out = torch.zeros((n, engine.padded_ld(n)), dtype=torch.float32, device="cuda")
            for weighted in (False, True):
                ww = w if weighted else None
                ref, te = gpu_time(lambda: engine.whd_map(cm, -1, ww, S, dtype="f32", method="exact", out=out))
                ref = ref[:min(n, 2048), :n].double().clone()
                D, tt = gpu_time(lambda: engine.whd_map(cm, -1, ww, S, dtype="f32", method="tensor", out=out))
                err = ((D[:min(n, 2048), :n].double() - ref).abs() / ref.abs().clamp(min=1.0)).max().item()
                pairs = n * (n - 1) / 2
                f_pair = 2.0 * m * (S + 3)
                # both tensor paths run on the int8 MMAs (unweighted: exact one-hot; weighted: fixed-point digits)
                bound = min(PEAK_I8 / f_pair, PEAK_HBM / 8.0)
                # CPU baseline on a bounded row block of the same input
                rows = max(8, min(n, int(2e7 / (n * m)) + 1))
                cmi = cm.astype(np.int64)
                po.whd_map(cmi[:64], -1, ww if ww is None else w, threads=threads)
                t0 = time.perf_counter(); po.whd_map(cmi, -1, w if weighted else None, rows=(0, rows), threads=threads)
                tc = time.perf_counter() - t0
                cpu = rows * n / tc / 1e6 / 2  # unordered-pair equivalents
                print(f"| {n} | {m} | {S} | {'yes' if weighted else 'no'} | {te:.2f} | {pairs/te/1e6:.1f} | {tt:.2f} | "
                      f"{pairs/tt/1e6:.1f} | {pairs*f_pair/tt/1e9:.1f} | {bound/1e9:.0f} | {pairs/(tt*1e-3)/bound:.3f} | "
                      f"{cpu:.1f} ({threads}) | {err:.1e} |", flush=True)
            del out, D, ref
            torch.cuda.empty_cache()
