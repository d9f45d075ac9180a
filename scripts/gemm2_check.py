"""A/B of the single-SM (CAS_GEMM_CTA_GROUP=1) and SM-pair (=2) tensor-core map kernels:
correctness vs the exact CUDA-core kernel, then timing.  Usage: gemm2_check.py [check|time] """
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from cassiopeia_b200 import engine
from cassiopeia_b200.synthetic import simulate_character_matrix


def inputs(n, m, S, seed=1, miss=0.15):
    cm, priors, table = simulate_character_matrix(n, m, S, seed=seed, mutation_rate=0.7, missing_fraction=miss)
    with np.errstate(invalid="ignore"):
        w = np.nan_to_num(-np.log(table))
    return cm, w


def check():
    bad = 0
    for n, m, S in ((5, 3, 2), (129, 7, 3), (255, 20, 9), (256, 20, 9), (257, 20, 9), (300, 33, 17), (1000, 40, 50),
                    (1537, 70, 20), (5001, 40, 50), (9000, 130, 30)):
        cm, w = inputs(n, m, S, seed=n)
        for ww in (None, w):
            for dt in ("f64", "f32"):
                ref = engine.whd_map(cm, -1, ww, S, dtype=dt, method="exact")[:, :n]
                os.environ["CAS_GEMM_CTA_GROUP"] = "2"
                D = engine.whd_map(cm, -1, ww, S, dtype=dt, method="tensor")[:, :n]
                torch.cuda.synchronize()
                if ww is None:
                    ok = torch.equal(D, ref)
                    msg = "bit-exact" if ok else f"MISMATCH {(D != ref).sum().item()} elements"
                else:
                    err = ((D.double() - ref.double()).abs() / ref.double().abs().clamp(min=1.0)).max().item()
                    ok = err <= 1e-6 and torch.equal(D, D.T)
                    msg = f"max rel err {err:.2e} sym={torch.equal(D, D.T)}"
                bad += not ok
                print(f"n={n} m={m} S={S} weighted={ww is not None} {dt}: {msg}", flush=True)
    print("CHECK", "FAILED" if bad else "OK", flush=True)
    return bad


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record(); torch.cuda.synchronize()
    return r, e0.elapsed_time(e1) / reps


def time_all():
    for n, m, S in ((20000, 40, 50), (20000, 60, 100), (30000, 200, 100), (50000, 60, 100)):
        cm, w = inputs(n, m, S)
        for ww in (None, w):
            line = f"n={n} m={m} S={S} weighted={ww is not None}:"
            for g in ("1", "2"):
                os.environ["CAS_GEMM_CTA_GROUP"] = g
                D, t = timeit(lambda: engine.whd_map(cm, -1, ww, S, dtype="f32", method="tensor"))
                pairs = n * (n - 1) / 2
                line += f"  cta_group::{g} {t:.2f} ms {pairs / t / 1e6:.1f} Gpairs/s {pairs * 2 * m * (S + 3) / t / 1e9:.0f} alg Tops"
                del D
            print(line, flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "check"
    sys.exit(check() if what == "check" else time_all())
