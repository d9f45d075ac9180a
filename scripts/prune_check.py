"""Pruned scan (default) vs dense scan (CAS_PRUNE=0): identical join lists and diagnostics, timing, and the
fraction of the dense scan's elements that were actually read.  Usage: prune_check.py [check|time]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from cassiopeia_b200 import engine
from cassiopeia_b200.synthetic import simulate_character_matrix


def build_map(n, m, S, dtype, rate, miss, weighted=True, seed=3):
    cm, priors, table = simulate_character_matrix(n, m, S, seed=seed, mutation_rate=rate, missing_fraction=miss)
    cm = np.vstack([cm, np.zeros((1, m), dtype=cm.dtype)])
    with np.errstate(invalid="ignore", divide="ignore"):
        w = np.nan_to_num(-np.log(table), nan=0.0, posinf=0.0)
    return engine.whd_map(cm, -1, w if weighted else None, S, dtype=dtype, method="exact"), n + 1


def run(D, n, algo, mode, prune):
    os.environ["CAS_PRUNE"] = "1" if prune else "0"
    if os.environ.get("PRUNE_CHECK_DEFAULT_MIN_K") is None:
        os.environ["CAS_PRUNE_MIN_K"] = "0"
    Dc = D.clone()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    joins, last2 = engine.agglomerate(Dc, n, algo, mode)
    e1.record(); torch.cuda.synchronize()
    return joins, last2, dict(engine.last_stats), e0.elapsed_time(e1)


def check():
    bad = 0
    for n, m, S, rate, miss in ((200, 20, 5, 0.7, 0.1), (1000, 40, 50, 0.1, 0.2), (3000, 40, 50, 0.7, 0.2),
                                (6001, 60, 100, 0.1, 0.2), (9000, 30, 10, 0.4, 0.3)):
        for algo, mode, dtype, weighted in (("nj", "fast", "f32", True), ("nj", "fast", "f64", True),
                                            ("upgma", "fast", "f32", True), ("upgma", "strict", "f64", False)):
            D, nn = build_map(n, m, S, dtype, rate, miss, weighted)
            jd, ld, sd, td = run(D, nn, algo, mode, False)
            jp, lp, sp, tp = run(D, nn, algo, mode, True)
            dense = sum(k * (k - 1) // 2 for k in range(3, nn + 1))
            ok = np.array_equal(jd, jp) and np.array_equal(ld, lp) and sd["near_ties"] == sp["near_ties"] \
                and sd["exact_ties"] == sp["exact_ties"]
            bad += not ok
            print(f"n={nn} m={m} S={S} rate={rate} {algo}/{mode}/{dtype}: {'identical' if ok else 'MISMATCH'} "
                  f"dense {td:.1f} ms pruned {tp:.1f} ms read {sp['scanned_elements'] / dense:.3f} of dense "
                  f"near_ties {sd['near_ties']}/{sp['near_ties']} exact {sd['exact_ties']}/{sp['exact_ties']}", flush=True)
            if not ok:
                diff = np.nonzero((jd != jp).any(axis=1))[0]
                print("  first differing join", diff[:3], jd[diff[:3]], jp[diff[:3]])
    print("CHECK", "FAILED" if bad else "OK", flush=True)
    return bad


def time_all():
    for n, m, S, rate, miss in ((20000, 60, 100, 0.1, 0.2), (20000, 60, 100, 0.7, 0.2), (50000, 60, 100, 0.1, 0.2)):
        for algo in ("nj", "upgma"):
            D, nn = build_map(n, m, S, "f32", rate, miss)
            dense = sum(k * (k - 1) // 2 for k in range(3, nn + 1))
            line = f"n={nn} rate={rate} {algo}:"
            res = {}
            for prune in ((True, False) if n <= 20000 else (True,)):
                j, l, s, t = run(D, nn, algo, "fast", prune)
                res[prune] = j
                line += f"  {'pruned' if prune else 'dense'} {t / 1e3:.2f} s"
                if prune:
                    line += f" (read {s['scanned_elements'] / dense:.3f} of dense)"
            if len(res) == 2:
                line += f"  identical={np.array_equal(res[True], res[False])}"
            print(line, flush=True)
            del D


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "check"
    sys.exit(check() if what == "check" else time_all())
