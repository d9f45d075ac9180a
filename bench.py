#!/usr/bin/env python
"""Benchmark of the distance-solver hot path on B200 (driver contract: one JSON line on stdout).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config c1|c2|c3|c4]

metric  : "NJ solve seconds @ n cells" (BASELINE.json) -- lower is better ("UPGMA solve seconds" for --config c2).
workload: default --config c3 = BASELINE.json configs[2]: NeighborJoiningSolver, 50k cells x 60 characters, mutation
          priors, 20 % missing data, implicit root (50 001 rows).  One step = one pass of the hot path:
          dissimilarity map of the device-resident character matrix + the complete loop (n - 2 joins) -> join list.
          --config c1 / c2 / c4 select the other BASELINE configs (1024 x 40 NJ; 10k x 40 UPGMA plain Hamming;
          200k x 40 NJ), --cells/--chars/--states/--missing/--algo/--no-priors override single fields.
mode    : default "exact": float64 map, the loop reproduces the REFERENCE's join list (sequential row sums where
          the maintained ones cannot decide).  "fast" = float32 map (reference tree absent near-ties),
          "strict" = every row sum re-accumulated each iteration (verification).
value   : device time per step, inputs resident in HBM (CUDA events, max over ranks).
e2e     : the same through the C-ABI call ``cas_solve_host`` from pinned HOST buffers (H2D of the character matrix +
          weights and D2H of the join list inside the timed region); ``e2e.solver_api`` = one solve through the
          drop-in class (pandas in, populated tree out).
parity  : join-list hash, near-tie / exact-tie counts and the exact mode's resolution statistics of the timed solve.
roofline: the TIMED loop.  The pruned scan is exact and skips ~99.98 % of SURVEY 8(d)'s algorithmic bytes, so the
          loop is latency-bound: ``achieved`` = algorithmic bytes / loop time (above the HBM peak by construction),
          ``traffic`` = bytes the scan kernels actually read per iteration (kernel counter) and ``latency_model``
          puts the measured microseconds per iteration next to the measured launch floor of the same launch
          pattern.  ``dense_scan`` is the HBM-bound kernel (CAS_PRUNE=0) measured live on a bounded sample.
--impl reference: the CPU restatement of the reference algorithm (oracle/, all host threads) on a bounded sample
          (a full solve at fewer cells), scaled to the workload by its n^2 / n^3 work (factors in the JSON).
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONFIGS = {
    # BASELINE.json configs[0..3]
    "c1": dict(algo="nj", cells=1024, chars=40, states=50, missing=0.1, priors=True),
    "c2": dict(algo="upgma", cells=10000, chars=40, states=50, missing=0.1, priors=False),
    "c3": dict(algo="nj", cells=50000, chars=60, states=100, missing=0.2, priors=True),
    "c4": dict(algo="nj", cells=200000, chars=40, states=50, missing=0.1, priors=True),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--algo", default=None, choices=["nj", "upgma"])
    ap.add_argument("--cells", type=int, default=None)
    ap.add_argument("--chars", type=int, default=None)
    ap.add_argument("--states", type=int, default=None)
    ap.add_argument("--missing", type=float, default=None)
    ap.add_argument("--no-priors", action="store_true")
    ap.add_argument("--mutation-rate", type=float, default=0.1)
    ap.add_argument("--mode", default="exact", choices=["exact", "fast", "strict"])
    ap.add_argument("--map-method", default=None, choices=["exact", "tensor"])
    ap.add_argument("--cpu-sample-cells", type=int, default=7000)
    ap.add_argument("--sharded-loop", action="store_true",
                    help="N > 1: shard the loop even when the map fits one GPU (default: only beyond one GPU's memory)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    for key in ("algo", "cells", "chars", "states", "missing"):
        if getattr(args, key) is None:
            setattr(args, key, cfg[key])
    args.priors = cfg["priors"] and not args.no_priors
    return args


def metric_name(args):
    return "NJ solve seconds @ n cells" if args.algo == "nj" else "UPGMA solve seconds @ n cells"


def make_workload(args, cells=None):
    """Synthetic character matrix of the workload (NJ: with the implicit root row), in the reference's map row
    order, as compact int32 codes + dense float64 weight table (negative_log priors; None without priors)."""
    from cassiopeia_b200 import host_prep
    from cassiopeia_b200.solver import solver_utilities
    from cassiopeia_b200.synthetic import simulate_character_matrix

    n = cells or args.cells
    cm, priors, _ = simulate_character_matrix(n, args.chars, args.states, seed=args.seed,
                                              mutation_rate=args.mutation_rate,
                                              missing_fraction=args.missing)
    rows = cm.astype(np.int64)
    if getattr(args, "algo", "nj") == "nj":
        rows = np.vstack([rows, np.zeros((1, args.chars), dtype=np.int64)])
    order = host_prep.reference_row_order(rows)
    weights = solver_utilities.transform_priors(priors, "negative_log") if getattr(args, "priors", True) else None
    codes, missing_code, table, n_states = host_prep.compact_states(rows[order], -1, weights)
    return codes, missing_code, table, n_states


def scan_bytes(n_rows: int, esz: int) -> float:
    """Algorithmic HBM bytes of the loop's scans: every live unordered pair once per iteration."""
    k = np.arange(3, n_rows + 1, dtype=np.float64)
    return float((esz * k * (k - 1) / 2).sum())


def loop_work(n_rows: int) -> float:
    k = np.arange(3, n_rows + 1, dtype=np.float64)
    return float((k * k).sum())


def join_hash(joins, last2) -> str:
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(joins, dtype=np.int32).tobytes())
    h.update(np.ascontiguousarray(last2, dtype=np.int32).tobytes())
    return h.hexdigest()[:16]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------- CPU arm
def cpu_sample(args, cells):
    """Full solve of the oracle port on `cells` cells with all host threads; returns
    (map seconds, loop seconds, threads)."""
    from oracle import pyoracle as po

    codes, missing_code, table, _ = make_workload(args, cells)
    cm = codes.astype(np.int64)
    threads = po.max_threads()
    po.whd_map(cm[:64], missing_code, table, threads=threads)  # spin the thread pool up
    t0 = time.perf_counter()
    D = po.whd_map(cm, missing_code, table, threads=threads)
    t1 = time.perf_counter()
    (po.nj if args.algo == "nj" else po.upgma)(D, threads=threads)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1, threads


def cpu_scale_factors(args, cells):
    extra = 1 if args.algo == "nj" else 0
    n_s, n_w = cells + extra, args.cells + extra
    return (n_w * (n_w - 1.0)) / (n_s * (n_s - 1.0)), loop_work(n_w) / loop_work(n_s)


def scale_cpu(args, t_map, t_loop, cells):
    f_map, f_loop = cpu_scale_factors(args, cells)
    return t_map * f_map + t_loop * f_loop


def cpu_baseline_object(args, t_map, t_loop, threads, cells):
    f_map, f_loop = cpu_scale_factors(args, cells)
    full = cells >= args.cells
    return {
        "value": scale_cpu(args, t_map, t_loop, cells), "unit": "s", "cores": threads, "kind": "port",
        "sample": (f"full {args.algo.upper()} solve of oracle/oracle.c (float64, OpenMP) at {cells} cells x {args.chars} "
                   f"chars: map {t_map:.2f} s, loop {t_loop:.2f} s on {threads} threads"
                   + ("" if full else f"; scaled to {args.cells} cells by pair count (map x{f_map:.1f}) and "
                                      f"sum k^2 (loop x{f_loop:.1f})")),
        "sample_cells": cells, "sample_map_s": t_map, "sample_loop_s": t_loop,
        "scale_map": f_map, "scale_loop": f_loop, "extrapolated": not full,
        "reference_numba": "profiles/r2_reference_numba_cpu.md (the reference's own numba path timed at n <= 4096)",
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cells = min(args.cpu_sample_cells, args.cells)
    for _ in range(args.warmup):
        cpu_sample(args, min(cells, 500))
    vals, last = [], None
    t_begin = time.perf_counter()
    for _ in range(args.steps):
        t_map, t_loop, threads = cpu_sample(args, cells)
        vals.append(scale_cpu(args, t_map, t_loop, cells))
        last = (t_map, t_loop, threads)
    wall = time.perf_counter() - t_begin
    value = float(np.mean(vals))
    base = cpu_baseline_object(args, last[0], last[1], last[2], cells)
    base["value"] = value
    out = {
        "impl": "reference", "metric": metric_name(args), "value": value, "unit": "s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1),
        "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def workload_config(args, n_gpus):
    """Identical for the b200 and the reference arm (the arithmetic mode is reported under "parity" / "dtype")."""
    esz = 8 if args.mode != "fast" else 4
    rows = args.cells + (1 if args.algo == "nj" else 0)
    solver = "NeighborJoiningSolver(add_root=True)" if args.algo == "nj" else "UPGMASolver()"
    pri = "mutation priors (negative_log)" if args.priors else "no priors (plain Hamming costs 0/1/2)"
    return {
        "workload": f"{solver}, {args.cells} cells x {args.chars} characters, {args.states} states, {pri}, "
                    f"{int(round(args.missing * 100))}% missing (BASELINE.json configs[{int(args.config[1]) - 1}])",
        "cells": args.cells, "characters": args.chars, "states": args.states, "missing_fraction": args.missing,
        "algo": args.algo, "priors": bool(args.priors),
        "parallelism": f"x{n_gpus} GPUs" if n_gpus > 1 else "single GPU",
        "l2": "map (n^2 * %d B = %.1f GB) %s L2 (126 MB)" % (esz, rows * rows * esz / 1e9,
                                                              "is far larger than" if rows * rows * esz > 4e8 else
                                                              "fits in; every step rewrites the whole map first"),
    }


# --------------------------------------------------------------------------- GPU arm
def solver_api_solve(args, joins):
    """One solve of the same workload through the drop-in class (the call a Cassiopeia user makes):
    pandas character matrix + priors in, populated CassiopeiaTree out -- host preparation, H2D, map, loop, D2H
    and the tree-building tail (rooting, populate_tree, collapse_unifurcations).  Reported next to the C-ABI
    `e2e`; never fails the bench."""
    try:
        import pandas as pd
        import cassiopeia_b200 as cb
        from cassiopeia_b200.synthetic import simulate_character_matrix

        cm, priors, _ = simulate_character_matrix(args.cells, args.chars, args.states, seed=args.seed,
                                                  mutation_rate=args.mutation_rate, missing_fraction=args.missing)
        frame = pd.DataFrame(cm, index=[f"c{i}" for i in range(cm.shape[0])])
        tree = cb.CassiopeiaTree(character_matrix=frame, priors=priors if args.priors else None)
        impl = {"fast": "b200_fast", "strict": "b200_strict"}.get(args.mode, "b200_exact")
        if args.algo == "nj":
            solver = cb.NeighborJoiningSolver(add_root=True, fast=True, implementation=impl)
        else:
            solver = cb.UPGMASolver(fast=True, implementation=impl)
        t0 = time.perf_counter()
        solver.solve(tree)
        seconds = time.perf_counter() - t0
        same = bool(np.array_equal(solver.last_join_list[0], joins))
        return {"value": seconds, "unit": "s", "joins_identical_to_device_path": same,
                "leaves": len(tree.leaves), "call": f"{type(solver).__name__}(implementation='{impl}').solve(CassiopeiaTree)"}
    except Exception as e:  # noqa: BLE001 - a reporting extra must not take the bench line down
        return {"error": f"{type(e).__name__}: {e}"}


def dense_scan_sample(args, engine, torch, cm_dev, missing_code, table, n_states, D, dt, method, esz, peak):
    """The HBM-bound kernel of the path: the dense scan (CAS_PRUNE=0), first iterations at full size."""
    n = D.shape[0]
    iters = min(300, n - 2)
    os.environ["CAS_PRUNE"] = "0"
    try:
        engine.whd_map(cm_dev, missing_code, table, n_states, dtype=dt, method=method, out=D)
        engine.agglomerate(D, n, args.algo, args.mode, max_iters=8)  # warm-up
        engine.whd_map(cm_dev, missing_code, table, n_states, dtype=dt, method=method, out=D)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        engine.agglomerate(D, n, args.algo, args.mode, max_iters=iters)
        e1.record()
        torch.cuda.synchronize()
    finally:
        os.environ.pop("CAS_PRUNE", None)
    kk = np.arange(n - iters + 1, n + 1, dtype=np.float64)
    d_bytes = float((esz * kk * (kk - 1) / 2).sum())
    d_ms = e0.elapsed_time(e1)
    d_gbs = d_bytes / (d_ms * 1e-3) / 1e9
    tname = "float" if esz == 4 else "double"
    return {"kernel": f"scan_kernel<{tname},{'NJ' if args.algo == 'nj' else 'UPGMA'}>", "bound": "hbm",
            "achieved": d_gbs, "peak": peak, "unit": "GB/s", "frac": d_gbs / peak,
            "traffic": 1.0005 * d_bytes / iters,
            "traffic_source": "algorithmic bytes x 1.0005 = dram__bytes_read / algorithmic of the ncu --set full "
                              "capture (profiles/r1_scan_kernel.md)",
            "sample": f"dense scan (CAS_PRUNE=0), first {iters} iterations at full size, {d_ms:.1f} ms incl. the "
                      "merge launches" + (" and the strict mode's row-sum launches" if args.mode == "strict" else ""),
            "algorithmic_bytes_per_launch_mean": d_bytes / iters, "launch_ms_mean_incl_merge": d_ms / iters}


def run_b200(args):
    import torch

    from cassiopeia_b200 import engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        from cassiopeia_b200 import sharded
        return sharded.bench_entry(args, sys.modules[__name__])

    torch.cuda.set_device(local_rank)
    dt = "f32" if args.mode == "fast" else "f64"
    esz = 4 if dt == "f32" else 8
    codes, missing_code, table, n_states = make_workload(args)
    # unweighted maps: the int8 tcgen05 kernel (bit-exact); prior-weighted maps: the exact CUDA-core kernel
    method = args.map_method or ("tensor" if (table is None or (args.mode == "fast" and n_states <= 64)) else "exact")
    n = codes.shape[0]
    cm_dev = torch.from_numpy(codes).cuda()
    ld = engine.padded_ld(n)
    D = torch.zeros((n, ld), dtype=torch.float32 if dt == "f32" else torch.float64, device="cuda")

    launches = [0]
    stats = {}

    def step():
        engine.whd_map(cm_dev, missing_code, table, n_states, dtype=dt, method=method, out=D)
        launches[0] += engine.last_launch_count()
        ev_mid.record()
        joins, last2 = engine.agglomerate(D, n, args.algo, args.mode)
        launches[0] += engine.last_launch_count()
        stats.update(engine.last_stats)
        return joins, last2

    ev_mid = torch.cuda.Event(enable_timing=True)
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches[0] = 0
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    mids = []
    torch.cuda.synchronize()
    ev[0].record()
    for s in range(args.steps):
        ev_mid = torch.cuda.Event(enable_timing=True)
        joins, last2 = step()
        mids.append(ev_mid)
        ev[s + 1].record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    total_ms = ev[0].elapsed_time(ev[-1])
    map_ms = float(np.mean([ev[s].elapsed_time(mids[s]) for s in range(args.steps)]))
    loop_ms = float(np.mean([mids[s].elapsed_time(ev[s + 1]) for s in range(args.steps)]))
    ms_per_step = total_ms / args.steps
    assert joins.shape == (n - 2, 2) and int(joins.max()) < 2 * n - 2

    peak, peak_src = peaks()
    b_alg = scan_bytes(n, esz)
    achieved = b_alg / (loop_ms * 1e-3) / 1e9
    n_scans = n - 2
    scanned = int(stats.get("scanned_elements", 0))
    pruned = scanned > 0
    tname = "float" if esz == 4 else "double"
    aname = "NJ" if args.algo == "nj" else "UPGMA"
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": peak_src, "map_ms": map_ms, "loop_ms": loop_ms,
                "map_gpairs_per_s": n * (n - 1) / 2 / (map_ms * 1e-3) / 1e9,
                "map_kernel": "whd_exact_kernel (CUDA cores, float64 in character order)" if method == "exact"
                              else ("whd_gemm2_kernel (tcgen05 int8, SM pair)" if table is None else
                                    "whd_gemm_kernel<2> (tcgen05 int8, four 7-bit digit planes of the weights)"),
                "algorithmic_bytes_per_step": b_alg, "algorithmic_bytes_per_launch_mean": b_alg / n_scans}
    if pruned:
        per_iter_us = 1e3 * loop_ms / n_scans
        floor_us = engine.loop_launch_floor_us(n, 2000, 3)
        roofline.update({
            "kernel": (f"scan_pruned_kernel<{tname},NJ> (+ prune_rowtest_kernel, merge_kernel): the timed loop"
                       if args.algo == "nj" else
                       f"keyed_scan_kernel<{tname}> (+ keyed_rowtest_kernel, merge_kernel<{tname},UPGMA>): the timed "
                       "loop, tie-aware bounds (DESIGN 4.5)"),
            "traffic": scanned * esz / n_scans,
            "read_fraction_of_algorithmic": scanned * esz / b_alg,
            "note": "exact pruned scan (DESIGN 4.1): bounds on the criterion prove almost every algorithmic byte "
                    "cannot hold the minimum, so frac = algorithmic bytes / time / HBM peak exceeds 1 by construction; "
                    "the loop is bound by its per-iteration latency, see latency_model",
            "latency_model": {
                "launches_per_iteration": 3, "us_per_iteration": per_iter_us,
                "launch_floor_us_per_iteration": floor_us, "floor_over_measured": floor_us / per_iter_us,
                "floor": "the same three programmatically chained launches (grids of row test, pruned scan and merge "
                         "at full size) with empty kernels, measured live (cas_loop_launch_floor)",
                "dependent_round_trips": ("row test 3, scan 7 (work list, bounds, map, exact re-evaluation, CTA "
                                          "reduction, grid reduction, publish), merge 5 (selection, rows, partial "
                                          "sums, grid reduction, warm start): ~15 x 0.6-1 us of L2 / HBM latency"
                                          if args.algo == "nj" else
                                          "row test 3 (+ one warp per row listed in the previous iteration: its "
                                          "row-level bound rebuilt from the sub-segment bounds), scan 7 (work list, "
                                          "bounds + partner ids, map + column ids, CTA reduction, grid reduction, "
                                          "publish), merge 4 (selection, rows, grid sync, warm start)"),
                "bytes_read_per_iteration": scanned * esz / n_scans}})
        roofline["dense_scan"] = dense_scan_sample(args, engine, torch, cm_dev, missing_code, table, n_states, D, dt,
                                                   method, esz, peak)
    else:
        roofline.update({"kernel": f"scan_kernel<{tname},{aname}>", "traffic": 1.0005 * b_alg / n_scans,
                         "traffic_source": "algorithmic bytes x 1.0005 (ncu, profiles/r1_scan_kernel.md)",
                         "launch_ms_mean_incl_merge": loop_ms / n_scans,
                         "note": "dense scan: every live pair read every iteration; below ~4k live rows the map is "
                                 "L2-resident and the iteration is launch-bound"})

    out = {
        "metric": metric_name(args), "value": ms_per_step / 1e3, "unit": "s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": dt, "data": "synthetic",
        "config": workload_config(args, 1),
        "mode": f"{args.mode}/{method}-map",
        "clocks": clocks, "gpu_launches": int(launches[0]), "roofline": roofline,
        "parity": {
            "mode": {"exact": "reference join list (maintained sums + sequential sums where ambiguous)",
                     "strict": "reference join list (row sums re-accumulated every iteration)",
                     "fast": "float32 map: reference tree only absent near-ties"}[args.mode],
            "join_hash": join_hash(joins, last2), "joins": int(joins.shape[0]),
            "near_ties": int(stats.get("near_ties", 0)), "exact_ties": int(stats.get("exact_ties", 0)),
            "ambiguous_iterations": int(stats.get("ambiguous_iterations", 0)),
            "rows_resolved": int(stats.get("rows_resolved", 0)),
            "max_rows_resolved": int(stats.get("max_rows_resolved", 0)),
            "pairs_changed_by_resolution": int(stats.get("pairs_changed", 0)),
            "checked_against": "tests/test_gpu_exact.py (oracle join lists up to 4096 cells x 3 seeds, GPU strict at "
                               "20k); profiles/r2_exact_time.md"},
    }

    if not args.no_e2e:
        del D
        torch.cuda.empty_cache()
        cm_pin = torch.from_numpy(codes).pin_memory()
        w_pin = torch.from_numpy(table).pin_memory() if table is not None else None
        # cas_solve_host reads the pinned buffers directly (ctypes pointers), synchronous call; one untimed call first
        # (the device-timed value above had its warm-up steps too: first-use costs of the library's own stream and
        # allocations are not part of a step)
        engine.solve_host(cm_pin.numpy(), missing_code, w_pin.numpy() if w_pin is not None else None, n_states,
                          args.algo, args.mode, method)
        t0 = time.perf_counter()
        j_e2e, l_e2e, tm = engine.solve_host(cm_pin.numpy(), missing_code, w_pin.numpy() if w_pin is not None else None,
                                             n_states, args.algo, args.mode, method)
        e2e_s = time.perf_counter() - t0
        engine.solve_host_release()  # the library keeps its device buffers between calls: free the 20 GB map
        assert np.array_equal(j_e2e, joins), "host-buffer path and device path disagree"
        out["e2e"] = {"value": e2e_s, "unit": "s",
                      "h2d_bytes_per_step": int(codes.nbytes + (table.nbytes if table is not None else 0)),
                      "d2h_bytes_per_step": int(j_e2e.nbytes + l_e2e.nbytes),
                      "device_ms": {"h2d": tm[0], "map": tm[1], "loop": tm[2], "d2h": tm[3]},
                      "api": "cas_solve_host (C ABI, pinned host buffers)"}
        out["e2e"]["solver_api"] = solver_api_solve(args, joins)
    if not args.no_cpu_baseline:
        cells = min(args.cpu_sample_cells, args.cells)
        t_map, t_loop, threads = cpu_sample(args, cells)
        out["cpu_baseline"] = cpu_baseline_object(args, t_map, t_loop, threads, cells)
    print(json.dumps(out), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
