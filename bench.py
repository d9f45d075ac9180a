#!/usr/bin/env python
"""Benchmark of the distance-solver hot path on B200 (driver contract: one JSON line on stdout).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

metric  : "NJ solve seconds @ n cells" (BASELINE.json) -- lower is better.
workload: BASELINE.json configs[2] = NeighborJoiningSolver, 50k cells x 60 characters, mutation
          priors, 20 % missing data, implicit root (50 001 rows), float32 "fast" mode.  One step
          = one pass of the hot path: dissimilarity map of the device-resident character matrix +
          the complete NJ loop (50 000 - 1 joins) -> join list.
value   : device time per step, inputs resident in HBM (CUDA events, max over ranks).
e2e     : the same through the C-ABI call ``cas_solve_host`` from pinned HOST buffers (H2D of the
          character matrix + weights and D2H of the join list inside the timed region).
roofline: NJ scan kernel, HBM-bound; algorithmic bytes = 4 * k(k-1)/2 per launch divided by
          the device time of the launches (scan + merge), against the measured copy bandwidth of
          MEASURED_PEAKS.json.  The timed loop uses the exact PRUNED scan, which is no HBM-bound kernel (it
          reads ~0.1 % of those bytes), so the roofline object describes the DENSE scan kernel measured live on
          a bounded sample (first 300 iterations at full size, CAS_PRUNE=0) and ``roofline.pruned_loop`` holds
          the timed loop's own figures (bytes actually read per iteration from the kernel's counter, the
          algorithmic-equivalent rate).
--impl reference: the CPU restatement of the reference algorithm (oracle/, all host threads) on a
          bounded sample (a full solve at fewer cells), scaled to the workload by its n^2 / n^3 work.
"""
from __future__ import annotations

import argparse
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

METRIC = "NJ solve seconds @ n cells"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cells", type=int, default=50000)
    ap.add_argument("--chars", type=int, default=60)
    ap.add_argument("--states", type=int, default=100)
    ap.add_argument("--missing", type=float, default=0.2)
    ap.add_argument("--mutation-rate", type=float, default=0.1)
    ap.add_argument("--mode", default="fast", choices=["fast", "strict"])
    ap.add_argument("--map-method", default=None, choices=["exact", "tensor"])
    ap.add_argument("--cpu-sample-cells", type=int, default=7000)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--seed", type=int, default=0)
    return ap.parse_args()


def make_workload(args, cells=None):
    """Synthetic character matrix of the workload with the implicit root row, in the reference's
    map row order, as compact int32 codes + dense float64 weight table (negative_log priors)."""
    from cassiopeia_b200 import host_prep
    from cassiopeia_b200.solver import solver_utilities
    from cassiopeia_b200.synthetic import simulate_character_matrix

    n = cells or args.cells
    cm, priors, _ = simulate_character_matrix(n, args.chars, args.states, seed=args.seed,
                                              mutation_rate=args.mutation_rate,
                                              missing_fraction=args.missing)
    rooted = np.vstack([cm, np.zeros((1, args.chars), dtype=cm.dtype)]).astype(np.int64)
    order = host_prep.reference_row_order(rooted)
    weights = solver_utilities.transform_priors(priors, "negative_log")
    codes, missing_code, table, n_states = host_prep.compact_states(rooted[order], -1, weights)
    return codes, missing_code, table, n_states


def scan_bytes(n_rows: int, esz: int) -> float:
    """Algorithmic HBM bytes of the loop's scans: every live unordered pair once per iteration."""
    k = np.arange(3, n_rows + 1, dtype=np.float64)
    return float((esz * k * (k - 1) / 2).sum())


def loop_work(n_rows: int) -> float:
    k = np.arange(3, n_rows + 1, dtype=np.float64)
    return float((k * k).sum())


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
    """Full NJ solve of the oracle port on `cells` cells with all host threads; returns
    (map seconds, loop seconds, threads)."""
    from oracle import pyoracle as po

    codes, missing_code, table, _ = make_workload(args, cells)
    cm = codes.astype(np.int64)
    threads = po.max_threads()
    po.whd_map(cm[:64], missing_code, table, threads=threads)  # spin the thread pool up
    t0 = time.perf_counter()
    D = po.whd_map(cm, missing_code, table, threads=threads)
    t1 = time.perf_counter()
    po.nj(D, threads=threads)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1, threads


def scale_cpu(args, t_map, t_loop, cells):
    n_s, n_w = cells + 1, args.cells + 1
    return (t_map * (n_w * (n_w - 1.0)) / (n_s * (n_s - 1.0))
            + t_loop * loop_work(n_w) / loop_work(n_s))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cells = min(args.cpu_sample_cells, args.cells)
    for _ in range(args.warmup):
        cpu_sample(args, min(cells, 500))
    vals = []
    t_begin = time.perf_counter()
    for _ in range(args.steps):
        t_map, t_loop, threads = cpu_sample(args, cells)
        vals.append(scale_cpu(args, t_map, t_loop, cells))
    wall = time.perf_counter() - t_begin
    value = float(np.mean(vals))
    sample = (f"full NJ solve (oracle/oracle.c, float64, OpenMP) at {cells} cells x {args.chars} chars; "
              f"map time scaled by pairs ratio, loop time by sum k^2 ratio to {args.cells} cells")
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1),
        "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args, args.gpus, "cpu"),
        "cpu_baseline": {"value": value, "unit": "s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def workload_config(args, n_gpus, mode):
    return {
        "workload": f"NeighborJoiningSolver(add_root=True), {args.cells} cells x {args.chars} characters, "
                    f"{args.states} states, mutation priors (negative_log), {int(args.missing * 100)}% missing "
                    "(BASELINE.json configs[2])",
        "cells": args.cells, "characters": args.chars, "states": args.states,
        "missing_fraction": args.missing, "mode": mode, "parallelism": f"row-sharded x{n_gpus}" if n_gpus > 1 else "single GPU",
        "l2": "map (n^2 * 4 B = %.1f GB) is far larger than L2; no flush needed" % ((args.cells + 1) ** 2 * 4 / 1e9),
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
        tree = cb.CassiopeiaTree(character_matrix=frame, priors=priors)
        fast = args.mode == "fast"
        solver = cb.NeighborJoiningSolver(add_root=True, fast=fast, implementation="b200_fast" if fast else "b200_strict")
        t0 = time.perf_counter()
        solver.solve(tree)
        seconds = time.perf_counter() - t0
        same = bool(np.array_equal(solver.last_join_list[0], joins))
        return {"value": seconds, "unit": "s", "joins_identical_to_device_path": same,
                "leaves": len(tree.leaves), "call": "NeighborJoiningSolver(add_root=True, fast=%s).solve(CassiopeiaTree)"
                % fast}
    except Exception as e:  # noqa: BLE001 - a reporting extra must not take the bench line down
        return {"error": f"{type(e).__name__}: {e}"}


def run_b200(args):
    import torch

    from cassiopeia_b200 import engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        from cassiopeia_b200 import sharded
        return sharded.bench_entry(args, METRIC, workload_config, scan_bytes, ClockSampler, peaks)

    torch.cuda.set_device(local_rank)
    method = args.map_method or "exact"
    dt = "f32" if args.mode == "fast" else "f64"
    esz = 4 if dt == "f32" else 8
    codes, missing_code, table, n_states = make_workload(args)
    n = codes.shape[0]
    cm_dev = torch.from_numpy(codes).cuda()
    ld = engine.padded_ld(n)
    D = torch.zeros((n, ld), dtype=torch.float32 if dt == "f32" else torch.float64, device="cuda")

    launches = [0]

    scanned = [0]

    def step():
        engine.whd_map(cm_dev, missing_code, table, n_states, dtype=dt, method=method, out=D)
        launches[0] += engine.last_launch_count()
        ev_mid.record()
        joins, last2 = engine.agglomerate(D, n, "nj", args.mode)
        launches[0] += engine.last_launch_count()
        scanned[0] = engine.last_stats["scanned_elements"]
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
    pruned = scanned[0] > 0
    tname = "float" if esz == 4 else "double"
    common = {"peak": peak, "unit": "GB/s", "peak_source": peak_src, "map_ms": map_ms, "loop_ms": loop_ms,
              "map_gpairs_per_s": n * (n - 1) / 2 / (map_ms * 1e-3) / 1e9}
    if not pruned:
        roofline = {"bound": "hbm", "kernel": f"scan_kernel<{tname},NJ>", "achieved": achieved,
                    "frac": achieved / peak, "traffic": None,
                    "algorithmic_bytes_per_launch_mean": b_alg / n_scans,
                    "launch_ms_mean_incl_merge": loop_ms / n_scans, **common}
    else:
        # The timed loop uses the exact pruned scan, which is not an HBM-bound kernel any more (it reads a
        # fraction of a percent of the map).  The roofline object therefore describes the HBM-bound DENSE scan
        # kernel, measured live on a bounded sample (the first iterations at full size, CAS_PRUNE=0); the
        # pruned loop's own figures follow in "pruned_loop".
        iters = min(300, n - 2)
        os.environ["CAS_PRUNE"] = "0"
        try:
            engine.whd_map(cm_dev, missing_code, table, n_states, dtype=dt, method=method, out=D)
            engine.agglomerate(D, n, "nj", args.mode, max_iters=8)  # warm-up
            engine.whd_map(cm_dev, missing_code, table, n_states, dtype=dt, method=method, out=D)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            engine.agglomerate(D, n, "nj", args.mode, max_iters=iters)
            e1.record()
            torch.cuda.synchronize()
        finally:
            os.environ.pop("CAS_PRUNE", None)
        kk = np.arange(n - iters + 1, n + 1, dtype=np.float64)
        d_bytes = float((esz * kk * (kk - 1) / 2).sum())
        d_ms = e0.elapsed_time(e1)
        d_gbs = d_bytes / (d_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": f"scan_kernel<{tname},NJ>", "achieved": d_gbs, "frac": d_gbs / peak,
                    "traffic": None,
                    "sample": f"dense scan (CAS_PRUNE=0), first {iters} iterations at full size, {d_ms:.1f} ms incl. "
                              "the merge launches",
                    "algorithmic_bytes_per_launch_mean": d_bytes / iters, "launch_ms_mean_incl_merge": d_ms / iters,
                    **common,
                    "pruned_loop": {
                        "kernels": [f"prune_rowtest_kernel<{tname},NJ>", f"scan_pruned_kernel<{tname},NJ>",
                                    f"merge_kernel<{tname},NJ>"],
                        "loop_ms": loop_ms, "launch_ms_mean": loop_ms / n_scans,
                        "traffic": scanned[0] * esz / n_scans,
                        "read_fraction_of_algorithmic": scanned[0] * esz / b_alg,
                        "algorithmic_equivalent_gbs": achieved, "algorithmic_equivalent_over_peak": achieved / peak,
                        "note": "exact pruned scan (DESIGN 4.1): bounds on the criterion let it skip almost all of "
                                "SURVEY 8(d)'s algorithmic bytes; traffic = bytes it actually read per iteration "
                                "(kernel counter); identical joins to the dense scan"}}

    out = {
        "metric": METRIC, "value": ms_per_step / 1e3, "unit": "s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": dt, "data": "synthetic",
        "config": workload_config(args, 1, f"{args.mode}/{method}-map"),
        "clocks": clocks, "gpu_launches": int(launches[0]), "roofline": roofline,
    }

    if not args.no_e2e:
        del D
        torch.cuda.empty_cache()
        cm_pin = torch.from_numpy(codes).pin_memory()
        w_pin = torch.from_numpy(table).pin_memory()
        # cas_solve_host reads the pinned buffers directly (ctypes pointers), synchronous call
        t0 = time.perf_counter()
        j_e2e, l_e2e, tm = engine.solve_host(cm_pin.numpy(), missing_code, w_pin.numpy(), n_states, "nj",
                                             args.mode, method)
        e2e_s = time.perf_counter() - t0
        assert np.array_equal(j_e2e, joins), "host-buffer path and device path disagree"
        out["e2e"] = {"value": e2e_s, "unit": "s",
                      "h2d_bytes_per_step": int(codes.nbytes + table.nbytes),
                      "d2h_bytes_per_step": int(j_e2e.nbytes + l_e2e.nbytes),
                      "device_ms": {"h2d": tm[0], "map": tm[1], "loop": tm[2], "d2h": tm[3]},
                      "api": "cas_solve_host (C ABI, pinned host buffers)"}
    if not args.no_e2e and "e2e" in out:
        out["e2e"]["solver_api"] = solver_api_solve(args, joins)
    if not args.no_cpu_baseline:
        cells = min(args.cpu_sample_cells, args.cells)
        t_map, t_loop, threads = cpu_sample(args, cells)
        out["cpu_baseline"] = {
            "value": scale_cpu(args, t_map, t_loop, cells), "unit": "s", "cores": threads, "kind": "port",
            "sample": f"full NJ solve of oracle/oracle.c at {cells} cells x {args.chars} chars "
                      f"(map {t_map:.2f} s, loop {t_loop:.2f} s on {threads} threads), scaled to "
                      f"{args.cells} cells by pair count (map) and sum k^2 (loop)"}
    print(json.dumps(out), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
