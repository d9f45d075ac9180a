#!/usr/bin/env python
"""Phase timeline of the pruned loop's kernels (development aid).

Build the instrumented library first (`make -C cassiopeia_b200/csrc timeline`), then
    CAS_B200_LIBRARY=cassiopeia_b200/libcas_b200_tl.so python scripts/timeline.py [--cells 50000] [--mode exact]
Every CTA's thread 0 stamps clock64 (its SM's cycle counter) and the global timer at the phase boundaries of
prune_rowtest_kernel / scan_pruned_kernel / merge_kernel for a window of iterations; this prints, per kernel and
phase, the median over iterations of (a) the phase duration on the CTA that finishes it last and (b) the global-timer
span from the first CTA's entry to the last CTA's stamp."""
import argparse
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

TL_CTAS, TL_SLOTS = 512, 12


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=50000)
    ap.add_argument("--mode", default="exact")
    ap.add_argument("--t0", type=int, default=5000)
    ap.add_argument("--window", type=int, default=64)
    a = ap.parse_args()
    import torch
    import bench
    from cassiopeia_b200 import _lib, engine

    L = _lib.load()
    args = argparse.Namespace(cells=a.cells, chars=60, states=100, missing=0.2, mutation_rate=0.1, seed=0, algo="nj",
                              priors=True)
    codes, missing_code, table, n_states = bench.make_workload(args)
    n = codes.shape[0]
    dt = "f32" if a.mode == "fast" else "f64"
    buf = torch.zeros(a.window * 3 * TL_CTAS * TL_SLOTS, dtype=torch.int64, device="cuda")
    L.cas_debug_timeline.restype = ctypes.c_int
    L.cas_debug_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    D = engine.whd_map(codes, missing_code, table, n_states, dtype=dt, method="exact")
    assert L.cas_debug_timeline(buf.data_ptr(), a.t0, a.t0 + a.window) == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    engine.agglomerate(D, n, "nj", a.mode, max_iters=a.t0 + a.window)
    e1.record()
    torch.cuda.synchronize()
    tl = buf.cpu().numpy().astype(np.int64).reshape(a.window, 3, TL_CTAS, TL_SLOTS)
    names = ["prune_rowtest", "scan_pruned", "merge"]
    labels = {
        0: ["entry", "after pdl_wait", "end"],
        1: ["entry", "after pdl_wait", "units done", "block_best done", "fence done", "atomic done",
            "last: fence", "last: reduced", "last: published", "last: E rows marked"],
        2: ["entry", "after pdl_wait", "body done", "partials stored", "fence done", "atomic done", "last: entered",
            "last: sums", "last: warm start done"],
    }
    print(f"n={n} mode={a.mode} iterations {a.t0}..{a.t0 + a.window}: "
          f"{1e3 * e0.elapsed_time(e1) / (a.t0 + a.window):.1f} us/iteration over the whole run (instrumented build)")
    # global-timer picture: per iteration, time from rowtest's first entry to each kernel's first entry / last stamp
    g = tl[:, :, :, TL_SLOTS - 1].astype(np.float64)
    g[g == 0] = np.nan
    first_entry = np.nanmin(g, axis=2)  # [window, 3]
    it_span = np.diff(first_entry[:, 0])
    print(f"global timer: iteration period (rowtest entry to next rowtest entry) median {np.nanmedian(it_span) / 1e3:.2f} us")
    for kn in range(3):
        print(f"  {names[kn]}: first CTA entry at +{np.nanmedian(first_entry[:, kn] - first_entry[:, 0]) / 1e3:.2f} us "
              f"after the row test's first entry")
    for kn in range(3):
        print(f"\n{names[kn]} (cycles of the CTA's own SM clock, 1.9 GHz; median over {a.window} iterations)")
        for s in range(1, len(labels[kn])):
            d = tl[:, kn, :, s] - tl[:, kn, :, 0]
            valid = (tl[:, kn, :, s] != 0) & (tl[:, kn, :, 0] != 0)
            d = np.where(valid, d, -1)
            mx = d.max(axis=1).astype(np.float64)
            med = np.array([np.median(d[i][valid[i]]) if valid[i].any() else np.nan for i in range(a.window)])
            nct = valid.sum(axis=1)
            print(f"  {labels[kn][s]:>24s}: slowest CTA {np.median(mx) / 1.9e3:7.2f} us, median CTA "
                  f"{np.nanmedian(med) / 1.9e3:7.2f} us after its own entry   (CTAs stamping: {int(np.median(nct))})")


if __name__ == "__main__":
    main()
