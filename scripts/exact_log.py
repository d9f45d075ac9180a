"""Where do the exact mode's ambiguous iterations occur and how many rows do they involve?"""
import argparse, sys
import numpy as np
sys.path.insert(0, ".")
import bench
from cassiopeia_b200 import engine
ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=100000)
ap.add_argument("--chars", type=int, default=40)
ap.add_argument("--states", type=int, default=50)
ap.add_argument("--missing", type=float, default=0.1)
a = ap.parse_args()
args = argparse.Namespace(cells=a.cells, chars=a.chars, states=a.states, missing=a.missing, mutation_rate=0.1, seed=0,
                          algo="nj", priors=True)
codes, missing_code, table, n_states = bench.make_workload(args)
n = codes.shape[0]
D = engine.whd_map(codes, missing_code, table, n_states, dtype="f64", method="exact")
engine.agglomerate(D, n, "nj", "exact")
st = engine.last_stats
log = st["ambiguous_log"]
W = st["ambiguous_window"]
print({k: v for k, v in st.items() if not k.startswith("ambiguous_")})
edges = [0, 0.1, 0.25, 0.5, 0.75, 0.9, 0.97, 0.99, 1.0]
for lo, hi in zip(edges[:-1], edges[1:]):
    sel = log[(log[:, 0] >= lo * n) & (log[:, 0] < hi * n)]
    if len(sel):
        w = W[(log[:, 0] >= lo * n) & (log[:, 0] < hi * n)]
        print(f"iterations {lo:.2f}n..{hi:.2f}n: {len(sel)} ambiguous, rows mean {sel[:, 2].mean():.1f} max {sel[:, 2].max()}, "
              f"window W median {np.median(w):.3g} max {w.max():.3g}")
big = log[np.argsort(-log[:, 2])[:12]]
print("largest:", big.tolist())
