"""Time of one rank's block-cyclic shard of the map (cas_whd_map_cyclic) on one GPU: tcgen05 kernel in shard mode vs the
exact CUDA-core kernel.  Usage: shard_map_time.py [--cells N] [--world G] [--priors]"""
import argparse, sys
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from cassiopeia_b200 import sharded

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=100000)
ap.add_argument("--world", type=int, default=8)
ap.add_argument("--priors", action="store_true")
ap.add_argument("--dtype", default="f64")
a = ap.parse_args()
args = argparse.Namespace(cells=a.cells, chars=40, states=50, missing=0.1, mutation_rate=0.1, seed=0, algo="nj", priors=a.priors)
codes, missing_code, table, n_states = bench.make_workload(args)
cm = torch.from_numpy(codes).cuda()
n = codes.shape[0]
for method in ("tensor", "exact", "tensor", "exact"):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    D = sharded.build_shard(cm, missing_code, table, n_states, 1, a.world, a.dtype, method)
    e1.record(); torch.cuda.synchronize()
    rows = len(sharded.shard_rows(n, 1, a.world))
    ms = e0.elapsed_time(e1)
    print(f"{method}: shard of {rows} rows x {n} columns ({a.dtype}, {'priors' if a.priors else 'plain'}): {ms:.2f} ms, "
          f"{rows * n / ms / 1e6:.1f} G ordered pairs/s", flush=True)
    del D
