"""BASELINE config 4's workload (NJ, 200k cells x 40 characters) on ONE B200 in fast mode: the float32 map is 160 GB and
just fits 180 GB of HBM (the reference-exact mode needs the float64 map, 320 GB, hence >= 4 GPUs).  Answers the round-1
verdict's open question "a single B200 would probably solve 200k faster than 8 do -- nobody measured it"."""
import argparse, json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from cassiopeia_b200 import engine

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=200000)
a = ap.parse_args()
args = argparse.Namespace(cells=a.cells, chars=40, states=50, missing=0.1, mutation_rate=0.1, seed=0, algo="nj", priors=True)
codes, missing_code, table, n_states = bench.make_workload(args)
n = codes.shape[0]
free_b, total_b = torch.cuda.mem_get_info()
out = {"config": f"NJ {n} rows x 40 x 50, priors, 10% missing, fast mode (float32 map) on one B200",
       "free_bytes_before": int(free_b), "total_bytes": int(total_b), "map_bytes": int(n) * engine.padded_ld(n) * 4}
try:
    cm = torch.from_numpy(codes).cuda()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    D = engine.whd_map(cm, missing_code, table, n_states, dtype="f32", method="exact")
    torch.cuda.synchronize(); t1 = time.perf_counter()
    joins, last2 = engine.agglomerate(D, n, "nj", "fast")
    torch.cuda.synchronize(); t2 = time.perf_counter()
    st = dict(engine.last_stats)
    used = np.concatenate([joins.ravel(), last2])
    out.update({"map_seconds": t1 - t0, "loop_seconds": t2 - t1, "us_per_iteration": 1e6 * (t2 - t1) / (n - 2),
                "join_list_is_a_valid_tree": bool(np.array_equal(np.sort(used), np.arange(2 * n - 2))),
                "join_hash": bench.join_hash(joins, last2), "near_ties": int(st.get("near_ties", 0)),
                "scanned_elements": int(st.get("scanned_elements", 0))})
except Exception as ex:  # out of memory: report, do not crash the box
    out["error"] = str(ex)[:300]
print(json.dumps(out), flush=True)
