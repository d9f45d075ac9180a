"""torchrun check of the multi-process NCCL path: sharded joins == single-GPU joins."""
import argparse, os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
import bench
from cassiopeia_b200 import engine, sharded

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=6000)
a = ap.parse_args()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
group = sharded.NcclGroup()
args = argparse.Namespace(cells=a.cells, chars=60, states=100, missing=0.2, mutation_rate=0.1, seed=0)
codes, missing_code, table, n_states = bench.make_workload(args)
n = codes.shape[0]
cm_dev = torch.from_numpy(codes).cuda()
D = engine.whd_map(cm_dev, missing_code, table, n_states, dtype="f32")
single, last_single = engine.agglomerate(D, n, "nj", "fast")
del D
peers = sharded.PeerGroup(n, "f32")
for name, fn, grp in (("nccl", sharded.sharded_solve, group), ("p2p", sharded.sharded_solve_p2p, peers),
                      ("p2p", sharded.sharded_solve_p2p, peers)):
    shard = sharded.build_shard(cm_dev, missing_code, table, n_states, rank, world, "f32")
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    joins, last2 = fn(shard, n, "nj", grp)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    ok = np.array_equal(joins, single) and np.array_equal(last2, last_single)
    print(f"rank {rank}/{world}: n={n} {name} sharded solve {t1 - t0:.3f}s ({(t1 - t0) / (n - 2) * 1e6:.1f} us/iter), "
          f"identical to single GPU: {ok}", flush=True)
    assert ok
peers.close()
group.close()
dist.destroy_process_group()
