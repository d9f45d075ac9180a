"""Row-sharded NJ / UPGMA over several B200s (one process per GPU, NCCL over NVLink).

The n x n map is dealt to the ranks in 64-row blocks, block-cyclically; every rank builds its
own rows from the replicated character matrix (no exchange) and the loop exchanges, per
iteration, the ranks' arg-min candidates and the three rows (lo, hi, last) with two NCCL
all-gathers issued from the C++ host loop of ``cas_sharded_solve``.  ``torch.distributed`` is
used only to hand the NCCL unique id to the other ranks.
"""
from __future__ import annotations

import ctypes
import json
import os
from typing import Optional

import numpy as np

from . import _lib, engine
from ._lib import CAS_ALGO_NJ, CAS_ALGO_UPGMA, CAS_F32, CAS_F64, CAS_MAP_EXACT

BLOCK = 64  # rows per ownership block (SHB in csrc/sharded.cu)


# ---- partition arithmetic (mirrors sh_owner / sh_local / sh_count; tested on CPU) -------------
def owner_of(slot: int, world: int) -> int:
    return (slot // BLOCK) % world


def local_of(slot: int, world: int) -> int:
    return (slot // BLOCK // world) * BLOCK + slot % BLOCK


def slot_of(local: int, rank: int, world: int) -> int:
    return ((local // BLOCK) * world + rank) * BLOCK + local % BLOCK


def local_rows(k: int, rank: int, world: int) -> int:
    nb, rem = divmod(k, BLOCK)
    rows = (nb // world + (1 if rank < nb % world else 0)) * BLOCK
    if rem > 0 and nb % world == rank:
        rows += rem
    return rows


def local_capacity(n: int, world: int) -> int:
    nblocks = (n + BLOCK - 1) // BLOCK
    return ((nblocks + world - 1) // world) * BLOCK


def shard_rows(n: int, rank: int, world: int) -> np.ndarray:
    """Global slots of this rank's local rows, in local order."""
    return np.array([slot_of(l, rank, world) for l in range(local_rows(n, rank, world))], dtype=np.int64)


# ---- device side ----------------------------------------------------------------------------
def build_shard(cm_dev, missing: int, table, n_states: int, rank: int, world: int, dtype: str = "f32"):
    """This rank's rows of the map, [local_capacity, ld], from the replicated character matrix."""
    torch = engine._torch()
    L = _lib.load()
    n, m = int(cm_dev.shape[0]), int(cm_dev.shape[1])
    ld = engine.padded_ld(n)
    cap = local_capacity(n, world)
    tdtype = torch.float64 if dtype == "f64" else torch.float32
    D = torch.zeros((cap, ld), dtype=tdtype, device="cuda")
    w_t = torch.from_numpy(np.ascontiguousarray(table, dtype=np.float64)).cuda() if table is not None else None
    ws_bytes = int(L.cas_whd_workspace_bytes(n, m, int(n_states), 1 if w_t is not None else 0, CAS_MAP_EXACT))
    ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device="cuda")
    rc = L.cas_whd_map_cyclic(cm_dev.data_ptr(), n, m, int(missing), w_t.data_ptr() if w_t is not None else None,
                              int(n_states), D.data_ptr(), ld, CAS_F64 if dtype == "f64" else CAS_F32,
                              int(rank), int(world), CAS_MAP_EXACT, ws.data_ptr(), ws_bytes,
                              engine._stream_ptr(torch))
    _lib.check(rc, "cas_whd_map_cyclic")
    return D


def emulated_solve(codes: np.ndarray, missing: int, table, n_states: int, algo: str, world: int,
                   dtype: str = "f32", max_iters: int = -1, p2p: bool = False):
    """All `world` virtual ranks in this process on the current GPU (no collective): the same
    kernels as the multi-process path, used to check that sharding does not change the joins."""
    torch = engine._torch()
    L = _lib.load()
    cm_dev = torch.from_numpy(np.ascontiguousarray(codes, dtype=np.int32)).cuda()
    n = int(cm_dev.shape[0])
    ld = engine.padded_ld(n)
    code = CAS_F64 if dtype == "f64" else CAS_F32
    shards = [build_shard(cm_dev, missing, table, n_states, r, world, dtype) for r in range(world)]
    ws_bytes = int(L.cas_sharded_workspace_bytes(n, world, code))
    wss = [torch.empty(ws_bytes, dtype=torch.uint8, device="cuda") for _ in range(world)]
    njoin = max(n - 2, 0)
    joins = torch.full((njoin + 1, 2), -1, dtype=torch.int32, device="cuda")
    last2 = torch.full((2,), -1, dtype=torch.int32, device="cuda")
    dptr = (ctypes.c_void_p * world)(*[s.data_ptr() for s in shards])
    wptr = (ctypes.c_void_p * world)(*[w.data_ptr() for w in wss])
    algo_code = CAS_ALGO_NJ if algo == "nj" else CAS_ALGO_UPGMA
    if p2p:
        # every virtual rank gets its own exchange buffer; "peer" stores are local in this process
        ex_bytes = int(L.cas_exchange_bytes(n, world, code))
        exs = [torch.zeros(ex_bytes, dtype=torch.uint8, device="cuda") for _ in range(world)]
        eptr = (ctypes.c_void_p * world)(*[e.data_ptr() for e in exs])
        rc = L.cas_sharded_solve_emulated_p2p(dptr, n, ld, code, algo_code, int(max_iters), int(world), eptr, 1,
                                              joins.data_ptr(), last2.data_ptr(), wptr, ws_bytes,
                                              engine._stream_ptr(torch))
        _lib.check(rc, "cas_sharded_solve_emulated_p2p")
        for e in exs:
            err = ctypes.c_int32(0)
            _lib.check(L.cas_exchange_error(e.data_ptr(), engine._stream_ptr(torch), ctypes.byref(err)),
                       "cas_exchange_error")
            if err.value:
                raise _lib.B200LibraryError("a P2P flag wait timed out")
    else:
        rc = L.cas_sharded_solve_emulated(dptr, n, ld, code, algo_code, int(max_iters), int(world),
                                          joins.data_ptr(), last2.data_ptr(), wptr, ws_bytes,
                                          engine._stream_ptr(torch))
        _lib.check(rc, "cas_sharded_solve_emulated")
    done = njoin if max_iters < 0 else min(njoin, max_iters)
    return joins[:done].cpu().numpy(), last2.cpu().numpy()


class NcclGroup:
    """NCCL communicator owned by libcas_b200; the unique id travels over torch.distributed."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised (one process per GPU)")
        L = _lib.load()
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        buf = (ctypes.c_char * 128)()
        if self.rank == 0:
            _lib.check(L.cas_nccl_unique_id(buf), "cas_nccl_unique_id")
        device = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device=device)
        dist.broadcast(t, src=0)
        ident = bytes(t.cpu().tolist())
        self.handle = ctypes.c_void_p()
        _lib.check(L.cas_nccl_comm_create(ident, self.rank, self.world, ctypes.byref(self.handle)),
                   "cas_nccl_comm_create")

    def close(self):
        if self.handle:
            _lib.load().cas_nccl_comm_destroy(self.handle)
            self.handle = ctypes.c_void_p()


class PeerGroup:
    """Exchange buffers for the P2P loop: every rank cudaMallocs one, exports a cudaIpc handle,
    the handles travel over torch.distributed and each rank maps its peers' buffers."""

    def __init__(self, n: int, dtype: str = "f32"):
        import torch
        import torch.distributed as dist

        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised (one process per GPU)")
        L = _lib.load()
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.n = int(n)
        self.code = CAS_F64 if dtype == "f64" else CAS_F32
        self.bytes = int(L.cas_exchange_bytes(self.n, self.world, self.code))
        self.own = ctypes.c_void_p()
        handle = (ctypes.c_char * 64)()
        _lib.check(L.cas_exchange_alloc(self.bytes, ctypes.byref(self.own), handle), "cas_exchange_alloc")
        device = "cuda" if dist.get_backend() == "nccl" else "cpu"
        mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device=device)
        gathered = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(gathered, mine)
        self.ptrs = []
        for r in range(self.world):
            if r == self.rank:
                self.ptrs.append(self.own.value)
            else:
                ptr = ctypes.c_void_p()
                _lib.check(L.cas_exchange_open(bytes(gathered[r].cpu().tolist()), ctypes.byref(ptr)),
                           "cas_exchange_open")
                self.ptrs.append(ptr.value)
        self.epoch = 0
        dist.barrier()

    def array(self):
        return (ctypes.c_void_p * self.world)(*self.ptrs)

    def check(self):
        import torch

        err = ctypes.c_int32(0)
        _lib.check(_lib.load().cas_exchange_error(self.own, engine._stream_ptr(torch), ctypes.byref(err)),
                   "cas_exchange_error")
        if err.value:
            raise _lib.B200LibraryError("a P2P flag wait timed out (peer rank stalled or died)")

    def close(self):
        import torch.distributed as dist

        L = _lib.load()
        if dist.is_initialized():
            dist.barrier()
        for r, ptr in enumerate(self.ptrs):
            if r != self.rank and ptr:
                L.cas_exchange_close(ptr)
        if self.own:
            L.cas_exchange_free(self.own)
            self.own = ctypes.c_void_p()
        self.ptrs = []


#: map elements this rank's pruned scans read during its most recent sharded solve (0: dense scan)
last_scanned_elements = 0


def _record_scanned(L, torch, ws, n, world, code):
    import ctypes

    global last_scanned_elements
    out = ctypes.c_int64(0)
    _lib.check(L.cas_sharded_scanned_elements(ws.data_ptr(), int(n), int(world), code, engine._stream_ptr(torch),
                                              ctypes.byref(out)), "cas_sharded_scanned_elements")
    last_scanned_elements = int(out.value)


def sharded_solve_p2p(shard, n: int, algo: str, group: PeerGroup, max_iters: int = -1):
    """P2P variant of ``sharded_solve``: NVLink stores + flags from inside the kernels, no NCCL."""
    torch = engine._torch()
    L = _lib.load()
    ld = int(shard.stride(0))
    code = CAS_F64 if shard.dtype == torch.float64 else CAS_F32
    assert code == group.code and int(n) <= group.n
    ws_bytes = int(L.cas_sharded_workspace_bytes(int(n), group.world, code))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    njoin = max(int(n) - 2, 0)
    joins = torch.full((njoin + 1, 2), -1, dtype=torch.int32, device="cuda")
    last2 = torch.full((2,), -1, dtype=torch.int32, device="cuda")
    group.epoch += 1
    rc = L.cas_sharded_solve_p2p(shard.data_ptr(), int(n), ld, code, CAS_ALGO_NJ if algo == "nj" else CAS_ALGO_UPGMA,
                                 int(max_iters), group.rank, group.world, group.array(), group.epoch,
                                 joins.data_ptr(), last2.data_ptr(), ws.data_ptr(), ws_bytes,
                                 engine._stream_ptr(torch))
    _lib.check(rc, "cas_sharded_solve_p2p")
    _record_scanned(L, torch, ws, n, group.world, code)
    done = njoin if max_iters < 0 else min(njoin, int(max_iters))
    out = joins[:done].cpu().numpy(), last2.cpu().numpy()
    group.check()
    return out


def sharded_solve(shard, n: int, algo: str, group: NcclGroup, max_iters: int = -1):
    """Runs the sharded loop on this rank's shard ([capacity, ld] CUDA tensor, destroyed).
    Every rank returns the same (joins, last2)."""
    torch = engine._torch()
    L = _lib.load()
    ld = int(shard.stride(0))
    code = CAS_F64 if shard.dtype == torch.float64 else CAS_F32
    ws_bytes = int(L.cas_sharded_workspace_bytes(int(n), group.world, code))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    njoin = max(int(n) - 2, 0)
    joins = torch.full((njoin + 1, 2), -1, dtype=torch.int32, device="cuda")
    last2 = torch.full((2,), -1, dtype=torch.int32, device="cuda")
    rc = L.cas_sharded_solve(shard.data_ptr(), int(n), ld, code, CAS_ALGO_NJ if algo == "nj" else CAS_ALGO_UPGMA,
                             int(max_iters), group.rank, group.world, group.handle, joins.data_ptr(),
                             last2.data_ptr(), ws.data_ptr(), ws_bytes, engine._stream_ptr(torch))
    _lib.check(rc, "cas_sharded_solve")
    _record_scanned(L, torch, ws, n, group.world, code)
    done = njoin if max_iters < 0 else min(njoin, int(max_iters))
    return joins[:done].cpu().numpy(), last2.cpu().numpy()


# ---- bench.py entry for --gpus N > 1 ------------------------------------------------------------
def bench_entry(args, metric, workload_config, scan_bytes, ClockSampler, peaks):
    import torch
    import torch.distributed as dist

    import bench

    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dt = "f32"
    codes, missing_code, table, n_states = bench.make_workload(args)
    n = codes.shape[0]
    exchange = os.environ.get("CAS_SHARDED_EXCHANGE", "p2p")
    if exchange == "p2p":
        try:
            group = PeerGroup(n, dt)
            solve_fn = sharded_solve_p2p
        except Exception as e:  # cudaIpc unavailable in this container
            if rank == 0:
                print(f"# P2P exchange unavailable ({e}); using NCCL all-gathers", flush=True)
            exchange = "nccl"
    if exchange != "p2p":
        group = NcclGroup()
        solve_fn = sharded_solve
    cm_dev = torch.from_numpy(codes).cuda()
    launches = [0]

    def step():
        shard = build_shard(cm_dev, missing_code, table, n_states, rank, world, dt)
        launches[0] += engine.last_launch_count()
        mid = torch.cuda.Event(enable_timing=True)
        mid.record()
        joins, last2 = solve_fn(shard, n, "nj", group)
        launches[0] += engine.last_launch_count()
        return joins, last2, mid

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches[0] = 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    mids, ends = [], []
    for _ in range(args.steps):
        joins, last2, mid = step()
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        mids.append(mid)
        ends.append(e)
    ev1.record()
    torch.cuda.synchronize()
    dist.barrier()
    clocks = sampler.stop()
    total_ms = ev0.elapsed_time(ev1)
    loop_ms = float(np.mean([m.elapsed_time(e) for m, e in zip(mids, ends)]))
    t = torch.tensor([total_ms, loop_ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, loop_ms = float(t[0]), float(t[1])
    # every rank must hold the same join list
    h = torch.tensor([int(np.int64(joins.astype(np.int64).sum())), int(joins[-1, 0])], dtype=torch.int64, device="cuda")
    hmax, hmin = h.clone(), h.clone()
    dist.all_reduce(hmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(hmin, op=dist.ReduceOp.MIN)
    assert torch.equal(hmax, hmin), "ranks disagree on the join list"
    # end to end: pinned host character matrix + weights -> device, shard build, loop, join list -> host
    cm_pin = torch.from_numpy(codes).pin_memory()
    torch.cuda.synchronize()
    dist.barrier()
    import time as _time
    t_e2e0 = _time.perf_counter()
    cm_e2e = cm_pin.to("cuda", non_blocking=True)
    shard = build_shard(cm_e2e, missing_code, table, n_states, rank, world, dt)  # uploads the weight table
    j_e2e, l_e2e = solve_fn(shard, n, "nj", group)
    torch.cuda.synchronize()
    e2e_s = torch.tensor([_time.perf_counter() - t_e2e0], dtype=torch.float64, device="cuda")
    dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    assert np.array_equal(j_e2e, joins)
    scanned = torch.tensor([last_scanned_elements], dtype=torch.int64, device="cuda")
    dist.all_reduce(scanned, op=dist.ReduceOp.SUM)
    scanned = int(scanned.item())
    # the HBM-bound kernel of this path is the DENSE sharded scan: measure it live on a bounded sample (first
    # iterations at full size, CAS_PRUNE=0) -- the timed loop above uses the exact pruned scan
    dense_iters = min(300, n - 2)
    os.environ["CAS_PRUNE"] = "0"
    try:
        shard = build_shard(cm_dev, missing_code, table, n_states, rank, world, dt)
        solve_fn(shard, n, "nj", group, max_iters=8)
        shard = build_shard(cm_dev, missing_code, table, n_states, rank, world, dt)
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        dist.barrier()
        d0.record()
        solve_fn(shard, n, "nj", group, max_iters=dense_iters)
        d1.record()
        torch.cuda.synchronize()
    finally:
        os.environ.pop("CAS_PRUNE", None)
    dms = torch.tensor([d0.elapsed_time(d1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(dms, op=dist.ReduceOp.MAX)
    dense_ms = float(dms[0])
    if rank == 0:
        peak, peak_src = peaks()
        ms_per_step = total_ms / args.steps
        b_alg = scan_bytes(n, 4)
        achieved = b_alg / (loop_ms * 1e-3) / 1e9
        exch = (("P2P NVLink stores + tagged flags from inside the scan/cols kernels" if exchange == "p2p"
                 else "2 NCCL all-gathers") + " (24 B x G candidates; 3 rows x k x 4 B)")
        kk = np.arange(n - dense_iters + 1, n + 1, dtype=np.float64)
        d_bytes = float((4 * kk * (kk - 1) / 2).sum())
        d_gbs = d_bytes / (dense_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "sh_scan_kernel<float,NJ>", "achieved": d_gbs, "peak": peak * world,
                    "unit": "GB/s", "frac": d_gbs / (peak * world), "traffic": None,
                    "peak_source": peak_src + f" x {world} GPUs", "loop_ms": loop_ms,
                    "sample": f"dense sharded scan (CAS_PRUNE=0), first {dense_iters} iterations at full size, "
                              f"{dense_ms:.1f} ms incl. the exchange and merge launches",
                    "exchange_per_iteration": exch}
        if scanned:
            roofline["pruned_loop"] = {
                "kernels": ["sh_rowtest_kernel", "sh_scan_pruned_kernel", "sh_cols_kernel", "sh_merge_kernel"],
                "loop_ms": loop_ms, "launch_ms_mean": loop_ms / max(n - 2, 1),
                "traffic": scanned * 4 / max(n - 2, 1), "read_fraction_of_algorithmic": scanned * 4 / b_alg,
                "algorithmic_equivalent_gbs": achieved, "algorithmic_equivalent_over_peak": achieved / (peak * world),
                "note": "exact pruned scan (DESIGN 4.1): the timed loop; traffic = bytes all ranks actually read "
                        "per iteration (kernel counters)"}
        else:
            roofline.update({"achieved": achieved, "frac": achieved / (peak * world)})
        out = {
            "metric": metric, "value": ms_per_step / 1e3, "unit": "s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": dt, "data": "synthetic",
            "config": workload_config(args, world, "fast/exact-map"),
            "clocks": clocks, "gpu_launches": int(launches[0]),
            "roofline": roofline,
            "e2e": {"value": float(e2e_s[0]), "unit": "s",
                    "h2d_bytes_per_step": int(codes.nbytes + table.nbytes) * world,
                    "d2h_bytes_per_step": int(joins.nbytes + last2.nbytes) * world,
                    "note": "per rank: pinned host character matrix + weight table -> device, shard build, "
                            "sharded loop, join list -> host (wall clock, max over ranks)"},
        }
        print(json.dumps(out), flush=True)
    group.close()
    dist.destroy_process_group()
