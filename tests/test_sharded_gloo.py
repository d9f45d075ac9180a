"""world_size-2 CPU tests (gloo) of the N>1 path: the exchange protocol of the row-sharded loop, with the
product's block-cyclic partition arithmetic, reproduces the oracle / the single-rank result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def _worker(rank, world, port, D, algo, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sharded_model

    joins = sharded_model.run(D, algo, rank, world)
    np.save(os.path.join(out_dir, f"joins_{rank}.npy"), joins)
    dist.barrier()
    dist.destroy_process_group()


def _run_world(D, algo, world, tmp_path, port):
    if world == 1:
        import sharded_model

        return [sharded_model.run(D, algo, 0, 1)]
    mp.spawn(_worker, args=(world, port, D, algo, str(tmp_path)), nprocs=world, join=True)
    return [np.load(os.path.join(str(tmp_path), f"joins_{r}.npy")) for r in range(world)]


def _map(n, seed):
    from oracle import pyoracle as po

    rng = np.random.default_rng(seed)
    cm = rng.integers(-1, 4, size=(n, 12)).astype(np.int64)
    return po.whd_map(cm, -1, None)


def test_gloo_two_ranks_upgma_reproduces_the_oracle(tmp_path):
    from oracle import pyoracle as po

    D = _map(150, 5)  # three 64-row blocks: both ranks own live rows until the end
    want, _ = po.upgma(D)
    got = _run_world(D, "upgma", 2, tmp_path, 29631)
    assert np.array_equal(got[0], got[1])
    assert np.array_equal(got[0], want.astype(np.int64))


def test_gloo_two_ranks_nj_equals_single_rank(tmp_path):
    D = _map(140, 9)
    one = _run_world(D, "nj", 1, tmp_path, 0)[0]
    two = _run_world(D, "nj", 2, tmp_path, 29632)
    assert np.array_equal(two[0], two[1])
    assert np.array_equal(two[0], one)
    used = set(one.ravel().tolist())
    assert len(used) == one.size and max(used) < 2 * 140 - 2


def test_unique_id_style_byte_broadcast_over_gloo(tmp_path):
    """The NCCL id / cudaIpc handles travel as uint8 tensors over torch.distributed."""
    mp.spawn(_bytes_worker, args=(2, 29633), nprocs=2, join=True)


def _bytes_worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    payload = bytes(range(128)) if rank == 0 else bytes(128)
    t = torch.tensor(list(payload), dtype=torch.uint8)
    dist.broadcast(t, src=0)
    assert bytes(t.tolist()) == bytes(range(128))
    mine = torch.tensor([rank] * 64, dtype=torch.uint8)
    got = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(got, mine)
    assert [int(g[0]) for g in got] == [0, 1]
    dist.destroy_process_group()
