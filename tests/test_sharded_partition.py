"""CPU tests of the block-cyclic partition arithmetic shared by cassiopeia_b200/sharded.py and
csrc/sharded.cu (sh_owner / sh_local / sh_slot / sh_count / sh_chunk)."""
import ctypes

import numpy as np
import pytest

from cassiopeia_b200 import _lib, sharded


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_partition_is_a_bijection(world):
    n = 1000
    seen = {}
    for s in range(n):
        r, l = sharded.owner_of(s, world), sharded.local_of(s, world)
        assert sharded.slot_of(l, r, world) == s
        assert (r, l) not in seen
        seen[(r, l)] = s
    for r in range(world):
        rows = sharded.shard_rows(n, r, world)
        assert np.all(np.diff(rows) > 0)  # local order is slot order
        assert len(rows) == sharded.local_rows(n, r, world) <= sharded.local_capacity(n, world)
    assert sum(sharded.local_rows(n, r, world) for r in range(world)) == n


@pytest.mark.parametrize("world", [1, 2, 5, 8])
def test_live_rows_are_a_prefix_of_the_local_rows(world):
    # swap compaction keeps live slots = [0, k): on every rank they must be the first local rows
    for k in (1, 63, 64, 65, 129, 500, 777):
        for r in range(world):
            cnt = sharded.local_rows(k, r, world)
            live = [l for l in range(sharded.local_capacity(800, world)) if sharded.slot_of(l, r, world) < k]
            assert live == list(range(cnt))


def test_python_arithmetic_matches_the_library():
    L = _lib.load()
    for world in (1, 2, 3, 8):
        for n in (2, 64, 65, 1000, 50001):
            assert L.cas_sharded_local_capacity(n, world) == sharded.local_capacity(n, world)
            for r in range(world):
                assert L.cas_sharded_local_rows(n, r, world) == sharded.local_rows(n, r, world)
    assert L.cas_sharded_workspace_bytes(1000, 4, 0) > 0
