"""Multi-GPU host logic (N > 1 path) on CPU: world_size-2 gloo process group.
The per-rank compute is stood in for by the CPU oracle (test infrastructure); what is
checked is the partition, the map re-indexing and the gather order."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from visibility_heuristic_path_planner_b200.sharding import gather_blocks, shard_batch, shard_bounds  # noqa: E402


def test_shard_bounds_cover_exactly_once():
    for n in (0, 1, 2, 7, 8, 9, 4096, 16384 * 16 + 3):
        for world in (1, 2, 3, 4, 8):
            b = [shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 4, 4)


def test_shard_batch_reindexes_maps():
    g = np.random.default_rng(0)
    maps = g.integers(0, 2, (9, 6, 5)).astype(np.uint8)
    items = g.integers(0, 5, (40, 2)).astype(np.int32)
    imap = np.sort(g.integers(0, 9, 40)).astype(np.int32)
    seen = 0
    for r in range(4):
        lm, li, lim, (lo, hi) = shard_batch(maps, items, imap, r, 4)
        assert np.array_equal(li, items[lo:hi])
        assert all(np.array_equal(lm[lim[k]], maps[imap[lo + k]]) for k in range(hi - lo))
        assert len(lm) == len(np.unique(imap[lo:hi]))
        seen += hi - lo
    assert seen == 40
    lm, li, lim, _ = shard_batch(maps, items, None, 1, 2)
    assert lim is None and len(lm) == 1 and len(li) == 20


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from conftest import rect_map
    from oracle_py import Oracle
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    ora = Oracle()
    nx, ny = 40, 28
    maps = np.stack([rect_map(nx, ny, 6, 100 + m, 2, 7) for m in range(5)]).astype(np.uint8)
    g = np.random.default_rng(3)
    n = 23
    src = np.stack([g.integers(0, nx, n), g.integers(0, ny, n)], axis=1).astype(np.int32)
    smap = np.sort(g.integers(0, 5, n)).astype(np.int32)
    lm, ls, lmap, (lo, hi) = shard_batch(maps, src, smap, rank, world)
    local = np.stack([ora.compute_visibility(lm[lmap[k]].astype(np.float64), *ls[k]) for k in range(hi - lo)])
    full = gather_blocks(local, n, dist)
    if rank == 0:
        ref = np.stack([ora.compute_visibility(maps[smap[k]].astype(np.float64), *src[k]) for k in range(n)])
        q.put(bool(np.array_equal(full, ref)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shard_and_gather():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
