"""Strip-partitioned planner, host logic on CPU (no GPU compute): the sweep schedule, the
halo-row plan (`vhp_strip_halo_rows`, a host function of the C-ABI) and the 2-rank
exchange over a gloo process group with a synthetic field standing in for the kernels."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from visibility_heuristic_path_planner_b200.giant import (decode_key, halo_rows, strip_layout,  # noqa: E402
                                                           sweep_schedule)


def test_schedule_and_layout():
    strips = strip_layout(1000, 8)
    assert strips[0] == (0, 125) and strips[-1] == (875, 1000)
    assert sweep_schedule(strips, 0) == [(0, None)] + [(k, k - 1) for k in range(1, 8)]
    order = sweep_schedule(strips, 500)
    assert order[0] == (4, None) and set(k for k, _ in order) == set(range(8))
    pos = {k: n for n, (k, _) in enumerate(order)}
    assert all(nb is None or pos[nb] < pos[k] for k, nb in order)  # neighbour sweeps first
    try:
        strip_layout(100, 4)
        assert False, "strips lower than one tile row must be rejected"
    except ValueError:
        pass
    assert decode_key((1 << 40) | (7 << 20) | 3, 50, 60) == (43, 63)
    assert decode_key((3 << 40) | (7 << 20) | 3, 50, 60) == (57, 57)


def test_halo_rows_geometry():
    import visibility_heuristic_path_planner_b200 as vhp
    lib = vhp.load_library()
    nx, ny = 777, 640
    strips = strip_layout(ny, 6)
    g = np.random.default_rng(0)
    for _ in range(300):
        sx, sy = int(g.integers(0, nx)), int(g.integers(0, ny))
        for k, nb in sweep_schedule(strips, sy):
            rows = halo_rows(lib, nx, ny, sx, sy, *strips[k])
            if nb is None:
                assert rows == [-1, -1, -1, -1]
                continue
            lo, hi = strips[nb]
            assert all(lo <= r < hi for r in rows if r >= 0), (sx, sy, k, rows)
            # the first tile row is 1..32 rows high, every other 32: the halo row is at most
            # 32 rows behind the strip's first row on the source side
            edge = strips[k][0] if k > nb else strips[k][1] - 1
            assert all(abs(r - edge) <= 32 for r in rows if r >= 0)


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from visibility_heuristic_path_planner_b200.giant import StripPlanner
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)

    class FakePlanner(StripPlanner):
        """CPU stand-in: vis[y, x] = 1000 y + x, the 'sweep' only records the halo it got."""
        def __init__(self, nx, ny, nstrips):
            import visibility_heuristic_path_planner_b200 as vhp
            self.torch, self.lib, self.dist = torch, vhp.load_library(), dist
            self.rank, self.world, self.dev = rank, world, torch.device("cpu")
            self.nx, self.ny = nx, ny
            self.strips = strip_layout(ny, nstrips)
            self.owner_of = [k % world for k in range(nstrips)]
            self.mine = [k for k in range(nstrips) if self.owner_of[k] == rank]
            xs = torch.arange(nx, dtype=torch.float64)
            self.f = {k: dict(vis=torch.arange(lo, hi, dtype=torch.float64)[:, None] * 1000 + xs)
                      for k, (lo, hi) in enumerate(self.strips) if k in self.mine}
            self.halo_bytes, self.log = 0, []

        def _sweep_strip(self, k, sx, sy, rows, halo):
            got = [None if r < 0 else float(halo[q][0]) / 1000 for q, r in enumerate(rows)]
            self.log.append((k, rows, got))

    fp = FakePlanner(300, 400, 5)
    ok = True
    for sx, sy in ((10, 5), (150, 200), (299, 399), (31, 81), (64, 160)):
        fp.log.clear()
        fp._sweep(sx, sy)
        for k, rows, got in fp.log:  # every halo row arrived from the right grid row
            ok &= all((r < 0 and g is None) or (r >= 0 and g == r) for r, g in zip(rows, got))
        ok &= sorted(k for k, _, _ in fp.log) == fp.mine
    q.put((rank, bool(ok), fp.halo_bytes))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_halo_exchange():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert sum(b for _, _, b in res) > 0  # rows really crossed ranks
