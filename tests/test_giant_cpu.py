"""Strip-partitioned planner, host logic on CPU (no GPU compute): the strip layout, the halo-row
plan (`vhp_strip_halo_rows`, a host function of the C-ABI), the handle's argument checks, and a
world-size-2 gloo run of the set-up every rank performs before `vhp_giant_create` (rank 0's NCCL
id reaches the other rank; both derive the same strips and the same halo rows for either side of
their boundary).  The exchange itself is NCCL inside the library and needs GPUs
(tests/test_gpu_giant_multirank.py)."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from visibility_heuristic_path_planner_b200.giant import halo_rows, strip_layout  # noqa: E402


def test_layout():
    strips = strip_layout(1000, 8)
    assert strips[0] == (0, 125) and strips[-1] == (875, 1000)
    assert strip_layout(8192, 8)[3] == (3072, 4096)
    s = strip_layout(1001, 3)
    assert [hi - lo for lo, hi in s] == [334, 334, 333] and s[0][0] == 0 and s[-1][1] == 1001
    try:
        strip_layout(100, 4)
        assert False, "strips lower than one tile row must be rejected"
    except ValueError:
        pass


def test_halo_rows_geometry():
    import visibility_heuristic_path_planner_b200 as vhp
    lib = vhp.load_library()
    nx, ny = 777, 640
    strips = strip_layout(ny, 6)
    g = np.random.default_rng(0)
    for _ in range(300):
        sx, sy = int(g.integers(0, nx)), int(g.integers(0, ny))
        for k, (lo, hi) in enumerate(strips):
            rows = halo_rows(lib, nx, ny, sx, sy, lo, hi)
            if lo <= sy < hi:
                assert rows == [-1, -1, -1, -1]
                continue
            upper = lo > sy
            nlo, nhi = strips[k - 1] if upper else strips[k + 1]
            assert all(nlo <= r < nhi for r in rows if r >= 0), (sx, sy, k, rows)
            # the first tile row is 1..32 rows high, every other 32: the halo row is at most
            # 32 rows behind the strip's first row on the source side
            edge = lo if upper else hi - 1
            assert all(abs(r - edge) <= 32 for r in rows if r >= 0)


def test_create_needs_a_device_and_checks_arguments():
    """No GPU here: the handle cannot be created (no CPU fallback), bad arguments are refused."""
    import ctypes as C
    import visibility_heuristic_path_planner_b200 as vhp
    lib = vhp.load_library()
    h = C.c_void_p()
    occ = np.ones((64, 64), np.uint8)
    st = lib.vhp_giant_create(None, occ.ctypes.data_as(C.c_void_p), 64, 64, 0, 1, None, 1, C.byref(h))
    assert st == -1 and not h.value
    assert lib.vhp_giant_unique_id(None) == -1
    assert lib.vhp_giant_local_rows(None, None, None) == -1


def _worker(rank, world, port, q):
    import torch.distributed as dist
    import visibility_heuristic_path_planner_b200 as vhp
    from visibility_heuristic_path_planner_b200 import giant
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lib = vhp.load_library()
    try:
        ident = giant.unique_id() if rank == 0 else None
        have_nccl = True
    except vhp.VhpError:
        ident, have_nccl = bytes(giant.ID_BYTES) if rank == 0 else None, False
    box = [ident]
    dist.broadcast_object_list(box, src=0)
    nx, ny = 300, 400
    strips = strip_layout(ny, world)
    lo, hi = strips[rank]
    ok = len(box[0]) == giant.ID_BYTES
    # what this rank would receive (consumer side) and what its neighbour would send it
    mine = []
    for sx, sy in ((10, 5), (150, 200), (299, 399), (31, 81), (64, 160), (96, 300)):
        mine.append(halo_rows(lib, nx, ny, sx, sy, lo, hi))
    allrows = [None] * world
    dist.all_gather_object(allrows, (rank, strips, mine, box[0]))
    for r, st, _, ident_r in allrows:
        ok &= st == strips and ident_r == box[0]
    for n, (sx, sy) in enumerate(((10, 5), (150, 200), (299, 399), (31, 81), (64, 160), (96, 300))):
        for r, _, rows_r, _ in allrows:
            rl, rh = strips[r]
            for qd, row in enumerate(rows_r[n]):
                if row < 0:
                    continue
                owner = next(k for k, (a, b) in enumerate(strips) if a <= row < b)
                ok &= abs(owner - r) == 1 and (owner < r) == (qd < 2)  # +y rows come from below
    q.put((rank, bool(ok), have_nccl))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_setup():
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
