"""Planner on ONE giant map, row-strip partitioned over the GPUs of a box (SURVEY 8e,
BASELINE configs[4]): the Python view of the C-ABI `vhp_giant_*` (include/vhp.h).

Everything runs in the library (csrc/giant.cu): strip sweeps as two chains (+y quadrants
upwards, -y quadrants downwards), halo rows by ncclSend / ncclRecv, the arg-min exchange by
ncclAllGather, the loop of solve() and reconstructPath on the device.  This module only creates
the handle -- with torch.distributed it broadcasts rank 0's 128-byte NCCL id -- and wraps the
outputs in numpy arrays.  There is no Python orchestration of the loop and no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import F64, Context, PlannerOut, VhpError, _np_ptr, load_library
from .sharding import shard_bounds

ID_BYTES = 128


class GiantStats(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("loop_mode", C.c_int32), ("solve_ms", C.c_double),
                ("loop_ms", C.c_double), ("nccl_ms", C.c_double), ("nccl_ops", C.c_int64),
                ("nccl_ops_timed", C.c_int64), ("halo_bytes_sent", C.c_int64), ("launches", C.c_int64),
                ("peer_handover", C.c_int32), ("reserved", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def strip_layout(ny: int, nstrips: int):
    """Row ranges of the strips (contiguous blocks, sizes differ by at most one row): the same
    rule as the library (`vhp_giant_strip_bounds`)."""
    strips = [shard_bounds(ny, k, nstrips) for k in range(nstrips)]
    if nstrips > 1 and min(hi - lo for lo, hi in strips) < 32:
        raise ValueError("strips must be at least 32 rows high (one tile row)")
    return strips


def halo_rows(lib, nx, ny, sx, sy, y0, y1):
    """Grid rows strip [y0, y1) needs per quadrant Q1..Q4 for a sweep from (sx, sy) (-1: none)."""
    rows = (C.c_int32 * 4)()
    lib.vhp_strip_halo_rows(nx, ny, sx, sy, y0, y1, C.byref(rows))
    return [int(r) for r in rows]


def unique_id() -> bytes:
    """Rank 0: the 128-byte id every rank passes to GiantPlanner (vhp_giant_unique_id)."""
    lib = load_library()
    buf = (C.c_uint8 * ID_BYTES)()
    st = lib.vhp_giant_unique_id(buf)
    if st != 0:
        raise VhpError(st, lib.vhp_last_error(None).decode())
    return bytes(buf)


class GiantPlanner:
    """One handle per rank (vhp_giant).  `solve` is collective: every rank calls it with the
    same arguments."""

    def __init__(self, occ, device=0, rank=0, world=1, nccl_id=None, strips_per_rank=1, ctx=None):
        self.lib = load_library()
        self.ctx = ctx or Context(device)
        self._own_ctx = ctx is None
        occ = np.ascontiguousarray(np.asarray(occ) != 0, dtype=np.uint8)
        self.ny, self.nx = occ.shape
        self.rank, self.world = rank, world
        h = C.c_void_p()
        idbuf = None
        if world > 1:
            if nccl_id is None or len(nccl_id) != ID_BYTES:
                raise ValueError("world > 1 needs the 128-byte id of rank 0 (giant.unique_id())")
            idbuf = (C.c_uint8 * ID_BYTES).from_buffer_copy(nccl_id)
        st = self.lib.vhp_giant_create(self.ctx.h, _np_ptr(occ), self.nx, self.ny, rank, world, idbuf,
                                       int(strips_per_rank), C.byref(h))
        if st != 0:
            raise VhpError(st, self.lib.vhp_last_error(self.ctx.h).decode())
        self.h = h
        y0, y1 = C.c_int(0), C.c_int(0)
        self.lib.vhp_giant_local_rows(self.h, C.byref(y0), C.byref(y1))
        self.rows = (y0.value, y1.value)

    @classmethod
    def from_torch_dist(cls, occ, device, dist, strips_per_rank=1):
        """Handle for this rank of an initialised torch.distributed group (any backend): rank 0's
        NCCL id is broadcast through the group."""
        rank, world = dist.get_rank(), dist.get_world_size()
        box = [unique_id() if rank == 0 and world > 1 else None]
        if world > 1:
            dist.broadcast_object_list(box, src=0)
        return cls(occ, device=device, rank=rank, world=world, nccl_id=box[0], strips_per_rank=strips_per_rank)

    def close(self):
        if getattr(self, "h", None):
            self.lib.vhp_giant_destroy(self.h)
            self.h = None
        if self._own_ctx and self.ctx is not None:
            self.ctx.close()
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_loop_mode(self, mode: int, batch: int = 0):
        st = self.lib.vhp_giant_set_loop_mode(self.h, int(mode), int(batch))
        if st != 0:
            raise VhpError(st, self.lib.vhp_last_error(self.ctx.h).decode())

    def solve(self, start, end, threshold, max_iter, fields=True):
        """solve() + reconstructPath() of one problem.  Returns dict(status, nb_of_sources,
        light_sources, path, path_length, stats[, vg, came, vis = this rank's rows, rows])."""
        cap = int(max_iter) + 2
        se = np.array([start[0], start[1], end[0], end[1]], dtype=np.int32)
        r = dict(status=np.zeros(1, np.int32), nb_sources=np.zeros(1, np.int32),
                 light_sources=np.zeros((cap, 2), np.int32), path_len=np.zeros(1),
                 path_n=np.zeros(1, np.int32), path=np.zeros((cap, 2), np.int32))
        if fields:
            nrows = self.rows[1] - self.rows[0]
            r.update(vg=np.zeros((nrows, self.nx)), came=np.zeros((nrows, self.nx), np.int32),
                     vis=np.zeros((nrows, self.nx)))
        po = PlannerOut(*[_np_ptr(r.get(k)) for k in ("status", "nb_sources", "light_sources", "path_len",
                                                      "path_n", "path", "vg", "came", "vis")])
        stats = GiantStats()
        st = self.lib.vhp_giant_solve(self.h, _np_ptr(se), float(threshold), int(max_iter), cap, F64,
                                      C.byref(po), C.byref(stats))
        if st != 0:
            raise VhpError(st, self.lib.vhp_last_error(self.ctx.h).decode())
        nb, status = int(r["nb_sources"][0]), int(r["status"][0])
        out = dict(status=status, nb_of_sources=nb, light_sources=r["light_sources"][: nb + 1].copy(),
                   path=r["path"][: int(r["path_n"][0])].copy(), path_length=float(r["path_len"][0]),
                   stats=stats.as_dict(), rows=self.rows)
        if fields:
            out.update(vg=r["vg"], came=r["came"], vis=r["vis"])
        return out
