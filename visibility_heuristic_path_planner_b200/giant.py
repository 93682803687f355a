"""Planner on ONE giant map, row-strip partitioned over the GPUs of a box (SURVEY 8e,
BASELINE configs[4]).

Every rank keeps the whole occupancy map (its bit planes are a few MB) and one strip
[y0, y1) of the fp64 working fields (visibility, global visibility, cached heuristic,
parents: 28 bytes per cell).  One planner iteration is

  1. sweep: the strip that holds the light source sweeps first; the strips above and below
     follow in order of distance.  Before a strip starts it receives, per quadrant, one
     fp64 visibility row of its neighbour on the source side (NCCL send/recv over NVLink,
     2 x nx doubles per strip boundary) -- `vhp_strip_halo_rows` tells both sides which rows;
  2. epilogue + arg-min on every strip at once (`vhp_strip_epilogue_dev`);
  3. all_gather of (h bits, push-order key, vg(end) bits) per rank, 24 bytes each; every rank
     takes the lexicographic minimum, which reproduces priority_queue::top() of the
     reference (first-pushed element among equal h).

The loop control of solve() (src/visibilityBasedSolver.cpp:127-140) runs on the host of
every rank identically.  A process may also own several strips (world size 1 owning all
of them is how the single-GPU test exercises the window kernels and the exchange plan).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import F64, VhpError, load_library
from .sharding import shard_bounds

NO_PARENT = -1
U64_MAX = (1 << 64) - 1
H_INF_BITS = 0x7FF0000000000000


def strip_layout(ny: int, nstrips: int):
    """Row ranges of the strips (contiguous blocks, sizes differ by at most one row)."""
    strips = [shard_bounds(ny, k, nstrips) for k in range(nstrips)]
    if nstrips > 1 and min(hi - lo for lo, hi in strips) < 32:
        raise ValueError("strips must be at least 32 rows high (one tile row)")
    return strips


def halo_rows(lib, nx, ny, sx, sy, y0, y1):
    rows = (C.c_int32 * 4)()
    lib.vhp_strip_halo_rows(nx, ny, sx, sy, y0, y1, C.byref(rows))
    return [int(r) for r in rows]


def sweep_schedule(strips, sy):
    """Order in which the strips sweep for a source in row sy: [(strip, source-side
    neighbour or None)], the source strip first, then outwards (both directions interleaved)."""
    owner = next(k for k, (lo, hi) in enumerate(strips) if lo <= sy < hi)
    order = [(owner, None)]
    for d in range(1, len(strips)):
        if owner + d < len(strips):
            order.append((owner + d, owner + d - 1))
        if owner - d >= 0:
            order.append((owner - d, owner - d + 1))
    return order


def decode_key(key: int, sx: int, sy: int):
    """Push-order key (quadrant << 40 | i << 20 | j) relative to the source -> cell."""
    qd, i, j = key >> 40, (key >> 20) & 0xFFFFF, key & 0xFFFFF
    return (sx + i if qd in (0, 3) else sx - i), (sy + j if qd < 2 else sy - j)


class StripPlanner:
    """solve() + reconstructPath() of one problem on a strip-partitioned map.

    occ      uint8 (ny, nx) whole map, 1 = free (host array; uploaded once)
    nstrips  number of strips; strip k is owned by rank k % world (world = 1: all local)
    device   CUDA device index of this rank
    dist     torch.distributed (initialised) or None
    grid_sweep  None: strips of >= 2^20 cells are swept by many CTAs at once, smaller ones by
             one CTA (the library default); 0 / 2 force the single-CTA / the grid kernels

    All device work (the library's kernels, torch copies / fills, NCCL send/recv) is
    enqueued on one torch stream owned by the planner, so it is ordered without host syncs.
    """

    def __init__(self, occ, nstrips, device=0, dist=None, grid_sweep=None):
        import torch
        from . import torch_context
        self.torch, self.lib, self.dist = torch, load_library(), dist
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1
        self.dev = torch.device("cuda", device)
        self.stream = torch.cuda.Stream(self.dev)
        self.ctx = torch_context(device, self.stream)
        if grid_sweep is not None:
            self._check(self.lib.vhp_context_set_grid_sweep(self.ctx.h, int(grid_sweep)))
        with torch.cuda.stream(self.stream):
            self._alloc(occ, nstrips)

    def close(self):
        self.stream.synchronize()
        self.ctx.close()

    def _alloc(self, occ, nstrips):
        torch = self.torch
        occ = np.ascontiguousarray(occ, dtype=np.uint8)
        self.occ_host = occ
        self.ny, self.nx = occ.shape
        self.strips = strip_layout(self.ny, nstrips)
        self.owner_of = [k % self.world for k in range(nstrips)]
        self.mine = [k for k in range(nstrips) if self.owner_of[k] == self.rank]
        self.occ = torch.from_numpy(occ).to(self.dev)
        # the map never changes: pack its bit planes once instead of before every sweep
        self._check(self.lib.vhp_prepare_maps_dev(self.ctx.h, self.occ.data_ptr(), 1, self.nx, self.ny))
        f = {}
        for k in self.mine:
            rows = self.strips[k][1] - self.strips[k][0]
            f[k] = dict(vis=torch.zeros((rows, self.nx), dtype=torch.float64, device=self.dev),
                        vg=torch.zeros((rows, self.nx), dtype=torch.float64, device=self.dev),
                        hc=torch.full((rows, self.nx), float("inf"), dtype=torch.float64, device=self.dev),
                        came=torch.full((rows, self.nx), NO_PARENT, dtype=torch.int32, device=self.dev),
                        best=torch.zeros(2, dtype=torch.int64, device=self.dev))
        self.f = f
        self.halo_bytes = 0

    # ---- helpers ---------------------------------------------------------------
    def _check(self, st):
        if st != 0:
            raise VhpError(st, self.lib.vhp_last_error(self.ctx.h).decode())

    def _strip_of_row(self, y):
        return next(k for k, (lo, hi) in enumerate(self.strips) if lo <= y < hi)

    def _cell(self, name, x, y):
        """Value of a field at (x, y) if this rank owns the row, else None."""
        k = self._strip_of_row(y)
        if k not in self.f:
            return None
        return self.f[k][name][y - self.strips[k][0], x]

    # ---- one sweep over all strips ------------------------------------------------
    def sweep(self, sx, sy):
        with self.torch.cuda.stream(self.stream):
            self._sweep(sx, sy)

    def _sweep(self, sx, sy):
        for k, nb in sweep_schedule(self.strips, sy):
            rows = halo_rows(self.lib, self.nx, self.ny, sx, sy, *self.strips[k])
            halo = self._fetch_halo(k, nb, rows)
            if k in self.f:
                self._sweep_strip(k, sx, sy, rows, halo)

    def _fetch_halo(self, k, nb, rows):
        """The (4, nx) fp64 halo rows strip k needs from its source-side neighbour nb: a local
        gather when this rank owns both, else NCCL send (owner of nb) / recv (owner of k).
        Both sides derive `rows` from the same geometry, so no negotiation is needed."""
        torch = self.torch
        if nb is None or all(r < 0 for r in rows):
            return None
        lo, hi = self.strips[nb]
        assert all(lo <= r < hi for r in rows if r >= 0), "halo row outside the neighbouring strip"
        src_rank, dst_rank = self.owner_of[nb], self.owner_of[k]
        halo = None
        if src_rank == self.rank:
            idx = torch.tensor([r - lo if r >= 0 else 0 for r in rows], device=self.dev)
            halo = self.f[nb]["vis"].index_select(0, idx).contiguous()
        if src_rank != dst_rank:
            op = None
            if self.rank == src_rank:
                op = self.dist.P2POp(self.dist.isend, halo, dst_rank)
                self.halo_bytes += halo.numel() * 8
            elif self.rank == dst_rank:
                halo = torch.empty((4, self.nx), dtype=torch.float64, device=self.dev)
                op = self.dist.P2POp(self.dist.irecv, halo, src_rank)
            if op is not None:
                for req in self.dist.batch_isend_irecv([op]):
                    req.wait()
        return halo

    def _sweep_strip(self, k, sx, sy, rows, halo):
        y0, y1 = self.strips[k]
        ptrs = (C.c_void_p * 4)(*[C.c_void_p(halo[q].data_ptr()) if (halo is not None and rows[q] >= 0)
                                  else None for q in range(4)])
        self._check(self.lib.vhp_strip_sweep_dev(self.ctx.h, self.occ.data_ptr(), self.nx, self.ny,
                                                 sx, sy, y0, y1, C.byref(ptrs), F64,
                                                 self.f[k]["vis"].data_ptr()))
        self._last_halo = halo  # keep it alive until the kernel that reads it has been enqueued

    # ---- epilogue + arg-min ----------------------------------------------------------
    def epilogue(self, sx, sy, ex, ey, thr, nb, ls_dev):
        with self.torch.cuda.stream(self.stream):
            return self._epilogue(sx, sy, ex, ey, thr, nb, ls_dev)

    def _epilogue(self, sx, sy, ex, ey, thr, nb, ls_dev):
        torch = self.torch
        for k in self.mine:
            y0, y1 = self.strips[k]
            f = self.f[k]
            self._check(self.lib.vhp_strip_epilogue_dev(
                self.ctx.h, self.nx, self.ny, y0, y1, sx, sy, ex, ey, float(thr), nb, ls_dev.data_ptr(),
                f["vis"].data_ptr(), f["vg"].data_ptr(), f["hc"].data_ptr(), f["came"].data_ptr(),
                f["best"].data_ptr()))
        vge = self._cell("vg", ex, ey)
        trip = torch.zeros(3, dtype=torch.int64, device=self.dev)
        if self.mine:
            bests = torch.stack([self.f[k]["best"] for k in self.mine]).cpu().numpy().view(np.uint64)
            h, key = min((int(b[0]), int(b[1])) for b in bests)
        else:
            h, key = U64_MAX, U64_MAX
        vbits = -1 if vge is None else int(np.float64(vge.item()).view(np.int64))
        mine = np.array([h, key], dtype=np.uint64).view(np.int64)
        trip = torch.tensor([int(mine[0]), int(mine[1]), vbits], dtype=torch.int64, device=self.dev)
        if self.dist is not None and self.world > 1:
            allt = [torch.empty_like(trip) for _ in range(self.world)]
            self.dist.all_gather(allt, trip)
            allt = torch.stack(allt).cpu().numpy()
        else:
            allt = trip.cpu().numpy()[None, :]
        keys = allt[:, :2].copy().view(np.uint64)
        h, key = min((int(a), int(b)) for a, b in keys)
        vg_end = max(float(np.int64(v).view(np.float64)) for v in allt[:, 2] if v != -1)
        return h, key, vg_end

    # ---- solve() ----------------------------------------------------------------------
    def solve(self, start, end, threshold, max_iter):
        """Returns dict(status, nb_of_sources, light_sources, path, path_length); the fields stay
        on the ranks (self.f[k]) and `gather_field` collects them."""
        with self.torch.cuda.stream(self.stream):
            return self._solve(start, end, threshold, max_iter)

    def _solve(self, start, end, threshold, max_iter):
        torch = self.torch
        (stx, sty), (ex, ey), thr = map(int, start), map(int, end), float(threshold)
        ex, ey = int(ex), int(ey)
        nx, ny = self.nx, self.ny
        for f in self.f.values():  # reset(), :42-60
            f["vis"].zero_(); f["vg"].zero_(); f["hc"].fill_(float("inf")); f["came"].fill_(NO_PARENT)
        occ = self.occ_host
        status = 0  # checks in the reference's order, :89-116
        if not (0 <= stx < nx and 0 <= sty < ny): status = 1
        elif not (0 <= ex < nx and 0 <= ey < ny): status = 2
        elif occ[sty, stx] == 0: status = 3
        elif occ[ey, ex] == 0: status = 4
        ls = np.zeros((max_iter + 2, 2), dtype=np.int32)
        nb = 0
        if status == 0:
            ls[0] = (stx, sty)
            k = self._strip_of_row(sty)
            if k in self.f:
                self.f[k]["came"][sty - self.strips[k][0], stx] = 0  # :122
            sx, sy = stx, sty
            ls_dev = torch.from_numpy(ls).to(self.dev)
            done = not (0.0 <= thr)  # visibility_global_(end) == 0 before the first sweep, :123,:127
            while not done:
                self._sweep(sx, sy)
                h, key, vg_end = self._epilogue(sx, sy, ex, ey, thr, nb, ls_dev)
                if h == U64_MAX:
                    raise VhpError(-3, "strip planner: no candidate cell (heap empty in the reference)")
                tx, ty = decode_key(key, sx, sy)
                nnb = nb + 1
                ls[nnb] = (tx, ty)
                if nnb > max_iter:
                    status, done = 5, True
                elif not (vg_end <= thr):
                    done = True
                elif (tx, ty) == (sx, sy):  # fixed point: the reference repeats it until max_iter
                    while nnb <= max_iter:
                        nnb += 1
                        ls[nnb] = (tx, ty)
                    status, done = 5, True
                nb = nnb
                sx, sy = tx, ty
                ls_dev[: nb + 1].copy_(torch.from_numpy(ls[: nb + 1]))
        path, length = np.zeros((0, 2), np.int32), 0.0
        if status == 0:
            ls[nb] = (ex, ey)  # :141
            pts = np.concatenate([ls[: nb + 1], [[ex, ey]]]).astype(np.int64)
            came_at = self._came_at(pts)
            # reconstructPath, :1183-1213
            x, y = ex, ey
            t, t_old, out = int(came_at[(x, y)]), -2, []
            while t != t_old and t >= 0 and len(out) < max_iter + 1:
                out.append((x, y))
                t_old = t
                x, y = int(ls[t][0]), int(ls[t][1])
                t = int(came_at[(x, y)])
            out.append((x, y))
            path = np.array(out[::-1], dtype=np.int32)
            for a, b in zip(path[:-1], path[1:]):
                dx, dy = float(a[0] - b[0]), int(a[1] - b[1])
                length = length + float(np.sqrt(np.float64(dx * dx) + np.float64(dy * dy)))
        return dict(status=status, nb_of_sources=nb, light_sources=ls[: nb + 1].copy(), path=path,
                    path_length=length)

    def _came_at(self, pts):
        """cameFrom_ at a list of cells, combined over the ranks."""
        torch = self.torch
        vals = torch.full((len(pts),), -(1 << 40), dtype=torch.int64, device=self.dev)
        for n, (x, y) in enumerate(pts):
            v = self._cell("came", int(x), int(y))
            if v is not None:
                vals[n] = v.to(torch.int64)
        if self.dist is not None and self.world > 1:
            self.dist.all_reduce(vals, op=self.dist.ReduceOp.MAX)
        vals = vals.cpu().numpy()
        return {(int(x), int(y)): int(v) for (x, y), v in zip(pts, vals)}

    def gather_field(self, name):
        """Whole field (numpy) assembled from the strips (every rank gets it)."""
        with self.torch.cuda.stream(self.stream):
            return self._gather_field(name)

    def _gather_field(self, name):
        torch = self.torch
        parts = []
        for k, (lo, hi) in enumerate(self.strips):
            if k in self.f:
                t = self.f[k][name]
            else:
                dt = torch.int32 if name == "came" else torch.float64
                t = torch.empty((hi - lo, self.nx), dtype=dt, device=self.dev)
            if self.dist is not None and self.world > 1:
                self.dist.broadcast(t, src=self.owner_of[k])
            parts.append(t.cpu().numpy())
        return np.concatenate(parts, axis=0)
