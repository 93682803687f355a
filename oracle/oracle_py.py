"""ctypes loaders for the CPU oracle and the compiled reference.

TEST INFRASTRUCTURE -- only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.  The product
package never does.

  Oracle()  -> oracle/_build/libvhp_oracle.so   (oracle/vhp_oracle.c, strict IEEE)
  Ref(kind) -> oracle/_ref/libvhp_ref_{strict,fast}.so (the unmodified reference
               behind oracle/ref_harness.cpp); `Ref.available(kind)` says whether
               the prebuilt file exists (it is built in the dev container from
               /root/reference and travels to the GPU box as a binary).

All fields are numpy arrays of shape (ny, nx), C-contiguous, i.e. the reference's
Field<T> layout index = x + y*nx (include/environment/field.h:26-29).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NO_PARENT = 1000000000000000
STATUS = ["OK", "START_OOB", "END_OOB", "START_OCCUPIED", "END_OCCUPIED", "MAX_ITER"]

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(ref: bool = True) -> None:
    """(Re)build the oracle and, if /root/reference exists, the reference."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if ref and os.path.isdir(os.environ.get("VHP_REF_SRC", "/root/reference")):
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class _Base:
    """Common numpy-level API over either library (prefix differs)."""

    prefix = ""

    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def _bind_common(self):
        f = self._fn("compute_visibility")
        f.argtypes = [_dp, C.c_int, C.c_int, C.c_int, C.c_int, _dp]
        f.restype = None
        f = self._fn("update_visibility")
        f.argtypes = [_dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                      C.c_double, _dp, _dp, _u64p, _ip, C.c_uint64, _ip,
                      C.POINTER(C.c_double)]
        f.restype = C.c_long
        f = self._fn("raycast_all")
        f.argtypes = [_dp, C.c_int, C.c_int, C.c_int, C.c_int, _dp]
        f.restype = None
        f = self._fn("generate_environment")
        f.argtypes = [_dp, C.c_int, C.c_int, C.c_long, C.c_long, C.c_long,
                      C.c_long, C.c_long, C.c_int]
        f.restype = None

    # -- a1 ---------------------------------------------------------------
    def compute_visibility(self, occ, sx, sy, vis_init=None):
        occ = _f64(occ)
        ny, nx = occ.shape
        vis = np.zeros((ny, nx)) if vis_init is None else _f64(vis_init).copy()
        self._fn("compute_visibility")(occ, nx, ny, int(sx), int(sy), vis)
        return vis

    # -- a2 ---------------------------------------------------------------
    def update_visibility(self, occ, src, end, thr, vg, came, ls_xy, nb):
        """One planner sweep.  vg/came are updated in place.  Returns
        (vis, (top_x, top_y), top_h, pushes)."""
        occ = _f64(occ)
        ny, nx = occ.shape
        vis = np.zeros((ny, nx))
        top = np.zeros(2, dtype=np.int32)
        h = C.c_double(0.0)
        ls = np.ascontiguousarray(ls_xy, dtype=np.int32).reshape(-1)
        pushes = self._fn("update_visibility")(
            occ, nx, ny, int(src[0]), int(src[1]), int(end[0]), int(end[1]),
            float(thr), vis, vg, came, ls, int(nb), top, C.byref(h))
        return vis, (int(top[0]), int(top[1])), h.value, int(pushes)

    # -- a5 ---------------------------------------------------------------
    def raycast_all(self, occ, sx, sy, ray_init=None):
        occ = _f64(occ)
        ny, nx = occ.shape
        ray = np.ones((ny, nx)) if ray_init is None else _f64(ray_init).copy()
        self._fn("raycast_all")(occ, nx, ny, int(sx), int(sy), ray)
        return ray

    def generate_environment(self, nx, ny, nb_of_obstacles, min_w, max_w, min_h,
                             max_h, seed):
        occ = np.empty((ny, nx))
        self._fn("generate_environment")(occ, nx, ny, nb_of_obstacles, min_w,
                                         max_w, min_h, max_h, seed)
        return occ


class Oracle(_Base):
    def accessibility_map(self, occ, sx, sy, alpha=1.0, fac=1.0, light_strength=1.0):
        """getAccessibilityMap.m restated (unpinned: MATLAB cannot run here)."""
        occ = _f64(occ)
        ny, nx = occ.shape
        vis = np.zeros((ny, nx))
        f = self.lib.vhp_oracle_accessibility_map
        f.argtypes = [_dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, _dp]
        f.restype = None
        f(occ, nx, ny, int(sx), int(sy), float(alpha), float(fac), float(light_strength), vis)
        return vis

    def visibility_cutoff(self, occ, sx, sy, cutoff=0.001):
        """computeVisibilityUsingQueue() as an order-free rule (vhp_oracle.h)."""
        occ = _f64(occ)
        ny, nx = occ.shape
        vis = np.zeros((ny, nx))
        f = self.lib.vhp_oracle_visibility_cutoff
        f.argtypes = [_dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _dp]
        f.restype = None
        f(occ, nx, ny, int(sx), int(sy), float(cutoff), vis)
        return vis

    prefix = "vhp_oracle_"

    def generate_environment_counter(self, nx, ny, nb_of_obstacles, min_w, max_w, min_h, max_h,
                                     seed, map_index):
        """Map `map_index` of the family `seed` of the batch generator (counter-based draws)."""
        f = self.lib.vhp_oracle_generate_environment_counter
        f.argtypes = [_dp, C.c_int, C.c_int, C.c_long, C.c_long, C.c_long, C.c_long, C.c_long,
                      C.c_ulonglong, C.c_ulonglong]
        f.restype = None
        occ = np.empty((ny, nx))
        f(occ, nx, ny, nb_of_obstacles, min_w, max_w, min_h, max_h, seed, map_index)
        return occ

    def __init__(self):
        path = os.path.join(HERE, "_build", "libvhp_oracle.so")
        if not os.path.exists(path):
            build(ref=False)
        self.lib = C.CDLL(path)
        self._bind_common()
        self.lib.vhp_oracle_env_draw.argtypes = [C.c_ulonglong, C.c_ulonglong, C.c_ulonglong, C.c_uint]
        self.lib.vhp_oracle_env_draw.restype = C.c_uint
        f = self.lib.vhp_oracle_solve
        f.argtypes = [_dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                      C.c_double, C.c_long, _dp, _dp, _u64p, _ip,
                      C.POINTER(C.c_long)]
        f.restype = C.c_int
        f = self.lib.vhp_oracle_reconstruct_path
        f.argtypes = [_u64p, C.c_int, _ip, C.c_int, C.c_int, _ip, C.c_long,
                      C.POINTER(C.c_double)]
        f.restype = C.c_long
        self.lib.vhp_oracle_eval_d.argtypes = [C.c_int] * 4
        self.lib.vhp_oracle_eval_d.restype = C.c_double

    def solve(self, occ, start, end, thr, max_iter):
        """solve() + reconstructPath().  Returns a dict like Ref.solve."""
        occ = _f64(occ)
        ny, nx = occ.shape
        vis = np.zeros((ny, nx))
        vg = np.zeros((ny, nx))
        came = np.zeros((ny, nx), dtype=np.uint64)
        ls = np.zeros(2 * (max_iter + 2), dtype=np.int32)
        nb = C.c_long(0)
        st = self.lib.vhp_oracle_solve(occ, nx, ny, int(start[0]), int(start[1]),
                                       int(end[0]), int(end[1]), float(thr),
                                       int(max_iter), vis, vg, came, ls,
                                       C.byref(nb))
        out = dict(status=st, nb_of_sources=nb.value, vis=vis, vg=vg, came=came,
                   light_sources=ls.reshape(-1, 2)[: nb.value + 1].copy(),
                   path=np.zeros((0, 2), np.int32), path_length=0.0)
        if st == 0:
            cap = nb.value + 2
            path = np.zeros(2 * cap, dtype=np.int32)
            length = C.c_double(0.0)
            n = self.lib.vhp_oracle_reconstruct_path(
                came, nx, ls, int(end[0]), int(end[1]), path, cap, C.byref(length))
            out["path"] = path.reshape(-1, 2)[:n].copy()
            out["path_length"] = length.value
        return out


class Ref(_Base):
    prefix = "ref_"

    def compute_visibility_queue(self, occ, sx, sy, vis_init=None):
        """computeVisibilityUsingQueue() (:701-893) of the compiled reference."""
        occ = _f64(occ)
        ny, nx = occ.shape
        vis = np.zeros((ny, nx)) if vis_init is None else _f64(vis_init).copy()
        f = self.lib.ref_compute_visibility_queue
        f.argtypes = [_dp, C.c_int, C.c_int, C.c_int, C.c_int, _dp]
        f.restype = None
        f(occ, nx, ny, int(sx), int(sy), vis)
        return vis

    @staticmethod
    def path(kind="strict"):
        return os.path.join(HERE, "_ref", f"libvhp_ref_{kind}.so")

    @staticmethod
    def available(kind="strict"):
        return os.path.exists(Ref.path(kind))

    @staticmethod
    def host_has_avx512():
        try:
            flags = next(l for l in open("/proc/cpuinfo") if l.startswith("flags")).split()
        except Exception:
            return False
        return all(f in flags for f in ("avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl"))

    def __init__(self, kind="strict"):
        # "fast" = the reference's own CMake flags (-O3 -Ofast -march=native -flto); built off the
        # box, so "native" is the best prebuilt variant this host's CPU can run
        if kind == "fast" and os.environ.get("VHP_REF_FAST", "") != "v3" and Ref.available("fast_v4") \
                and Ref.host_has_avx512():
            kind = "fast_v4"
        self.kind = kind
        self.lib = C.CDLL(Ref.path(kind))
        self._bind_common()
        self.lib.ref_build_flags.restype = C.c_char_p
        f = self.lib.ref_solve
        f.argtypes = [_dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                      C.c_double, C.c_long, _dp, _dp, _u64p, _ip,
                      C.POINTER(C.c_long), _ip, C.c_long, C.POINTER(C.c_long),
                      C.POINTER(C.c_double), C.POINTER(C.c_double)]
        f.restype = C.c_int
        f = self.lib.ref_time_compute_visibility
        f.argtypes = [_dp, C.c_int, C.c_int, _ip, C.c_int, C.c_int,
                      C.POINTER(C.c_double)]
        f.restype = C.c_double
        f = self.lib.ref_time_solve
        f.argtypes = [_dp, C.c_int, C.c_int, _ip, C.c_int, C.c_double, C.c_long,
                      C.c_int, C.POINTER(C.c_long)]
        f.restype = C.c_double

    def flags(self):
        return self.lib.ref_build_flags().decode()

    def solve(self, occ, start, end, thr, max_iter):
        occ = _f64(occ)
        ny, nx = occ.shape
        vis = np.zeros((ny, nx))
        vg = np.zeros((ny, nx))
        came = np.zeros((ny, nx), dtype=np.uint64)
        ls = np.zeros(2 * (max_iter + 2), dtype=np.int32)
        cap = max_iter + 3
        path = np.zeros(2 * cap, dtype=np.int32)
        nb, pn = C.c_long(0), C.c_long(0)
        plen, printed = C.c_double(0.0), C.c_double(0.0)
        st = self.lib.ref_solve(occ, nx, ny, int(start[0]), int(start[1]),
                                int(end[0]), int(end[1]), float(thr),
                                int(max_iter), vis, vg, came, ls, C.byref(nb),
                                path, cap, C.byref(pn), C.byref(plen),
                                C.byref(printed))
        return dict(status=st, nb_of_sources=nb.value, vis=vis, vg=vg, came=came,
                    light_sources=ls.reshape(-1, 2)[: nb.value + 1].copy(),
                    path=path.reshape(-1, 2)[: pn.value].copy(),
                    path_length=plen.value, printed_length=printed.value)

    def time_compute_visibility(self, occ, sources, nthreads=1):
        """Wall seconds for computeVisibility() over all sources (see harness)."""
        occ = _f64(occ)
        ny, nx = occ.shape
        src = np.ascontiguousarray(sources, dtype=np.int32).reshape(-1)
        chk = C.c_double(0.0)
        return self.lib.ref_time_compute_visibility(
            occ, nx, ny, src, len(src) // 2, int(nthreads), C.byref(chk))

    def time_solve(self, occ, start_end, thr, max_iter, nthreads=1):
        occ = _f64(occ)
        ny, nx = occ.shape
        se = np.ascontiguousarray(start_end, dtype=np.int32).reshape(-1)
        tot = C.c_long(0)
        secs = self.lib.ref_time_solve(occ, nx, ny, se, len(se) // 4, float(thr),
                                       int(max_iter), int(nthreads), C.byref(tot))
        return secs, tot.value
