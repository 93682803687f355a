// planner_common.cuh -- the per-cell epilogue of updateVisibility
// (reference src/visibilityBasedSolver.cpp:417-430) and the arg-min key that stands for
// heap_->top() (:130), shared by the single-CTA planner kernel and the strip epilogue
// kernel of the giant-map path.
#ifndef VHP_PLANNER_COMMON_CUH
#define VHP_PLANNER_COMMON_CUH

#include <cstdint>

#include "vhp.h"

namespace {

struct Best {
  unsigned long long h;   // IEEE bits of h (h >= 0, so the bit pattern is monotonic)
  unsigned long long key; // quadrant << 40 | i << 20 | j  (push order)
};

__device__ __forceinline__ bool better(const Best &a, const Best &b) {
  return a.h < b.h || (a.h == b.h && a.key < b.key);
}

__device__ __forceinline__ double eval_d(int ax, int ay, int bx, int by) {
  // include/solver/visibilityBasedSolver.h:112-115 (all operands are exact integers)
  const double dx = (double)(ax - bx);
  const int dy = ay - by;
  return __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), (double)(dy * dy)));
}

constexpr unsigned long long kHInf = 0x7ff0000000000000ull; // +inf: the cell has no parent yet

// One visited cell (X, Y) (index c into the fields): max-merge into vg (:417-418),
// first-writer parent (:419-423), cached heuristic h = scale*vg + (d_end + d_parent)
// (:424-430, un-contracted), and the running arg-min on (h bits, push order).  The push
// order is quadrant Q1..Q4, i outer, j inner, of the FIRST quadrant that visits the cell
// (axis cells are visited twice with the same h; the earlier push wins).
// The arithmetic of one visited cell with its four field values already loaded (v = vis, h = hc,
// g0 = vg, cf = came): callers that walk many cells per thread issue the loads of several
// cells first (memory-level parallelism), then call this.
__device__ __forceinline__ void epilogue_cell_loaded(const int X, const int Y, const size_t c,
                                                     const double v, double h, const double g0,
                                                     int cf, const int sx, const int sy,
                                                     const int ex, const int ey, const double thr,
                                                     const double scale, const int nb,
                                                     const int32_t *__restrict__ ls, double *vg,
                                                     double *hc, int32_t *came, Best &best) {
  if ((X == 0 && sx > 0) || (Y == 0 && sy > 0)) return; // never visited (loop bounds :434-527)
  if (v > 0.0 || thr <= 0.0) { // a dark cell cannot raise vg or gain a parent (thr > 0)
    const double g = v > g0 ? v : g0; // std::max(v, vg)
    if (g != g0) vg[c] = g;
    const bool fresh = v >= thr && cf == VHP_NO_PARENT;
    if (fresh) {
      cf = nb;
      came[c] = nb;
    }
    if (cf != VHP_NO_PARENT && (fresh || g != g0)) { // (a parent implies vg >= thr)
      const int px = __ldcg(ls + 2 * cf), py = __ldcg(ls + 2 * cf + 1);
      h = __dadd_rn(__dmul_rn(scale, g), __dadd_rn(eval_d(X, Y, ex, ey), eval_d(X, Y, px, py)));
      hc[c] = h;
    }
  }
  const unsigned long long hb = (unsigned long long)__double_as_longlong(h);
  if (hb <= best.h && hb != kHInf) {
    const int dx = X - sx, dy = Y - sy;
    unsigned long long qd, i, j;
    if (dx >= 0 && dy >= 0) { qd = 0; i = dx; j = dy; }
    else if (dx < 0 && dy >= 0) { qd = 1; i = -dx; j = dy; }
    else if (dx <= 0 && (dx < 0 || sx >= 1)) { qd = 2; i = -dx; j = -dy; }
    else { qd = 3; i = dx; j = -dy; }
    const Best cand{hb, (qd << 40) | (i << 20) | j};
    if (better(cand, best)) best = cand;
  }
}

// The same cell in the FIRST sweep of a problem (nb == 0), where the fields still hold whatever
// the workspace held: nothing is loaded -- vg is 0, the cached heuristic +inf and the parent
// "none" (0 at the start cell, :122) after reset() :42-60 -- and all three are stored for every
// cell, the never-visited border included.  This replaces a reset pass over the fields and the
// first sweep's loads of them (a problem runs 1.8 sweeps on average in the batch of bench.py).
__device__ __forceinline__ void epilogue_cell_first(const int X, const int Y, const size_t c,
                                                    const double v, const int sx, const int sy,
                                                    const int ex, const int ey, const double thr,
                                                    const double scale, const int32_t *__restrict__,
                                                    double *vg, double *hc, int32_t *came, Best &best) {
  int cf = (X == sx && Y == sy) ? 0 : VHP_NO_PARENT; // the first source is the start cell
  double h = __longlong_as_double((long long)kHInf), g = 0.0;
  const bool visited = !((X == 0 && sx > 0) || (Y == 0 && sy > 0));
  if (visited && (v > 0.0 || thr <= 0.0)) {
    g = v > 0.0 ? v : 0.0; // std::max(v, 0)
    if (v >= thr && cf == VHP_NO_PARENT) cf = 0;
    if (cf != VHP_NO_PARENT) // the parent is light source 0 = the start cell = this sweep's source
      h = __dadd_rn(__dmul_rn(scale, g), __dadd_rn(eval_d(X, Y, ex, ey), eval_d(X, Y, sx, sy)));
  }
  vg[c] = g;
  came[c] = cf;
  hc[c] = h;
  if (!visited) return;
  const unsigned long long hb = (unsigned long long)__double_as_longlong(h);
  if (hb <= best.h && hb != kHInf) {
    const int dx = X - sx, dy = Y - sy;
    unsigned long long qd, i, j;
    if (dx >= 0 && dy >= 0) { qd = 0; i = dx; j = dy; }
    else if (dx < 0 && dy >= 0) { qd = 1; i = -dx; j = dy; }
    else if (dx <= 0 && (dx < 0 || sx >= 1)) { qd = 2; i = -dx; j = -dy; }
    else { qd = 3; i = dx; j = -dy; }
    const Best cand{hb, (qd << 40) | (i << 20) | j};
    if (better(cand, best)) best = cand;
  }
}

__device__ __forceinline__ void epilogue_cell(const int X, const int Y, const size_t c, const int sx,
                                              const int sy, const int ex, const int ey,
                                              const double thr, const double scale, const int nb,
                                              const int32_t *__restrict__ ls,
                                              const double *__restrict__ vis, double *vg,
                                              double *hc, int32_t *came, Best &best) {
  if ((X == 0 && sx > 0) || (Y == 0 && sy > 0)) return; // never visited (loop bounds :434-527)
  const double v = __ldcg(vis + c);
  double h = __ldcg(hc + c);
  if (v > 0.0 || thr <= 0.0) { // a dark cell cannot raise vg or gain a parent (thr > 0)
    const double g0 = __ldcg(vg + c);
    const double g = v > g0 ? v : g0; // std::max(v, vg)
    if (g != g0) vg[c] = g;
    int cf = __ldcg(came + c);
    const bool fresh = v >= thr && cf == VHP_NO_PARENT;
    if (fresh) {
      cf = nb;
      came[c] = nb;
    }
    if (cf != VHP_NO_PARENT && (fresh || g != g0)) { // (a parent implies vg >= thr)
      const int px = __ldcg(ls + 2 * cf), py = __ldcg(ls + 2 * cf + 1);
      h = __dadd_rn(__dmul_rn(scale, g), __dadd_rn(eval_d(X, Y, ex, ey), eval_d(X, Y, px, py)));
      hc[c] = h;
    }
  }
  const unsigned long long hb = (unsigned long long)__double_as_longlong(h);
  if (hb <= best.h && hb != kHInf) {
    const int dx = X - sx, dy = Y - sy;
    unsigned long long qd, i, j;
    if (dx >= 0 && dy >= 0) { qd = 0; i = dx; j = dy; }
    else if (dx < 0 && dy >= 0) { qd = 1; i = -dx; j = dy; }
    else if (dx <= 0 && (dx < 0 || sx >= 1)) { qd = 2; i = -dx; j = -dy; }
    else { qd = 3; i = dx; j = -dy; }
    const Best cand{hb, (qd << 40) | (i << 20) | j};
    if (better(cand, best)) best = cand;
  }
}

// ---- loop control of solve(), shared by the single-CTA planner kernel (one thread of the
// persistent CTA) and the control kernels of the many-CTA path for large single problems.

// Validity checks in the reference's order (:89-116); rowbits = forward row bit plane of the map.
__device__ __forceinline__ int planner_validate(const uint32_t *rowbits, const int wx, const int nx,
                                                const int ny, const int stx, const int sty,
                                                const int ex, const int ey) {
  auto free_cell = [&](int x, int y) { return (rowbits[(size_t)y * wx + (x >> 5)] >> (x & 31)) & 1u; };
  if ((unsigned)stx >= (unsigned)nx || (unsigned)sty >= (unsigned)ny) return VHP_START_OOB;
  if ((unsigned)ex >= (unsigned)nx || (unsigned)ey >= (unsigned)ny) return VHP_END_OOB;
  if (!free_cell(stx, sty)) return VHP_START_OCCUPIED;
  if (!free_cell(ex, ey)) return VHP_END_OCCUPIED;
  return VHP_OK;
}

// heap_->top() -> next light source (:130-139) after the sweep from (sx, sy) as source nb.
// Appends to ls, returns the new nb_of_sources; tx/ty = the next source; done / status as the
// loop test (:127), max_iter and the fixed point decide.
__device__ __forceinline__ int planner_next_source(const Best b, const int sx, const int sy,
                                                   const int nb, const int max_iter,
                                                   const double thr, const double vg_end,
                                                   int32_t *ls, int &tx, int &ty, int &done,
                                                   int &status) {
  const int qd = (int)(b.key >> 40), i = (int)((b.key >> 20) & 0xFFFFF), j = (int)(b.key & 0xFFFFF);
  tx = (qd == 0 || qd == 3) ? sx + i : sx - i;
  ty = (qd < 2) ? sy + j : sy - j;
  int nnb = nb + 1;
  ls[2 * nnb] = tx;
  ls[2 * nnb + 1] = ty;
  done = 0;
  if (nnb > max_iter) { done = 1; status = VHP_MAX_ITER; }
  else if (!(vg_end <= thr)) done = 1; // loop test :127
  else if (tx == sx && ty == sy) {
    // fixed point: the same source again -> every further iteration is identical
    while (nnb <= max_iter) {
      ++nnb;
      ls[2 * nnb] = tx;
      ls[2 * nnb + 1] = ty;
    }
    done = 1;
    status = VHP_MAX_ITER;
  }
  return nnb;
}

// lightSources_[nb] = end (:141) and reconstructPath (:1183-1213): walk end -> start through
// cameFrom_ / lightSources_, reverse, sum the segment lengths.  Returns the number of points.
// came_of(t, x, y) = cameFrom_(x, y), where (x, y) is lightSources_[t] (t == nb: the end point);
// the single-GPU kernels read the field, the strip-partitioned planner a table gathered from
// the ranks that own the rows.
template <typename CameOf>
__device__ __forceinline__ long planner_reconstruct_with(const int status, const int nb, const int ex,
                                                         const int ey, const int ls_cap, int32_t *ls,
                                                         int32_t *path, double &total,
                                                         CameOf came_of) {
  long n = 0;
  total = 0.0;
  if (status != VHP_OK) return 0;
  ls[2 * nb] = ex; ls[2 * nb + 1] = ey;
  int x = ex, y = ey;
  int t = came_of(nb, x, y), t_old = -2;
  while (t != t_old && t >= 0 && n < ls_cap - 1) {
    path[2 * n] = x; path[2 * n + 1] = y; ++n;
    t_old = t;
    x = ls[2 * t]; y = ls[2 * t + 1];
    t = came_of(t, x, y);
  }
  path[2 * n] = x; path[2 * n + 1] = y; ++n;
  for (long a = 0, b = n - 1; a < b; ++a, --b) {
    const int tx = path[2 * a], ty = path[2 * a + 1];
    path[2 * a] = path[2 * b]; path[2 * a + 1] = path[2 * b + 1];
    path[2 * b] = tx; path[2 * b + 1] = ty;
  }
  for (long k = 0; k + 1 < n; ++k)
    total = __dadd_rn(total, eval_d(path[2 * k], path[2 * k + 1], path[2 * k + 2], path[2 * k + 3]));
  return n;
}

__device__ __forceinline__ long planner_reconstruct(const int status, const int nb, const int ex,
                                                    const int ey, const int nx, const int ls_cap,
                                                    int32_t *ls, const int32_t *came, int32_t *path,
                                                    double &total) {
  return planner_reconstruct_with(status, nb, ex, ey, ls_cap, ls, path, total,
                                  [&](int, int x, int y) { return __ldcg(came + (size_t)y * nx + x); });
}

// lane 0 of the warp ends up with the warp's best
__device__ __forceinline__ Best warp_best(Best b) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    Best o;
    o.h = __shfl_xor_sync(0xffffffffu, b.h, off);
    o.key = __shfl_xor_sync(0xffffffffu, b.key, off);
    if (better(o, b)) b = o;
  }
  return b;
}

} // namespace
#endif
