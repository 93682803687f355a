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
