// kernels_planner.cu -- K2/K3/K5: the visibility-heuristic planner loop on the device.
//
// Replaces, per problem (map, start, end):
//   visibilityBasedSolver::solve            src/visibilityBasedSolver.cpp:76-160
//   visibilityBasedSolver::updateVisibility src/visibilityBasedSolver.cpp:379-565
//   resetQueue / heap_->top()               :65-71, :130
//   visibilityBasedSolver::reconstructPath  src/visibilityBasedSolver.cpp:1183-1213
//
// One persistent CTA per problem runs the whole data-dependent loop without any
// host round trip:
//   1. sweep from the current light source (the tile-wavefront K1 body of
//      sweep_tile_body.cuh, fp64 store) into the local visibility field `vis`
//      -- the DP part of updateVisibility;
//   2. fused per-cell epilogue over the grid (coalesced): global visibility
//      max-merge (:417-418), first-writer parent (:419-423), heuristic
//      h = scale*vg + (d(cell,end) + d(cell,parent)) (:424-430) and the arg-min
//      that stands for the heap: the reference pushes every visible cell and only
//      ever reads top(), which is the FIRST pushed element attaining the minimum h
//      (Node::operator< is strict).  The push order is quadrant Q1..Q4, i outer,
//      j inner, so the reduction key is (h, quadrant, i, j) with the first
//      quadrant that visits the cell;
//   3. next source = arg-min cell, ++nb_of_sources, max_iter check, end-visible
//      test (:127-140).
// Two exact shortcuts keep the loop cheap:
//   * h of a cell only depends on (vg, end, parent) and the parent never changes once
//     set (:419-423), so h is cached per cell (`hc`) and recomputed only where vg rose or
//     the parent was just assigned; the arg-min then reads 16 bytes per cell and the
//     push-order key is only evaluated for cells that tie or beat the running minimum.
//   * the reference has no closed set: once the arg-min cell IS the current source it is
//     picked again forever (same sweep -> same fields -> same h), until max_iter
//     (SURVEY A.2 item 7).  That fixed point is detected and the remaining identical
//     iterations are not executed: the light-source list is filled with the repeated
//     cell and the status is VHP_MAX_ITER, exactly as the reference ends.
// Afterwards thread 0 walks cameFrom_/lightSources_ from the end point
// (reconstructPath) and sums the segment lengths.
//
// Cells the reference never visits (column 0 / row 0 unless the source lies on
// them, loop bounds :434-438,:478-483,:522-527) get no epilogue.
#include <cstdint>

#include "sweep_tile_body.cuh"

namespace {

struct PlannerParams {
  TileArgs fp;
  const int32_t *se_xy, *prob_map;
  double thr;
  int max_iter, ls_cap;
  // working fields, fp64, [nprob][ny][nx]
  double *vis, *vg;
  double *hc;      // cached heuristic h per cell, +inf where the cell has no parent yet
  int32_t *came;
  // small outputs, [nprob]...
  int32_t *status, *nb, *ls, *path_n, *path;
  double *path_len;
  // optional fp32 exports of vg / vis (NULL: none)
  float *vg32, *vis32;
};

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

__global__ void __launch_bounds__(kTileWarps * 32) planner_kernel(const PlannerParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_ctl[5];              // {done, next x, next y, status, nb_of_sources}
  __shared__ Best s_best[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;
  const int64_t q = blockIdx.x;
  const int nx = p.fp.nx, ny = p.fp.ny;
  const size_t cells = (size_t)nx * ny;
  const int stx = p.se_xy[4 * q], sty = p.se_xy[4 * q + 1];
  const int ex = p.se_xy[4 * q + 2], ey = p.se_xy[4 * q + 3];
  const int map = p.prob_map ? p.prob_map[q] : 0;
  const uint32_t *rowbits = p.fp.pl.rowF + (size_t)map * p.fp.pl.row_plane;
  double *vis = p.vis + q * cells, *vg = p.vg + q * cells, *hc = p.hc + q * cells;
  int32_t *came = p.came + q * cells;
  int32_t *ls = p.ls + q * (size_t)p.ls_cap * 2;
  const double thr = p.thr;

  // reset(): :42-60 (every sweep writes all of `vis`, border cells included)
  for (size_t c = tid; c < cells; c += blockDim.x) {
    vg[c] = 0.0;
    hc[c] = __longlong_as_double(0x7ff0000000000000ll);
    came[c] = VHP_NO_PARENT;
  }
  if (tid == 0) {
    auto free_cell = [&](int x, int y) {
      return (rowbits[(size_t)y * p.fp.pl.wx + (x >> 5)] >> (x & 31)) & 1u;
    };
    int st = VHP_OK; // checks in the reference's order, :89-116
    if ((unsigned)stx >= (unsigned)nx || (unsigned)sty >= (unsigned)ny) st = VHP_START_OOB;
    else if ((unsigned)ex >= (unsigned)nx || (unsigned)ey >= (unsigned)ny) st = VHP_END_OOB;
    else if (!free_cell(stx, sty)) st = VHP_START_OCCUPIED;
    else if (!free_cell(ex, ey)) st = VHP_END_OCCUPIED;
    s_ctl[3] = st;
    s_ctl[1] = stx;
    s_ctl[2] = sty;
    s_ctl[0] = 0;
  }
  __syncthreads();
  int status = s_ctl[3];
  int nb = 0;
  if (status == VHP_OK) {
    if (tid == 0) {
      ls[0] = stx; ls[1] = sty;                 // lightSources_[0] = start, :121
      came[(size_t)sty * nx + stx] = 0;         // :122
      // visibility_global_(end) = 0 (:123) holds after the reset; loop test :127
      s_ctl[0] = !(0.0 <= thr);
    }
    __syncthreads();
    const double scale =
        __dsqrt_rn((double)((unsigned long long)ny * ny + (unsigned long long)nx * nx)); // :49
    int sx = stx, sy = sty;
    bool done = s_ctl[0] != 0;
    while (!done) {
      // ---- 1. sweep (visibility_.reset() + the DP of updateVisibility)
      tile_sweep_cta<double, kTileWarps>(p.fp, map, sx, sy, vis, smem_raw);
      // ---- 2. per-cell epilogue + arg-min
      Best best{~0ull, ~0ull};
      {
        int X = tid % nx, Y = tid / nx; // cell c = tid + k * blockDim.x, walked incrementally
        const int bdx = blockDim.x % nx, bdy = blockDim.x / nx;
        for (size_t c = tid; c < cells; c += blockDim.x) {
          const bool visited = !((X == 0 && sx > 0) || (Y == 0 && sy > 0));
          if (visited) {
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
              if (cf != VHP_NO_PARENT && (fresh || g != g0)) { // (parent set implies vg >= thr)
                const int px = __ldcg(ls + 2 * cf), py = __ldcg(ls + 2 * cf + 1);
                h = __dadd_rn(__dmul_rn(scale, g),
                              __dadd_rn(eval_d(X, Y, ex, ey), eval_d(X, Y, px, py)));
                hc[c] = h;
              }
            }
            const unsigned long long hb = (unsigned long long)__double_as_longlong(h);
            if (hb <= best.h && hb != 0x7ff0000000000000ull) {
              const int dx = X - sx, dy = Y - sy;
              unsigned long long qd, i, j; // first quadrant that visits the cell, :388-564
              if (dx >= 0 && dy >= 0) { qd = 0; i = dx; j = dy; }
              else if (dx < 0 && dy >= 0) { qd = 1; i = -dx; j = dy; }
              else if (dx <= 0 && (dx < 0 || sx >= 1)) { qd = 2; i = -dx; j = -dy; }
              else { qd = 3; i = dx; j = -dy; }
              const Best cand{hb, (qd << 40) | (i << 20) | j};
              if (better(cand, best)) best = cand;
            }
          }
          X += bdx;
          Y += bdy;
          if (X >= nx) { X -= nx; ++Y; }
        }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        Best o;
        o.h = __shfl_xor_sync(0xffffffffu, best.h, off);
        o.key = __shfl_xor_sync(0xffffffffu, best.key, off);
        if (better(o, best)) best = o;
      }
      if (lane == 0) s_best[warp] = best;
      __syncthreads();
      if (warp == 0) {
        Best b = lane < NW ? s_best[lane] : Best{~0ull, ~0ull};
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          Best o;
          o.h = __shfl_xor_sync(0xffffffffu, b.h, off);
          o.key = __shfl_xor_sync(0xffffffffu, b.key, off);
          if (better(o, b)) b = o;
        }
        if (lane == 0) {
          // ---- 3. heap_->top() -> next light source, :130-139
          const int qd = (int)(b.key >> 40), i = (int)((b.key >> 20) & 0xFFFFF), j = (int)(b.key & 0xFFFFF);
          const int tx = (qd == 0 || qd == 3) ? sx + i : sx - i;
          const int ty = (qd < 2) ? sy + j : sy - j;
          int nnb = nb + 1;
          ls[2 * nnb] = tx;
          ls[2 * nnb + 1] = ty;
          s_ctl[1] = tx;
          s_ctl[2] = ty;
          int d = 0;
          if (nnb > p.max_iter) { d = 1; s_ctl[3] = VHP_MAX_ITER; }
          else if (!(__ldcg(vg + (size_t)ey * nx + ex) <= thr)) d = 1; // loop test :127
          else if (tx == sx && ty == sy) {
            // fixed point: the same source again -> every further iteration is identical
            while (nnb <= p.max_iter) {
              ++nnb;
              ls[2 * nnb] = tx;
              ls[2 * nnb + 1] = ty;
            }
            d = 1;
            s_ctl[3] = VHP_MAX_ITER;
          }
          s_ctl[0] = d;
          s_ctl[4] = nnb;
        }
      }
      __syncthreads();
      nb = s_ctl[4];
      sx = s_ctl[1];
      sy = s_ctl[2];
      done = s_ctl[0] != 0;
      status = s_ctl[3];
      __syncthreads();
    }
  }

  if (nb == 0) // no sweep ran: visibility_ keeps the zeros of reset() (:43)
    for (size_t c = tid; c < cells; c += blockDim.x) vis[c] = 0.0;
  if (tid == 0) {
    p.status[q] = status;
    p.nb[q] = nb;
    int32_t *path = p.path + q * (size_t)p.ls_cap * 2;
    long n = 0;
    double total = 0.0;
    if (status == VHP_OK) {
      ls[2 * nb] = ex; ls[2 * nb + 1] = ey; // lightSources_[nb] = end, :141
      // reconstructPath, :1183-1213: walk end -> start, then reverse
      int x = ex, y = ey;
      int t = __ldcg(came + (size_t)y * nx + x), t_old = -2;
      while (t != t_old && t >= 0 && n < p.ls_cap - 1) {
        path[2 * n] = x; path[2 * n + 1] = y; ++n;
        t_old = t;
        x = ls[2 * t]; y = ls[2 * t + 1];
        t = __ldcg(came + (size_t)y * nx + x);
      }
      path[2 * n] = x; path[2 * n + 1] = y; ++n;
      for (long a = 0, b = n - 1; a < b; ++a, --b) {
        const int tx = path[2 * a], ty = path[2 * a + 1];
        path[2 * a] = path[2 * b]; path[2 * a + 1] = path[2 * b + 1];
        path[2 * b] = tx; path[2 * b + 1] = ty;
      }
      for (long k = 0; k + 1 < n; ++k)
        total = __dadd_rn(total, eval_d(path[2 * k], path[2 * k + 1], path[2 * k + 2], path[2 * k + 3]));
    }
    p.path_n[q] = (int32_t)n;
    p.path_len[q] = total;
  }
  __syncthreads();
  if (p.vg32 || p.vis32) {
    for (size_t c = tid; c < cells; c += blockDim.x) {
      if (p.vg32) p.vg32[q * cells + c] = __double2float_rn(__ldcg(vg + c));
      if (p.vis32) p.vis32[q * cells + c] = __double2float_rn(__ldcg(vis + c));
    }
  }
}

} // namespace

bool vhp_planner_supported(int nx, int ny) { return vhp_sweep_tile_supported(nx, ny); }

cudaError_t vhp_launch_planner(const VhpTilePlanes &pl, int nx, int ny, const int32_t *d_se_xy,
                               const int32_t *d_prob_map, int64_t nprob, double threshold,
                               int32_t max_iter, int32_t ls_cap, const double *d_rcp2,
                               double *d_vis, double *d_vg, double *d_hc, int32_t *d_came,
                               int32_t *d_status,
                               int32_t *d_nb, int32_t *d_ls, double *d_path_len, int32_t *d_path_n,
                               int32_t *d_path, float *d_vg32, float *d_vis32, int *d_err,
                               cudaStream_t st, int64_t *launches) {
  if (!vhp_sweep_tile_supported(nx, ny)) return cudaErrorInvalidConfiguration;
  PlannerParams p;
  p.fp.pl = pl;
  p.fp.nx = nx; p.fp.ny = ny;
  p.fp.src_xy = nullptr; p.fp.src_map = nullptr; p.fp.out = nullptr;
  p.fp.rtab = reinterpret_cast<const double2 *>(d_rcp2);
  p.fp.err = d_err;
  p.fp.vec = ((uintptr_t)d_vis % 16 == 0 && nx % 2 == 0) ? 1 : 0;
  p.se_xy = d_se_xy; p.prob_map = d_prob_map;
  p.thr = threshold; p.max_iter = max_iter; p.ls_cap = ls_cap;
  p.vis = d_vis; p.vg = d_vg; p.hc = d_hc; p.came = d_came;
  p.status = d_status; p.nb = d_nb; p.ls = d_ls;
  p.path_len = d_path_len; p.path_n = d_path_n; p.path = d_path;
  p.vg32 = d_vg32; p.vis32 = d_vis32;
  const size_t smem = tile_smem_bytes<double>(nx, ny);
  cudaError_t e = cudaFuncSetAttribute(planner_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return e;
  planner_kernel<<<(unsigned)nprob, kTileWarps * 32, smem, st>>>(p);
  if (launches) *launches += 1;
  return cudaGetLastError();
}
