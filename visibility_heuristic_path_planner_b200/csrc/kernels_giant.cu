// kernels_giant.cu -- control kernels of the strip-partitioned planner (csrc/giant.cu): the loop
// of solve() (reference src/visibilityBasedSolver.cpp:76-160) with ALL of its state on the
// device, so that a whole planner iteration -- strip sweeps, halo rows, epilogue, arg-min
// exchange, next-source selection -- is a fixed sequence of launches with fixed arguments.
// That sequence is either captured once as the body of a CUDA-graph WHILE node (one process)
// or enqueued a few iterations ahead of the host (several ranks, NCCL in between); every
// kernel returns at once when ctl[0] ("done") is set.
//
//   ctl = int[12] {done, source x, source y, status, nb_of_sources, iterations run,
//                  end x, end y, threshold (a double in ints 8-9), max_iter, -}
// The query (end point, threshold, max_iter) lives there too, so the captured graph depends
// only on the geometry and the buffers and is reused for every query on the same map.
#include <climits>
#include <cstdint>

#include "planner_common.cuh"
#include "sweep_tile_body.cuh"
#include "giant_internal.h"

namespace {

__global__ void giant_reset_kernel(double *vis, double *vg, double *hc, int32_t *came, size_t cells) {
  for (size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x; c < cells;
       c += (size_t)gridDim.x * blockDim.x) { // reset(), :42-60
    vis[c] = 0.0;
    vg[c] = 0.0;
    hc[c] = __longlong_as_double(0x7ff0000000000000ll);
    came[c] = VHP_NO_PARENT;
  }
}

__device__ __forceinline__ const GiantStripDev *strip_of_row(const GiantLocal &loc, const int y) {
  for (int s = 0; s < loc.n; ++s)
    if (y >= loc.s[s].y0 && y < loc.s[s].y1) return &loc.s[s];
  return nullptr;
}

__global__ void giant_begin_kernel(const uint32_t *rowbits, int wx, int nx, int ny, int stx, int sty,
                                   int ex, int ey, double thr, int max_iter, int32_t *ls,
                                   const GiantLocal loc, int *ctl) {
  const int st = planner_validate(rowbits, wx, nx, ny, stx, sty, ex, ey); // :89-116
  int done = 1;
  if (st == VHP_OK) {
    ls[0] = stx; ls[1] = sty;                 // lightSources_[0] = start, :121
    if (const GiantStripDev *s = strip_of_row(loc, sty))
      s->came[(size_t)(sty - s->y0) * nx + stx] = 0; // :122
    done = !(0.0 <= thr);                     // visibility_global_(end) = 0 (:123), loop test :127
  }
  ctl[0] = done; ctl[1] = stx; ctl[2] = sty; ctl[3] = st; ctl[4] = 0; ctl[5] = 0;
  ctl[6] = ex; ctl[7] = ey;
  *reinterpret_cast<double *>(ctl + 8) = thr;
  ctl[10] = max_iter; ctl[11] = 0;
}

// The fp64 visibility rows strip [cy0, cy1) needs below its first tile rows of quadrants qfirst
// and qfirst + 1 (the two quadrants of one y direction), taken from the strip [y0, y1) that
// holds them: dst[qq][x].  Both sides of a strip boundary derive the row numbers from the same
// geometry (tile_window_of), so nothing is negotiated; a quadrant without a halo row (it starts
// at the source, or has no rows in the consumer's window) leaves dst[qq] untouched and the
// sweep kernel ignores it.
__global__ void giant_halo_gather_kernel(const int *ctl, int nx, int ny, const double *vis, int y0,
                                         int y1, int cy0, int cy1, int qfirst, double *dst) {
  if (ctl[0]) return;
  const int sx = ctl[1], sy = ctl[2];
  for (int qq = 0; qq < 2; ++qq) {
    int jw0, jw1, Jlo, Jhi, hy;
    tile_window_of(qfirst + qq, nx, ny, sx, sy, cy0, cy1, &jw0, &jw1, &Jlo, &Jhi, &hy);
    if (hy < y0 || hy >= y1) continue;
    const double *row = vis + (size_t)(hy - y0) * nx;
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < nx; x += gridDim.x * blockDim.x)
      dst[(size_t)qq * nx + x] = __ldcg(row + x);
  }
}

// this process's entry of the arg-min exchange: {h bits, push-order key, vg(end) bits or 0, 0}
__global__ void giant_pack_key_kernel(const int *ctl, const Best *bests, const GiantLocal loc, int nx,
                                      unsigned long long *out) {
  if (ctl[0]) return;
  const int ex = ctl[6], ey = ctl[7];
  Best b{~0ull, ~0ull};
  for (int s = 0; s < loc.n; ++s)
    if (better(bests[s], b)) b = bests[s];
  unsigned long long vge = 0ull; // vg >= 0: its bit pattern is monotonic, the owner's value is the max
  if (const GiantStripDev *s = strip_of_row(loc, ey))
    vge = (unsigned long long)__double_as_longlong(__ldcg(s->vg + (size_t)(ey - s->y0) * nx + ex));
  out[0] = b.h; out[1] = b.key; out[2] = vge; out[3] = 0ull;
}

// heap_->top() over all ranks = lexicographic minimum of (h bits, push order) -- an all-reduce
// (min) of h alone would lose the first-pushed tie-break -- then the loop
// control of solve() (:127-140), identically on every rank.
__global__ void giant_step_kernel(const unsigned long long *all, int world, int *ctl, int32_t *ls,
                                  cudaGraphConditionalHandle cond, int use_cond) {
  if (ctl[0]) {
    if (use_cond) cudaGraphSetConditional(cond, 0);
    return;
  }
  const double thr = *reinterpret_cast<const double *>(ctl + 8);
  const int max_iter = ctl[10];
  Best b{~0ull, ~0ull};
  unsigned long long vge = 0ull;
  for (int r = 0; r < world; ++r) {
    const Best o{all[4 * r], all[4 * r + 1]};
    if (better(o, b)) b = o;
    vge = all[4 * r + 2] > vge ? all[4 * r + 2] : vge;
  }
  int tx = ctl[1], ty = ctl[2], d = 1, st = ctl[3], nnb = ctl[4];
  if (b.h == ~0ull) {
    st = VHP_MAX_ITER; // no candidate at all (the reference would read an empty heap)
  } else {
    nnb = planner_next_source(b, ctl[1], ctl[2], ctl[4], max_iter, thr,
                              __longlong_as_double((long long)vge), ls, tx, ty, d, st);
  }
  ctl[1] = tx; ctl[2] = ty; ctl[3] = st; ctl[4] = nnb; ctl[5] += 1;
  __threadfence();
  ctl[0] = d;
  if (use_cond) cudaGraphSetConditional(cond, d ? 0 : 1);
}

// cameFrom_ at lightSources_[0 .. nb-1] and at the end point (entry nb), INT_MIN where this
// process does not own the row: an all-reduce (max) over the ranks completes the table.
__global__ void giant_came_at_kernel(const int *ctl, const int32_t *ls, int nx, const GiantLocal loc,
                                     int cap, int32_t *came_at) {
  const int nb = ctl[4], ok = ctl[3] == VHP_OK, ex = ctl[6], ey = ctl[7];
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < cap; t += gridDim.x * blockDim.x) {
    int v = INT_MIN;
    if (ok && t <= nb) {
      const int x = t < nb ? ls[2 * t] : ex, y = t < nb ? ls[2 * t + 1] : ey;
      if (const GiantStripDev *s = strip_of_row(loc, y)) v = __ldcg(s->came + (size_t)(y - s->y0) * nx + x);
    }
    came_at[t] = v;
  }
}

// tail of solve(): lightSources_[nb] = end (:141), reconstructPath (:1183-1213)
__global__ void giant_finish_kernel(const int *ctl, const int32_t *came_at, int ls_cap, int32_t *ls,
                                    int32_t *status, int32_t *nb_out, double *path_len,
                                    int32_t *path_n, int32_t *path, int32_t *iters) {
  const int st = ctl[3], nb = ctl[4], ex = ctl[6], ey = ctl[7];
  *status = st;
  *nb_out = nb;
  *iters = ctl[5];
  double total;
  const long n = planner_reconstruct_with(st, nb, ex, ey, ls_cap, ls, path, total,
                                          [&](int t, int, int) { return came_at[t]; });
  *path_n = (int32_t)n;
  *path_len = total;
}

} // namespace

cudaError_t vhp_launch_giant_reset(const GiantStripDev &s, int nx, cudaStream_t st, int64_t *launches) {
  giant_reset_kernel<<<148 * 8, 256, 0, st>>>(s.vis, s.vg, s.hc, s.came, (size_t)(s.y1 - s.y0) * nx);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

cudaError_t vhp_launch_giant_begin(const VhpTilePlanes &pl, int nx, int ny, const int32_t se[4],
                                   double thr, int max_iter, int32_t *d_ls, const GiantLocal &loc,
                                   int *d_ctl, cudaStream_t st, int64_t *launches) {
  giant_begin_kernel<<<1, 1, 0, st>>>(pl.rowF, pl.wx, nx, ny, se[0], se[1], se[2], se[3], thr, max_iter,
                                      d_ls, loc, d_ctl);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

cudaError_t vhp_launch_giant_halo_gather(const int *d_ctl, int nx, int ny, const GiantStripDev &src,
                                         int cy0, int cy1, int qfirst, double *d_dst, cudaStream_t st,
                                         int64_t *launches) {
  const int blocks = (nx + 255) / 256 < 64 ? (nx + 255) / 256 : 64;
  giant_halo_gather_kernel<<<blocks, 256, 0, st>>>(d_ctl, nx, ny, src.vis, src.y0, src.y1, cy0, cy1,
                                                   qfirst, d_dst);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

cudaError_t vhp_launch_giant_pack_key(const int *d_ctl, const unsigned long long *d_bests,
                                      const GiantLocal &loc, int nx, unsigned long long *d_out,
                                      cudaStream_t st, int64_t *launches) {
  giant_pack_key_kernel<<<1, 1, 0, st>>>(d_ctl, reinterpret_cast<const Best *>(d_bests), loc, nx, d_out);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

cudaError_t vhp_launch_giant_step(const unsigned long long *d_all, int world, int *d_ctl, int32_t *d_ls,
                                  unsigned long long cond_handle, int use_cond, cudaStream_t st,
                                  int64_t *launches) {
  giant_step_kernel<<<1, 1, 0, st>>>(d_all, world, d_ctl, d_ls, (cudaGraphConditionalHandle)cond_handle,
                                     use_cond);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

cudaError_t vhp_launch_giant_came_at(const int *d_ctl, const int32_t *d_ls, int nx,
                                     const GiantLocal &loc, int cap, int32_t *d_came_at,
                                     cudaStream_t st, int64_t *launches) {
  giant_came_at_kernel<<<(cap + 255) / 256, 256, 0, st>>>(d_ctl, d_ls, nx, loc, cap, d_came_at);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

cudaError_t vhp_launch_giant_finish(const int *d_ctl, const int32_t *d_came_at, int ls_cap,
                                    int32_t *d_ls, int32_t *d_status, int32_t *d_nb,
                                    double *d_path_len, int32_t *d_path_n, int32_t *d_path,
                                    int32_t *d_iters, cudaStream_t st, int64_t *launches) {
  giant_finish_kernel<<<1, 1, 0, st>>>(d_ctl, d_came_at, ls_cap, d_ls, d_status, d_nb, d_path_len,
                                       d_path_n, d_path, d_iters);
  if (launches) *launches += 1;
  return cudaGetLastError();
}
