// kernels_sweep_variant.cu -- opt-in variants of the visibility sweep:
//
//   model 1  getAccessibilityMap.m (reference MATLAB_code/visibility/getAccessibilityMap.m:1-118,
//            the paper's Algorithm 1): decay `alpha` per cell, curve factor `fac` (octant boundary
//            i == j*fac, c = (j*fac)/i resp. i/(j*fac)), diagonal from (i-1, j-1), every cell of the
//            grid computed (MATLAB's 1-based loops have no never-written border);
//   model 2  computeVisibilityUsingQueue() (reference src/visibilityBasedSolver.cpp:701-893) as the
//            order-free rule stated in include/vhp.h (VHP_VARIANT_QUEUE): early termination -- a
//            free cell is computed only if a cell that pushes it holds more than `cutoff` (0.001)
//            -- diagonal from (i-1, j-1), source = lightStrength_ whatever its occupancy.
//
// With fac != 1 a cell of the row octant may depend on its neighbour in the SAME L-front, so
// these variants advance by anti-diagonals d = i + j (every dependency, value or pusher, has a
// smaller d): one CTA per (pair, quadrant), a block barrier per diagonal, the fp64 field itself
// (global memory, L2-resident for one quadrant) as the only state.  Every quadrant computes its
// own two axes (identical values from identical inputs), so the four CTAs of a pair never read
// each other's cells.  These are secondary modes: correct and parallel, not tuned -- the
// default sweep (kernels_sweep_tile.cu) is the throughput path.
//
// Arithmetic: IEEE binary64, one rounding per operation, in the order the sources write it.
#include <algorithm>
#include <cstdint>

#include "vhp_internal.h"
#include "sweep_common.cuh"

namespace {

struct VariantArgs {
  const uint8_t *occ;
  int nx, ny;
  const int32_t *src_xy, *src_map;
  double *buf; // [pair][ny][nx], zero-filled
  int model;
  double alpha, fac, ls, cutoff;
  int *err;
};

__global__ void __launch_bounds__(256) sweep_variant_kernel(const VariantArgs p) {
  const int q = blockIdx.x & 3;
  const int64_t pair = blockIdx.x >> 2;
  const int nx = p.nx, ny = p.ny;
  const int sx = p.src_xy[2 * pair], sy = p.src_xy[2 * pair + 1];
  if ((unsigned)sx >= (unsigned)nx || (unsigned)sy >= (unsigned)ny) { // CTA-uniform
    if (threadIdx.x == 0) atomicOr(p.err, 1);
    return;
  }
  const int map = p.src_map ? p.src_map[pair] : 0;
  const uint8_t *occ = p.occ + (size_t)map * nx * ny;
  double *vis = p.buf + (size_t)pair * nx * ny;
  const int dx = (q == 0 || q == 3) ? 1 : -1, dy = (q < 2) ? 1 : -1;
  const int Ex = dx > 0 ? nx - 1 - sx : sx, Ey = dy > 0 ? ny - 1 - sy : sy;
  const bool matlab = p.model == 1;
  auto at = [&](int i, int j) -> double * { return vis + (size_t)(sx + dx * i) + (size_t)(sy + dy * j) * nx; };
  auto free_cell = [&](int i, int j) { return occ[(size_t)(sx + dx * i) + (size_t)(sy + dy * j) * nx] != 0; };

  if (threadIdx.x == 0) {
    // model 1: lightStrength * obstacle (.m:17-18, :34); model 2: lightStrength_ unconditionally (:707)
    *at(0, 0) = matlab ? __dmul_rn(p.ls, free_cell(0, 0) ? 1.0 : 0.0) : 1.0;
  }
  __syncthreads();
  for (int d = 1; d <= Ex + Ey; ++d) {
    const int ilo = d > Ey ? d - Ey : 0, ihi = d < Ex ? d : Ex;
    for (int i = ilo + (int)threadIdx.x; i <= ihi; i += blockDim.x) {
      const int j = d - i;
      const bool fr = free_cell(i, j);
      double v;
      if (matlab) {
        const double jf = __dmul_rn((double)j, p.fac);
        if (i == 0) {
          v = __dmul_rn(p.alpha, __ldcg(at(0, j - 1)));
        } else if (j == 0) {
          v = __dmul_rn(p.alpha, __ldcg(at(i - 1, 0)));
        } else if ((double)i == jf) {
          v = __dmul_rn(p.alpha, __ldcg(at(i - 1, j - 1)));
        } else if ((double)i > jf) {
          const double c = __ddiv_rn(jf, (double)i);
          v = __dmul_rn(p.alpha, lerp_rn(__ldcg(at(i - 1, j)), __ldcg(at(i - 1, j - 1)), c));
        } else {
          const double c = __ddiv_rn((double)i, jf);
          v = __dmul_rn(p.alpha, lerp_rn(__ldcg(at(i, j - 1)), __ldcg(at(i - 1, j - 1)), c));
        }
        *at(i, j) = fr ? v : __dmul_rn(v, 0.0);
      } else {
        if (!fr) continue; // occupied cells are never visited (:732-734): they keep 0
        bool pushed = i <= 1 && j <= 1; // the eight neighbours of the source (:709-716)
        if (!pushed && i - 1 >= 1) pushed = __ldcg(at(i - 1, j)) > p.cutoff;
        if (!pushed && j - 1 >= 1) pushed = __ldcg(at(i, j - 1)) > p.cutoff;
        if (!pushed && i == j) pushed = __ldcg(at(i - 1, j - 1)) > p.cutoff;
        if (!pushed) continue;
        if (i == 0) v = __ldcg(at(0, j - 1));
        else if (j == 0) v = __ldcg(at(i - 1, 0));
        else if (i == j) v = __ldcg(at(i - 1, j - 1));
        else if (i > j) v = lerp_rn(__ldcg(at(i - 1, j)), __ldcg(at(i - 1, j - 1)), __ddiv_rn((double)j, (double)i));
        else v = lerp_rn(__ldcg(at(i, j - 1)), __ldcg(at(i - 1, j - 1)), __ddiv_rn((double)i, (double)j));
        *at(i, j) = v;
      }
    }
    __syncthreads();
  }
}

__global__ void variant_to_f32_kernel(const double *__restrict__ in, float *__restrict__ out, size_t n) {
  for (size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x; c < n; c += (size_t)gridDim.x * blockDim.x)
    out[c] = __double2float_rn(in[c]);
}

} // namespace

// d_buf: fp64 field of every pair (the output itself for VHP_F64); it is zero-filled here.
cudaError_t vhp_launch_sweep_variant(const uint8_t *d_occ, int nx, int ny, const int32_t *d_src_xy,
                                     const int32_t *d_src_map, int64_t npairs, int model, double alpha,
                                     double fac, double light_strength, double cutoff, double *d_buf,
                                     float *d_out32, int *d_err, cudaStream_t st, int64_t *launches) {
  const size_t n = (size_t)npairs * nx * ny;
  cudaError_t e = cudaMemsetAsync(d_buf, 0, n * sizeof(double), st);
  if (e != cudaSuccess) return e;
  VariantArgs p;
  p.occ = d_occ; p.nx = nx; p.ny = ny; p.src_xy = d_src_xy; p.src_map = d_src_map; p.buf = d_buf;
  p.model = model; p.alpha = alpha; p.fac = fac; p.ls = light_strength; p.cutoff = cutoff; p.err = d_err;
  sweep_variant_kernel<<<(unsigned)(npairs * 4), 256, 0, st>>>(p);
  if (launches) *launches += 1;
  if (d_out32) {
    variant_to_f32_kernel<<<148 * 8, 256, 0, st>>>(d_buf, d_out32, n);
    if (launches) *launches += 1;
  }
  return cudaGetLastError();
}
