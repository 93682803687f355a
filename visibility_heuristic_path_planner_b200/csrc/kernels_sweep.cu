// kernels_sweep.cu -- K1, straightforward form: the in-library cross-check of the tile
// kernel and the fallback for grids whose tile boundaries do not fit shared memory.
//
// Replaces visibilityBasedSolver::computeVisibility
// (reference src/visibilityBasedSolver.cpp:570-696).  The reference walks each of
// the four quadrants around the light source cell by cell (i outer, j inner).
// Here the sweep is the "L-front" dynamic program: all cells at
// Chebyshev distance k from the source depend only on cells at distance k-1, so
// a front advances one ring per step: one CTA per (pair, quadrant), fronts in
// global scratch, byte occupancy, IEEE division, direct stores.
//
// Arithmetic contract (bit parity with the strict-IEEE reference build):
//   v = a - c*(a - b), c = min(i,j)/max(i,j), each operation rounded once
//   (__dsub_rn/__dmul_rn, never contracted), then v*occ with occ in {0,1}.
#include <algorithm>
#include <cstdint>

#include "vhp_internal.h"
#include "sweep_common.cuh"

namespace {

// CTA = (pair, quadrant).  Quadrant-local cell (i,j) <-> (sx + dx*i, sy + dy*j).
// Fronts col[t] = q[k][t] and row[t] = q[t][k] live in global scratch.  The
// quadrants are extended by one cell to the X=0 / Y=0 border with occupancy
// forced to 0 there, which yields the zeros the reference leaves in cells it
// never visits (loop bounds :434-438, :478-483, :522-527).
template <typename OutT>
__global__ void __launch_bounds__(256)
sweep_naive_kernel(const uint8_t *__restrict__ occ, int nx, int ny,
                   const int32_t *__restrict__ src_xy, const int32_t *__restrict__ src_map,
                   OutT *__restrict__ out_all, double *__restrict__ scratch, int M,
                   int *__restrict__ err) {
  const int q = blockIdx.x & 3;
  const int64_t pair = blockIdx.x >> 2;
  const int sx = src_xy[2 * pair], sy = src_xy[2 * pair + 1];
  if ((unsigned)sx >= (unsigned)nx || (unsigned)sy >= (unsigned)ny) { // CTA-uniform
    if (threadIdx.x == 0) atomicOr(err, 1);
    return;
  }
  const int map = src_map ? src_map[pair] : 0;
  const uint8_t *o = occ + (size_t)map * nx * ny;
  OutT *out = out_all + (size_t)pair * nx * ny;
  const int dx = (q == 0 || q == 3) ? 1 : -1;
  const int dy = (q < 2) ? 1 : -1;
  const int Ex = dx > 0 ? nx - sx : sx + 1;
  const int Ey = dy > 0 ? ny - sy : sy + 1;
  // shared axes are stored by exactly one quadrant (values are identical)
  const bool store_i0 = (q == 0 || q == 3); // i == 0 column: +y axis by Q1, -y axis by Q4
  const bool store_j0 = (q == 0 || q == 1); // j == 0 row:    +x axis by Q1, -x axis by Q2
  double *colc = scratch + (size_t)blockIdx.x * 4 * M;
  double *coln = colc + M, *rowc = coln + M, *rown = rowc + M;

  auto occ_eff = [&](int X, int Y) -> double {
    if ((X == 0 && sx > 0) || (Y == 0 && sy > 0)) return 0.0;
    return o[(size_t)Y * nx + X] != 0 ? 1.0 : 0.0;
  };

  if (threadIdx.x == 0) {
    const double s0 = __dmul_rn(1.0, occ_eff(sx, sy));
    colc[0] = s0;
    rowc[0] = s0;
    if (q == 0) out[(size_t)sy * nx + sx] = to_out<OutT>(s0);
  }
  __syncthreads();
  const int K = Ex > Ey ? Ex : Ey;
  for (int k = 1; k < K; ++k) {
    const double fk = (double)k;
    for (int t = threadIdx.x; t < k; t += blockDim.x) {
      const double c = __ddiv_rn((double)t, fk);
      if (k < Ex && t < Ey) { // column part: cell (k, t)
        const int X = sx + dx * k, Y = sy + dy * t;
        const double a = colc[t];
        double v = (t == 0) ? a : lerp_rn(a, colc[t - 1], c);
        v = __dmul_rn(v, occ_eff(X, Y));
        coln[t] = v;
        if (t > 0 || store_j0) out[(size_t)Y * nx + X] = to_out<OutT>(v);
      }
      if (k < Ey && t < Ex) { // row part: cell (t, k)
        const int X = sx + dx * t, Y = sy + dy * k;
        const double a = rowc[t];
        double v = (t == 0) ? a : lerp_rn(a, rowc[t - 1], c);
        v = __dmul_rn(v, occ_eff(X, Y));
        rown[t] = v;
        if (t > 0 || store_i0) out[(size_t)Y * nx + X] = to_out<OutT>(v);
      }
    }
    __syncthreads();
    // diagonal (reference quirk, no i==j branch): q[k][k] = q[k][k-1]*occ(k,k)
    if (threadIdx.x == 0 && k < Ex && k < Ey) {
      const int X = sx + dx * k, Y = sy + dy * k;
      const double d = __dmul_rn(coln[k - 1], occ_eff(X, Y));
      coln[k] = d;
      rown[k] = d;
      out[(size_t)Y * nx + X] = to_out<OutT>(d);
    }
    __syncthreads();
    double *t1 = colc; colc = coln; coln = t1;
    double *t2 = rowc; rowc = rown; rown = t2;
  }
}

} // namespace

// ------------------------------- launchers -----------------------------------
size_t vhp_sweep_naive_scratch_bytes(int nx, int ny, int64_t npairs) {
  const size_t M = (size_t)std::max(nx, ny) + 2;
  return (size_t)npairs * 4 * 4 * M * sizeof(double);
}

cudaError_t vhp_launch_sweep_naive(const uint8_t *d_occ, int nx, int ny, const int32_t *d_src_xy,
                                   const int32_t *d_src_map, int64_t npairs, vhp_dtype dtype,
                                   void *d_out, double *d_scratch, int *d_err, cudaStream_t st,
                                   int64_t *launches) {
  const int M = std::max(nx, ny) + 2;
  const unsigned grid = (unsigned)(npairs * 4);
  if (dtype == VHP_F32)
    sweep_naive_kernel<float><<<grid, 256, 0, st>>>(d_occ, nx, ny, d_src_xy, d_src_map,
                                                    (float *)d_out, d_scratch, M, d_err);
  else
    sweep_naive_kernel<double><<<grid, 256, 0, st>>>(d_occ, nx, ny, d_src_xy, d_src_map,
                                                     (double *)d_out, d_scratch, M, d_err);
  if (launches) *launches += 1;
  return cudaGetLastError();
}
