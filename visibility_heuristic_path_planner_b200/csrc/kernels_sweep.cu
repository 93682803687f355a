// kernels_sweep.cu -- K1: batched stand-alone visibility sweep for sm_100a.
//
// Replaces visibilityBasedSolver::computeVisibility
// (reference src/visibilityBasedSolver.cpp:570-696).  The reference walks each of
// the four quadrants around the light source cell by cell (i outer, j inner).
// Here the sweep is the "L-front" dynamic program: all cells at
// Chebyshev distance k from the source depend only on cells at distance k-1, so
// a front advances one ring per step.
//
// Two kernels:
//   sweep_naive_kernel  one CTA per (pair, quadrant), fronts in global scratch,
//                       byte occupancy, direct stores.  Simple; used as the
//                       in-library cross-check and for maps the front kernel does
//                       not cover.
//   sweep_front_kernel  the tuned kernel (see the block comment above it).
//
// Arithmetic contract (bit parity with the strict-IEEE reference build):
//   v = a - c*(a - b), c = min(i,j)/max(i,j), each operation rounded once
//   (__dsub_rn/__dmul_rn, never contracted), then v*occ with occ in {0,1}.
#include <cstdint>

#include "vhp_internal.h"
#include "sweep_common.cuh"

namespace {

// ---------------------------------------------------------------------------
// map packing: uint8 occupancy -> row-major and column-major bit planes
// ---------------------------------------------------------------------------
__global__ void pack_rows_kernel(const uint8_t *__restrict__ occ, int nmaps, int nx, int ny,
                                 uint32_t *__restrict__ rowbits, int wpr) {
  const size_t total = (size_t)nmaps * ny * wpr;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int w = (int)(idx % wpr);
    const size_t my = idx / wpr; // m*ny + y
    const uint8_t *row = occ + my * nx;
    const int x0 = w * 32;
    uint32_t bits = 0;
#pragma unroll 8
    for (int b = 0; b < 32; ++b) {
      const int x = x0 + b;
      if (x < nx && row[x] != 0) bits |= 1u << b;
    }
    rowbits[idx] = bits;
  }
}

__global__ void pack_cols_kernel(const uint8_t *__restrict__ occ, int nmaps, int nx, int ny,
                                 uint32_t *__restrict__ colbits, int wpc) {
  // thread index runs over x fastest so the strided byte reads coalesce
  const size_t total = (size_t)nmaps * wpc * nx;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(idx % nx);
    const size_t mw = idx / nx;
    const int w = (int)(mw % wpc);
    const size_t m = mw / wpc;
    const uint8_t *base = occ + m * (size_t)nx * ny + x;
    const int y0 = w * 32;
    uint32_t bits = 0;
#pragma unroll 8
    for (int b = 0; b < 32; ++b) {
      const int y = y0 + b;
      if (y < ny && base[(size_t)y * nx] != 0) bits |= 1u << b;
    }
    colbits[(m * nx + x) * wpc + w] = bits;
  }
}

// diagnostic: count (i, k) pairs, 0 <= i < k <= kmax, where ratio_rn != IEEE i/k
__global__ void ratio_selftest_kernel(const double *__restrict__ rcp, int kmax,
                                      unsigned long long *mismatches) {
  const int k = blockIdx.x + 1;
  if (k > kmax) return;
  const double fk = (double)k, r = rcp[2 * k];
  unsigned long long bad = 0;
  for (int i = threadIdx.x; i < k; i += blockDim.x)
    if (ratio_rn((double)i, fk, r) != __ddiv_rn((double)i, fk)) ++bad;
  if (bad) atomicAdd(mismatches, bad);
}

// table[k] = {RN(1/k), (double)k}
__global__ void rcp_table_kernel(double *table, int len) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < len) {
    table[2 * k] = (k == 0) ? 0.0 : __drcp_rn((double)k);
    table[2 * k + 1] = (double)k;
  }
}

// ---------------------------------------------------------------------------
// naive L-front kernel
// ---------------------------------------------------------------------------
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
cudaError_t vhp_launch_pack_maps(const uint8_t *d_occ, int nmaps, int nx, int ny,
                                 uint32_t *d_rowbits, uint32_t *d_colbits, int wpr, int wpc,
                                 cudaStream_t st, int64_t *launches) {
  const size_t tr = (size_t)nmaps * ny * wpr, tc = (size_t)nmaps * nx * wpc;
  const int bs = 256;
  const unsigned gr = (unsigned)std::min<size_t>((tr + bs - 1) / bs, 148u * 32u);
  const unsigned gc = (unsigned)std::min<size_t>((tc + bs - 1) / bs, 148u * 32u);
  pack_rows_kernel<<<gr, bs, 0, st>>>(d_occ, nmaps, nx, ny, d_rowbits, wpr);
  pack_cols_kernel<<<gc, bs, 0, st>>>(d_occ, nmaps, nx, ny, d_colbits, wpc);
  if (launches) *launches += 2;
  return cudaGetLastError();
}

cudaError_t vhp_launch_rcp_table(double *d_table, int len, cudaStream_t st, int64_t *launches) {
  rcp_table_kernel<<<(len + 255) / 256, 256, 0, st>>>(d_table, len);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

cudaError_t vhp_launch_ratio_selftest(const double *d_rcp, int kmax,
                                      unsigned long long *d_mismatches, cudaStream_t st,
                                      int64_t *launches) {
  ratio_selftest_kernel<<<kmax, 128, 0, st>>>(d_rcp, kmax, d_mismatches);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

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
