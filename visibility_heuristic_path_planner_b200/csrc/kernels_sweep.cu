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

namespace {

__device__ __forceinline__ double lerp_rn(double a, double b, double c) {
  return __dsub_rn(a, __dmul_rn(c, __dsub_rn(a, b)));
}

// Correctly rounded i/k for integers 0 <= i < k <= 16384 from r = RN(1/k):
// one Newton/Markstein correction.  Verified exhaustively against IEEE division
// (tests/test_gpu_sweep.py::test_ratio_exact_exhaustive and the host-side check
// described in DESIGN.md).
__device__ __forceinline__ double ratio_rn(double fi, double fk, double r) {
  const double q0 = __dmul_rn(fi, r);
  const double rem = __fma_rn(-q0, fk, fi);
  return __fma_rn(rem, r, q0);
}

template <typename OutT> __device__ __forceinline__ OutT to_out(double v);
template <> __device__ __forceinline__ float to_out<float>(double v) { return __double2float_rn(v); }
template <> __device__ __forceinline__ double to_out<double>(double v) { return v; }

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
  const double fk = (double)k, r = rcp[k];
  unsigned long long bad = 0;
  for (int i = threadIdx.x; i < k; i += blockDim.x)
    if (ratio_rn((double)i, fk, r) != __ddiv_rn((double)i, fk)) ++bad;
  if (bad) atomicAdd(mismatches, bad);
}

__global__ void rcp_table_kernel(double *table, int len) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < len) table[k] = (k == 0) ? 0.0 : __drcp_rn((double)k);
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

// ---------------------------------------------------------------------------
// tuned front kernel
// ---------------------------------------------------------------------------
// One CTA per (map, source) pair.  Thread `tid` owns the four ABSOLUTE
// coordinates u = 4*tid + e (e = 0..3), used both as x for the two "row fronts"
// and as y for the two "column fronts":
//     RU: cells (x, sy+k)   RD: cells (x, sy-k)     |x-sx| < k   (row parts)
//     CR: cells (sx+k, y)   CL: cells (sx-k, y)     |y-sy| < k   (column parts)
// Each of the 16 front values a thread owns stays in an fp64 register for the
// whole sweep; the upstream neighbour b comes from the adjacent register, a warp
// shuffle at thread boundaries and a double-buffered shared-memory slot at warp
// boundaries.  c = i/k is one table read of RN(1/k) plus three fp64 ops and is
// shared by the two fronts of a type.  Occupancy is read as one 32-bit word per
// front per step from the bit planes (prefetched one step ahead).
//
// Stores: row fronts write 4 consecutive x per thread (one 128-bit store when
// the row pitch allows it).  Column fronts produce one x per step, so each warp
// stages S = 32 B / sizeof(OutT) steps of its 128 rows in shared memory
// (tile[kk][y], pitch 132 -> conflict-free both ways) and flushes sector-sized
// row segments aligned on absolute X.
//
// The diagonal cell (k,k) of a quadrant equals q[k][k-1]*occ(k,k) (reference
// quirk).  q[k][k-1] is a column-front value: the column owner of |y-sy| == k
// derives it from its own neighbour register, the row owner of |x-sx| == k gets
// it through a 4-entry shared slot and also stores the diagonal cell.
constexpr int kT = 4;                 // coordinates per thread
constexpr int kWSpan = 32 * kT;       // coordinates per warp
constexpr int kPitch = kWSpan + 4;    // staging tile pitch (elements)
constexpr int kBig = 0x3fffffff;

struct FrontParams {
  const uint32_t *rowbits, *colbits;
  int wpr, wpc;
  size_t row_plane, col_plane;
  int nx, ny;
  const int32_t *src_xy, *src_map;
  void *out;
  const double *rcp;
  int vec_ok;
  int *err;
};

template <typename OutT> struct Vec4;
template <> struct Vec4<float> {
  static __device__ __forceinline__ void store(float *p, double a, double b, double c, double d) {
    __stcs(reinterpret_cast<float4 *>(p),
           make_float4(__double2float_rn(a), __double2float_rn(b), __double2float_rn(c),
                       __double2float_rn(d)));
  }
  static __device__ __forceinline__ void store_shared(float *p, double a, double b, double c,
                                                      double d) {
    *reinterpret_cast<float4 *>(p) = make_float4(__double2float_rn(a), __double2float_rn(b),
                                                 __double2float_rn(c), __double2float_rn(d));
  }
};
template <> struct Vec4<double> {
  static __device__ __forceinline__ void store(double *p, double a, double b, double c, double d) {
    __stcs(reinterpret_cast<double2 *>(p), make_double2(a, b));
    __stcs(reinterpret_cast<double2 *>(p) + 1, make_double2(c, d));
  }
  static __device__ __forceinline__ void store_shared(double *p, double a, double b, double c,
                                                      double d) {
    reinterpret_cast<double2 *>(p)[0] = make_double2(a, b);
    reinterpret_cast<double2 *>(p)[1] = make_double2(c, d);
  }
};

// all four elements active, none of them new on the front, thread entirely on one
// side of the source: b comes from the neighbour towards the source
template <bool PLUS>
__device__ __forceinline__ void full_update(double (&F)[4], double nb, const double (&c)[4],
                                            uint32_t nib) {
  double b0, b1, b2, b3;
  if (PLUS) { b0 = nb; b1 = F[0]; b2 = F[1]; b3 = F[2]; }
  else      { b0 = F[1]; b1 = F[2]; b2 = F[3]; b3 = nb; }
  const double v0 = lerp_rn(F[0], b0, c[0]);
  const double v1 = lerp_rn(F[1], b1, c[1]);
  const double v2 = lerp_rn(F[2], b2, c[2]);
  const double v3 = lerp_rn(F[3], b3, c[3]);
  F[0] = (nib & 1u) ? v0 : 0.0;
  F[1] = (nib & 2u) ? v1 : 0.0;
  F[2] = (nib & 4u) ? v2 : 0.0;
  F[3] = (nib & 8u) ? v3 : 0.0;
}

// flush one staged block of a column front: rows [wy0, wy0+128) x X in [Xb, Xb+S)
template <typename OutT, int DIR>
__device__ __forceinline__ void flush_block(const OutT *tile, OutT *out, int nx, int ny, int sx,
                                            int sy, int k, int X, int wy0, int lane, int vec_ok) {
  constexpr int S = 32 / (int)sizeof(OutT);
  constexpr int V = 16 / (int)sizeof(OutT);
  __syncwarp();
  const int Xb = X & ~(S - 1);
  const int h = lane & 1, r = lane >> 1;
  const int Xc = Xb + V * h;
#pragma unroll
  for (int pass = 0; pass < kWSpan / 16; ++pass) {
    const int yl = pass * 16 + r;
    const int y = wy0 + yl;
    const int j = y > sy ? y - sy : sy - y;
    OutT vals[V];
#pragma unroll
    for (int m = 0; m < V; ++m) vals[m] = tile[(V * h + m) * kPitch + yl];
    if (y < ny) {
      bool all_ok, any_ok;
      if (DIR > 0) {
        all_ok = (j < Xc - sx) && (Xc + V - 1 <= sx + k);
        any_ok = (j < Xc + V - 1 - sx) && (Xc <= sx + k);
      } else {
        all_ok = (j < sx - (Xc + V - 1)) && (Xc >= sx - k);
        any_ok = (j < sx - Xc) && (Xc + V - 1 >= sx - k);
      }
      OutT *dst = out + (size_t)y * nx + Xc;
      if (all_ok && vec_ok) {
        if (sizeof(OutT) == 4)
          __stcs(reinterpret_cast<float4 *>(dst),
                 make_float4((float)vals[0], (float)vals[1], (float)vals[V > 2 ? 2 : 0],
                             (float)vals[V > 3 ? 3 : 0]));
        else
          __stcs(reinterpret_cast<double2 *>(dst), make_double2((double)vals[0], (double)vals[1]));
      } else if (any_ok) {
#pragma unroll
        for (int m = 0; m < V; ++m) {
          const int Xm = Xc + m;
          const bool ok = DIR > 0 ? (j < Xm - sx && Xm <= sx + k) : (j < sx - Xm && Xm >= sx - k);
          if (ok) __stcs(dst + m, vals[m]);
        }
      }
    }
  }
  __syncwarp();
}

template <typename OutT, int MAXNT, int MINB>
__global__ void __launch_bounds__(MAXNT, MINB) sweep_front_kernel(const FrontParams p) {
  constexpr int S = 32 / (int)sizeof(OutT);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NW = blockDim.x >> 5;
  const int64_t pair = blockIdx.x;
  const int nx = p.nx, ny = p.ny;
  const int sx = __ldg(p.src_xy + 2 * pair), sy = __ldg(p.src_xy + 2 * pair + 1);
  if ((unsigned)sx >= (unsigned)nx || (unsigned)sy >= (unsigned)ny) { // CTA-uniform
    if (tid == 0) atomicOr(p.err, 1);
    return;
  }
  const int map = p.src_map ? __ldg(p.src_map + pair) : 0;
  const uint32_t *__restrict__ rowbits = p.rowbits + (size_t)map * p.row_plane;
  const uint32_t *__restrict__ colbits = p.colbits + (size_t)map * p.col_plane;
  OutT *__restrict__ out = reinterpret_cast<OutT *>(p.out) + (size_t)pair * nx * ny;
  const int vec_ok = p.vec_ok;

  // shared memory carve-up
  OutT *tiles = reinterpret_cast<OutT *>(smem_raw);                 // [2][NW][S][kPitch]
  const size_t tile_elems = (size_t)S * kPitch;
  double *edge = reinterpret_cast<double *>(tiles + 2 * (size_t)NW * tile_elems); // [2][4][NW][2]
  double *slot = edge + 2 * 4 * NW * 2;                              // [2][4]
  OutT *tileR = tiles + (size_t)warp * tile_elems;
  OutT *tileL = tiles + ((size_t)NW + warp) * tile_elems;
  for (int i = tid; i < 2 * 4 * NW * 2 + 8; i += blockDim.x) edge[i] = 0.0;

  const int u0 = kT * tid;
  const int wu0 = kWSpan * warp;
  const int widx = tid >> 3;            // 32-bit word holding this thread's nibble
  const int nsh = 4 * (tid & 7);
  // border forcing: X == 0 (sx > 0) and Y == 0 (sy > 0) stay dark
  const uint32_t keep_x = (tid == 0 && sx > 0) ? 0xEu : 0xFu;
  const uint32_t keep_y = (tid == 0 && sy > 0) ? 0xEu : 0xFu;

  // per-thread geometry for both front types (x: row fronts, y: column fronts)
  auto geom = [&](int s, int n, int &imin, int &ifull, int &side, double &f0) {
    const int lo = u0, hi = u0 + kT - 1;
    const bool allvalid = hi < n;
    if (lo > s) side = 1; else if (hi < s) side = -1; else side = 0;
    imin = kBig;
    int imax = 0;
#pragma unroll
    for (int e = 0; e < kT; ++e) {
      const int u = u0 + e;
      const int i = u > s ? u - s : s - u;
      if (u < n) { imin = min(imin, i); imax = max(imax, i); }
    }
    ifull = (allvalid && side != 0) ? imax + 1 : kBig;
    f0 = (double)(lo > s ? lo - s : s - lo);
  };
  int imin_x, ifull_x, side_x, imin_y, ifull_y, side_y;
  double f0x, f0y;
  geom(sx, nx, imin_x, ifull_x, side_x, f0x);
  geom(sy, ny, imin_y, ifull_y, side_y, f0y);
  const int wmin_x = __reduce_min_sync(0xffffffffu, imin_x);
  const int wmin_y = __reduce_min_sync(0xffffffffu, imin_y);
  const bool wplus_x = wu0 + kWSpan - 1 > sx, wminus_x = wu0 < sx;
  const bool wplus_y = wu0 + kWSpan - 1 > sy, wminus_y = wu0 < sy;

  double RU[4] = {0, 0, 0, 0}, RD[4] = {0, 0, 0, 0}, CR[4] = {0, 0, 0, 0}, CL[4] = {0, 0, 0, 0};

  // source cell
  {
    const uint32_t w = __ldg(rowbits + (size_t)sy * p.wpr + (sx >> 5));
    const double s0 = ((w >> (sx & 31)) & 1u) ? 1.0 : 0.0;
    if ((unsigned)(sx - u0) < (unsigned)kT) {
      const int e = sx - u0;
#pragma unroll
      for (int q = 0; q < kT; ++q)
        if (q == e) { RU[q] = s0; RD[q] = s0; }
      out[(size_t)sy * nx + sx] = to_out<OutT>(s0);
    }
    if ((unsigned)(sy - u0) < (unsigned)kT) {
      const int e = sy - u0;
#pragma unroll
      for (int q = 0; q < kT; ++q)
        if (q == e) { CR[q] = s0; CL[q] = s0; }
    }
  }

  auto nib_row = [&](int Y) -> uint32_t {
    if (Y < 0 || Y >= ny || (Y == 0 && sy > 0)) return 0u;
    return (__ldg(rowbits + (size_t)Y * p.wpr + widx) >> nsh) & keep_x;
  };
  auto nib_col = [&](int X) -> uint32_t {
    if (X < 0 || X >= nx || (X == 0 && sx > 0)) return 0u;
    return (__ldg(colbits + (size_t)X * p.wpc + widx) >> nsh) & keep_y;
  };
  // nibbles for the current step (n*), the previous step (p*) and the next (x*)
  uint32_t nRU = nib_row(sy + 1) & 0xFu, nRD = nib_row(sy - 1) & 0xFu;
  uint32_t nCR = nib_col(sx + 1) & 0xFu, nCL = nib_col(sx - 1) & 0xFu;
  uint32_t pRU = 0, pRD = 0, pCR = 0, pCL = 0;

  const int Krow = max(sy, ny - 1 - sy), Kcol = max(sx, nx - 1 - sx);
  const int Kmax = max(Krow, Kcol);
  __syncthreads();

  for (int k = 1; k <= Kmax + 1; ++k) {
    const int par = k & 1;
    // prefetch next step's occupancy
    const uint32_t xRU = nib_row(sy + k + 1), xRD = nib_row(sy - k - 1);
    const uint32_t xCR = nib_col(sx + k + 1), xCL = nib_col(sx - k - 1);
    const double fk = (double)k;
    const double r = __ldg(p.rcp + k);
    double *edge_rd = edge + (size_t)(par ^ 1) * 4 * NW * 2;
    double *edge_wr = edge + (size_t)par * 4 * NW * 2;
    const double *slot_rd = slot + (par ^ 1) * 4;
    double *slot_wr = slot + par * 4;

    // ------------------------------ row fronts ------------------------------
    if (k > wmin_x && k <= Krow + 1) {
      const bool ru_on = sy + k < ny, rd_on = sy - k >= 0;
      double LU = 0, LD = 0, RUn = 0, RDn = 0;
      if (wplus_x) {
        LU = __shfl_up_sync(0xffffffffu, RU[3], 1);
        LD = __shfl_up_sync(0xffffffffu, RD[3], 1);
        if (lane == 0 && warp > 0) {
          LU = edge_rd[(0 * NW + warp - 1) * 2 + 1];
          LD = edge_rd[(1 * NW + warp - 1) * 2 + 1];
        }
      }
      if (wminus_x) {
        RUn = __shfl_down_sync(0xffffffffu, RU[0], 1);
        RDn = __shfl_down_sync(0xffffffffu, RD[0], 1);
        if (lane == 31 && warp < NW - 1) {
          RUn = edge_rd[(0 * NW + warp + 1) * 2 + 0];
          RDn = edge_rd[(1 * NW + warp + 1) * 2 + 0];
        }
      }
      if (k > imin_x) {
        if (k > ifull_x) {
          // ---- fast path
          double c[4];
          if (side_x > 0) {
            c[0] = ratio_rn(f0x, fk, r);       c[1] = ratio_rn(f0x + 1.0, fk, r);
            c[2] = ratio_rn(f0x + 2.0, fk, r); c[3] = ratio_rn(f0x + 3.0, fk, r);
            if (ru_on) full_update<true>(RU, LU, c, nRU);
            if (rd_on) full_update<true>(RD, LD, c, nRD);
          } else {
            c[0] = ratio_rn(f0x, fk, r);       c[1] = ratio_rn(f0x - 1.0, fk, r);
            c[2] = ratio_rn(f0x - 2.0, fk, r); c[3] = ratio_rn(f0x - 3.0, fk, r);
            if (ru_on) full_update<false>(RU, RUn, c, nRU);
            if (rd_on) full_update<false>(RD, RDn, c, nRD);
          }
          if (ru_on) {
            OutT *dst = out + (size_t)(sy + k) * nx + u0;
            if (vec_ok) Vec4<OutT>::store(dst, RU[0], RU[1], RU[2], RU[3]);
            else {
#pragma unroll
              for (int e = 0; e < kT; ++e) __stcs(dst + e, to_out<OutT>(RU[e]));
            }
          }
          if (rd_on) {
            OutT *dst = out + (size_t)(sy - k) * nx + u0;
            if (vec_ok) Vec4<OutT>::store(dst, RD[0], RD[1], RD[2], RD[3]);
            else {
#pragma unroll
              for (int e = 0; e < kT; ++e) __stcs(dst + e, to_out<OutT>(RD[e]));
            }
          }
        } else {
          // ---- generic path: front edge, the thread holding sx, grid edge
          const double oU[4] = {RU[0], RU[1], RU[2], RU[3]};
          const double oD[4] = {RD[0], RD[1], RD[2], RD[3]};
#pragma unroll
          for (int e = 0; e < kT; ++e) {
            const int x = u0 + e;
            const int d = x - sx;
            const int i = d < 0 ? -d : d;
            if (x < nx && i < k) {
              const bool joined = (i == k - 1) && (k >= 2);
              double aU = oU[e], aD = oD[e];
              if (joined) {
                // diagonal cells (x, sy +- (k-1)) of the previous ring
                if (sy + k - 1 < ny) {
                  aU = ((pRU >> e) & 1u) ? slot_rd[d > 0 ? 0 : 1] : 0.0;
                  __stcs(out + (size_t)(sy + k - 1) * nx + x, to_out<OutT>(aU));
                }
                if (sy - (k - 1) >= 0) {
                  aD = ((pRD >> e) & 1u) ? slot_rd[d > 0 ? 3 : 2] : 0.0;
                  __stcs(out + (size_t)(sy - (k - 1)) * nx + x, to_out<OutT>(aD));
                }
              }
              const double c = ratio_rn((double)i, fk, r);
              double bU, bD;
              if (d > 0) {
                bU = e > 0 ? oU[e > 0 ? e - 1 : 0] : LU;
                bD = e > 0 ? oD[e > 0 ? e - 1 : 0] : LD;
              } else {
                bU = e < 3 ? oU[e < 3 ? e + 1 : 3] : RUn;
                bD = e < 3 ? oD[e < 3 ? e + 1 : 3] : RDn;
              }
              if (ru_on) {
                double v = (i == 0) ? aU : lerp_rn(aU, bU, c);
                v = ((nRU >> e) & 1u) ? v : 0.0;
                RU[e] = v;
                __stcs(out + (size_t)(sy + k) * nx + x, to_out<OutT>(v));
              } else {
                RU[e] = aU;
              }
              if (rd_on) {
                double v = (i == 0) ? aD : lerp_rn(aD, bD, c);
                v = ((nRD >> e) & 1u) ? v : 0.0;
                RD[e] = v;
                __stcs(out + (size_t)(sy - k) * nx + x, to_out<OutT>(v));
              } else {
                RD[e] = aD;
              }
            }
          }
        }
      }
      if (lane == 31) {
        edge_wr[(0 * NW + warp) * 2 + 1] = RU[3];
        edge_wr[(1 * NW + warp) * 2 + 1] = RD[3];
      }
      if (lane == 0) {
        edge_wr[(0 * NW + warp) * 2 + 0] = RU[0];
        edge_wr[(1 * NW + warp) * 2 + 0] = RD[0];
      }
    }

    // ----------------------------- column fronts ----------------------------
    if (k > wmin_y && k <= Kcol) {
      const bool cr_on = sx + k < nx, cl_on = sx - k >= 0;
      const int kkR = (sx + k) & (S - 1), kkL = (sx - k) & (S - 1);
      double LR = 0, LL = 0, RRn = 0, RLn = 0;
      if (wplus_y) {
        LR = __shfl_up_sync(0xffffffffu, CR[3], 1);
        LL = __shfl_up_sync(0xffffffffu, CL[3], 1);
        if (lane == 0 && warp > 0) {
          LR = edge_rd[(2 * NW + warp - 1) * 2 + 1];
          LL = edge_rd[(3 * NW + warp - 1) * 2 + 1];
        }
      }
      if (wminus_y) {
        RRn = __shfl_down_sync(0xffffffffu, CR[0], 1);
        RLn = __shfl_down_sync(0xffffffffu, CL[0], 1);
        if (lane == 31 && warp < NW - 1) {
          RRn = edge_rd[(2 * NW + warp + 1) * 2 + 0];
          RLn = edge_rd[(3 * NW + warp + 1) * 2 + 0];
        }
      }
      if (k > imin_y) {
        if (k > ifull_y) {
          double c[4];
          if (side_y > 0) {
            c[0] = ratio_rn(f0y, fk, r);       c[1] = ratio_rn(f0y + 1.0, fk, r);
            c[2] = ratio_rn(f0y + 2.0, fk, r); c[3] = ratio_rn(f0y + 3.0, fk, r);
            if (cr_on) full_update<true>(CR, LR, c, nCR);
            if (cl_on) full_update<true>(CL, LL, c, nCL);
          } else {
            c[0] = ratio_rn(f0y, fk, r);       c[1] = ratio_rn(f0y - 1.0, fk, r);
            c[2] = ratio_rn(f0y - 2.0, fk, r); c[3] = ratio_rn(f0y - 3.0, fk, r);
            if (cr_on) full_update<false>(CR, RRn, c, nCR);
            if (cl_on) full_update<false>(CL, RLn, c, nCL);
          }
          if (cr_on)
            Vec4<OutT>::store_shared(tileR + kkR * kPitch + kT * lane, CR[0], CR[1], CR[2], CR[3]);
          if (cl_on)
            Vec4<OutT>::store_shared(tileL + kkL * kPitch + kT * lane, CL[0], CL[1], CL[2], CL[3]);
        } else {
          const double oR[4] = {CR[0], CR[1], CR[2], CR[3]};
          const double oL[4] = {CL[0], CL[1], CL[2], CL[3]};
#pragma unroll
          for (int e = 0; e < kT; ++e) {
            const int y = u0 + e;
            const int d = y - sy;
            const int j = d < 0 ? -d : d;
            if (y < ny && j < k) {
              const bool joined = (j == k - 1) && (k >= 2);
              double bR, bL;
              if (d > 0) {
                bR = e > 0 ? oR[e > 0 ? e - 1 : 0] : LR;
                bL = e > 0 ? oL[e > 0 ? e - 1 : 0] : LL;
              } else {
                bR = e < 3 ? oR[e < 3 ? e + 1 : 3] : RRn;
                bL = e < 3 ? oL[e < 3 ? e + 1 : 3] : RLn;
              }
              double aR = oR[e], aL = oL[e];
              if (joined) {
                // diagonal of the previous ring: q[k-1][k-2] * occ(k-1,k-1)
                aR = ((pCR >> e) & 1u) ? bR : 0.0;
                aL = ((pCL >> e) & 1u) ? bL : 0.0;
              }
              const double c = ratio_rn((double)j, fk, r);
              if (cr_on) {
                double v = (j == 0) ? aR : lerp_rn(aR, bR, c);
                v = ((nCR >> e) & 1u) ? v : 0.0;
                CR[e] = v;
                tileR[kkR * kPitch + kT * lane + e] = to_out<OutT>(v);
              } else {
                CR[e] = aR;
              }
              if (cl_on) {
                double v = (j == 0) ? aL : lerp_rn(aL, bL, c);
                v = ((nCL >> e) & 1u) ? v : 0.0;
                CL[e] = v;
                tileL[kkL * kPitch + kT * lane + e] = to_out<OutT>(v);
              } else {
                CL[e] = aL;
              }
              // hand q[k][k-1] to the row owners of |x-sx| == k
              if (j == k - 1) {
                if (d >= 0) { slot_wr[0] = CR[e]; slot_wr[1] = CL[e]; }
                if (d <= 0) { slot_wr[3] = CR[e]; slot_wr[2] = CL[e]; }
              }
            }
          }
        }
      }
      if (lane == 31) {
        edge_wr[(2 * NW + warp) * 2 + 1] = CR[3];
        edge_wr[(3 * NW + warp) * 2 + 1] = CL[3];
      }
      if (lane == 0) {
        edge_wr[(2 * NW + warp) * 2 + 0] = CR[0];
        edge_wr[(3 * NW + warp) * 2 + 0] = CL[0];
      }
      if (cr_on && (kkR == S - 1 || sx + k == nx - 1))
        flush_block<OutT, +1>(tileR, out, nx, ny, sx, sy, k, sx + k, wu0, lane, vec_ok);
      if (cl_on && kkL == 0)
        flush_block<OutT, -1>(tileL, out, nx, ny, sx, sy, k, sx - k, wu0, lane, vec_ok);
    }

    pRU = nRU; pRD = nRD; pCR = nCR; pCL = nCL;
    nRU = xRU; nRD = xRD; nCR = xCR; nCL = xCL;
    __syncthreads();
  }
}

template <typename OutT, int MAXNT, int MINB>
cudaError_t launch_front(const FrontParams &p, int64_t npairs, int nt, size_t smem,
                         cudaStream_t st) {
  auto kern = sweep_front_kernel<OutT, MAXNT, MINB>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<(unsigned)npairs, nt, smem, st>>>(p);
  return cudaGetLastError();
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

bool vhp_sweep_front_supported(int nx, int ny) {
  const int n = std::max(nx, ny);
  return n >= 1 && n <= 1024 * kT && n <= 16384;
}

cudaError_t vhp_launch_sweep_front(const VhpPackedMaps &maps, int nx, int ny,
                                   const int32_t *d_src_xy, const int32_t *d_src_map,
                                   int64_t npairs, vhp_dtype dtype, void *d_out,
                                   const double *d_rcp, int *d_err, cudaStream_t st,
                                   int64_t *launches) {
  FrontParams p;
  p.err = d_err;
  p.rowbits = maps.rowbits; p.colbits = maps.colbits;
  p.wpr = maps.wpr; p.wpc = maps.wpc;
  p.row_plane = maps.row_plane; p.col_plane = maps.col_plane;
  p.nx = nx; p.ny = ny;
  p.src_xy = d_src_xy; p.src_map = d_src_map;
  p.out = d_out; p.rcp = d_rcp;
  const int n = std::max(nx, ny);
  const int nt = ((n + kT - 1) / kT + 31) / 32 * 32;
  const int nw = nt / 32;
  const size_t esz = dtype == VHP_F32 ? 4 : 8;
  const int S = 32 / (int)esz;
  const size_t smem = 2 * (size_t)nw * S * kPitch * esz + (2 * 4 * nw * 2 + 8) * sizeof(double);
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
  // 128-bit stores need 16-byte aligned rows: nx % 4 == 0 (f32) / nx % 2 == 0 (f64)
  p.vec_ok = (dtype == VHP_F32) ? (nx % 4 == 0) : (nx % 2 == 0);
  cudaError_t e;
  if (nt <= 256) {
    e = dtype == VHP_F32 ? launch_front<float, 256, 3>(p, npairs, nt, smem, st)
                         : launch_front<double, 256, 3>(p, npairs, nt, smem, st);
  } else {
    e = dtype == VHP_F32 ? launch_front<float, 1024, 1>(p, npairs, nt, smem, st)
                         : launch_front<double, 1024, 1>(p, npairs, nt, smem, st);
  }
  if (launches) *launches += 1;
  return e;
}
