// sweep_front_body.cuh -- the tuned K1 sweep as a CTA-wide device function, shared by
// the batched sweep kernel (kernels_sweep_front.cu) and the planner kernel
// (kernels_planner.cu).  See kernels_sweep_front.cu for the design notes.
#ifndef VHP_SWEEP_FRONT_BODY_CUH
#define VHP_SWEEP_FRONT_BODY_CUH

#include <cstdint>

#include "vhp_internal.h"
#include "sweep_common.cuh"

namespace {

constexpr int kT = 4;                 // coordinates per thread
constexpr int kWSpan = 32 * kT;       // coordinates per warp
constexpr int kPitch = kWSpan + 4;    // staging tile pitch (elements)
constexpr int kBig = 0x3fffffff;

enum { SIDE_PLUS = 0, SIDE_MINUS = 1, SIDE_MIXED = 2 };

struct FrontParams {
  const uint32_t *rowbits, *colbits;
  int wpr, wpc;
  size_t row_plane, col_plane;
  int nx, ny;
  const int32_t *src_xy, *src_map;
  void *out;
  const double *rcp;
  uint32_t edge_p2; // bytes per parity of the edge-slot region (power of two)
  int *err;
};

template <typename OutT> struct Vec4;
template <> struct Vec4<float> {
  static __device__ __forceinline__ void store(float *p, const double (&F)[4]) {
    __stcs(reinterpret_cast<float4 *>(p),
           make_float4(__double2float_rn(F[0]), __double2float_rn(F[1]),
                       __double2float_rn(F[2]), __double2float_rn(F[3])));
  }
  static __device__ __forceinline__ void store_shared(float *p, const double (&F)[4]) {
    *reinterpret_cast<float4 *>(p) =
        make_float4(__double2float_rn(F[0]), __double2float_rn(F[1]), __double2float_rn(F[2]),
                    __double2float_rn(F[3]));
  }
};
template <> struct Vec4<double> {
  static __device__ __forceinline__ void store(double *p, const double (&F)[4]) {
    __stcs(reinterpret_cast<double2 *>(p), make_double2(F[0], F[1]));
    __stcs(reinterpret_cast<double2 *>(p) + 1, make_double2(F[2], F[3]));
  }
  static __device__ __forceinline__ void store_shared(double *p, const double (&F)[4]) {
    reinterpret_cast<double2 *>(p)[0] = make_double2(F[0], F[1]);
    reinterpret_cast<double2 *>(p)[1] = make_double2(F[2], F[3]);
  }
};

// One ring step of one front for the thread's 4 elements.  b = the neighbour
// towards the source: left (lower u) on the plus side, right on the minus side;
// lm marks per element which one applies in the warp that contains the source.
template <int SIDE>
__device__ __forceinline__ void front_update(double (&F)[4], double L, double R,
                                             const double (&c)[4], uint32_t nib, uint32_t lm) {
  double b0, b1, b2, b3;
  if (SIDE == SIDE_PLUS) {
    b0 = L; b1 = F[0]; b2 = F[1]; b3 = F[2];
  } else if (SIDE == SIDE_MINUS) {
    b0 = F[1]; b1 = F[2]; b2 = F[3]; b3 = R;
  } else {
    b0 = (lm & 1u) ? L : F[1];
    b1 = (lm & 2u) ? F[0] : F[2];
    b2 = (lm & 4u) ? F[1] : F[3];
    b3 = (lm & 8u) ? F[2] : R;
  }
  const double v0 = lerp_rn(F[0], b0, c[0]);
  const double v1 = lerp_rn(F[1], b1, c[1]);
  const double v2 = lerp_rn(F[2], b2, c[2]);
  const double v3 = lerp_rn(F[3], b3, c[3]);
  F[0] = (nib & 1u) ? v0 : 0.0;
  F[1] = (nib & 2u) ? v1 : 0.0;
  F[2] = (nib & 4u) ? v2 : 0.0;
  F[3] = (nib & 8u) ? v3 : 0.0;
}

__device__ __forceinline__ double lds64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts64(uint32_t a, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}

// bits e = 0..3 with |d0 + e| < k
__device__ __forceinline__ uint32_t active_mask(int d0, int k) {
  const int lo = max(0, 1 - k - d0), hi = min(3, k - 1 - d0);
  return lo <= hi ? ((2u << hi) - 1u) & ~((1u << lo) - 1u) : 0u;
}

// flush one staged block of a column front: rows [wy0, wy0+128) x X in [Xb, Xb+S)
template <typename OutT, int DIR>
__device__ __forceinline__ void flush_block(const OutT *__restrict__ tile, OutT *__restrict__ out,
                                            int nx, int ny, int sx, int sy, int k, int X, int wy0,
                                            int lane, bool fast) {
  constexpr int S = 32 / (int)sizeof(OutT);
  constexpr int V = 16 / (int)sizeof(OutT);
  __syncwarp();
  const int Xb = X & ~(S - 1);
  const int h = lane & 1, r = lane >> 1;
  const int Xc = Xb + V * h;
  const OutT *tp = tile + (V * h) * kPitch + r;
  OutT *dst = out + (size_t)(wy0 + r) * nx + Xc;
  const size_t dstep = (size_t)16 * nx;
  if (fast) {
#pragma unroll
    for (int pass = 0; pass < kWSpan / 16; ++pass) {
      if constexpr (sizeof(OutT) == 4)
        __stcs(reinterpret_cast<float4 *>(dst),
               make_float4(tp[pass * 16], tp[kPitch + pass * 16], tp[2 * kPitch + pass * 16],
                           tp[3 * kPitch + pass * 16]));
      else
        __stcs(reinterpret_cast<double2 *>(dst),
               make_double2(tp[pass * 16], tp[kPitch + pass * 16]));
      dst += dstep;
    }
  } else {
#pragma unroll 2
    for (int pass = 0; pass < kWSpan / 16; ++pass) {
      const int y = wy0 + pass * 16 + r;
      const int j = y > sy ? y - sy : sy - y;
      if (y < ny) {
#pragma unroll
        for (int m = 0; m < V; ++m) {
          const int Xm = Xc + m;
          const bool ok = DIR > 0 ? (j < Xm - sx && Xm <= sx + k) : (j < sx - Xm && Xm >= sx - k);
          if (ok) __stcs(dst + m, tp[m * kPitch + pass * 16]);
        }
      }
      dst += dstep;
    }
  }
  __syncwarp();
}

// Shared memory layout (dynamic):
//   [0, tiles_bytes)            staging tiles [2 dirs][NW][S][kPitch] of OutT
//   edge region, aligned to 2*P2 (P2 = power of two >= (NW+1)*64 bytes), two
//   parities of P2 bytes each: per warp 64 B = [front RU,RD,CR,CL][lo,hi] doubles,
//   followed by the 4 diagonal hand-off slots.  Parity toggles by XOR with P2.
// One complete sweep from light source (sx, sy) by the whole CTA (blockDim.x =
// front_geometry().nt threads, all of them must call).  Writes every cell of
// out[ny][nx] exactly once.  Ends with a block barrier: the result is visible to
// the CTA and the shared memory may be reused.
template <typename OutT, bool VEC>
__device__ __forceinline__ void sweep_front_body(const FrontParams &p, const int sx, const int sy,
                                                 const uint32_t *__restrict__ rowbits,
                                                 const uint32_t *__restrict__ colbits,
                                                 OutT *__restrict__ out,
                                                 unsigned char *smem_raw) {
  constexpr int S = 32 / (int)sizeof(OutT);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NW = blockDim.x >> 5;
  const int nx = p.nx, ny = p.ny;
  const int wpr = p.wpr, wpc = p.wpc;

  OutT *tiles = reinterpret_cast<OutT *>(smem_raw);
  constexpr int tile_elems = S * kPitch;
  OutT *tileR = tiles + warp * tile_elems;
  OutT *tileL = tiles + (NW + warp) * tile_elems;
  const uint32_t P2 = p.edge_p2;
  const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(smem_raw);
  const uint32_t ereg = (smem0 + 2u * NW * tile_elems * (uint32_t)sizeof(OutT) + 2u * P2 - 1u) & ~(2u * P2 - 1u);
  for (uint32_t i = tid; i < 2u * P2 / 8u; i += blockDim.x) sts64(ereg + 8u * i, 0.0);

  const int u0 = kT * tid, wu0 = kWSpan * warp, wu1 = wu0 + kWSpan - 1;
  const int nsh = 4 * (tid & 7);
  const int widx_x = min(tid >> 3, wpr - 1), widx_y = min(tid >> 3, wpc - 1);
  const int dx0 = u0 - sx, dy0 = u0 - sy;
  // in-grid masks, and the same with the never-written border (X==0 / Y==0) forced dark
  const uint32_t ing_x = u0 + 3 < nx ? 0xFu : (u0 < nx ? (1u << (nx - u0)) - 1u : 0u);
  const uint32_t ing_y = u0 + 3 < ny ? 0xFu : (u0 < ny ? (1u << (ny - u0)) - 1u : 0u);
  const uint32_t keep_x = (tid == 0 && sx > 0) ? (ing_x & 0xEu) : ing_x;
  const uint32_t keep_y = (tid == 0 && sy > 0) ? (ing_y & 0xEu) : ing_y;
  // elements on the plus side of the source (upstream neighbour = lower u)
  const uint32_t lm_x = dx0 > 0 ? 0xFu : (dx0 < -2 ? 0u : (0xFu & ~((2u << (-dx0)) - 1u)));
  const uint32_t lm_y = dy0 > 0 ? 0xFu : (dy0 < -2 ? 0u : (0xFu & ~((2u << (-dy0)) - 1u)));
  // warp-uniform geometry
  const int side_x = wu0 > sx ? SIDE_PLUS : (wu1 < sx ? SIDE_MINUS : SIDE_MIXED);
  const int side_y = wu0 > sy ? SIDE_PLUS : (wu1 < sy ? SIDE_MINUS : SIDE_MIXED);
  const int wmin_x = wu0 >= nx ? kBig : (side_x == SIDE_PLUS ? wu0 - sx : (side_x == SIDE_MINUS ? sx - wu1 : 0));
  const int wmin_y = wu0 >= ny ? kBig : (side_y == SIDE_PLUS ? wu0 - sy : (side_y == SIDE_MINUS ? sy - wu1 : 0));
  const int wimax_x = max(abs(wu0 - sx), abs(wu1 - sx));
  const int wimax_y = max(abs(wu0 - sy), abs(wu1 - sy));
  const double fdx[4] = {(double)dx0, (double)(dx0 + 1), (double)(dx0 + 2), (double)(dx0 + 3)};
  const double fdy[4] = {(double)dy0, (double)(dy0 + 1), (double)(dy0 + 2), (double)(dy0 + 3)};
  // this warp's edge slots (parity 0); neighbours are at +-64 bytes
  uint32_t eb = ereg + 64u * warp;
  const uint32_t slot_off = 64u * (NW - warp); // from eb to the diagonal slots
  const bool lane_lo = lane == 0, lane_hi = lane == 31;
  const bool fix_lo = lane_lo && warp > 0, fix_hi = lane_hi && warp < NW - 1;

  double RU[4] = {0, 0, 0, 0}, RD[4] = {0, 0, 0, 0}, CR[4] = {0, 0, 0, 0}, CL[4] = {0, 0, 0, 0};

  // source cell
  {
    const uint32_t w = __ldg(rowbits + sy * wpr + (sx >> 5));
    const double s0 = ((w >> (sx & 31)) & 1u) ? 1.0 : 0.0;
    if ((unsigned)(-dx0) < (unsigned)kT) {
#pragma unroll
      for (int q = 0; q < kT; ++q)
        if (q == -dx0) { RU[q] = s0; RD[q] = s0; }
      out[(size_t)sy * nx + sx] = to_out<OutT>(s0);
    }
    if ((unsigned)(-dy0) < (unsigned)kT) {
#pragma unroll
      for (int q = 0; q < kT; ++q)
        if (q == -dy0) { CR[q] = s0; CL[q] = s0; }
    }
  }

  // occupancy nibble of row Y / column X (callers keep Y, X inside the grid)
  auto ld_row = [&](int Y) -> uint32_t { return (__ldg(rowbits + (unsigned)(Y * wpr + widx_x)) >> nsh) & keep_x; };
  auto ld_col = [&](int X) -> uint32_t { return (__ldg(colbits + (unsigned)(X * wpc + widx_y)) >> nsh) & keep_y; };
  uint32_t nRU = 0, nRD = 0, nCR = 0, nCL = 0, pRU = 0, pRD = 0, pCR = 0, pCL = 0;

  const int Krow = max(sy, ny - 1 - sy), Kcol = max(sx, nx - 1 - sx);
  const int Kmax = max(Krow, Kcol);
  const int kx0 = wmin_x + 1, ky0 = wmin_y + 1; // first ring that touches this warp
  const int kstart = min(min(kx0, ky0), Kmax + 2);
  OutT *const out_u0 = out + u0;
  __syncthreads();

  int k = 1;
  for (; k < kstart; ++k) __syncthreads(); // not reached by any front yet
  if (k & 1) eb ^= P2;                      // parity of the first real step
  double fk = (double)k;

#pragma unroll 1
  for (; k <= Kmax + 1; ++k, fk += 1.0, eb ^= P2) {
    // eb: slots written this step; (eb ^ P2): slots written by the previous step
    const uint32_t er = eb ^ P2;
    const double r = __ldg(p.rcp + 2 * k);

    // ------------------------------ row fronts ------------------------------
    if (k >= kx0 && k <= Krow + 1) {
      const bool ru_on = sy + k < ny, rd_on = sy - k >= 0;
      if (k == kx0) { // first ring in this warp: fetch current and previous nibbles
        nRU = ru_on ? ld_row(sy + k) : 0u;
        pRU = (sy + k - 1 < ny) ? ld_row(sy + k - 1) : 0u;
        nRD = (sy - k >= 1) ? ld_row(sy - k) : 0u;
        pRD = (sy - k + 1 >= 1) ? ld_row(sy - k + 1) : 0u;
      }
      uint32_t xRU = ld_row(min(sy + k + 1, ny - 1));
      uint32_t xRD = ld_row(max(sy - k - 1, 0));
      xRD = (sy - k - 1 >= 1) ? xRD : 0u;
      double LU = 0, LD = 0, RUn = 0, RDn = 0;
      if (side_x != SIDE_MINUS) {
        LU = __shfl_up_sync(0xffffffffu, RU[3], 1);
        LD = __shfl_up_sync(0xffffffffu, RD[3], 1);
        if (fix_lo) { LU = lds64(er - 64u + 8u); LD = lds64(er - 64u + 24u); }
      }
      if (side_x != SIDE_PLUS) {
        RUn = __shfl_down_sync(0xffffffffu, RU[0], 1);
        RDn = __shfl_down_sync(0xffffffffu, RD[0], 1);
        if (fix_hi) { RUn = lds64(er + 64u); RDn = lds64(er + 64u + 16u); }
      }
      uint32_t act = 0xFu;
      const bool interior = k > wimax_x + 1;
      if (!interior) { // the front edge is inside this warp
        act = active_mask(dx0, k);
        const int tP = k - 1 - dx0, tM = 1 - k - dx0; // elements with x = sx +- (k-1)
        if (k >= 2 && ((unsigned)tP < 4u || (unsigned)tM < 4u)) {
          // diagonal cells of ring k-1 join the row fronts: value = q[k-1][k-2]*occ
#pragma unroll
          for (int e = 0; e < kT; ++e) {
            if ((e == tP || e == tM) && ((ing_x >> e) & 1u)) {
              const uint32_t qU = e == tP ? 0u : 8u, qD = e == tP ? 24u : 16u;
              if (sy + k - 1 < ny) {
                RU[e] = ((pRU >> e) & 1u) ? lds64(er + slot_off + qU) : 0.0;
                __stcs(out_u0 + (size_t)(sy + k - 1) * nx + e, to_out<OutT>(RU[e]));
              }
              if (sy - (k - 1) >= 0) {
                RD[e] = ((pRD >> e) & 1u) ? lds64(er + slot_off + qD) : 0.0;
                __stcs(out_u0 + (size_t)(sy - (k - 1)) * nx + e, to_out<OutT>(RD[e]));
              }
            }
          }
        }
      }
      const double c[4] = {ratio_rn(fabs(fdx[0]), fk, r), ratio_rn(fabs(fdx[1]), fk, r),
                           ratio_rn(fabs(fdx[2]), fk, r), ratio_rn(fabs(fdx[3]), fk, r)};
      const uint32_t mU = nRU & act, mD = nRD & act;
      if (side_x == SIDE_PLUS) {
        if (ru_on) front_update<SIDE_PLUS>(RU, LU, RUn, c, mU, lm_x);
        if (rd_on) front_update<SIDE_PLUS>(RD, LD, RDn, c, mD, lm_x);
      } else if (side_x == SIDE_MINUS) {
        if (ru_on) front_update<SIDE_MINUS>(RU, LU, RUn, c, mU, lm_x);
        if (rd_on) front_update<SIDE_MINUS>(RD, LD, RDn, c, mD, lm_x);
      } else {
        if (ru_on) front_update<SIDE_MIXED>(RU, LU, RUn, c, mU, lm_x);
        if (rd_on) front_update<SIDE_MIXED>(RD, LD, RDn, c, mD, lm_x);
      }
      const uint32_t sm = act & ing_x;
      const bool vec = VEC && (interior ? wu1 < nx : sm == 0xFu);
      if (ru_on) {
        OutT *dst = out_u0 + (size_t)(sy + k) * nx;
        if (vec) Vec4<OutT>::store(dst, RU);
        else {
#pragma unroll
          for (int e = 0; e < kT; ++e)
            if ((sm >> e) & 1u) __stcs(dst + e, to_out<OutT>(RU[e]));
        }
      }
      if (rd_on) {
        OutT *dst = out_u0 + (size_t)(sy - k) * nx;
        if (vec) Vec4<OutT>::store(dst, RD);
        else {
#pragma unroll
          for (int e = 0; e < kT; ++e)
            if ((sm >> e) & 1u) __stcs(dst + e, to_out<OutT>(RD[e]));
        }
      }
      if (lane_hi) { sts64(eb + 8u, RU[3]); sts64(eb + 24u, RD[3]); }
      if (lane_lo) { sts64(eb, RU[0]); sts64(eb + 16u, RD[0]); }
      pRU = nRU; nRU = xRU; pRD = nRD; nRD = xRD;
    }

    // ----------------------------- column fronts ----------------------------
    if (k >= ky0 && k <= Kcol) {
      const bool cr_on = sx + k < nx, cl_on = sx - k >= 0;
      if (k == ky0) {
        nCR = cr_on ? ld_col(sx + k) : 0u;
        pCR = (sx + k - 1 < nx) ? ld_col(sx + k - 1) : 0u;
        nCL = (sx - k >= 1) ? ld_col(sx - k) : 0u;
        pCL = (sx - k + 1 >= 1) ? ld_col(sx - k + 1) : 0u;
      }
      uint32_t xCR = ld_col(min(sx + k + 1, nx - 1));
      uint32_t xCL = ld_col(max(sx - k - 1, 0));
      xCL = (sx - k - 1 >= 1) ? xCL : 0u;
      double LR = 0, LL = 0, RRn = 0, RLn = 0;
      if (side_y != SIDE_MINUS) {
        LR = __shfl_up_sync(0xffffffffu, CR[3], 1);
        LL = __shfl_up_sync(0xffffffffu, CL[3], 1);
        if (fix_lo) { LR = lds64(er - 64u + 40u); LL = lds64(er - 64u + 56u); }
      }
      if (side_y != SIDE_PLUS) {
        RRn = __shfl_down_sync(0xffffffffu, CR[0], 1);
        RLn = __shfl_down_sync(0xffffffffu, CL[0], 1);
        if (fix_hi) { RRn = lds64(er + 64u + 32u); RLn = lds64(er + 64u + 48u); }
      }
      uint32_t act = 0xFu;
      const int tP = k - 1 - dy0, tM = 1 - k - dy0; // elements with y = sy +- (k-1)
      const bool interior = k > wimax_y + 1;
      const bool at_edge = !interior && ((unsigned)tP < 4u || (unsigned)tM < 4u);
      if (!interior) {
        act = active_mask(dy0, k);
        if (k >= 2 && at_edge) {
          // the element joining the front starts from the diagonal of ring k-1:
          // q[k-1][k-1] = q[k-1][k-2]*occ = (its upstream neighbour) * (its previous bit)
#pragma unroll
          for (int e = 0; e < kT; ++e) {
            if (e == tP) {
              const double bR = e > 0 ? CR[e > 0 ? e - 1 : 0] : LR;
              const double bL = e > 0 ? CL[e > 0 ? e - 1 : 0] : LL;
              CR[e] = ((pCR >> e) & 1u) ? bR : 0.0;
              CL[e] = ((pCL >> e) & 1u) ? bL : 0.0;
            } else if (e == tM) {
              const double bR = e < 3 ? CR[e < 3 ? e + 1 : 3] : RRn;
              const double bL = e < 3 ? CL[e < 3 ? e + 1 : 3] : RLn;
              CR[e] = ((pCR >> e) & 1u) ? bR : 0.0;
              CL[e] = ((pCL >> e) & 1u) ? bL : 0.0;
            }
          }
        }
      }
      const double c[4] = {ratio_rn(fabs(fdy[0]), fk, r), ratio_rn(fabs(fdy[1]), fk, r),
                           ratio_rn(fabs(fdy[2]), fk, r), ratio_rn(fabs(fdy[3]), fk, r)};
      const uint32_t mR = nCR & act, mL = nCL & act;
      if (side_y == SIDE_PLUS) {
        if (cr_on) front_update<SIDE_PLUS>(CR, LR, RRn, c, mR, lm_y);
        if (cl_on) front_update<SIDE_PLUS>(CL, LL, RLn, c, mL, lm_y);
      } else if (side_y == SIDE_MINUS) {
        if (cr_on) front_update<SIDE_MINUS>(CR, LR, RRn, c, mR, lm_y);
        if (cl_on) front_update<SIDE_MINUS>(CL, LL, RLn, c, mL, lm_y);
      } else {
        if (cr_on) front_update<SIDE_MIXED>(CR, LR, RRn, c, mR, lm_y);
        if (cl_on) front_update<SIDE_MIXED>(CL, LL, RLn, c, mL, lm_y);
      }
      if (at_edge) { // hand q[k][k-1] to the row owners of |x-sx| == k
#pragma unroll
        for (int e = 0; e < kT; ++e) {
          if (e == tP) { sts64(eb + slot_off, CR[e]); sts64(eb + slot_off + 8u, CL[e]); }
          if (e == tM) { sts64(eb + slot_off + 24u, CR[e]); sts64(eb + slot_off + 16u, CL[e]); }
        }
      }
      const int kkR = (sx + k) & (S - 1), kkL = (sx - k) & (S - 1);
      if (cr_on) Vec4<OutT>::store_shared(tileR + kkR * kPitch + kT * lane, CR);
      if (cl_on) Vec4<OutT>::store_shared(tileL + kkL * kPitch + kT * lane, CL);
      if (lane_hi) { sts64(eb + 40u, CR[3]); sts64(eb + 56u, CL[3]); }
      if (lane_lo) { sts64(eb + 32u, CR[0]); sts64(eb + 48u, CL[0]); }
      if (cr_on && (kkR == S - 1 || sx + k == nx - 1)) {
        const int Xb = (sx + k) & ~(S - 1);
        const bool fast = VEC && kkR == S - 1 && wu1 < ny && wimax_y < Xb - sx;
        flush_block<OutT, +1>(tileR, out, nx, ny, sx, sy, k, sx + k, wu0, lane, fast);
      }
      if (cl_on && kkL == 0) {
        const int Xe = sx - k + S - 1; // last column of the block
        const bool fast = VEC && wu1 < ny && wimax_y < sx - Xe;
        flush_block<OutT, -1>(tileL, out, nx, ny, sx, sy, k, sx - k, wu0, lane, fast);
      }
      pCR = nCR; nCR = xCR; pCL = nCL; nCL = xCL;
    }
    __syncthreads();
  }
}

// launch geometry of sweep_front_body for an nx x ny grid
inline void front_geometry(int nx, int ny, vhp_dtype dtype, int &nt, uint32_t &p2, size_t &smem) {
  const int n = nx > ny ? nx : ny;
  nt = ((n + kT - 1) / kT + 31) / 32 * 32;
  const int nw = nt / 32;
  const size_t esz = dtype == VHP_F32 ? 4 : 8;
  const int S = 32 / (int)esz;
  p2 = 64;
  while (p2 < (uint32_t)(nw + 1) * 64u) p2 <<= 1;
  // tiles + alignment slack + two parities of the edge region
  smem = 2 * (size_t)nw * S * kPitch * esz + 2 * (size_t)p2 + 2 * (size_t)p2;
}

} // namespace
#endif
