// kernels_sweep_ring.cu -- K1, batched visibility sweep, front-specialised warps.
//
// Replaces visibilityBasedSolver::computeVisibility
// (reference src/visibilityBasedSolver.cpp:570-696) for grids up to 1024 x 1024,
// the BASELINE configurations.  Same L-front dynamic program and arithmetic
// contract as the other sweep kernels (sweep_common.cuh).
//
// One CTA per (map, source) pair, all warps advance ring k = 1, 2, ... in lock
// step (one block barrier per ring).  Ring k consists of four fronts
//     RU: cells (x, sy+k)   RD: cells (x, sy-k)     |x-sx| < k   ("row" fronts)
//     CR: cells (sx+k, y)   CL: cells (sx-k, y)     |y-sy| < k   ("column" fronts)
// plus four diagonal cells.  Every warp serves ONE front and one 128-wide slice of
// its "along" coordinate (x for row fronts, y for column fronts): lane l owns the
// four absolute coordinates u = 128*slice + 4*l + e.  The front values stay in
// fp64 registers for the whole sweep; a step costs one shuffle for the upstream
// neighbour (plus one shared-memory slot at slice boundaries), three fp64 ops
// per element for c = |u-s|/k from a table of RN(1/k), the lerp, one occupancy
// word from the bit planes (prefetched a step ahead) and the store.  Warps whose
// slice the front has not reached yet, or whose front has left the grid, only
// wait on the barrier.  Compared with serving all four fronts in every thread
// (kernels_sweep_front.cu) this keeps the per-ring overhead proportional to the
// fronts that are actually active: for a random source typically one front of
// each kind is alive for most of the sweep.
//
// Stores: row fronts write 4 consecutive x per thread (one 128-bit store when the
// row pitch allows it).  Column fronts produce one x per step, so each warp
// stages S = 32 B / sizeof(OutT) steps of its 128 rows in shared memory
// (tile[kk][y], pitch 132: conflict-free both ways) and flushes sector-sized row
// segments aligned on absolute X.
//
// Diagonal cell (k,k) of a quadrant = q[k][k-1]*occ(k,k) (the reference has no
// i==j branch, `v` keeps the previous inner-loop value).  q[k][k-1] belongs to a
// column front: the column-front element |y-sy| == k starts from (its upstream
// neighbour) * (its own previous occupancy bit); the row-front element
// |x-sx| == k receives q[k][k-1] through a shared slot and also stores the
// diagonal cell.
#include <cstdint>

#include "vhp_internal.h"
#include "sweep_common.cuh"

namespace {

constexpr int kT = 4;                 // coordinates per thread
constexpr int kWSpan = 32 * kT;       // coordinates per warp
constexpr int kPitch = kWSpan + 4;    // staging tile pitch (elements)
constexpr int kBig = 0x3fffffff;

enum { SIDE_PLUS = 0, SIDE_MINUS = 1, SIDE_MIXED = 2 };

struct RingParams {
  const uint32_t *rowbits, *colbits;
  const uint32_t *rowbits_map; // rowbits of this CTA's map (set inside the kernel)
  int wpr, wpc;
  size_t row_plane, col_plane;
  int nx, ny;
  const int32_t *src_xy, *src_map;
  void *out;
  const double *rcp;
  uint32_t edge_p2; // bytes per parity of the edge-slot region (power of two)
  int *err;
};

__device__ __forceinline__ double lds64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts64(uint32_t a, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}

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

// One ring step of a front for the thread's 4 elements.  b = the neighbour towards
// the source: lower u on the plus side, higher u on the minus side; lm marks per
// element which one applies in the slice that contains the source.
template <int SIDE>
__device__ __forceinline__ void front_update(double (&F)[4], double nb_lo, double nb_hi,
                                             const double (&c)[4], uint32_t nib, uint32_t lm) {
  double b0, b1, b2, b3;
  if (SIDE == SIDE_PLUS) {
    b0 = nb_lo; b1 = F[0]; b2 = F[1]; b3 = F[2];
  } else if (SIDE == SIDE_MINUS) {
    b0 = F[1]; b1 = F[2]; b2 = F[3]; b3 = nb_hi;
  } else {
    b0 = (lm & 1u) ? nb_lo : F[1];
    b1 = (lm & 2u) ? F[0] : F[2];
    b2 = (lm & 4u) ? F[1] : F[3];
    b3 = (lm & 8u) ? F[2] : nb_hi;
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

// bits e = 0..3 with |d0 + e| < k
__device__ __forceinline__ uint32_t active_mask(int d0, int k) {
  const int lo = max(0, 1 - k - d0), hi = min(3, k - 1 - d0);
  return lo <= hi ? ((2u << hi) - 1u) & ~((1u << lo) - 1u) : 0u;
}

// flush one staged block of a column front: rows [wy0, wy0+128) x X in [Xb, Xb+S)
template <typename OutT>
__device__ __forceinline__ void flush_block(const OutT *__restrict__ tile, OutT *__restrict__ out,
                                            int nx, int ny, int sx, int sy, int dir, int k, int X,
                                            int wy0, int lane, bool fast) {
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
          const int dist = dir > 0 ? Xc + m - sx : sx - (Xc + m); // ring index of that column
          if (j < dist && dist <= k) __stcs(dst + m, tp[m * kPitch + pass * 16]);
        }
      }
      dst += dstep;
    }
  }
  __syncwarp();
}

// Shared memory layout (dynamic):
//   [0, tiles_bytes)            staging tiles [2 column fronts][NS][S][kPitch] of OutT
//   edge region, aligned to 2*P2 (P2 = power of two), two parities of P2 bytes:
//   per (front, slice) 16 B = {lo, hi} doubles, then the 4 diagonal hand-off slots
//   [column dir][row dir].  Parity toggles by XOR with P2.

// per-warp constants of one front slice
struct FrontGeom {
  int dir;        // +1 / -1 across rings
  int sa, na;     // source coordinate and grid size along the front
  int st;         // source coordinate across rings
  int Kf, Kend;   // last ring the front is on / last ring this warp has work
  int wu0, wu1;   // slice range along the front
  int wimax;      // largest |u - sa| in the slice
  int d0;         // u0 - sa of this thread
  uint32_t ing, keep, lm;
  int nsh;
};

// Rings kfirst..Kend of one front slice.  SIDE is a per-warp constant, so the loop
// is instantiated per side and the neighbour selection costs nothing.
template <typename OutT, bool VEC, bool ISROW, int SIDE>
__device__ __forceinline__ void front_rings(const RingParams &p, const FrontGeom &g, const int sx,
                                            const int sy, const int kfirst, const int Kmax,
                                            const uint32_t *__restrict__ pp, const int pstride,
                                            OutT *__restrict__ out, OutT *op, const int ostride,
                                            OutT *tile, uint32_t eb, const uint32_t slots,
                                            double (&F)[4], const bool fix_lo, const bool fix_hi) {
  constexpr int S = 32 / (int)sizeof(OutT);
  const int lane = threadIdx.x & 31;
  const int nx = p.nx, ny = p.ny;
  const uint32_t P2 = p.edge_p2;
  const int dir = g.dir;
  const int Kld = dir > 0 ? g.Kf : g.Kf - 1; // last ring with a real occupancy line (ring at
                                             // coordinate 0 is the forced-dark border)
  const double fd0 = (double)g.d0;
  const double2 *__restrict__ tab = reinterpret_cast<const double2 *>(p.rcp);
  // occupancy words: `nw` holds the line of the current ring (raw, extracted at use so
  // the load issued one ring earlier has a full step to land), pnib the previous nibble
  uint32_t nw, pnib;
  {
    const int T = g.st + dir * kfirst; // pp points at the line of ring kfirst
    nw = (kfirst <= Kld) ? __ldg(pp) : 0u;
    const int Tp = T - dir;
    pnib = (dir > 0 || Tp >= 1) ? ((__ldg(pp - pstride) >> g.nsh) & g.keep) : 0u;
    pp += pstride;
  }
  int T = g.st + dir * kfirst;

#pragma unroll 1
  for (int k = kfirst; k <= Kmax + 1; ++k) {
    if (k <= g.Kend) {
      // eb: slots written in this step; er: slots written in the previous step
      const uint32_t er = eb ^ P2;
      const bool on = k <= g.Kf;
      const double2 rk = __ldg(tab + k); // {RN(1/k), (double)k}
      const uint32_t nib = (nw >> g.nsh) & g.keep;
      nw = 0u;
      if (k < Kld) nw = __ldg(pp); // line of ring k+1
      pp += pstride;
      double nb_lo = 0, nb_hi = 0;
      if (SIDE != SIDE_MINUS) {
        nb_lo = __shfl_up_sync(0xffffffffu, F[3], 1);
        if (fix_lo) nb_lo = lds64(er - 16u + 8u);
      }
      if (SIDE != SIDE_PLUS) {
        nb_hi = __shfl_down_sync(0xffffffffu, F[0], 1);
        if (fix_hi) nb_hi = lds64(er + 16u);
      }
      uint32_t act = 0xFu;
      const bool interior = k > g.wimax + 1;
      bool at_edge = false;
      int tP = 0, tM = 0;
      if (!interior) { // the front edge is inside this slice
        act = active_mask(g.d0, k);
        tP = k - 1 - g.d0; tM = 1 - k - g.d0; // elements with u = sa +- (k-1)
        at_edge = (unsigned)tP < 4u || (unsigned)tM < 4u;
        if (k >= 2 && at_edge) {
          // the element that joins the front starts from the diagonal cell of ring k-1
#pragma unroll
          for (int e = 0; e < kT; ++e) {
            if ((e == tP || e == tM) && ((g.ing >> e) & 1u)) {
              if (ISROW) {
                // q[k-1][k-2] from the column front on this element's side; slot [col dir][row dir]
                const uint32_t sl = (e == tP ? 0u : 16u) + (dir > 0 ? 0u : 8u);
                F[e] = ((pnib >> e) & 1u) ? lds64((slots ^ (er & P2)) + sl) : 0.0;
                __stcs(op - ostride + e, to_out<OutT>(F[e]));
              } else {
                double b;
                if (e == tP) b = e > 0 ? F[e > 0 ? e - 1 : 0] : nb_lo;
                else         b = e < 3 ? F[e < 3 ? e + 1 : 3] : nb_hi;
                F[e] = ((pnib >> e) & 1u) ? b : 0.0;
              }
            }
          }
        }
      }
      if (on) {
        const double c[4] = {ratio_rn(fabs(fd0), rk.y, rk.x), ratio_rn(fabs(fd0 + 1.0), rk.y, rk.x),
                             ratio_rn(fabs(fd0 + 2.0), rk.y, rk.x),
                             ratio_rn(fabs(fd0 + 3.0), rk.y, rk.x)};
        front_update<SIDE>(F, nb_lo, nb_hi, c, nib & act, g.lm);
        if (ISROW) {
          const uint32_t sm = act & g.ing;
          if (VEC && (interior ? g.wu1 < g.na : sm == 0xFu)) Vec4<OutT>::store(op, F);
          else {
#pragma unroll
            for (int e = 0; e < kT; ++e)
              if ((sm >> e) & 1u) __stcs(op + e, to_out<OutT>(F[e]));
          }
        } else {
          if (at_edge) { // hand q[k][k-1] to the row fronts: slot [col dir][row dir]
            const uint32_t sl = (slots ^ (eb & P2)) + (dir > 0 ? 0u : 16u);
#pragma unroll
            for (int e = 0; e < kT; ++e) {
              if (e == tP) sts64(sl, F[e]);       // y = sy + (k-1): row front RU
              if (e == tM) sts64(sl + 8u, F[e]);  // y = sy - (k-1): row front RD
            }
          }
          const int kk = T & (S - 1);
          Vec4<OutT>::store_shared(tile + kk * kPitch + kT * lane, F);
          const bool last = dir > 0 ? (kk == S - 1 || T == nx - 1) : kk == 0;
          if (last) {
            // all rows of the slice strictly inside the cone for the whole block?
            const int near = dir > 0 ? (T & ~(S - 1)) - sx : sx - (T + S - 1);
            const bool fast = VEC && (dir > 0 ? kk == S - 1 : true) && g.wu1 < ny && g.wimax < near;
            flush_block<OutT>(tile, out, nx, ny, sx, sy, dir, k, T, g.wu0, lane, fast);
          }
        }
      }
      if (lane == 31) sts64(eb + 8u, F[3]);
      if (lane == 0) sts64(eb, F[0]);
      pnib = nib;
      op += ostride;
      T += dir;
    }
    eb ^= P2;
    __syncthreads();
  }
}

template <typename OutT, bool VEC, bool ISROW>
__device__ __forceinline__ void run_front(const RingParams &p, const int sx, const int sy,
                                          const uint32_t *__restrict__ plane, const int pitch,
                                          OutT *__restrict__ out, const int f, const int slice,
                                          const int NS, OutT *tile, const uint32_t ereg) {
  const int lane = threadIdx.x & 31;
  const int nx = p.nx, ny = p.ny;
  FrontGeom g;
  g.dir = (f & 1) ? -1 : 1;
  g.sa = ISROW ? sx : sy;
  g.na = ISROW ? nx : ny;
  g.st = ISROW ? sy : sx;
  const int ntm = ISROW ? ny : nx;
  g.Kf = g.dir > 0 ? ntm - 1 - g.st : g.st;
  g.Kend = ISROW ? g.Kf + 1 : g.Kf; // row fronts also store the last diagonal
  const int Kmax = max(max(sx, nx - 1 - sx), max(sy, ny - 1 - sy));
  g.wu0 = kWSpan * slice;
  g.wu1 = g.wu0 + kWSpan - 1;
  const int u0 = g.wu0 + kT * lane;
  g.d0 = u0 - g.sa;
  g.nsh = 4 * (lane & 7);
  const int widx = min(4 * slice + (lane >> 3), pitch - 1);
  // in-grid mask, and the same with the never-written border (coordinate 0) forced dark
  g.ing = u0 + 3 < g.na ? 0xFu : (u0 < g.na ? (1u << (g.na - u0)) - 1u : 0u);
  g.keep = (u0 == 0 && g.sa > 0) ? (g.ing & 0xEu) : g.ing;
  g.lm = g.d0 > 0 ? 0xFu : (g.d0 < -2 ? 0u : (0xFu & ~((2u << (-g.d0)) - 1u)));
  const int side = g.wu0 > g.sa ? SIDE_PLUS : (g.wu1 < g.sa ? SIDE_MINUS : SIDE_MIXED);
  const int wmin = g.wu0 >= g.na ? kBig
                 : (side == SIDE_PLUS ? g.wu0 - g.sa : (side == SIDE_MINUS ? g.sa - g.wu1 : 0));
  g.wimax = max(abs(g.wu0 - g.sa), abs(g.wu1 - g.sa));
  const bool fix_lo = lane == 0 && slice > 0, fix_hi = lane == 31 && slice < NS - 1;
  uint32_t eb = ereg + 16u * (uint32_t)(f * NS + slice);
  const uint32_t slots = ereg + 16u * (uint32_t)(4 * NS); // parity 0

  double F[4] = {0, 0, 0, 0};
  // source cell value s0 = 1*occ(source); it seeds the element on the front's axis
  if ((unsigned)(-g.d0) < (unsigned)kT) {
    const uint32_t w = __ldg(p.rowbits_map + sy * p.wpr + (sx >> 5));
    const double s0 = ((w >> (sx & 31)) & 1u) ? 1.0 : 0.0;
#pragma unroll
    for (int q = 0; q < kT; ++q)
      if (q == -g.d0) F[q] = s0;
    if (ISROW && f == 0) out[(size_t)sy * nx + sx] = to_out<OutT>(s0);
  }

  const int k0 = wmin + 1;                           // first ring that touches this slice
  const int kfirst = k0 <= g.Kend ? k0 : Kmax + 2;   // never active otherwise
  int k = 1;
  for (; k < kfirst && k <= Kmax + 1; ++k) __syncthreads(); // not reached yet
  if (kfirst > Kmax + 1) return;
  if (kfirst & 1) eb ^= p.edge_p2; // parity of the first real step

  const int pstride = g.dir * pitch;
  // line / row of ring kfirst; one step outside the grid when this warp only has the
  // final diagonal to store (such a pointer is formed but never dereferenced)
  const int T0 = g.st + g.dir * kfirst;
  const uint32_t *pp = plane + ((ptrdiff_t)T0 * pitch + widx);
  const int ostride = ISROW ? g.dir * nx : 0;
  OutT *op = ISROW ? out + ((ptrdiff_t)T0 * nx + u0) : out;
  if (side == SIDE_PLUS)
    front_rings<OutT, VEC, ISROW, SIDE_PLUS>(p, g, sx, sy, kfirst, Kmax, pp, pstride, out, op,
                                             ostride, tile, eb, slots, F, fix_lo, fix_hi);
  else if (side == SIDE_MINUS)
    front_rings<OutT, VEC, ISROW, SIDE_MINUS>(p, g, sx, sy, kfirst, Kmax, pp, pstride, out, op,
                                              ostride, tile, eb, slots, F, fix_lo, fix_hi);
  else
    front_rings<OutT, VEC, ISROW, SIDE_MIXED>(p, g, sx, sy, kfirst, Kmax, pp, pstride, out, op,
                                              ostride, tile, eb, slots, F, fix_lo, fix_hi);
}

template <typename OutT, bool VEC>
__global__ void __launch_bounds__(1024, 1) sweep_ring_kernel(const RingParams p) {
  constexpr int S = 32 / (int)sizeof(OutT);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int NS = blockDim.x >> 7; // slices per front (4 fronts)
  const int64_t pair = blockIdx.x;
  const int nx = p.nx, ny = p.ny;
  const int sx = __ldg(p.src_xy + 2 * pair), sy = __ldg(p.src_xy + 2 * pair + 1);
  if ((unsigned)sx >= (unsigned)nx || (unsigned)sy >= (unsigned)ny) { // CTA-uniform
    if (tid == 0) atomicOr(p.err, 1);
    return;
  }
  const int map = p.src_map ? __ldg(p.src_map + pair) : 0;
  RingParams q = p;
  q.rowbits_map = p.rowbits + (size_t)map * p.row_plane;
  const uint32_t *colbits = p.colbits + (size_t)map * p.col_plane;
  OutT *out = reinterpret_cast<OutT *>(p.out) + (size_t)pair * nx * ny;

  OutT *tiles = reinterpret_cast<OutT *>(smem_raw);
  constexpr int tile_elems = S * kPitch;
  const uint32_t P2 = p.edge_p2;
  const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(smem_raw);
  const uint32_t ereg =
      (smem0 + 2u * NS * tile_elems * (uint32_t)sizeof(OutT) + 2u * P2 - 1u) & ~(2u * P2 - 1u);
  for (uint32_t i = tid; i < 2u * P2 / 8u; i += blockDim.x) sts64(ereg + 8u * i, 0.0);
  __syncthreads();

  const int f = warp / NS, slice = warp - f * NS; // 0 RU, 1 RD, 2 CR, 3 CL
  if (f < 2)
    run_front<OutT, VEC, true>(q, sx, sy, q.rowbits_map, p.wpr, out, f, slice, NS,
                               (OutT *)nullptr, ereg);
  else
    run_front<OutT, VEC, false>(q, sx, sy, colbits, p.wpc, out, f, slice, NS,
                                tiles + ((f - 2) * NS + slice) * tile_elems, ereg);
}

template <typename OutT, bool VEC>
cudaError_t launch_ring(const RingParams &p, int64_t npairs, int nt, size_t smem, cudaStream_t st) {
  auto kern = sweep_ring_kernel<OutT, VEC>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<(unsigned)npairs, nt, smem, st>>>(p);
  return cudaGetLastError();
}

void ring_geometry(int nx, int ny, vhp_dtype dtype, int &nt, uint32_t &p2, size_t &smem) {
  const int n = std::max(nx, ny);
  const int ns = (n + kWSpan - 1) / kWSpan;
  nt = 4 * ns * 32;
  const size_t esz = dtype == VHP_F32 ? 4 : 8;
  const int S = 32 / (int)esz;
  p2 = 64;
  while (p2 < (uint32_t)(4 * ns) * 16u + 32u) p2 <<= 1;
  // staging tiles of the two column fronts + alignment slack + two parities of edge slots
  smem = 2 * (size_t)ns * S * kPitch * esz + 2 * (size_t)p2 + 2 * (size_t)p2;
}

} // namespace

bool vhp_sweep_ring_supported(int nx, int ny) { return std::max(nx, ny) <= 8 * kWSpan; }

cudaError_t vhp_launch_sweep_ring(const VhpPackedMaps &maps, int nx, int ny,
                                  const int32_t *d_src_xy, const int32_t *d_src_map,
                                  int64_t npairs, vhp_dtype dtype, void *d_out,
                                  const double *d_rcp, int *d_err, cudaStream_t st,
                                  int64_t *launches) {
  RingParams p;
  p.err = d_err;
  p.rowbits = maps.rowbits; p.colbits = maps.colbits;
  p.rowbits_map = nullptr;
  p.wpr = maps.wpr; p.wpc = maps.wpc;
  p.row_plane = maps.row_plane; p.col_plane = maps.col_plane;
  p.nx = nx; p.ny = ny;
  p.src_xy = d_src_xy; p.src_map = d_src_map;
  p.out = d_out; p.rcp = d_rcp;
  int nt; size_t smem;
  ring_geometry(nx, ny, dtype, nt, p.edge_p2, smem);
  if (nt > 1024) return cudaErrorInvalidConfiguration;
  // 128-bit stores need 16-byte aligned rows: nx % 4 == 0 (f32) / nx % 2 == 0 (f64)
  const bool vec = (dtype == VHP_F32) ? (nx % 4 == 0) : (nx % 2 == 0);
  cudaError_t e;
  if (dtype == VHP_F32)
    e = vec ? launch_ring<float, true>(p, npairs, nt, smem, st)
            : launch_ring<float, false>(p, npairs, nt, smem, st);
  else
    e = vec ? launch_ring<double, true>(p, npairs, nt, smem, st)
            : launch_ring<double, false>(p, npairs, nt, smem, st);
  if (launches) *launches += 1;
  return e;
}
