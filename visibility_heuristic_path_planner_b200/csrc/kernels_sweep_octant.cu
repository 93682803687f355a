// kernels_sweep_octant.cu -- K1, batched visibility sweep: one warp per octant,
// no block barriers.
//
// Replaces visibilityBasedSolver::computeVisibility
// (reference src/visibilityBasedSolver.cpp:570-696) for grids up to 1021 x 1021.
// Same L-front dynamic program and arithmetic contract as the other sweep
// kernels (sweep_common.cuh): v = a - c*(a - b), c = d/k, every operation rounded
// once, times the occupancy bit.
//
// Decomposition.  One CTA (8 warps) per (map, source) pair, one warp per octant:
//     rows     RU+ RU- RD+ RD-   cells (sx +- d, sy +- k)   warps 0..3
//     columns  CR+ CR- CL+ CL-   cells (sx +- k, sy +- d)   warps 4..7
// k = ring index ("step", the coordinate across the front), d = offset along the
// front, 0 <= d < k.  Inside an octant cell (d, k) depends on (d, k-1) and
// (d-1, k-1) only, so a warp keeps its whole front in fp64 registers (NS slices
// of 128 offsets, lane l owns offsets 128m + 4l + e - phi), advances one ring
// per iteration and needs one 64-bit shuffle per slice for the upstream
// neighbour.  Warps never wait on a block barrier: the only cross-warp
// dependency is the diagonal cell q[k][k] = q[k][k-1] * occ(k,k) (the reference
// has no i == j branch), which the column octant of a
// quadrant computes and hands to the row octant of the same quadrant through a
// small shared-memory ring (the value itself is the full/empty flag).
//
// Occupancy comes from four bit planes per map (forward / mirrored, row / column
// major), one 32-bit word per lane per ring, with the reference's never-written
// border baked in as "occupied".  The mirrored planes let
// the "-" octants run the same code as the "+" octants.
//
// Stores.  Row octants own 4 consecutive, 16-byte aligned x per thread (phi
// shifts the ownership so that the groups are aligned) and write one 128-bit
// store per slice and ring.  Column octants produce one x per ring: every thread
// parks its converted values in a private shared-memory slot for S = 32 B /
// sizeof(OutT) rings and then writes one full 32-byte sector per row.
//
// c = d/k is computed as fma(d, rh, d*rl) with (rh, rl) a double-double 1/k from
// a table: correctly rounded for all 0 <= d < k <= 16384 (verified exhaustively,
// vhp_selftest_ratio / tests) at two fp64 operations instead of a division.
#include <cstdint>

#include "vhp_internal.h"
#include "sweep_common.cuh"

namespace {

#ifndef VHP_OCT_MINB
#define VHP_OCT_MINB (NS >= 8 ? 12 : 16)
#endif
constexpr unsigned kFull = 0xffffffffu;
constexpr unsigned long long kEmpty = ~0ull;    // a NaN pattern no visibility value has

struct OctArgs {
  const uint32_t *row_f, *row_r, *col_f, *col_r; // bit planes [map][line][4*NS words]
  size_t row_plane, col_plane;                   // words per map
  const uint8_t *occ;                            // byte maps (source cell only)
  int nx, ny;
  const int32_t *src_xy, *src_map;
  void *out;
  const double2 *rtab;                           // {RN(1/k), 1/k - RN(1/k)}
  int *err;
};

__device__ __forceinline__ double rot_up(double v, int lane) {
  // lane l receives the value of lane (l - 1) & 31
  return __shfl_sync(kFull, v, (lane + 31) & 31);
}

__device__ __forceinline__ unsigned long long ld_slot(uint32_t a) {
  unsigned long long v;
  asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void st_slot(uint32_t a, unsigned long long v) {
  asm volatile("st.volatile.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory");
}

// one ring of one slice for the thread's 4 offsets; `mask` = occupancy & activity
// bits.  Descending e keeps F[e-1] at its previous-ring value.
template <int OFF>
__device__ __forceinline__ void slice_update(double (&F)[4], const double nb, const uint32_t mask,
                                             const double fd0, const double rh, const double rl) {
#pragma unroll
  for (int e = 3; e >= 0; --e) {
    const double b = e ? F[e > 0 ? e - 1 : 0] : nb;
    const double fd = __dadd_rn(fd0, (double)(OFF + e));
    const double c = __fma_rn(fd, rh, __dmul_rn(fd, rl));
    const double v = lerp_rn(F[e], b, c);
    F[e] = ((mask >> e) & 1u) ? v : 0.0;
  }
}

// 16-byte global store of 4 floats / 2 doubles with the streaming hint
__device__ __forceinline__ void stg16(float *p, float a, float b, float c, float d) {
  __stcs(reinterpret_cast<float4 *>(p), make_float4(a, b, c, d));
}
__device__ __forceinline__ void stg16(double *p, double a, double b) {
  __stcs(reinterpret_cast<double2 *>(p), make_double2(a, b));
}

// store the thread's 4 values of one row-octant slice: p points at offset e = 0,
// estep = +1 (forward) / -1 (mirrored octant); `m` = store-eligible elements
template <typename OutT, bool VEC>
__device__ __forceinline__ void row_store(OutT *p, const int estep, const double (&F)[4],
                                          const uint32_t m) {
  if (VEC && m == 0xFu) {
    if (estep > 0) {
      if constexpr (sizeof(OutT) == 4) {
        stg16(p, to_out<OutT>(F[0]), to_out<OutT>(F[1]), to_out<OutT>(F[2]), to_out<OutT>(F[3]));
      } else {
        stg16(p, to_out<OutT>(F[0]), to_out<OutT>(F[1]));
        stg16(p + 2, to_out<OutT>(F[2]), to_out<OutT>(F[3]));
      }
    } else {
      if constexpr (sizeof(OutT) == 4) {
        stg16(p - 3, to_out<OutT>(F[3]), to_out<OutT>(F[2]), to_out<OutT>(F[1]), to_out<OutT>(F[0]));
      } else {
        stg16(p - 3, to_out<OutT>(F[3]), to_out<OutT>(F[2]));
        stg16(p - 1, to_out<OutT>(F[1]), to_out<OutT>(F[0]));
      }
    }
  } else if (m) {
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if ((m >> e) & 1u) __stcs(p + e * estep, to_out<OutT>(F[e]));
  }
}

// park the thread's 4 converted values of one column-octant slice (private slot)
template <typename OutT>
__device__ __forceinline__ void col_park(uint32_t slot, const double (&F)[4]) {
  if constexpr (sizeof(OutT) == 4) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(slot), "f"(to_out<float>(F[0])),
                 "f"(to_out<float>(F[1])), "f"(to_out<float>(F[2])), "f"(to_out<float>(F[3]))
                 : "memory");
  } else {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(slot), "d"(F[0]), "d"(F[1]) : "memory");
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(slot + 512u), "d"(F[2]), "d"(F[3])
                 : "memory");
  }
}

// geometry of one octant (warp-uniform)
struct Oct {
  int dir_ac;     // +1 / -1 across rings
  int rev;        // 1 for the mirrored ("-") side along the front
  int s_ac;       // source coordinate across
  int K;          // rings 1..K
  int Dlim;       // largest offset d inside the grid
  int phi;        // alignment shift: d = D - phi
  int base;       // plane coordinate of D = 0
  int nsl;        // slices that hold grid cells
};

// ---------------------------------------------------------------------------------
// One octant, all rings.  ISROW: front runs along x (rows sy +- k); otherwise along y.
//
// Register layout: F[j] is the slice j places below the edge slice (the slice that
// holds the newest offset); when the edge moves into the next slice the array is
// renamed (F[j+1] = F[j]) so the edge code exists once (j == 0) and the interior
// slices are one fall-through chain entered at j = me.
// ---------------------------------------------------------------------------------
// 4 ones (fp32) / 2 ones (fp64) per 16-byte store
template <typename OutT> __device__ __forceinline__ void stg16_ones(OutT *p) {
  if constexpr (sizeof(OutT) == 4) stg16(p, 1.0f, 1.0f, 1.0f, 1.0f);
  else { stg16(p, 1.0, 1.0); stg16(p + 2, 1.0, 1.0); }
}

template <typename OutT, int NS, bool VEC, bool ISROW>
__device__ __forceinline__ void octant_sweep(const OctArgs &p, const Oct &g, const int sx,
                                             const double s0, const uint32_t *__restrict__ plane,
                                             OutT *__restrict__ out, const uint32_t diag,
                                             const uint32_t stage) {
  constexpr int S = 4;                            // rings parked per flush (columns)
  constexpr int WP = 4 * NS;                      // plane words per line
  constexpr uint32_t kSlotBytes = sizeof(OutT) == 4 ? 512u : 1024u; // one ring of one slice
  constexpr uint32_t kSliceBytes = S * kSlotBytes;
  const int lane = threadIdx.x & 31;
  const int nx = p.nx;
  const int K = g.K, phi = g.phi, Dlim = g.Dlim, dir = g.dir_ac;
  const int estep = g.rev ? -1 : 1;

  // occupancy nibble of slice m = bits of plane coordinate base + 128m + 4l .. +3
  const int nbl = (g.base >> 2) + lane;
  const int nshift = (nbl & 7) * 4;
  int src_top = nbl >> 3;                         // + 4*me
  double fd_top = (double)(4 * lane - phi);       // + 128*me

  // store-eligible elements per slice (4 bits each): 0 <= d <= Dlim, and the axis
  // d == 0 belongs to the "+" octant
  uint32_t sm = 0;
#pragma unroll
  for (int m = 0; m < NS; ++m)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int d = 128 * m + 4 * lane + e - phi;
      if (d >= (g.rev ? 1 : 0) && d <= Dlim) sm |= 1u << (4 * m + e);
    }
  uint32_t smr = sm & 0xFu;                       // nibble j = slice me - j

  double F[NS][4];
#pragma unroll
  for (int j = 0; j < NS; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) F[j][e] = (j == 0 && lane == 0 && e == phi) ? s0 : 0.0;

  // real coordinate along the front of element e = 0 of slice 0
  const int al0 = g.rev ? (128 * NS - 1) - (g.base + 4 * lane) : g.base + 4 * lane;
  // rows: (al0, source row); columns: row of al0
  OutT *const org = ISROW ? out + ((ptrdiff_t)g.s_ac * nx + al0) : out + (ptrdiff_t)al0 * nx;
  // rows: pointer to (element 0 of the edge slice, current row); columns: row of that element
  OutT *ptop = ISROW ? org + (ptrdiff_t)dir * nx : org;
  const ptrdiff_t rstep = ISROW ? (ptrdiff_t)dir * nx : 0;
  const uint32_t slot0 = stage + 16u * (uint32_t)lane;
  uint32_t slot_lane = slot0;                     // + me * kSliceBytes

  const uint32_t *pl = plane + (ptrdiff_t)(g.s_ac + dir) * WP + lane;
  uint32_t wnext = (lane < WP) ? __ldg(pl) : 0u;
  int t = g.s_ac;
  int me_cur = 0;
  int blk_lo = 1;   // first ring parked in the current block (columns)
  int avail = 0;    // rows: diagonal entries known to be published
  // last diagonal entry that is handed over: the row octant reads entries 1..min(Ky-1, Ex)
  const int jlast = ISROW ? min(K - 1, Dlim) : min(Dlim - 1, K);
  // "lit" mode: every cell of the front so far is exactly 1.0 (no obstacle met yet), so a
  // ring whose occupancy bits are all set reproduces 1.0 everywhere: a - c*(a - b) with
  // a == b == 1 is exactly 1.  The front registers are only materialised when it ends.
  bool lit = s0 == 1.0;
  const int sp = g.base + phi; // plane coordinate of the source

#define VHP_SLICE(j)                                                                         \
  {                                                                                          \
    const double r_ = rot_up(F[j][3], lane);                                                 \
    const double nb_ = lane ? r_ : rprev;                                                    \
    rprev = r_;                                                                              \
    const uint32_t nib_ = __shfl_sync(kFull, wcur, src_top - 4 * (j)) >> nshift;             \
    slice_update<-128 * (j)>(F[j], nb_, nib_, fd_top, rh, rl);                               \
    if (ISROW)                                                                               \
      row_store<OutT, VEC>(ptop - estep * (128 * (j)), estep, F[j], (smr >> (4 * (j))) & 0xFu); \
    else                                                                                     \
      col_park<OutT>(slot - (uint32_t)(j) * kSliceBytes, F[j]);                              \
  }

#pragma unroll 1
  for (int k = 1; k <= K; ++k) {
    t += dir;
    const uint32_t wcur = wnext;
    pl += dir * WP;
    if (k < K && lane < WP) wnext = __ldg(pl);

    // edge slice: rows -> slice of the last interpolated offset d = k-1;
    // columns -> slice of the diagonal cell d = k
    const int dstar = ISROW ? k - 1 + phi : k + phi;
    const bool has_edge = ISROW ? (k - 1 <= Dlim) : (k <= Dlim);
    const int me = has_edge ? dstar >> 7 : g.nsl - 1; // top slice of this ring

    double dval = 0.0;
    if (ISROW && has_edge && k >= 2) { // diagonal cell of ring k-1 from the column octant
      if (k - 1 > avail) {             // wait in batches: the producer publishes in order
        const int target = min(k + 6, jlast);
        const uint32_t a = diag + 8u * (uint32_t)target;
        unsigned ns = 256;
        while (ld_slot(a) == kEmpty) {
          __nanosleep(ns);
          ns = min(2 * ns, 4096u);
        }
        avail = target;
      }
      const uint32_t a = diag + 8u * (uint32_t)(k - 1);
      unsigned long long raw = ld_slot(a);
      while (raw == kEmpty) raw = ld_slot(a);
      dval = __longlong_as_double((long long)raw);
    }

    bool general = true;
    if (lit) {
      // all occupancy bits of the active offsets 0..dmax set?  (lane j holds plane word j)
      const int dmax = min(ISROW ? k - 1 : k, Dlim);
      const int lo = max(sp - 32 * lane, 0), hi = min(sp + dmax - 32 * lane, 31);
      const uint32_t need = lo <= hi ? ((2u << hi) - 1u) & ~((1u << lo) - 1u) : 0u;
      bool ok = (wcur & need) == need;
      if (ISROW && has_edge && k >= 2) ok = ok && dval == 1.0;
      if (__all_sync(kFull, ok)) {
        general = false;
        const int r0 = dstar - 128 * me - 4 * lane;
        if (ISROW) {
          OutT *q = org + (ptrdiff_t)(dir * k) * nx;
          const uint32_t wedge = !has_edge || r0 >= 3 ? 0xFu : (r0 < 0 ? 0u : (2u << r0) - 1u);
#pragma unroll 1
          for (int m = 0; m <= me; ++m, q += estep * 128) {
            uint32_t msk = (sm >> (4 * m)) & 0xFu;
            if (m == me) msk &= wedge;
            if (VEC && msk == 0xFu) {
              stg16_ones<OutT>(estep > 0 ? q : q - 3);
            } else if (msk) {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if ((msk >> e) & 1u) __stcs(q + e * estep, (OutT)1);
            }
          }
        } else {
          uint32_t a = slot0 + (uint32_t)(t & (S - 1)) * kSlotBytes;
#pragma unroll 1
          for (int m = 0; m <= me; ++m, a += kSliceBytes) {
            if constexpr (sizeof(OutT) == 4) {
              asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(a), "f"(1.0f) : "memory");
            } else {
              asm volatile("st.shared.v2.f64 [%0], {%1, %1};" ::"r"(a), "d"(1.0) : "memory");
              asm volatile("st.shared.v2.f64 [%0], {%1, %1};" ::"r"(a + 512u), "d"(1.0) : "memory");
            }
          }
          if (has_edge && k <= jlast && r0 >= 0 && r0 < 4)
            asm volatile("st.relaxed.cluster.shared::cluster.u64 [%0], %1;" ::"r"(diag + 8u * (uint32_t)k),
                         "l"(__double_as_longlong(1.0)) : "memory");
        }
      } else {
        // leave lit mode: build the front as it stands after ring k-1 (offsets 0..dprev are
        // 1.0, everything else 0) in the renamed layout the general code expects
        lit = false;
        const int dsp = ISROW ? k - 2 + phi : k - 1 + phi;             // dstar of ring k-1
        const bool edge_prev = ISROW ? (k - 2 <= Dlim) : (k - 1 <= Dlim);
        me_cur = edge_prev ? max(dsp, 0) >> 7 : g.nsl - 1;
        const int dprev = ISROW ? max(min(k - 2, Dlim), 0) : min(k - 1, Dlim);
        smr = 0;
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          const int m = me_cur - j;
          if (m >= 0) smr |= ((sm >> (4 * m)) & 0xFu) << (4 * j);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int d = 128 * m + 4 * lane + e - phi;
            F[j][e] = (m >= 0 && d >= 0 && d <= dprev) ? 1.0 : 0.0;
          }
        }
        src_top = (nbl >> 3) + 4 * me_cur;
        fd_top = (double)(4 * lane - phi + 128 * me_cur);
        ptop = ISROW ? org + (ptrdiff_t)(dir * k) * nx + estep * 128 * me_cur
                     : org + (ptrdiff_t)(estep * 128 * me_cur) * nx;
        slot_lane = slot0 + (uint32_t)me_cur * kSliceBytes;
      }
    }

    if (general) {
      const double2 rr = __ldg(p.rtab + k);
      const double rh = rr.x, rl = rr.y;
      if (has_edge && me != me_cur) { // the edge enters the next slice: rename
#pragma unroll
        for (int j = NS - 1; j > 0; --j)
#pragma unroll
          for (int e = 0; e < 4; ++e) F[j][e] = F[j - 1][e];
#pragma unroll
        for (int e = 0; e < 4; ++e) F[0][e] = 0.0;
        ++me_cur;
        smr = (smr << 4) | ((sm >> (4 * me_cur)) & 0xFu);
        src_top += 4;
        fd_top = __dadd_rn(fd_top, 128.0);
        ptop += estep * 128 * (ISROW ? 1 : nx);
        slot_lane += kSliceBytes;
      }

      double rprev = 0.0;
      const uint32_t slot = slot_lane + (uint32_t)(t & (S - 1)) * kSlotBytes;
      // ---- interior slices, ascending offsets (fall-through chain) ---------------
      switch (me_cur) {
        case 7: if (NS > 7) VHP_SLICE(NS > 7 ? 7 : 0)
        case 6: if (NS > 6) VHP_SLICE(NS > 6 ? 6 : 0)
        case 5: if (NS > 5) VHP_SLICE(NS > 5 ? 5 : 0)
        case 4: if (NS > 4) VHP_SLICE(NS > 4 ? 4 : 0)
        case 3: if (NS > 3) VHP_SLICE(NS > 3 ? 3 : 0)
        case 2: if (NS > 2) VHP_SLICE(NS > 2 ? 2 : 0)
        case 1: if (NS > 1) VHP_SLICE(NS > 1 ? 1 : 0)
        default: break;
      }
      if (!has_edge) {
        VHP_SLICE(0)
      } else {
        // ---- edge slice ----------------------------------------------------------
        const double r = rot_up(F[0][3], lane);
        const double nb = lane ? r : rprev;
        const uint32_t nib = __shfl_sync(kFull, wcur, src_top) >> nshift;
        const int r0 = dstar - 128 * me_cur - 4 * lane;
        if (ISROW) {
          // offsets d <= k-1 are interpolated; d == k-1 starts from the diagonal cell
          const uint32_t wedge = r0 >= 3 ? 0xFu : (r0 < 0 ? 0u : (2u << r0) - 1u);
          if (k >= 2) {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (e == r0) F[0][e] = dval;
          }
          slice_update<0>(F[0], nb, nib & wedge, fd_top, rh, rl);
          row_store<OutT, VEC>(ptop, estep, F[0], smr & wedge);
        } else {
          // offsets d <= k-1 are interpolated, d == k is the diagonal cell
          const uint32_t wedge = r0 >= 4 ? 0xFu : (r0 <= 0 ? 0u : (1u << r0) - 1u);
          slice_update<0>(F[0], nb, nib & wedge, fd_top, rh, rl);
          double pz = 0.0;
          if ((dstar & 3) == 0) { // the diagonal offset is some lane's e == 0
            const double pr = rot_up(F[0][3], lane);
            const double p0 = __shfl_sync(kFull, F[NS > 1 ? 1 : 0][3], 31);
            pz = lane ? pr : p0;
          }
          if ((unsigned)r0 < 4u) {
            const double pv = r0 == 0 ? pz : (r0 == 1 ? F[0][0] : (r0 == 2 ? F[0][1] : F[0][2]));
            const double dk = ((nib >> r0) & 1u) ? pv : 0.0;
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (e == r0) F[0][e] = dk;
            // hand the diagonal cell to the row octant of this quadrant (its shared memory)
            if (k <= jlast)
              asm volatile("st.relaxed.cluster.shared::cluster.u64 [%0], %1;" ::"r"(diag + 8u * (uint32_t)k),
                           "l"(__double_as_longlong(dk)) : "memory");
          }
          col_park<OutT>(slot, F[0]);
        }
      }
      if (ISROW) ptop += rstep;
    }
    if (!ISROW) {
      // ---- flush the parked block when it is complete ----------------------------
      const int kk = t & (S - 1);
      const bool last = (dir > 0 ? kk == S - 1 : kk == 0) || k == K;
      if (last) {
        const int xb = t & ~(S - 1);
        const bool full = (k - blk_lo + 1) == S;
        OutT *rowp = org + xb; // slice 0, element 0
#pragma unroll 1
        for (int m = 0; m <= me; ++m, rowp += (ptrdiff_t)(estep * 128) * nx) {
          const uint32_t sl = stage + (uint32_t)m * kSliceBytes + 16u * (uint32_t)lane;
          const uint32_t smm = (sm >> (4 * m)) & 0xFu;
          const int d0 = 128 * m + 4 * lane - phi;
          if (d0 > k || smm == 0u) continue; // nothing of this thread is inside the wedge yet
          OutT *q = rowp;
          if (VEC && full && smm == 0xFu && d0 + 3 <= blk_lo) {
            if constexpr (sizeof(OutT) == 4) {
              float4 v[4];
#pragma unroll
              for (int j = 0; j < 4; ++j)
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(v[j].x), "=f"(v[j].y), "=f"(v[j].z), "=f"(v[j].w)
                             : "r"(sl + (uint32_t)j * kSlotBytes));
              stg16(q, v[0].x, v[1].x, v[2].x, v[3].x);
              stg16(q + (ptrdiff_t)estep * nx, v[0].y, v[1].y, v[2].y, v[3].y);
              stg16(q + (ptrdiff_t)estep * 2 * nx, v[0].z, v[1].z, v[2].z, v[3].z);
              stg16(q + (ptrdiff_t)estep * 3 * nx, v[0].w, v[1].w, v[2].w, v[3].w);
            } else {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                double2 lo[2], hi[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                  const uint32_t a = sl + (uint32_t)(2 * h + j) * kSlotBytes;
                  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];"
                               : "=d"(lo[j].x), "=d"(lo[j].y) : "r"(a));
                  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];"
                               : "=d"(hi[j].x), "=d"(hi[j].y) : "r"(a + 512u));
                }
                stg16(q + 2 * h, lo[0].x, lo[1].x);
                stg16(q + (ptrdiff_t)estep * nx + 2 * h, lo[0].y, lo[1].y);
                stg16(q + (ptrdiff_t)estep * 2 * nx + 2 * h, hi[0].x, hi[1].x);
                stg16(q + (ptrdiff_t)estep * 3 * nx + 2 * h, hi[0].y, hi[1].y);
              }
            }
          } else {
            // slow path: block not full, diagonal band, or grid edge
#pragma unroll 1
            for (int j = 0; j < S; ++j) {
              const int kj = dir * (xb + j - sx);
              if (kj < blk_lo || kj > k) continue;
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                if (((smm >> e) & 1u) && d0 + e <= kj) {
                  OutT val;
                  const uint32_t a = sl + (uint32_t)j * kSlotBytes;
                  if constexpr (sizeof(OutT) == 4)
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(val) : "r"(a + 4u * e));
                  else
                    asm volatile("ld.shared.f64 %0, [%1];"
                                 : "=d"(val) : "r"(a + (e >> 1) * 512u + (e & 1) * 8u));
                  __stcs(q + (ptrdiff_t)(estep * e) * nx + j, val);
                }
              }
            }
          }
        }
        blk_lo = k + 1;
      }
    }
  }
#undef VHP_SLICE
}

// One warp-sized CTA per octant; the two octants of a quadrant form a cluster:
// rank 0 = column octant (produces the diagonal cells), rank 1 = row octant
// (consumes them; the hand-off array lives in ITS shared memory and the column
// octant writes it through the cluster's distributed shared memory).  Everything
// that selects the role derives from blockIdx, so the compiler can prove the
// control flow warp-uniform.
template <typename OutT, int NS, bool VEC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(32, VHP_OCT_MINB)
sweep_octant_kernel(const OctArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int W = 128 * NS;
  const int lane = threadIdx.x;
  const bool isrow = blockIdx.x & 1;
  const int qx = (blockIdx.x >> 1) & 1, qy = (blockIdx.x >> 2) & 1; // 1 = the "-" side
  const int64_t pair = blockIdx.x >> 3;
  const int nx = p.nx, ny = p.ny;
  const int sx = __ldg(p.src_xy + 2 * pair), sy = __ldg(p.src_xy + 2 * pair + 1);
  if ((unsigned)sx >= (unsigned)nx || (unsigned)sy >= (unsigned)ny) { // uniform over the cluster
    if (lane == 0) atomicOr(p.err, 1);
    return;
  }
  const int Ex = qx ? sx : nx - 1 - sx, Ey = qy ? sy : ny - 1 - sy; // quadrant extents
  const int map = p.src_map ? __ldg(p.src_map + pair) : 0;
  OutT *out = reinterpret_cast<OutT *>(p.out) + (size_t)pair * nx * ny;
  const double s0 = __ldg(p.occ + ((size_t)map * ny + sy) * nx + sx) ? 1.0 : 0.0;
  if ((blockIdx.x & 7) == 1 && lane == 0) out[(size_t)sy * nx + sx] = to_out<OutT>(s0);

  const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(smem_raw);
  if (isrow) {
    const int nd = min(Ex, Ey) + 1;
    for (int i = lane; i < nd; i += 32) st_slot(smem0 + 8u * i, kEmpty);
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;"
               ::: "memory");

  Oct g;
  g.dir_ac = (isrow ? qy : qx) ? -1 : 1;
  g.rev = isrow ? qx : qy;
  g.s_ac = isrow ? sy : sx;
  const int s_al = isrow ? sx : sy;
  g.K = isrow ? Ey : Ex;
  g.Dlim = isrow ? Ex : Ey;
  const int sp = g.rev ? W - 1 - s_al : s_al; // plane coordinate of the source
  g.phi = sp & 3;
  g.base = sp & ~3;
  g.nsl = ((g.Dlim + g.phi) >> 7) + 1;
  if (g.K == 0 || (g.rev && g.Dlim == 0)) return;

  if (isrow) {
    const uint32_t *plane = (g.rev ? p.row_r : p.row_f) + (size_t)map * p.row_plane;
    octant_sweep<OutT, NS, VEC, true>(p, g, sx, s0, plane, out, smem0, 0u);
  } else {
    // the row octant's copy of smem0 (cluster rank 1)
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem0), "r"(1));
    const uint32_t *plane = (g.rev ? p.col_r : p.col_f) + (size_t)map * p.col_plane;
    octant_sweep<OutT, NS, VEC, false>(p, g, sx, s0, plane, out, remote, smem0);
  }
}

// ---------------------------------------------------------------------------------
// bit planes: forward / mirrored, row / column major, W = 128*NS bits per line.
// Baked border (the cells the reference never writes are treated as occupied):
//   row planes: line y = 0 is all zero; mirrored plane also x = 0
//   col planes: line x = 0 is all zero; mirrored plane also y = 0
// (the forward planes keep x = 0 / y = 0 because a source on the border uses them
// as its axis).
// ---------------------------------------------------------------------------------
__global__ void pack_oct_rows_kernel(const uint8_t *__restrict__ occ, int nmaps, int nx, int ny,
                                     int wp, uint32_t *__restrict__ row_f,
                                     uint32_t *__restrict__ row_r) {
  const size_t total = (size_t)nmaps * ny * wp;
  const int W = 32 * wp;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int w = (int)(idx % wp);
    const size_t my = idx / wp;
    const int y = (int)(my % ny);
    const uint8_t *row = occ + my * nx;
    uint32_t f = 0, r = 0;
    if (y != 0) {
      for (int b = 0; b < 32; ++b) {
        const int xf = 32 * w + b, xr = W - 1 - xf;
        if (xf < nx && row[xf] != 0) f |= 1u << b;
        if (xr < nx && xr > 0 && row[xr] != 0) r |= 1u << b;
      }
    }
    row_f[idx] = f;
    row_r[idx] = r;
  }
}

__global__ void pack_oct_cols_kernel(const uint8_t *__restrict__ occ, int nmaps, int nx, int ny,
                                     int wp, uint32_t *__restrict__ col_f,
                                     uint32_t *__restrict__ col_r) {
  // x fastest across threads so the strided byte reads coalesce
  const size_t total = (size_t)nmaps * wp * nx;
  const int W = 32 * wp;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(idx % nx);
    const size_t mw = idx / nx;
    const int w = (int)(mw % wp);
    const size_t m = mw / wp;
    const uint8_t *base = occ + m * (size_t)nx * ny + x;
    uint32_t f = 0, r = 0;
    if (x != 0) {
      for (int b = 0; b < 32; ++b) {
        const int yf = 32 * w + b, yr = W - 1 - yf;
        if (yf < ny && base[(size_t)yf * nx] != 0) f |= 1u << b;
        if (yr < ny && yr > 0 && base[(size_t)yr * nx] != 0) r |= 1u << b;
      }
    }
    const size_t o = (m * nx + x) * wp + w;
    col_f[o] = f;
    col_r[o] = r;
  }
}

// table[k] = {rh, rl}: rh = RN(1/k), rl = RN((1 - rh*k) * rh)
__global__ void rcp2_table_kernel(double2 *table, int len) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < len) {
    double rh = 0.0, rl = 0.0;
    if (k > 0) {
      const double fk = (double)k;
      rh = __drcp_rn(fk);
      rl = __dmul_rn(__fma_rn(-rh, fk, 1.0), rh);
    }
    table[k] = make_double2(rh, rl);
  }
}

// diagnostic: count (d, k), 0 <= d < k <= kmax, where fma(d, rh, d*rl) != d/k
__global__ void ratio2_selftest_kernel(const double2 *__restrict__ tab, int kmax,
                                       unsigned long long *mismatches) {
  const int k = blockIdx.x + 1;
  if (k > kmax) return;
  const double fk = (double)k;
  const double2 rr = tab[k];
  unsigned long long bad = 0;
  for (int d = threadIdx.x; d < k; d += blockDim.x) {
    const double fd = (double)d;
    if (__fma_rn(fd, rr.x, __dmul_rn(fd, rr.y)) != __ddiv_rn(fd, fk)) ++bad;
  }
  if (bad) atomicAdd(mismatches, bad);
}

int oct_ns(int nx, int ny) {
  const int n = std::max(nx, ny);
  for (int ns : {2, 4, 8})
    if (n + 3 <= 128 * ns) return ns;
  return 0;
}

template <typename OutT, int NS, bool VEC>
cudaError_t launch_oct(const OctArgs &p, int64_t npairs, size_t smem, cudaStream_t st) {
  auto kern = sweep_octant_kernel<OutT, NS, VEC>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<(unsigned)(npairs * 8), 32, smem, st>>>(p);
  return cudaGetLastError();
}

template <typename OutT, bool VEC>
cudaError_t launch_oct_ns(int ns, const OctArgs &p, int64_t npairs, size_t smem, cudaStream_t st) {
  switch (ns) {
    case 2: return launch_oct<OutT, 2, VEC>(p, npairs, smem, st);
    case 4: return launch_oct<OutT, 4, VEC>(p, npairs, smem, st);
    case 8: return launch_oct<OutT, 8, VEC>(p, npairs, smem, st);
  }
  return cudaErrorInvalidConfiguration;
}

} // namespace

bool vhp_sweep_octant_supported(int nx, int ny) { return oct_ns(nx, ny) != 0; }

int vhp_oct_words_per_line(int nx, int ny) { return 4 * oct_ns(nx, ny); }

cudaError_t vhp_launch_pack_oct(const uint8_t *d_occ, int nmaps, int nx, int ny, uint32_t *row_f,
                                uint32_t *row_r, uint32_t *col_f, uint32_t *col_r,
                                cudaStream_t st, int64_t *launches) {
  const int wp = vhp_oct_words_per_line(nx, ny);
  const size_t tr = (size_t)nmaps * ny * wp, tc = (size_t)nmaps * nx * wp;
  const int bs = 256;
  const unsigned gr = (unsigned)std::min<size_t>((tr + bs - 1) / bs, 148u * 32u);
  const unsigned gc = (unsigned)std::min<size_t>((tc + bs - 1) / bs, 148u * 32u);
  pack_oct_rows_kernel<<<gr, bs, 0, st>>>(d_occ, nmaps, nx, ny, wp, row_f, row_r);
  pack_oct_cols_kernel<<<gc, bs, 0, st>>>(d_occ, nmaps, nx, ny, wp, col_f, col_r);
  if (launches) *launches += 2;
  return cudaGetLastError();
}

cudaError_t vhp_launch_rcp2_table(double *d_table, int len, cudaStream_t st, int64_t *launches) {
  rcp2_table_kernel<<<(len + 255) / 256, 256, 0, st>>>(reinterpret_cast<double2 *>(d_table), len);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

cudaError_t vhp_launch_ratio2_selftest(const double *d_rcp2, int kmax,
                                       unsigned long long *d_mismatches, cudaStream_t st,
                                       int64_t *launches) {
  ratio2_selftest_kernel<<<kmax, 128, 0, st>>>(reinterpret_cast<const double2 *>(d_rcp2), kmax,
                                               d_mismatches);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

cudaError_t vhp_launch_sweep_octant(const VhpOctPlanes &pl, const uint8_t *d_occ, int nx, int ny,
                                    const int32_t *d_src_xy, const int32_t *d_src_map,
                                    int64_t npairs, vhp_dtype dtype, void *d_out,
                                    const double *d_rcp2, int *d_err, cudaStream_t st,
                                    int64_t *launches) {
  const int ns = oct_ns(nx, ny);
  if (!ns) return cudaErrorInvalidConfiguration;
  OctArgs p;
  p.row_f = pl.row_f; p.row_r = pl.row_r; p.col_f = pl.col_f; p.col_r = pl.col_r;
  p.row_plane = pl.row_plane; p.col_plane = pl.col_plane;
  p.occ = d_occ;
  p.nx = nx; p.ny = ny;
  p.src_xy = d_src_xy; p.src_map = d_src_map;
  p.out = d_out;
  p.rtab = reinterpret_cast<const double2 *>(d_rcp2);
  p.err = d_err;
  // diagonal hand-off array + the column octant's parking slots (S = 4 rings per slice)
  const size_t slice_bytes = dtype == VHP_F32 ? 2048 : 4096;
  // (the row octant uses the block as hand-off array, the column octant as parking slots)
  const size_t smem = std::max(8 * (size_t)(std::min(nx, ny) + 2),
                               (size_t)(((ny + 2) >> 7) + 1) * slice_bytes);
  const bool vec = (dtype == VHP_F32) ? (nx % 4 == 0) : (nx % 2 == 0);
  cudaError_t e;
  if (dtype == VHP_F32)
    e = vec ? launch_oct_ns<float, true>(ns, p, npairs, smem, st)
            : launch_oct_ns<float, false>(ns, p, npairs, smem, st);
  else
    e = vec ? launch_oct_ns<double, true>(ns, p, npairs, smem, st)
            : launch_oct_ns<double, false>(ns, p, npairs, smem, st);
  if (launches) *launches += 1;
  return e;
}
