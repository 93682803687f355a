// result_transport.cu -- device half of the packed result transport.
//
// The host-buffer entry points are PCIe-bound: 4096 sweeps of a 1000 x 1000 grid are 16.4 GB
// of fp32 results per call against a kernel that produces them in 2.5 ms.  Visibility fields
// are mostly flat (lit 1.0, shadow 0.0), so before a chunk of results leaves the device this
// kernel classifies every 128-byte unit of it as uniform (all elements 0.0 or all elements 1.0,
// bit for bit) or literal, and compacts the literal units into one stream.  Only the stream and
// two bits of meta data per unit cross PCIe; host threads rebuild the exact bytes
// (host_expand.cpp).  (A unit of equal elements of any other value is literal: they are rare, and
// shipping one element per unit to cover them doubled the bytes of the empty-grid batch.)
//
// Units are 128 bytes (32 fp32 / 16 fp64 cells): one warp-wide 16-byte load covers four units,
// eight loads make a mask word (32 units = 4 KB), kept in registers.  A unit is uniform when
// the eight lanes that hold it see one element value, 0.0 or 1.0, bit for bit.  The literal units of a word
// get consecutive slots of the stream from one atomicAdd on the chunk's cursor.
//
// Direct mode (the caller's buffer is pinned, mapped and 16-byte aligned): the literal units
// are not compacted but stored by this kernel straight to their final place in host memory, so
// the host threads only write the uniform units and no byte of the result is written to host
// memory twice.  A partial last unit of the chunk cannot be stored whole: it goes to the meta
// block and the host copies its valid bytes.  To balance PCIe against the host threads the
// device can also take whole mask words: in direct mode the words w with w % 16 < gpu_share
// (except the last word of the chunk) are stored completely by this kernel, uniform units
// included, and skipped by the host.
#include <algorithm>
#include <cstdint>

#include "vhp_internal.h"

namespace {

constexpr unsigned kAllLanes = 0xffffffffu;
constexpr int kLanesPerUnit = kVhpPackUnit / 16; // 8
constexpr int kUnitsPerLoad = 32 / kLanesPerUnit; // 4
constexpr int kLoadsPerWord = 32 / kUnitsPerLoad; // 8

// meta block: [cursor u64, pad to 16][tail unit][mask u32 x nwords][word_base u32 x nwords]
//             [vmask u32 x nwords]   (vmask bit u: the uniform unit u of the word is 1.0, else 0.0)
template <int ELEM> // element size in bytes: 4 or 8
__global__ void __launch_bounds__(256)
pack_results_kernel(const uint4 *__restrict__ in, const int64_t nunits, unsigned char *__restrict__ meta,
                    uint4 *__restrict__ lit, uint4 *__restrict__ host_dst, const int tail_partial,
                    const int gpu_share) {
  const int64_t nwords = (nunits + 31) / 32;
  unsigned long long *cursor = reinterpret_cast<unsigned long long *>(meta);
  uint4 *tail = reinterpret_cast<uint4 *>(meta + 16);
  uint32_t *mask = reinterpret_cast<uint32_t *>(meta + kVhpPackMetaHead);
  uint32_t *word_base = mask + nwords;
  uint32_t *vmask = word_base + nwords;
  const int lane = threadIdx.x & 31;
  const int grp = lane / kLanesPerUnit;        // which of the four units of a load this lane holds
  const int lead = grp * kLanesPerUnit;        // first lane of that unit
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n16 = nunits * kLanesPerUnit;  // readable 16-byte pieces
  for (int64_t w = warp0; w < nwords; w += nwarps) {
    const int64_t u0 = w * 32;
    const int64_t q0 = u0 * kLanesPerUnit + lane; // this lane's 16-byte piece of load 0
    uint4 v[kLoadsPerWord];
    uint32_t lit_bits = 0;                       // bit k: this lane's unit of load k is literal
    uint32_t one_bits = 0;                       // bit k: ... is uniform 1.0
#pragma unroll
    for (int k = 0; k < kLoadsPerWord; ++k) {
      const int64_t q = q0 + 32 * k;
      v[k] = q < n16 ? __ldg(in + q) : make_uint4(0u, 0u, 0u, 0u);
      const uint32_t f0 = __shfl_sync(kAllLanes, v[k].x, lead), f1 = __shfl_sync(kAllLanes, v[k].y, lead);
      const bool same = ELEM == 4 ? (v[k].x == f0 && v[k].y == f0 && v[k].z == f0 && v[k].w == f0)
                                  : (v[k].x == f0 && v[k].y == f1 && v[k].z == f0 && v[k].w == f1);
      const uint32_t b = __ballot_sync(kAllLanes, same);
      const bool zero = ELEM == 4 ? f0 == 0u : (f0 | f1) == 0u;
      const bool one = ELEM == 4 ? f0 == 0x3f800000u : (f0 == 0u && f1 == 0x3ff00000u);
      const bool uni = ((b >> lead) & 0xffu) == 0xffu && (zero || one);
      lit_bits |= (uni ? 0u : 1u) << k;
      one_bits |= ((uni && one) ? 1u : 0u) << k;
    }
    // mask bit of unit u0 + 4k + g  <-  lit_bits bit k of any lane of group g (vm, one_bits likewise)
    uint32_t m = 0, vm = 0;
#pragma unroll
    for (int k = 0; k < kLoadsPerWord; ++k) {
      const uint32_t b = __ballot_sync(kAllLanes, (lit_bits >> k) & 1u);
      const uint32_t bo = __ballot_sync(kAllLanes, (one_bits >> k) & 1u);
#pragma unroll
      for (int g = 0; g < kUnitsPerLoad; ++g) {
        m |= ((b >> (g * kLanesPerUnit)) & 1u) << (k * kUnitsPerLoad + g);
        vm |= ((bo >> (g * kLanesPerUnit)) & 1u) << (k * kUnitsPerLoad + g);
      }
    }
    if (u0 + 32 > nunits) { // units past the end of the chunk
      m &= (1u << (int)(nunits - u0)) - 1u;
      vm &= (1u << (int)(nunits - u0)) - 1u;
    }
    if (host_dst && (int)(w & 15) < gpu_share && w != nwords - 1) {
      // a word the device delivers completely (mask 0: nothing left for the host to copy)
#pragma unroll
      for (int k = 0; k < kLoadsPerWord; ++k) host_dst[q0 + 32 * k] = v[k];
      if (lane == 0) {
        mask[w] = 0u;
        word_base[w] = 0u;
        vmask[w] = 0u;
      }
      continue;
    }
    uint32_t base = 0;
    if (lane == 0 && m) base = (uint32_t)atomicAdd(cursor, (unsigned long long)__popc(m));
    base = __shfl_sync(kAllLanes, base, 0);
    if (lane == 0) {
      mask[w] = m;
      word_base[w] = base;
      vmask[w] = vm;
    }
#pragma unroll
    for (int k = 0; k < kLoadsPerWord; ++k) {
      const int u = k * kUnitsPerLoad + grp; // this lane's unit of load k
      if ((m >> u) & 1u) {
        if (host_dst) {
          if (tail_partial && u0 + u == nunits - 1) tail[lane - lead] = v[k];
          else host_dst[q0 + 32 * k] = v[k];
        } else {
          const uint32_t rank = __popc(m & ((1u << u) - 1u));
          lit[(size_t)(base + rank) * kLanesPerUnit + (lane - lead)] = v[k];
        }
      }
    }
  }
}

} // namespace

cudaError_t vhp_launch_pack_results(const void *d_in, int64_t nunits, int elem_bytes, void *d_meta,
                                    void *d_literals, void *host_dst, int tail_partial,
                                    int gpu_share, int sm_count, cudaStream_t st,
                                    int64_t *launches) {
  if (nunits <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(d_meta, 0, 16, st);
  if (e != cudaSuccess) return e;
  const int64_t nwords = (nunits + 31) / 32;
  const int64_t want = (nwords + 7) / 8; // 8 warps per CTA
  const unsigned grid = (unsigned)std::min<int64_t>(want, (int64_t)sm_count * 16);
  if (elem_bytes == 4)
    pack_results_kernel<4><<<grid, 256, 0, st>>>(reinterpret_cast<const uint4 *>(d_in), nunits,
                                                 reinterpret_cast<unsigned char *>(d_meta),
                                                 reinterpret_cast<uint4 *>(d_literals),
                                                 reinterpret_cast<uint4 *>(host_dst), tail_partial,
                                                 gpu_share);
  else
    pack_results_kernel<8><<<grid, 256, 0, st>>>(reinterpret_cast<const uint4 *>(d_in), nunits,
                                                 reinterpret_cast<unsigned char *>(d_meta),
                                                 reinterpret_cast<uint4 *>(d_literals),
                                                 reinterpret_cast<uint4 *>(host_dst), tail_partial,
                                                 gpu_share);
  if (launches) *launches += 1;
  return cudaGetLastError();
}


// ---- thresholded binary visibility (VHP binary output) ----------------------------------------
// bits[p][y][w], w < ceil(nx / 32): bit b of word w = (vis[p][y][32w + b] >= thr), decided on the
// fp64 value the sweep computed (the reference's `visibility_ >= threshold`, bit-exact; an fp32
// round trip could flip cells that sit on the threshold).  One warp per row, U words at a time:
// U coalesced loads in flight per lane, U ballots, one store of U words.
namespace {

__device__ __forceinline__ uint32_t transitions_of(const uint32_t w, const uint32_t prev_msb) {
  return w ^ ((w << 1) | prev_msb); // bit b set: cell b differs from cell b - 1 (cell -1 of a row = 0)
}

// row_cnt != null: also the number of transition columns of every row (see "row runs" below)
template <typename T, int U>
__global__ void __launch_bounds__(256)
threshold_bits_kernel(const T *__restrict__ vis, const int64_t nrows, const int nx, const int wpr,
                      const T thr, uint32_t *__restrict__ bits, uint16_t *__restrict__ row_cnt) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = warp0; row < nrows; row += nwarps) {
    const T *r = vis + row * nx;
    uint32_t *o = bits + row * wpr;
    uint32_t carry = 0, mine = 0; // carry: the top bit of the word before this group
    int cnt = 0, w0 = 0;
    for (; w0 < wpr; w0 += U) {
      T v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) { // U coalesced loads in flight per lane
        const int x = 32 * (w0 + u) + lane;
        v[u] = x < nx ? __ldcs(r + x) : (T)-1;
      }
      mine = 0;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t m = __ballot_sync(kAllLanes, 32 * (w0 + u) + lane < nx && v[u] >= thr);
        if (lane == u) mine = m;
      }
      const bool have = lane < U && w0 + lane < wpr;
      if (have) o[w0 + lane] = mine;
      if (row_cnt) {
        const uint32_t below = __shfl_up_sync(kAllLanes, mine, 1);
        if (have) cnt += __popc(transitions_of(mine, lane ? below >> 31 : carry));
        carry = __shfl_sync(kAllLanes, mine, U - 1) >> 31;
      }
    }
    if (row_cnt) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) cnt += __shfl_xor_sync(kAllLanes, cnt, off);
      // a run that reaches the last cell is closed at nx: the zero padding bit of the last word does
      // that by itself unless the row fills its last word
      const uint32_t last = __shfl_sync(kAllLanes, mine, (wpr - 1) - (w0 - U));
      if ((nx & 31) == 0) cnt += (int)(last >> 31);
      if (lane == 0) row_cnt[row] = (uint16_t)cnt;
    }
  }
}

} // namespace

cudaError_t vhp_launch_threshold_bits(const double *d_vis, int64_t nrows, int nx, double thr,
                                      uint32_t *d_bits, int sm_count, cudaStream_t st, int64_t *launches,
    uint16_t *d_row_cnt) {
  const int wpr = (nx + 31) / 32;
  int64_t blocks = (nrows + 7) / 8;
  if (blocks > (int64_t)sm_count * 8) blocks = (int64_t)sm_count * 8;
  if (blocks < 1) blocks = 1;
  threshold_bits_kernel<double, 4><<<(unsigned)blocks, 256, 0, st>>>(d_vis, nrows, nx, wpr, thr, d_bits, d_row_cnt);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

// ---- thresholded binary visibility as row runs -------------------------------------------------
// The visible set of a row, {x : vis(x, y) >= thr}, as its sorted transition columns
// t0 < t1 < ...: visible on [t0, t1), [t2, t3), ... (an open last run ends at nx, so the count is
// always even).  A visibility polygon crosses a row a handful of times, so this is 10-20 x smaller
// than one bit per cell -- small enough that the host-buffer call is bound by the sweep again and
// not by the box's PCIe / host-memory path, whatever the number of GPUs.
//   pass 1  bits[row][wpr] (the sweep itself, or threshold_bits_kernel above) and transitions per
//           row (uint16; threshold_bits_kernel or runs_row_count_kernel)
//   pass 2  runs_rows_kernel: offset of every row inside its pair, transitions per pair (uint32)
//   pass 3  runs_scan_kernel: exclusive scan of the pair totals (one CTA; <= a few thousand pairs)
//   pass 4  runs_write_kernel: positions, one warp per row
namespace {

// transitions per row from the bits alone (the sweep wrote them itself): one warp per 4 rows
__global__ void __launch_bounds__(256)
runs_row_count_kernel(const uint32_t *__restrict__ bits, const int64_t nrows, const int nx, const int wpr,
                      uint16_t *__restrict__ row_cnt) {
  constexpr int R = 4;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row0 = warp0 * R; row0 < nrows; row0 += nwarps * R) {
    int cnt[R] = {};
    for (int w0 = 0; w0 < wpr; w0 += 32) {
      const int w = w0 + lane;
      uint32_t cur[R], carry[R];
#pragma unroll
      for (int u = 0; u < R; ++u) {
        const bool have = row0 + u < nrows && w < wpr;
        const uint32_t *r = bits + (row0 + u) * wpr;
        cur[u] = have ? __ldg(r + w) : 0u;
        carry[u] = (have && lane == 0 && w0 > 0) ? __ldg(r + w - 1) >> 31 : 0u;
      }
#pragma unroll
      for (int u = 0; u < R; ++u) {
        const uint32_t below = __shfl_up_sync(kAllLanes, cur[u], 1);
        if (w < wpr) {
          cnt[u] += __popc(transitions_of(cur[u], lane ? below >> 31 : carry[u]));
          // a run that reaches the last cell is closed at nx: the zero padding of the last word does
          // that by itself unless the row fills its last word
          if (w == wpr - 1 && (nx & 31) == 0) cnt[u] += (int)(cur[u] >> 31);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const int c = __reduce_add_sync(kAllLanes, cnt[u]);
      if (lane == 0 && row0 + u < nrows) row_cnt[row0 + u] = (uint16_t)c;
    }
  }
}

// The same two passes for rows of a multiple of four words (at most 128): G lanes per row (a power of
// two >= words / 4), ONE 16-byte load per lane, so a load instruction covers 32 / G whole rows and a
// warp keeps 4 of them in flight.  (One word per lane and one row at a time, both passes ran at a
// third of the HBM rate: too few bytes in flight.)
struct RowWords {
  uint32_t t[4]; // transitions of this lane's four words
  int n;         // how many
};
__device__ __forceinline__ RowWords row_words(const uint4 v, const int l, const int w4, const int G) {
  RowWords r;
  const uint32_t pw = __shfl_up_sync(kAllLanes, v.w, 1, G); // the word before this lane's first one
  const bool have = l < w4;
  r.t[0] = have ? transitions_of(v.x, l ? pw >> 31 : 0u) : 0u;
  r.t[1] = have ? transitions_of(v.y, v.x >> 31) : 0u;
  r.t[2] = have ? transitions_of(v.z, v.y >> 31) : 0u;
  r.t[3] = have ? transitions_of(v.w, v.z >> 31) : 0u;
  r.n = __popc(r.t[0]) + __popc(r.t[1]) + __popc(r.t[2]) + __popc(r.t[3]);
  return r;
}

__global__ void __launch_bounds__(256)
runs_row_count_v4_kernel(const uint4 *__restrict__ bits4, const int nrows, const int nx, const int w4,
                         const int G, uint16_t *__restrict__ row_cnt) {
  constexpr int R = 4;
  const int lane = threadIdx.x & 31, l = lane & (G - 1), rpi = 32 / G, sub = lane / G;
  const int warp0 = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), nwarps = (int)((gridDim.x * blockDim.x) >> 5);
  const bool closes = (nx & 31) == 0; // a row that fills its last word: a run reaching the end is closed at nx
  for (int64_t row0 = (int64_t)warp0 * rpi * R; row0 < nrows; row0 += (int64_t)nwarps * rpi * R) {
    uint4 v[R];
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const int64_t row = row0 + u * rpi + sub;
      v[u] = (row < nrows && l < w4) ? __ldg(bits4 + row * w4 + l) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const int64_t row = row0 + u * rpi + sub;
      int c = row_words(v[u], l, w4, G).n;
      if (closes && l == w4 - 1) c += (int)(v[u].w >> 31);
      for (int off = G >> 1; off > 0; off >>= 1) c += __shfl_xor_sync(kAllLanes, c, off, G);
      if (l == 0 && row < nrows) row_cnt[row] = (uint16_t)c;
    }
  }
}

__global__ void __launch_bounds__(256)
runs_write_v4_kernel(const uint4 *__restrict__ bits4, const int nrows, const int ny, const int nx, const int w4,
                     const int G, const uint32_t *__restrict__ row_off,
                     const unsigned long long *__restrict__ pair_ptr, const unsigned long long chunk_base,
                     uint16_t *__restrict__ trans) {
  constexpr int R = 4;
  const int lane = threadIdx.x & 31, l = lane & (G - 1), rpi = 32 / G, sub = lane / G;
  const int warp0 = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), nwarps = (int)((gridDim.x * blockDim.x) >> 5);
  const bool closes = (nx & 31) == 0;
  for (int64_t row0 = (int64_t)warp0 * rpi * R; row0 < nrows; row0 += (int64_t)nwarps * rpi * R) {
    uint4 v[R];
    unsigned long long off[R];
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const int64_t row = row0 + u * rpi + sub;
      const bool live = row < nrows;
      v[u] = (live && l < w4) ? __ldg(bits4 + row * w4 + l) : make_uint4(0u, 0u, 0u, 0u);
      off[u] = live ? __ldg(pair_ptr + (unsigned)row / (unsigned)ny) - chunk_base + __ldg(row_off + row) : 0ull;
    }
#pragma unroll
    for (int u = 0; u < R; ++u) {
      RowWords rw = row_words(v[u], l, w4, G);
      int inc = rw.n; // inclusive scan over the G lanes of the row
      for (int o = 1; o < G; o <<= 1) {
        const int t = __shfl_up_sync(kAllLanes, inc, o, G);
        if (l >= o) inc += t;
      }
      uint16_t *out = trans + off[u] + (inc - rw.n);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint32_t t = rw.t[k];
        while (t) {
          const int b = __ffs(t) - 1;
          t &= t - 1;
          *out++ = (uint16_t)(32 * (4 * l + k) + b);
        }
      }
      if (closes && l == w4 - 1 && (v[u].w >> 31)) *out = (uint16_t)nx; // (rows past nrows hold zeros)
    }
  }
}

// rows the 16-byte kernels take: whole 16-byte pieces, at most 32 lanes per row, 32-bit row numbers
bool runs_v4_ok(const void *bits, int64_t nrows, int wpr, int *w4, int *G) {
  if (wpr % 4 != 0 || wpr > 128 || nrows >= ((int64_t)1 << 31) || ((uintptr_t)bits & 15) != 0) return false;
  *w4 = wpr / 4;
  int g = 1;
  while (g < *w4) g <<= 1;
  *G = g;
  return true;
}

// one CTA per pair: row_off[row] = transitions of the pair's earlier rows, pair_tot[pair] = all of them
__global__ void __launch_bounds__(256)
runs_rows_kernel(const uint16_t *__restrict__ row_cnt, const int ny, uint32_t *__restrict__ row_off,
                 uint32_t *__restrict__ pair_tot) {
  __shared__ uint32_t s_warp[8];
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const uint16_t *c = row_cnt + (size_t)blockIdx.x * ny;
  uint32_t *o = row_off + (size_t)blockIdx.x * ny;
  const int per = (ny + 255) / 256, i0 = min(ny, t * per), i1 = min(ny, i0 + per);
  uint32_t sum = 0;
  for (int i = i0; i < i1; ++i) sum += c[i];
  uint32_t inc = sum; // inclusive scan over the CTA's 256 partial sums
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t v = __shfl_up_sync(kAllLanes, inc, d);
    if (lane >= d) inc += v;
  }
  if (lane == 31) s_warp[wid] = inc;
  __syncthreads();
  uint32_t before = 0;
  for (int w = 0; w < wid; ++w) before += s_warp[w];
  uint32_t run = before + inc - sum;
  for (int i = i0; i < i1; ++i) { o[i] = run; run += c[i]; }
  if (t == 255) pair_tot[blockIdx.x] = before + inc;
}

// pair_ptr[p] = base + sum of pair_tot[0 .. p), pair_ptr[npairs] = base + total (one CTA)
__global__ void __launch_bounds__(1024)
runs_scan_kernel(const uint32_t *__restrict__ pair_tot, const int npairs, const unsigned long long base,
                 unsigned long long *__restrict__ pair_ptr) {
  __shared__ unsigned long long s_part[1024];
  const int t = threadIdx.x, per = (npairs + 1023) / 1024;
  unsigned long long sum = 0;
  for (int i = t * per; i < min(npairs, (t + 1) * per); ++i) sum += pair_tot[i];
  s_part[t] = sum;
  __syncthreads();
  if (t == 0) {
    unsigned long long run = base;
    for (int i = 0; i < 1024; ++i) { const unsigned long long v = s_part[i]; s_part[i] = run; run += v; }
    pair_ptr[npairs] = run;
  }
  __syncthreads();
  unsigned long long run = s_part[t];
  for (int i = t * per; i < min(npairs, (t + 1) * per); ++i) { pair_ptr[i] = run; run += pair_tot[i]; }
}

// one warp per 4 rows (their loads issued together): write the transition columns of row `row` at
// trans[pair_ptr[pair] - chunk_base + row_off[row]]
__global__ void __launch_bounds__(256)
runs_write_kernel(const uint32_t *__restrict__ bits, const int64_t nrows, const int ny, const int nx,
                  const int wpr, const uint32_t *__restrict__ row_off,
                  const unsigned long long *__restrict__ pair_ptr, const unsigned long long chunk_base,
                  uint16_t *__restrict__ trans) {
  constexpr int R = 4;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row0 = warp0 * R; row0 < nrows; row0 += nwarps * R) {
    unsigned long long off[R];
    uint32_t first[R];
    int64_t pair = row0 / ny;
    int y = (int)(row0 - pair * ny);
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const int64_t row = min(row0 + u, nrows - 1);
      off[u] = __ldg(pair_ptr + pair) - chunk_base + __ldg(row_off + row);
      first[u] = lane < wpr ? __ldg(bits + row * wpr + lane) : 0u;
      if (++y == ny && row0 + u + 1 < nrows) { y = 0; ++pair; }
    }
#pragma unroll
    for (int u = 0; u < R; ++u) {
      if (row0 + u >= nrows) break;
      const uint32_t *r = bits + (row0 + u) * wpr;
      uint16_t *out = trans + off[u];
      int base = 0;
      uint32_t last = 0;
      for (int w0 = 0; w0 < wpr; w0 += 32) {
        const int w = w0 + lane;
        const uint32_t cur = w0 == 0 ? first[u] : (w < wpr ? __ldg(r + w) : 0u);
        const uint32_t below = __shfl_up_sync(kAllLanes, cur, 1);
        const uint32_t prev = lane ? below >> 31 : (w0 > 0 ? __ldg(r + w0 - 1) >> 31 : 0u);
        uint32_t t = w < wpr ? transitions_of(cur, prev) : 0u;
        const int n = __popc(t);
        // exclusive prefix of n over the lanes: a row has a handful of transitions, so walk the
        // lanes that hold any (a shuffle scan where they are many)
        uint32_t holders = __ballot_sync(kAllLanes, n != 0);
        int pre = 0, total = 0;
        if (__popc(holders) <= 6) {
          while (holders) {
            const int L = __ffs(holders) - 1;
            holders &= holders - 1;
            const int nL = __shfl_sync(kAllLanes, n, L);
            if (lane > L) pre += nL;
            total += nL;
          }
        } else {
          int inc = n;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(kAllLanes, inc, o);
            if (lane >= o) inc += v;
          }
          pre = inc - n;
          total = __shfl_sync(kAllLanes, inc, 31);
        }
        int pos = base + pre;
        while (t) {
          const int b = __ffs(t) - 1;
          t &= t - 1;
          out[pos++] = (uint16_t)(32 * w + b);
        }
        base += total;
        if (w == wpr - 1) last = cur;
      }
      // a run that reaches the last cell of a row that fills its last word is closed at nx
      if ((nx & 31) == 0 && lane == ((wpr - 1) & 31) && (last >> 31)) out[base] = (uint16_t)nx;
    }
  }
}

} // namespace

cudaError_t vhp_launch_runs_row_count(const uint32_t *d_bits, int64_t nrows, int nx, uint16_t *d_row_cnt,
                                      int sm_count, cudaStream_t st, int64_t *launches) {
  int w4, G;
  if (runs_v4_ok(d_bits, nrows, (nx + 31) / 32, &w4, &G)) {
    const int64_t per_block = (int64_t)8 * (32 / G) * 4; // rows one pass of a block covers
    const int64_t blocks4 = std::max<int64_t>(1, std::min<int64_t>((nrows + per_block - 1) / per_block, (int64_t)sm_count * 8));
    runs_row_count_v4_kernel<<<(unsigned)blocks4, 256, 0, st>>>((const uint4 *)d_bits, (int)nrows, nx, w4, G, d_row_cnt);
    if (launches) *launches += 1;
    return cudaGetLastError();
  }
  const int64_t blocks = std::max<int64_t>(1, std::min<int64_t>((nrows + 31) / 32, (int64_t)sm_count * 8));
  runs_row_count_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_bits, nrows, nx, (nx + 31) / 32, d_row_cnt);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

// d_row_cnt: the counts threshold_bits wrote.  Leaves row_off (per row, inside its pair) and pair_ptr
// (npairs + 1 entries, starting at `base`) on the device; the caller reads pair_ptr[npairs] to
// learn how many positions the write pass produces.
cudaError_t vhp_launch_runs_count(const uint16_t *d_row_cnt, int64_t npairs, int ny, uint32_t *d_row_off,
                                  uint32_t *d_pair_tot, unsigned long long base,
                                  unsigned long long *d_pair_ptr, cudaStream_t st, int64_t *launches) {
  runs_rows_kernel<<<(unsigned)npairs, 256, 0, st>>>(d_row_cnt, ny, d_row_off, d_pair_tot);
  runs_scan_kernel<<<1, 1024, 0, st>>>(d_pair_tot, (int)npairs, base, d_pair_ptr);
  if (launches) *launches += 2;
  return cudaGetLastError();
}

cudaError_t vhp_launch_runs_write(const uint32_t *d_bits, int64_t npairs, int ny, int nx,
                                  const uint32_t *d_row_off, const unsigned long long *d_pair_ptr,
                                  unsigned long long chunk_base, uint16_t *d_trans, int sm_count,
                                  cudaStream_t st, int64_t *launches) {
  const int wpr = (nx + 31) / 32;
  const int64_t nrows = npairs * ny;
  int w4, G;
  if (runs_v4_ok(d_bits, nrows, wpr, &w4, &G)) {
    const int64_t per_block = (int64_t)8 * (32 / G) * 4;
    const int64_t blocks4 = std::max<int64_t>(1, std::min<int64_t>((nrows + per_block - 1) / per_block, (int64_t)sm_count * 8));
    runs_write_v4_kernel<<<(unsigned)blocks4, 256, 0, st>>>((const uint4 *)d_bits, (int)nrows, ny, nx, w4, G, d_row_off,
                                                           d_pair_ptr, chunk_base, d_trans);
    if (launches) *launches += 1;
    return cudaGetLastError();
  }
  int64_t blocks = std::min<int64_t>((nrows + 31) / 32, (int64_t)sm_count * 8);
  runs_write_kernel<<<(unsigned)std::max<int64_t>(blocks, 1), 256, 0, st>>>(d_bits, nrows, ny, nx, wpr, d_row_off,
                                                                           d_pair_ptr, chunk_base, d_trans);
  if (launches) *launches += 1;
  return cudaGetLastError();
}
