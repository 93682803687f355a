// result_transport.cu -- device half of the packed result transport.
//
// The host-buffer entry points are PCIe-bound: 4096 sweeps of a 1000 x 1000 grid are 16.4 GB
// of fp32 results per call against a kernel that produces them in 2.5 ms.  Visibility fields
// are mostly flat (lit 1.0, shadow 0.0), so before a chunk of results leaves the device this
// kernel classifies every 512-byte unit of it as uniform (all elements equal, bit for bit) or
// literal, and compacts the literal units into one stream.  Only the stream and 8.25 bytes of
// meta data per unit cross PCIe; host threads rebuild the exact bytes (host_expand.cpp).
//
// One warp packs one mask word (32 consecutive units = 16 KB): a unit is one 16-byte load per
// lane; the literal units of the word get consecutive slots of the stream from one atomicAdd
// on the chunk's cursor and are copied in a second pass (an L1/L2 hit).
//
// Direct mode (the caller's buffer is pinned, mapped and 16-byte aligned): the literal units
// are not compacted but stored by this kernel straight to their final place in host memory
// (512 contiguous bytes per warp store over PCIe), so the host threads only write the uniform
// units and no byte of the result is written to host memory twice.  A partial last unit of the
// chunk cannot be stored whole: it goes to the meta block and the host copies its valid bytes.
#include <cstdint>

#include "vhp_internal.h"

namespace {

constexpr unsigned kAllLanes = 0xffffffffu;

// meta block: [cursor u64, pad to 16][tail unit 512 B][mask u32 x nwords][word_base u32 x nwords]
//             [desc u64 x 32*nwords]
template <int ELEM> // element size in bytes: 4 or 8
__global__ void __launch_bounds__(256)
pack_results_kernel(const uint4 *__restrict__ in, const int64_t nunits, unsigned char *__restrict__ meta,
                    uint4 *__restrict__ lit, uint4 *__restrict__ host_dst, const int tail_partial) {
  const int64_t nwords = (nunits + 31) / 32;
  unsigned long long *cursor = reinterpret_cast<unsigned long long *>(meta);
  uint4 *tail = reinterpret_cast<uint4 *>(meta + 16);
  uint32_t *mask = reinterpret_cast<uint32_t *>(meta + kVhpPackMetaHead);
  uint32_t *word_base = mask + nwords;
  uint2 *desc = reinterpret_cast<uint2 *>(meta + kVhpPackMetaHead + (size_t)nwords * 8);
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t w = warp0; w < nwords; w += nwarps) {
    const int64_t u0 = w * 32;
    const int nu = (int)min((int64_t)32, nunits - u0);
    const uint4 *src = in + u0 * 32 + lane;
    bool my_lit = false;
    uint2 my_desc = make_uint2(0u, 0u);
#pragma unroll 8
    for (int u = 0; u < 32; ++u) {
      if (u < nu) { // warp-uniform
        const uint4 v = __ldg(src + (size_t)u * 32);
        const uint32_t f0 = __shfl_sync(kAllLanes, v.x, 0), f1 = __shfl_sync(kAllLanes, v.y, 0);
        const bool same = ELEM == 4 ? (v.x == f0 && v.y == f0 && v.z == f0 && v.w == f0)
                                    : (v.x == f0 && v.y == f1 && v.z == f0 && v.w == f1);
        const bool uni = __all_sync(kAllLanes, same);
        if (lane == u) {
          my_lit = !uni;
          my_desc = make_uint2(f0, f1);
        }
      }
    }
    const uint32_t m = __ballot_sync(kAllLanes, my_lit);
    uint32_t base = 0;
    if (lane == 0 && m) base = (uint32_t)atomicAdd(cursor, (unsigned long long)__popc(m));
    base = __shfl_sync(kAllLanes, base, 0);
    if (lane == 0) {
      mask[w] = m;
      word_base[w] = base;
    }
    desc[u0 + lane] = my_desc; // the desc array is padded to whole words
    if (host_dst) {
      for (uint32_t rest = m; rest; rest &= rest - 1) {
        const int u = __ffs(rest) - 1;
        const uint4 v = __ldg(src + (size_t)u * 32);
        if (tail_partial && u0 + u == nunits - 1) tail[lane] = v;
        else host_dst[(u0 + u) * 32 + lane] = v;
      }
    } else {
      uint4 *dst = lit + (size_t)base * 32 + lane;
      for (uint32_t rest = m; rest; rest &= rest - 1) {
        const int u = __ffs(rest) - 1;
        *dst = __ldg(src + (size_t)u * 32);
        dst += 32;
      }
    }
  }
}

} // namespace

cudaError_t vhp_launch_pack_results(const void *d_in, int64_t nunits, int elem_bytes, void *d_meta,
                                    void *d_literals, void *host_dst, int tail_partial,
                                    int sm_count, cudaStream_t st, int64_t *launches) {
  if (nunits <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(d_meta, 0, 16, st);
  if (e != cudaSuccess) return e;
  const int64_t nwords = (nunits + 31) / 32;
  const int64_t want = (nwords + 7) / 8; // 8 warps per CTA
  const unsigned grid = (unsigned)std::min<int64_t>(want, (int64_t)sm_count * 8);
  if (elem_bytes == 4)
    pack_results_kernel<4><<<grid, 256, 0, st>>>(reinterpret_cast<const uint4 *>(d_in), nunits,
                                                 reinterpret_cast<unsigned char *>(d_meta),
                                                 reinterpret_cast<uint4 *>(d_literals),
                                                 reinterpret_cast<uint4 *>(host_dst), tail_partial);
  else
    pack_results_kernel<8><<<grid, 256, 0, st>>>(reinterpret_cast<const uint4 *>(d_in), nunits,
                                                 reinterpret_cast<unsigned char *>(d_meta),
                                                 reinterpret_cast<uint4 *>(d_literals),
                                                 reinterpret_cast<uint4 *>(host_dst), tail_partial);
  if (launches) *launches += 1;
  return cudaGetLastError();
}
