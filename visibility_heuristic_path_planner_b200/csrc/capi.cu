// capi.cu -- extern "C" entry points of include/vhp.h: context management and the
// batched sweep / ray-casting / planner calls.  Plain pointers in, status out.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "vhp_internal.h"

namespace {

thread_local std::string g_last_error;

vhp_status fail(vhp_context *ctx, vhp_status st, const std::string &msg) {
  g_last_error = msg;
  if (ctx) ctx->last_error = msg;
  return st;
}

vhp_status cuda_fail(vhp_context *ctx, cudaError_t e, const char *what) {
  return fail(ctx, VHP_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define VHP_CUDA(ctx, call)                                   \
  do {                                                        \
    cudaError_t e_ = (call);                                  \
    if (e_ != cudaSuccess) return cuda_fail(ctx, e_, #call);  \
  } while (0)

// Entry points run on the context's device and leave the caller's current device as they found it
// (a torch process with several GPUs keeps its own notion of the current device).
struct DeviceGuard {
  int prev = -1;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != device) err = cudaSetDevice(device);
    else prev = -1; // nothing to restore
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};
#define VHP_ON_DEVICE(ctx)                 \
  DeviceGuard device_guard_((ctx)->device); \
  if (device_guard_.err != cudaSuccess) return cuda_fail(ctx, device_guard_.err, "cudaSetDevice")

vhp_status ensure(vhp_context *ctx, VhpDevBuf &b, size_t bytes) {
  if (bytes <= b.cap) return VHP_OK;
  if (b.p) {
    VHP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    VHP_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    VHP_CUDA(ctx, cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
  }
  VHP_CUDA(ctx, cudaMalloc(&b.p, bytes));
  b.cap = bytes;
  return VHP_OK;
}

vhp_status ensure_rcp2(vhp_context *ctx, int len) {
  if (len <= ctx->rcp2_len) return VHP_OK;
  len = std::max(len, 4096 + 8);
  if (ctx->rcp2_table) {
    VHP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    VHP_CUDA(ctx, cudaFree(ctx->rcp2_table));
    ctx->rcp2_table = nullptr;
    ctx->rcp2_len = 0;
  }
  VHP_CUDA(ctx, cudaMalloc(&ctx->rcp2_table, (size_t)len * 2 * sizeof(double)));
  VHP_CUDA(ctx, vhp_launch_rcp2_table(ctx->rcp2_table, len, ctx->stream, &ctx->launches));
  ctx->rcp2_len = len;
  return VHP_OK;
}

// (re)build the bit planes of the tile kernel for d_occ unless they are cached for this pointer
vhp_status pack_tile(vhp_context *ctx, const uint8_t *d_occ, int nmaps, int nx, int ny,
                     bool force) {
  if (!force && ctx->planes_sticky && ctx->tile_src == d_occ && ctx->tile_nmaps == nmaps &&
      ctx->tile_nx == nx && ctx->tile_ny == ny)
    return VHP_OK;
  int wx, wy, nsum;
  vhp_tile_plane_geometry(nx, ny, &wx, &wy, &nsum);
  const size_t row_plane = (size_t)ny * wx, col_plane = (size_t)nx * wy;
  const size_t bytes = (size_t)nmaps * (2 * (row_plane + col_plane) + nsum) * sizeof(uint32_t);
  if (bytes > ctx->tile_bytes) {
    if (ctx->tile_buf) {
      VHP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      VHP_CUDA(ctx, cudaFree(ctx->tile_buf));
      ctx->tile_buf = nullptr;
      ctx->tile_bytes = 0;
      ctx->tile_src = nullptr;
      ctx->planes_sticky = false;
    }
    VHP_CUDA(ctx, cudaMalloc(&ctx->tile_buf, bytes));
    ctx->tile_bytes = bytes;
  }
  uint32_t *rowF = ctx->tile_buf, *rowR = rowF + (size_t)nmaps * row_plane;
  uint32_t *colF = rowR + (size_t)nmaps * row_plane, *colR = colF + (size_t)nmaps * col_plane;
  uint32_t *bsum = colR + (size_t)nmaps * col_plane;
  VHP_CUDA(ctx, vhp_launch_pack_tile(d_occ, nmaps, nx, ny, rowF, rowR, colF, colR, bsum,
                                     ctx->stream, &ctx->launches));
  ctx->tile.rowF = rowF; ctx->tile.rowR = rowR;
  ctx->tile.colF = colF; ctx->tile.colR = colR;
  ctx->tile.bsum = bsum;
  ctx->tile.wx = wx; ctx->tile.wy = wy;
  ctx->tile.row_plane = row_plane; ctx->tile.col_plane = col_plane;
  ctx->tile_src = d_occ; ctx->tile_nmaps = nmaps; ctx->tile_nx = nx; ctx->tile_ny = ny;
  return VHP_OK;
}

vhp_status check_common(vhp_context *ctx, const void *occ, int nmaps, int nx, int ny,
                        const void *xy, int64_t n, int dtype, const void *out) {
  if (!ctx) return fail(nullptr, VHP_ERR_INVALID_ARG, "null context");
  if (!occ || !xy || !out) return fail(ctx, VHP_ERR_INVALID_ARG, "null buffer");
  if (nmaps < 1 || nx < 1 || ny < 1 || n < 0)
    return fail(ctx, VHP_ERR_INVALID_ARG, "bad sizes");
  if ((int64_t)nx * ny > (int64_t)1 << 30 || nx > 16384 || ny > 16384)
    return fail(ctx, VHP_ERR_UNSUPPORTED, "grid larger than 16384 x 16384 is not supported");
  if (dtype != VHP_F32 && dtype != VHP_F64) return fail(ctx, VHP_ERR_INVALID_ARG, "bad dtype");
  return VHP_OK;
}

// host-side validation of (x, y[, map]) items
vhp_status check_points(vhp_context *ctx, const int32_t *xy, int per_item, const int32_t *maps,
                        int64_t n, int nmaps, int nx, int ny, const char *what) {
  for (int64_t i = 0; i < n; ++i) {
    if (maps && (maps[i] < 0 || maps[i] >= nmaps))
      return fail(ctx, VHP_ERR_INVALID_ARG, std::string(what) + ": map index out of range");
    if (per_item == 2) {
      const int x = xy[2 * i], y = xy[2 * i + 1];
      if (x < 0 || x >= nx || y < 0 || y >= ny)
        return fail(ctx, VHP_ERR_INVALID_ARG,
                    std::string(what) + ": source outside the grid at item " + std::to_string(i));
    }
  }
  return VHP_OK;
}

vhp_status check_device_error(vhp_context *ctx) {
  int flag = 0;
  VHP_CUDA(ctx, cudaMemcpyAsync(&flag, ctx->d_err, sizeof(int), cudaMemcpyDeviceToHost,
                                ctx->stream));
  VHP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (flag) {
    VHP_CUDA(ctx, cudaMemsetAsync(ctx->d_err, 0, sizeof(int), ctx->stream));
    return fail(ctx, VHP_ERR_INVALID_ARG, "a source / start / end point lies outside the grid");
  }
  return VHP_OK;
}

// Grid-mode strip sweep of rows [y0, y1) (see vhp_strip_sweep_dev): CTAs to spread one sweep over.
int grid_sweep_ctas(const vhp_context *ctx, int rows) {
  return std::max(2, std::min(2 * ctx->sm_count, (4 * (rows / 32 + 2) + 7) / 8));
}

enum class Op { Sweep, Raycast };

vhp_status run_dev(vhp_context *ctx, Op op, const uint8_t *d_occ, int nmaps, int nx, int ny,
                   const int32_t *d_xy, const int32_t *d_map, int64_t n, vhp_dtype dtype,
                   void *d_out) {
  if (n == 0) return VHP_OK;
  if (op == Op::Raycast) {
    VHP_CUDA(ctx, vhp_launch_raycast(d_occ, nx, ny, d_xy, d_map, n, dtype, d_out, ctx->d_err,
                                     ctx->stream, &ctx->launches));
    return VHP_OK;
  }
  const size_t esz = dtype == VHP_F32 ? 4 : 8;
  const size_t cells = (size_t)nx * ny;
  // A few sweeps of a large map: one sweep at a time on the whole GPU (grid mode, see
  // vhp_strip_sweep_dev) instead of one CTA per pair.  (1000^2: up to 4 pairs, 2048^2 and up: 16;
  // maps too wide for the one-CTA kernel's shared memory: up to 16 pairs whatever the size.)
  const bool tile_ok = vhp_sweep_tile_supported(nx, ny);
  const bool use_grid =
      ctx->sweep_impl == 0 && vhp_sweep_grid_supported(nx, ny) &&
      (ctx->grid_sweep == 2 ||
       (ctx->grid_sweep == 1 && n <= (tile_ok ? std::min<int64_t>(16, (int64_t)(cells / 250000)) : 16)));
  if (ctx->sweep_impl == 0 && (tile_ok || use_grid)) {
    vhp_status st = ensure_rcp2(ctx, std::max(nx, ny) + 64);
    if (st != VHP_OK) return st;
    if ((st = pack_tile(ctx, d_occ, nmaps, nx, ny, false)) != VHP_OK) return st;
    if (use_grid) {
      std::vector<int32_t> h_xy(2 * (size_t)n), h_map;
      VHP_CUDA(ctx, cudaMemcpyAsync(h_xy.data(), d_xy, h_xy.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
      if (d_map) {
        h_map.resize((size_t)n);
        VHP_CUDA(ctx, cudaMemcpyAsync(h_map.data(), d_map, h_map.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
      }
      VHP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      if ((st = ensure(ctx, ctx->b_grid, vhp_sweep_grid_ws_bytes(nx, ny))) != VHP_OK) return st;
      int wx, wy, nsum;
      vhp_tile_plane_geometry(nx, ny, &wx, &wy, &nsum);
      const int ctas = grid_sweep_ctas(ctx, ny);
      for (int64_t k = 0; k < n; ++k) {
        const int sx = h_xy[2 * k], sy = h_xy[2 * k + 1];
        const int64_t m = d_map ? h_map[k] : 0;
        if (sx < 0 || sx >= nx || sy < 0 || sy >= ny || m < 0 || m >= nmaps)
          return fail(ctx, VHP_ERR_INVALID_ARG, "a source / start / end point lies outside the grid");
        VhpTilePlanes pl = ctx->tile; // planes of map m (the grid kernels sweep "map 0")
        pl.rowF += m * pl.row_plane; pl.rowR += m * pl.row_plane;
        pl.colF += m * pl.col_plane; pl.colR += m * pl.col_plane;
        pl.bsum += m * (size_t)nsum;
        VHP_CUDA(ctx, vhp_launch_sweep_window(pl, nx, ny, sx, sy, 0, ny, nullptr, dtype,
                                              (char *)d_out + (size_t)k * cells * esz,
                                              ctx->rcp2_table, ctx->d_err, ctx->b_grid.p, ctas,
                                              ctx->stream, &ctx->launches));
      }
      return VHP_OK;
    }
    VHP_CUDA(ctx, vhp_launch_sweep_tile(ctx->tile, nx, ny, d_xy, d_map, n, dtype, d_out,
                                        ctx->rcp2_table, ctx->d_err, ctx->stream, &ctx->launches));
    return VHP_OK;
  }
  // naive kernel: chunk so that the scratch fronts stay small
  const int64_t chunk = 4096;
  vhp_status st = ensure(ctx, ctx->b_scratch, vhp_sweep_naive_scratch_bytes(nx, ny, std::min(n, chunk)));
  if (st != VHP_OK) return st;
  for (int64_t p0 = 0; p0 < n; p0 += chunk) {
    const int64_t np = std::min(chunk, n - p0);
    VHP_CUDA(ctx, vhp_launch_sweep_naive(d_occ, nx, ny, d_xy + 2 * p0, d_map ? d_map + p0 : nullptr,
                                         np, dtype, (char *)d_out + (size_t)p0 * nx * ny * esz,
                                         (double *)ctx->b_scratch.p, ctx->d_err, ctx->stream,
                                         &ctx->launches));
  }
  return VHP_OK;
}

// Plain transport of the host-buffer entry points: run in chunks of pairs from p_begin on and
// overlap the D2H of chunk c with the kernel of chunk c+1 (two device buffers, copy stream).
vhp_status run_host_plain(vhp_context *ctx, Op op, int nmaps, int nx, int ny, const int32_t *d_xy,
                          const int32_t *d_map, int64_t n, int64_t p_begin, vhp_dtype dtype,
                          void *out) {
  const size_t cells = (size_t)nx * ny, esz = dtype == VHP_F32 ? 4 : 8;
  vhp_status st;
  const size_t chunk_bytes_target = (size_t)1 << 30;
  int64_t chunk = std::max<int64_t>(1, (int64_t)(chunk_bytes_target / (cells * esz)));
  chunk = std::min(chunk, n - p_begin);
  const size_t buf_bytes = (size_t)chunk * cells * esz;
  if ((st = ensure(ctx, ctx->b_out[0], buf_bytes)) != VHP_OK) return st;
  if (n - p_begin > chunk && (st = ensure(ctx, ctx->b_out[1], buf_bytes)) != VHP_OK) return st;
  // pin the caller's buffer for full-rate async copies (ignore "already pinned").  Page-locking
  // costs milliseconds: small results (a single 101 x 101 sweep is 80 KB) are copied as they are.
  const size_t out_bytes = (size_t)(n - p_begin) * cells * esz;
  char *const out_first = (char *)out + (size_t)p_begin * cells * esz;
  const bool registered = out_bytes >= ((size_t)32 << 20) &&
                          cudaHostRegister(out_first, out_bytes, cudaHostRegisterDefault) == cudaSuccess;
  (void)cudaGetLastError();
  vhp_status result = VHP_OK;
  int it = 0;
  for (int64_t p0 = p_begin; p0 < n && result == VHP_OK; p0 += chunk, ++it) {
    const int b = it & 1;
    const int64_t np = std::min(chunk, n - p0);
    if (it >= 2) { // the copy that last read this buffer must be done
      cudaError_t e = cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[b], 0);
      if (e != cudaSuccess) { result = cuda_fail(ctx, e, "cudaStreamWaitEvent"); break; }
    }
    result = run_dev(ctx, op, (const uint8_t *)ctx->b_occ.p, nmaps, nx, ny, d_xy + 2 * p0,
                     d_map ? d_map + p0 : nullptr, np, dtype, ctx->b_out[b].p);
    if (result != VHP_OK) break;
    cudaError_t e = cudaEventRecord(ctx->ev_done[b], ctx->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[b], 0);
    if (e != cudaSuccess) { result = cuda_fail(ctx, e, "event hand-over to the copy stream"); break; }
    e = cudaMemcpyAsync((char *)out + (size_t)p0 * cells * esz, ctx->b_out[b].p,
                                    (size_t)np * cells * esz, cudaMemcpyDeviceToHost,
                                    ctx->copy_stream);
    if (e != cudaSuccess) { result = cuda_fail(ctx, e, "cudaMemcpyAsync D2H"); break; }
    e = cudaEventRecord(ctx->ev_copied[b], ctx->copy_stream);
    if (e != cudaSuccess) { result = cuda_fail(ctx, e, "cudaEventRecord"); break; }
    ctx->last_d2h_bytes += (int64_t)((size_t)np * cells * esz);
  }
  cudaError_t e1 = cudaStreamSynchronize(ctx->copy_stream);
  cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
  if (registered) cudaHostUnregister(out_first);
  if (result != VHP_OK) return result;
  if (e1 != cudaSuccess) return cuda_fail(ctx, e1, "sync copy stream");
  if (e2 != cudaSuccess) return cuda_fail(ctx, e2, "sync stream");
  return VHP_OK;
}

// Packed transport (result_transport.cu, host_expand.cpp): every chunk of results is packed on
// the device into uniform / literal 128-byte units; only the meta data and the literal units
// cross PCIe and a pool of host threads rebuilds the exact bytes in the caller's buffer.
// Three sets of buffers: while chunk c is computed and packed, the literals of chunk c-1 are
// copied and chunk c-2 is expanded.  When the caller's buffer is pinned, mapped and 16-byte
// aligned ("direct"), the pack kernel stores the literal units straight to their place in it
// and the host threads only write the uniform units.  In automatic mode the call gives up
// after the first chunks when they hardly compress; *resume_from is the first pair not
// delivered.
vhp_status run_host_packed(vhp_context *ctx, Op op, int nmaps, int nx, int ny, const int32_t *d_xy,
                           const int32_t *d_map, int64_t n, vhp_dtype dtype, void *out,
                           int64_t *resume_from) {
  constexpr int NS = vhp_context::kPackSets;
  const size_t cells = (size_t)nx * ny, esz = dtype == VHP_F32 ? 4 : 8;
  const size_t pair_bytes = cells * esz;
  int64_t chunk = std::max<int64_t>(1, (int64_t)(((size_t)256 << 20) / pair_bytes));
  chunk = std::min(chunk, n);
  const int64_t nchunks = (n + chunk - 1) / chunk;
  const int64_t units_max = (int64_t)(((size_t)chunk * pair_bytes + kVhpPackUnit - 1) / kVhpPackUnit);
  const size_t meta_max = vhp_pack_meta_bytes(units_max, (int)esz), lit_max = (size_t)units_max * kVhpPackUnit;
  vhp_status st;
  // direct mode: a device-accessible address of the caller's buffer, if it is pinned host memory
  char *out_dev = nullptr;
  {
    cudaPointerAttributes attr;
    void *dp = nullptr;
    if (ctx->result_direct && (((uintptr_t)out) & 15u) == 0 && pair_bytes % 16 == 0 &&
        cudaPointerGetAttributes(&attr, out) == cudaSuccess && attr.type == cudaMemoryTypeHost &&
        cudaHostGetDevicePointer(&dp, out, 0) == cudaSuccess)
      out_dev = (char *)dp;
    (void)cudaGetLastError();
  }
  const bool direct = out_dev != nullptr;
  if (!ctx->expand_pool) {
    int t = 0;
    if (const char *e = std::getenv("VHP_HOST_THREADS")) t = std::atoi(e);
    if (t <= 0) t = (int)std::min(32u, std::max(1u, std::thread::hardware_concurrency()));
    try {
      ctx->expand_pool = new VhpExpandPool(t);
    } catch (...) { // no host threads to be had: deliver everything by plain copies
      ctx->expand_pool = nullptr;
      *resume_from = 0;
      return VHP_OK;
    }
  }
  ctx->last_transport_packed = direct ? 2 : 1;
  const int nsets = (int)std::min<int64_t>(NS, nchunks);
  for (int s = 0; s < nsets; ++s) {
    if ((st = ensure(ctx, ctx->b_pack_out[s], lit_max)) != VHP_OK) return st;
    if ((st = ensure(ctx, ctx->b_pack_meta[s], meta_max)) != VHP_OK) return st;
    if (!direct && (st = ensure(ctx, ctx->b_pack_lit[s], lit_max)) != VHP_OK) return st;
  }
  const size_t lit_need = direct ? 0 : lit_max;
  if (meta_max > ctx->h_pack_meta_cap || lit_need > ctx->h_pack_lit_cap) {
    for (int s = 0; s < NS; ++s) {
      if (ctx->h_pack_meta[s]) cudaFreeHost(ctx->h_pack_meta[s]);
      if (ctx->h_pack_lit[s]) cudaFreeHost(ctx->h_pack_lit[s]);
      ctx->h_pack_meta[s] = ctx->h_pack_lit[s] = nullptr;
    }
    ctx->h_pack_meta_cap = ctx->h_pack_lit_cap = 0;
    for (int s = 0; s < NS; ++s) {
      VHP_CUDA(ctx, cudaHostAlloc(&ctx->h_pack_meta[s], meta_max, cudaHostAllocDefault));
      if (lit_need) VHP_CUDA(ctx, cudaHostAlloc(&ctx->h_pack_lit[s], lit_need, cudaHostAllocDefault));
    }
    ctx->h_pack_meta_cap = meta_max;
    ctx->h_pack_lit_cap = lit_need;
  }
  VhpExpandPool &pool = *ctx->expand_pool;
  int64_t tickets[NS] = {0, 0, 0}, last_ticket = 0;
  int64_t lit_units_total = 0, units_total = 0;
  bool bail = false;
  vhp_status result = VHP_OK;

  // Direct mode: the device can also deliver a share of the mask words completely (uniform
  // units included).  Off by default: on the B200 boxes every byte that arrives over PCIe
  // while 16 host threads stream into the same memory costs about as much time as 3.5 bytes
  // written by the host threads (share 0 / 2 / 4 / 6 sixteenths: 96 / 125 / 155 / 184 ms for
  // 16.4 GB), so the fastest split moves as few bytes over PCIe as possible.
  double trace_pack_ms = 0.0, trace_wait_s = 0.0, trace_ticket_s = 0.0;
  const double trace_busy0 = pool.busy_seconds();
  const auto trace_t0 = std::chrono::steady_clock::now();
  int share_of[NS] = {0, 0, 0};
  auto gpu_share_now = [&]() -> int { return std::max(0, std::min(ctx->result_gpu_share, 16)); };
  auto chunk_units = [&](int64_t it) {
    const int64_t np = std::min(chunk, n - it * chunk);
    return (int64_t)(((size_t)np * pair_bytes + kVhpPackUnit - 1) / kVhpPackUnit);
  };
  auto launch = [&](int64_t it) -> vhp_status {
    const int s = (int)(it % NS);
    const int64_t p0 = it * chunk, np = std::min(chunk, n - p0);
    vhp_status r = run_dev(ctx, op, (const uint8_t *)ctx->b_occ.p, nmaps, nx, ny, d_xy + 2 * p0,
                           d_map ? d_map + p0 : nullptr, np, dtype, ctx->b_pack_out[s].p);
    if (r != VHP_OK) return r;
    // packing (PCIe-bound in direct mode) runs on the copy stream, so that the sweeps of the
    // next chunk overlap it
    VHP_CUDA(ctx, cudaEventRecord(ctx->ev_pack_lit[s], ctx->stream));
    VHP_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_pack_lit[s], 0));
    VHP_CUDA(ctx, cudaEventRecord(ctx->ev_pack_t0[s], ctx->copy_stream));
    const int64_t nu = chunk_units(it);
    const int tail_partial = ((size_t)np * pair_bytes) % kVhpPackUnit != 0;
    share_of[s] = direct ? gpu_share_now() : 0;
    VHP_CUDA(ctx, vhp_launch_pack_results(ctx->b_pack_out[s].p, nu, (int)esz, ctx->b_pack_meta[s].p,
                                          ctx->b_pack_lit[s].p,
                                          direct ? out_dev + (size_t)p0 * pair_bytes : nullptr,
                                          tail_partial, share_of[s], ctx->sm_count,
                                          ctx->copy_stream, &ctx->launches));
    VHP_CUDA(ctx, cudaMemcpyAsync(ctx->h_pack_meta[s], ctx->b_pack_meta[s].p, vhp_pack_meta_bytes(nu, (int)esz),
                                  cudaMemcpyDeviceToHost, ctx->copy_stream));
    VHP_CUDA(ctx, cudaEventRecord(ctx->ev_pack_meta[s], ctx->copy_stream));
    return VHP_OK;
  };
  auto finish = [&](int64_t it) -> vhp_status {
    const int s = (int)(it % NS);
    const int64_t p0 = it * chunk, np = std::min(chunk, n - p0);
    const int64_t nu = chunk_units(it), nwords = (nu + 31) / 32;
    const auto tw0 = std::chrono::steady_clock::now();
    VHP_CUDA(ctx, cudaEventSynchronize(ctx->ev_pack_meta[s]));
    trace_wait_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - tw0).count();
    if (ctx->transport_trace) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, ctx->ev_pack_t0[s], ctx->ev_pack_meta[s]) == cudaSuccess)
        trace_pack_ms += ms;
    }
    const char *meta = (const char *)ctx->h_pack_meta[s];
    const uint64_t nlit = *(const uint64_t *)meta;
    if (nlit && !direct) {
      VHP_CUDA(ctx, cudaMemcpyAsync(ctx->h_pack_lit[s], ctx->b_pack_lit[s].p,
                                    (size_t)nlit * kVhpPackUnit, cudaMemcpyDeviceToHost,
                                    ctx->copy_stream));
      VHP_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    }
    // words the device delivered completely (their literal units are not in the cursor)
    const int64_t gpu_words = share_of[s] ? ((nwords - 1) / 16) * share_of[s] +
                                                std::min<int64_t>((nwords - 1) % 16, share_of[s])
                                          : 0;
    ctx->last_d2h_bytes += (int64_t)(vhp_pack_meta_bytes(nu, (int)esz) + (size_t)nlit * kVhpPackUnit +
                                     (size_t)gpu_words * 32 * kVhpPackUnit);
    lit_units_total += (int64_t)nlit;
    units_total += nu - gpu_words * 32; // the literal fraction is measured on the host's words
    VhpPackedChunk c;
    c.mask = (const uint32_t *)(meta + kVhpPackMetaHead);
    c.word_base = c.mask + nwords;
    c.vmask = c.word_base + nwords;
    c.elem_bytes = (int)esz;
    c.literals = direct ? nullptr : (const char *)ctx->h_pack_lit[s];
    c.tail = meta + 16;
    c.gpu_share = share_of[s];
    c.dst = (char *)out + (size_t)p0 * pair_bytes;
    c.nunits = nu;
    c.valid_bytes = (size_t)np * pair_bytes;
    tickets[s] = last_ticket = pool.submit(c);
    return VHP_OK;
  };

  int64_t launched = 0, finished = 0;
  for (int64_t it = 0; it < nchunks && result == VHP_OK && !bail; ++it) {
    const int s = (int)(it % NS);
    if (tickets[s]) { // staging set s is free again
      const auto tt0 = std::chrono::steady_clock::now();
      pool.wait(tickets[s]);
      trace_ticket_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - tt0).count();
    }
    if ((result = launch(it)) != VHP_OK) break;
    ++launched;
    if (it >= 1) {
      if ((result = finish(it - 1)) != VHP_OK) break;
      ++finished;
      // automatic mode: results that hardly compress are cheaper to copy directly
      if (ctx->result_transport == 1 && finished == 1 && lit_units_total * 10 > units_total * 6)
        bail = true;
    }
  }
  while (result == VHP_OK && finished < launched) {
    result = finish(finished);
    if (result == VHP_OK) ++finished;
  }
  if (last_ticket) pool.wait(last_ticket);
  cudaError_t e1 = cudaStreamSynchronize(ctx->copy_stream);
  cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
  if (e2 == cudaSuccess) e2 = e1;
  if (result != VHP_OK) return result;
  if (e2 != cudaSuccess) return cuda_fail(ctx, e2, "sync stream");
  *resume_from = std::min(n, launched * chunk);
  if (ctx->transport_trace)
    std::fprintf(stderr,
                 "[vhp transport] %s, %lld chunks: wall %.1f ms | packing on the GPU %.1f ms | host "
                 "expansion %.1f ms (%d threads) | main thread waited %.1f ms for the GPU, %.1f ms "
                 "for the host threads | %.2f of %.2f GB over PCIe\n",
                 direct ? "direct" : "staged", (long long)launched,
                 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - trace_t0).count(),
                 trace_pack_ms, 1e3 * (pool.busy_seconds() - trace_busy0), pool.threads(),
                 1e3 * trace_wait_s, 1e3 * trace_ticket_s, ctx->last_d2h_bytes / 1e9,
                 (double)(launched * chunk * pair_bytes) / 1e9);
  return VHP_OK;
}

// ---- packed handle: the lossless packed form of a batch of fields, kept on the host -------------
// vhp_visibility_batch_packed runs the same chunks as run_host_packed in staged form, but the meta
// blocks and literal streams stay in pinned memory owned by the handle instead of being expanded:
// the call moves 3-7 % of the field bytes over PCIe and writes nothing else to host memory.
// vhp_packed_expand rebuilds any range of pairs later (bit-identical to vhp_visibility_batch).
} // namespace

struct vhp_packed {
  struct Block { char *p; size_t cap, used; };
  struct Chunk { size_t meta_off, lit_off; int meta_blk, lit_blk; int64_t nunits, np; uint64_t nlit; };
  int device = 0;
  int nx = 0, ny = 0, elem = 4;
  int64_t npairs = 0, chunk_pairs = 0;
  size_t pair_bytes = 0;
  std::vector<Block> blocks; // pinned host memory, reused by the next call on this handle
  std::vector<Chunk> chunks;
  int64_t bytes_held = 0;

  // `bytes` of pinned memory, 256-byte aligned: from the first block with room, else a new block
  bool take(size_t bytes, int *blk, size_t *off) {
    bytes = (bytes + 255) & ~(size_t)255;
    for (size_t b = 0; b < blocks.size(); ++b)
      if (blocks[b].cap - blocks[b].used >= bytes) {
        *blk = (int)b; *off = blocks[b].used; blocks[b].used += bytes;
        return true;
      }
    Block nb{nullptr, std::max<size_t>(bytes, (size_t)128 << 20), 0};
    if (cudaHostAlloc((void **)&nb.p, nb.cap, cudaHostAllocDefault) != cudaSuccess) {
      (void)cudaGetLastError();
      return false;
    }
    nb.used = bytes;
    blocks.push_back(nb);
    *blk = (int)blocks.size() - 1; *off = 0;
    return true;
  }
  VhpPackedChunk view(const Chunk &c) const {
    const char *meta = blocks[c.meta_blk].p + c.meta_off;
    const int64_t nwords = (c.nunits + 31) / 32;
    VhpPackedChunk v;
    v.mask = (const uint32_t *)(meta + kVhpPackMetaHead);
    v.word_base = v.mask + nwords;
    v.vmask = v.word_base + nwords;
    v.elem_bytes = elem;
    v.literals = c.nlit ? blocks[c.lit_blk].p + c.lit_off : meta; // (never read when there are none)
    v.nunits = c.nunits;
    v.valid_bytes = (size_t)c.np * pair_bytes;
    return v;
  }
};

namespace {

vhp_status run_host_to_packed(vhp_context *ctx, int nmaps, int nx, int ny, const int32_t *d_xy,
                              const int32_t *d_map, int64_t n, vhp_dtype dtype, vhp_packed *h) {
  constexpr int NS = vhp_context::kPackSets;
  const size_t cells = (size_t)nx * ny, esz = dtype == VHP_F32 ? 4 : 8;
  const size_t pair_bytes = cells * esz;
  // chunks of 1 GB of fields (nothing is expanded here, so a chunk costs one event wait and two copies:
  // with 256 MB chunks those fixed costs were a third of the call)
  int64_t chunk = std::max<int64_t>(1, (int64_t)(((size_t)1 << 30) / pair_bytes));
  chunk = std::min(chunk, n);
  const int64_t nchunks = (n + chunk - 1) / chunk;
  const int64_t units_max = (int64_t)(((size_t)chunk * pair_bytes + kVhpPackUnit - 1) / kVhpPackUnit);
  const size_t meta_max = vhp_pack_meta_bytes(units_max, (int)esz), lit_max = (size_t)units_max * kVhpPackUnit;
  h->device = ctx->device; h->nx = nx; h->ny = ny; h->elem = (int)esz;
  h->npairs = n; h->chunk_pairs = chunk; h->pair_bytes = pair_bytes;
  h->chunks.assign((size_t)nchunks, vhp_packed::Chunk{});
  for (auto &b : h->blocks) b.used = 0;
  h->bytes_held = 0;
  vhp_status st;
  const int nsets = (int)std::min<int64_t>(NS, nchunks);
  for (int s = 0; s < nsets; ++s) {
    if ((st = ensure(ctx, ctx->b_pack_out[s], lit_max)) != VHP_OK) return st;
    if ((st = ensure(ctx, ctx->b_pack_meta[s], meta_max)) != VHP_OK) return st;
    if ((st = ensure(ctx, ctx->b_pack_lit[s], lit_max)) != VHP_OK) return st;
  }
  auto chunk_units = [&](int64_t it) {
    const int64_t np = std::min(chunk, n - it * chunk);
    return (int64_t)(((size_t)np * pair_bytes + kVhpPackUnit - 1) / kVhpPackUnit);
  };
  // chunk `it`: sweeps and packing on the compute stream, the copy of the meta block on the copy stream
  auto launch = [&](int64_t it) -> vhp_status {
    const int s = (int)(it % NS);
    const int64_t p0 = it * chunk, np = std::min(chunk, n - p0), nu = chunk_units(it);
    vhp_packed::Chunk &c = h->chunks[(size_t)it];
    c.nunits = nu; c.np = np;
    if (!h->take(vhp_pack_meta_bytes(nu, (int)esz), &c.meta_blk, &c.meta_off))
      return fail(ctx, VHP_ERR_CUDA, "vhp_visibility_batch_packed: out of pinned host memory");
    if (it >= NS) VHP_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_pack_t0[s], 0)); // set s copied out
    vhp_status r = run_dev(ctx, Op::Sweep, (const uint8_t *)ctx->b_occ.p, nmaps, nx, ny, d_xy + 2 * p0,
                           d_map ? d_map + p0 : nullptr, np, dtype, ctx->b_pack_out[s].p);
    if (r != VHP_OK) return r;
    // (packing is an HBM-speed pass here -- nothing goes to host memory -- so it stays on the compute
    // stream and the copy stream carries copies only)
    VHP_CUDA(ctx, vhp_launch_pack_results(ctx->b_pack_out[s].p, nu, (int)esz, ctx->b_pack_meta[s].p,
                                          ctx->b_pack_lit[s].p, nullptr, 0, 0, ctx->sm_count, ctx->stream,
                                          &ctx->launches));
    VHP_CUDA(ctx, cudaEventRecord(ctx->ev_pack_lit[s], ctx->stream));
    VHP_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_pack_lit[s], 0));
    VHP_CUDA(ctx, cudaMemcpyAsync(h->blocks[c.meta_blk].p + c.meta_off, ctx->b_pack_meta[s].p,
                                  vhp_pack_meta_bytes(nu, (int)esz), cudaMemcpyDeviceToHost, ctx->copy_stream));
    VHP_CUDA(ctx, cudaEventRecord(ctx->ev_pack_meta[s], ctx->copy_stream));
    return VHP_OK;
  };
  // ... once its meta block has arrived the literal count is known: copy that many units
  auto finish = [&](int64_t it) -> vhp_status {
    const int s = (int)(it % NS);
    vhp_packed::Chunk &c = h->chunks[(size_t)it];
    VHP_CUDA(ctx, cudaEventSynchronize(ctx->ev_pack_meta[s]));
    c.nlit = *(const uint64_t *)(h->blocks[c.meta_blk].p + c.meta_off);
    const size_t meta_bytes = vhp_pack_meta_bytes(c.nunits, (int)esz), lit_bytes = (size_t)c.nlit * kVhpPackUnit;
    if (c.nlit) {
      if (!h->take(lit_bytes, &c.lit_blk, &c.lit_off))
        return fail(ctx, VHP_ERR_CUDA, "vhp_visibility_batch_packed: out of pinned host memory");
      VHP_CUDA(ctx, cudaMemcpyAsync(h->blocks[c.lit_blk].p + c.lit_off, ctx->b_pack_lit[s].p, lit_bytes,
                                    cudaMemcpyDeviceToHost, ctx->copy_stream));
    }
    VHP_CUDA(ctx, cudaEventRecord(ctx->ev_pack_t0[s], ctx->copy_stream)); // the set's device buffers are free
    ctx->last_d2h_bytes += (int64_t)(meta_bytes + lit_bytes);
    h->bytes_held += (int64_t)(meta_bytes + lit_bytes);
    return VHP_OK;
  };
  vhp_status result = VHP_OK;
  int64_t launched = 0, finished = 0;
  for (int64_t it = 0; it < nchunks && result == VHP_OK; ++it) {
    if ((result = launch(it)) != VHP_OK) break;
    ++launched;
    if (it >= 1) {
      if ((result = finish(it - 1)) != VHP_OK) break;
      ++finished;
    }
  }
  while (result == VHP_OK && finished < launched) {
    result = finish(finished);
    if (result == VHP_OK) ++finished;
  }
  cudaError_t e1 = cudaStreamSynchronize(ctx->copy_stream);
  cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
  if (e2 == cudaSuccess) e2 = e1;
  if (result != VHP_OK) return result;
  if (e2 != cudaSuccess) return cuda_fail(ctx, e2, "sync stream");
  ctx->last_result_bytes = (int64_t)((size_t)n * pair_bytes);
  ctx->last_transport_packed = 1;
  return VHP_OK;
}

// host buffers: upload, run in chunks, deliver the results (packed or plain transport)
vhp_status run_host(vhp_context *ctx, Op op, const uint8_t *occ, int nmaps, int nx, int ny,
                    const int32_t *xy, const int32_t *maps, int64_t n, vhp_dtype dtype,
                    void *out) {
  vhp_status st = check_common(ctx, occ, nmaps, nx, ny, xy, n, dtype, out);
  if (st != VHP_OK) return st;
  st = check_points(ctx, xy, 2, maps, n, nmaps, nx, ny,
                    op == Op::Sweep ? "vhp_visibility_batch" : "vhp_raycast_batch");
  if (st != VHP_OK) return st;
  ctx->last_d2h_bytes = ctx->last_result_bytes = 0;
  ctx->last_transport_packed = 0;
  if (n == 0) return VHP_OK;
  VHP_ON_DEVICE(ctx);
  const size_t cells = (size_t)nx * ny, esz = dtype == VHP_F32 ? 4 : 8;
  const size_t occ_bytes = (size_t)nmaps * cells;
  if ((st = ensure(ctx, ctx->b_occ, occ_bytes)) != VHP_OK) return st;
  if ((st = ensure(ctx, ctx->b_src, (size_t)n * 2 * sizeof(int32_t))) != VHP_OK) return st;
  if (maps && (st = ensure(ctx, ctx->b_map, (size_t)n * sizeof(int32_t))) != VHP_OK) return st;
  VHP_CUDA(ctx, cudaMemcpyAsync(ctx->b_occ.p, occ, occ_bytes, cudaMemcpyHostToDevice, ctx->stream));
  VHP_CUDA(ctx, cudaMemcpyAsync(ctx->b_src.p, xy, (size_t)n * 2 * sizeof(int32_t),
                                cudaMemcpyHostToDevice, ctx->stream));
  if (maps)
    VHP_CUDA(ctx, cudaMemcpyAsync(ctx->b_map.p, maps, (size_t)n * sizeof(int32_t),
                                  cudaMemcpyHostToDevice, ctx->stream));
  ctx->planes_sticky = false; // b_occ content changed
  ctx->tile_src = nullptr;
  if (op == Op::Sweep && ctx->sweep_impl == 0 &&
      (vhp_sweep_tile_supported(nx, ny) || vhp_sweep_grid_supported(nx, ny))) {
    // pack once for all chunks
    if ((st = pack_tile(ctx, (const uint8_t *)ctx->b_occ.p, nmaps, nx, ny, true)) != VHP_OK)
      return st;
    ctx->planes_sticky = true;
  }
  const int32_t *d_xy = (const int32_t *)ctx->b_src.p;
  const int32_t *d_map = maps ? (const int32_t *)ctx->b_map.p : nullptr;
  const size_t out_bytes = (size_t)n * cells * esz;
  ctx->last_result_bytes = (int64_t)out_bytes;
  int64_t p_begin = 0;
  vhp_status result = VHP_OK;
  // small results are not worth the thread pool's wake-up
  if (ctx->result_transport == 2 || (ctx->result_transport == 1 && out_bytes >= ((size_t)64 << 20))) {
    result = run_host_packed(ctx, op, nmaps, nx, ny, d_xy, d_map, n, dtype, out, &p_begin);
  }
  if (result == VHP_OK && p_begin < n)
    result = run_host_plain(ctx, op, nmaps, nx, ny, d_xy, d_map, n, p_begin, dtype, out);
  ctx->planes_sticky = false;
  ctx->tile_src = nullptr;
  if (result != VHP_OK) return result;
  return check_device_error(ctx);
}

// ---- thresholded binary visibility ------------------------------------------------------------
// Chunks of pairs: fp64 sweep into a device buffer, threshold_bits_kernel, and for host callers
// the D2H of chunk c (copy stream, two bit buffers) under the sweeps of chunk c + 1.
int64_t bin_chunk_pairs(size_t cells, int64_t n) {
  return std::min<int64_t>(n, std::max<int64_t>(1, (int64_t)(((size_t)4 << 30) / (cells * 8))));
}

// Thresholded visibility of np pairs as bits (and, d_row_cnt != null, transition columns per row).
// Batches the one-CTA tile kernel takes are swept straight into bits (kFmtBits: the field itself is
// never stored); a non-positive threshold, a few pairs of a large map (grid mode) and the naive
// kernel go through the fp64 field in ctx->b_bin (at most bin_chunk_pairs pairs).
bool sweeps_into_bits(const vhp_context *ctx, int nx, int ny, int64_t np, double thr) {
  const size_t cells = (size_t)nx * ny;
  return ctx->sweep_impl == 0 && ctx->bin_direct && thr > 0.0 && vhp_sweep_tile_supported(nx, ny) &&
         !(ctx->grid_sweep == 2 && vhp_sweep_grid_supported(nx, ny)) &&
         !(ctx->grid_sweep == 1 && vhp_sweep_grid_supported(nx, ny) &&
           np <= std::min<int64_t>(16, (int64_t)(cells / 250000)));
}

vhp_status sweep_to_bits(vhp_context *ctx, const uint8_t *d_occ, int nmaps, int nx, int ny,
                         const int32_t *d_xy, const int32_t *d_map, int64_t np, double thr,
                         uint32_t *d_bits, uint16_t *d_row_cnt = nullptr) {
  if (sweeps_into_bits(ctx, nx, ny, np, thr)) {
    vhp_status st = ensure_rcp2(ctx, std::max(nx, ny) + 64);
    if (st != VHP_OK) return st;
    if ((st = pack_tile(ctx, d_occ, nmaps, nx, ny, false)) != VHP_OK) return st;
    VHP_CUDA(ctx, vhp_launch_sweep_tile(ctx->tile, nx, ny, d_xy, d_map, np, VHP_F32, d_bits, ctx->rcp2_table,
                                        ctx->d_err, ctx->stream, &ctx->launches, &thr, true));
    if (d_row_cnt)
      VHP_CUDA(ctx, vhp_launch_runs_row_count(d_bits, np * ny, nx, d_row_cnt, ctx->sm_count, ctx->stream,
                                              &ctx->launches));
    return VHP_OK;
  }
  vhp_status st = ensure(ctx, ctx->b_bin, (size_t)np * nx * ny * 8);
  if (st != VHP_OK) return st;
  if ((st = run_dev(ctx, Op::Sweep, d_occ, nmaps, nx, ny, d_xy, d_map, np, VHP_F64, ctx->b_bin.p)) != VHP_OK)
    return st;
  VHP_CUDA(ctx, vhp_launch_threshold_bits((const double *)ctx->b_bin.p, np * ny, nx, thr, d_bits,
                                          ctx->sm_count, ctx->stream, &ctx->launches, d_row_cnt));
  return VHP_OK;
}

vhp_status run_bin_dev(vhp_context *ctx, const uint8_t *d_occ, int nmaps, int nx, int ny,
                       const int32_t *d_xy, const int32_t *d_map, int64_t n, double thr,
                       uint32_t *d_bits) {
  const size_t cells = (size_t)nx * ny, wpr = (size_t)(nx + 31) / 32;
  // no intermediate field when the sweep writes bits itself: one launch for the whole batch
  const int64_t chunk = sweeps_into_bits(ctx, nx, ny, n, thr) ? n : bin_chunk_pairs(cells, n);
  vhp_status st;
  for (int64_t p0 = 0; p0 < n; p0 += chunk) {
    const int64_t np = std::min(chunk, n - p0);
    st = sweep_to_bits(ctx, d_occ, nmaps, nx, ny, d_xy + 2 * p0, d_map ? d_map + p0 : nullptr, np, thr,
                       d_bits + (size_t)p0 * ny * wpr);
    if (st != VHP_OK) return st;
  }
  return VHP_OK;
}

vhp_status run_bin_host(vhp_context *ctx, const uint8_t *occ, int nmaps, int nx, int ny,
                        const int32_t *xy, const int32_t *maps, int64_t n, double thr,
                        uint32_t *out_bits) {
  vhp_status st = check_common(ctx, occ, nmaps, nx, ny, xy, n, VHP_F64, out_bits);
  if (st != VHP_OK) return st;
  if ((st = check_points(ctx, xy, 2, maps, n, nmaps, nx, ny, "vhp_visibility_batch_bin")) != VHP_OK) return st;
  ctx->last_d2h_bytes = ctx->last_result_bytes = 0;
  ctx->last_transport_packed = 0;
  if (n == 0) return VHP_OK;
  VHP_ON_DEVICE(ctx);
  const size_t cells = (size_t)nx * ny, wpr = (size_t)(nx + 31) / 32, occ_bytes = (size_t)nmaps * cells;
  if ((st = ensure(ctx, ctx->b_occ, occ_bytes)) != VHP_OK) return st;
  if ((st = ensure(ctx, ctx->b_src, (size_t)n * 8)) != VHP_OK) return st;
  if (maps && (st = ensure(ctx, ctx->b_map, (size_t)n * 4)) != VHP_OK) return st;
  VHP_CUDA(ctx, cudaMemcpyAsync(ctx->b_occ.p, occ, occ_bytes, cudaMemcpyHostToDevice, ctx->stream));
  VHP_CUDA(ctx, cudaMemcpyAsync(ctx->b_src.p, xy, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
  if (maps) VHP_CUDA(ctx, cudaMemcpyAsync(ctx->b_map.p, maps, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
  ctx->planes_sticky = false;
  ctx->tile_src = nullptr;
  if (ctx->sweep_impl == 0 && (vhp_sweep_tile_supported(nx, ny) || vhp_sweep_grid_supported(nx, ny))) {
    if ((st = pack_tile(ctx, (const uint8_t *)ctx->b_occ.p, nmaps, nx, ny, true)) != VHP_OK) return st;
    ctx->planes_sticky = true;
  }
  const int32_t *d_xy = (const int32_t *)ctx->b_src.p;
  const int32_t *d_map = maps ? (const int32_t *)ctx->b_map.p : nullptr;
  // the sweep writes bits itself: four chunks (the D2H of one under the sweeps of the next), each
  // many waves of CTAs; else chunks bounded by the fp64 field
  const int64_t chunk = sweeps_into_bits(ctx, nx, ny, n, thr)
                            ? std::min<int64_t>(n, std::max<int64_t>((int64_t)ctx->sm_count * 6, (n + 3) / 4))
                            : bin_chunk_pairs(cells, n);
  const size_t pair_words = (size_t)ny * wpr, chunk_bytes = (size_t)chunk * pair_words * 4;
  vhp_status result = VHP_OK;
  for (int b = 0; b < 2 && result == VHP_OK; ++b)
    if (b == 0 || n > chunk) result = ensure(ctx, ctx->b_out[b], chunk_bytes);
  int it = 0;
  for (int64_t p0 = 0; p0 < n && result == VHP_OK; p0 += chunk, ++it) {
    const int b = it & 1;
    const int64_t np = std::min(chunk, n - p0);
    cudaError_t e = cudaSuccess;
    if (it >= 2) e = cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[b], 0); // bit buffer b is free again
    if (e != cudaSuccess) { result = cuda_fail(ctx, e, "cudaStreamWaitEvent"); break; }
    result = sweep_to_bits(ctx, (const uint8_t *)ctx->b_occ.p, nmaps, nx, ny, d_xy + 2 * p0,
                           d_map ? d_map + p0 : nullptr, np, thr, (uint32_t *)ctx->b_out[b].p);
    if (result != VHP_OK) break;
    e = cudaEventRecord(ctx->ev_done[b], ctx->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[b], 0);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(out_bits + (size_t)p0 * pair_words, ctx->b_out[b].p, (size_t)np * pair_words * 4,
                          cudaMemcpyDeviceToHost, ctx->copy_stream);
    if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_copied[b], ctx->copy_stream);
    if (e != cudaSuccess) { result = cuda_fail(ctx, e, "binary visibility: enqueue"); break; }
    ctx->last_d2h_bytes += (int64_t)((size_t)np * pair_words * 4);
  }
  cudaError_t e1 = cudaStreamSynchronize(ctx->copy_stream);
  cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
  ctx->planes_sticky = false;
  ctx->tile_src = nullptr;
  ctx->last_result_bytes = (int64_t)((size_t)n * pair_words * 4);
  if (result != VHP_OK) return result;
  if (e1 != cudaSuccess) return cuda_fail(ctx, e1, "sync copy stream");
  if (e2 != cudaSuccess) return cuda_fail(ctx, e2, "sync stream");
  return check_device_error(ctx);
}

// Row runs of the thresholded visibility (vhp_visibility_batch_runs): per chunk of pairs fp64 sweep,
// threshold bits, transitions per row / pair, exclusive scan, transition columns; the total of a
// chunk is read back (8 bytes) before its columns are written and copied.
vhp_status run_runs_host(vhp_context *ctx, const uint8_t *occ, int nmaps, int nx, int ny,
                         const int32_t *xy, const int32_t *maps, int64_t n, double thr,
                         uint16_t *row_count, uint64_t *pair_ptr, uint16_t *trans, int64_t trans_cap,
                         int64_t *trans_used) {
  vhp_status st = check_common(ctx, occ, nmaps, nx, ny, xy, n, VHP_F64, row_count);
  if (st != VHP_OK) return st;
  if (!pair_ptr || (!trans && trans_cap > 0) || trans_cap < 0)
    return fail(ctx, VHP_ERR_INVALID_ARG, "vhp_visibility_batch_runs: null buffer");
  if ((st = check_points(ctx, xy, 2, maps, n, nmaps, nx, ny, "vhp_visibility_batch_runs")) != VHP_OK) return st;
  ctx->last_d2h_bytes = ctx->last_result_bytes = 0;
  ctx->last_transport_packed = 0;
  if (trans_used) *trans_used = 0;
  pair_ptr[0] = 0;
  if (n == 0) return VHP_OK;
  VHP_ON_DEVICE(ctx);
  const size_t cells = (size_t)nx * ny, wpr = (size_t)(nx + 31) / 32, occ_bytes = (size_t)nmaps * cells;
  if ((st = ensure(ctx, ctx->b_occ, occ_bytes)) != VHP_OK) return st;
  if ((st = ensure(ctx, ctx->b_src, (size_t)n * 8)) != VHP_OK) return st;
  if (maps && (st = ensure(ctx, ctx->b_map, (size_t)n * 4)) != VHP_OK) return st;
  VHP_CUDA(ctx, cudaMemcpyAsync(ctx->b_occ.p, occ, occ_bytes, cudaMemcpyHostToDevice, ctx->stream));
  VHP_CUDA(ctx, cudaMemcpyAsync(ctx->b_src.p, xy, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
  if (maps) VHP_CUDA(ctx, cudaMemcpyAsync(ctx->b_map.p, maps, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
  ctx->planes_sticky = false;
  ctx->tile_src = nullptr;
  if (ctx->sweep_impl == 0 && (vhp_sweep_tile_supported(nx, ny) || vhp_sweep_grid_supported(nx, ny))) {
    if ((st = pack_tile(ctx, (const uint8_t *)ctx->b_occ.p, nmaps, nx, ny, true)) != VHP_OK) return st;
    ctx->planes_sticky = true;
  }
  const int32_t *d_xy = (const int32_t *)ctx->b_src.p;
  const int32_t *d_map = maps ? (const int32_t *)ctx->b_map.p : nullptr;
  // the sweep writes bits itself: a chunk is bounded by its bits (1 GB), not by an fp64 field
  const int64_t chunk = sweeps_into_bits(ctx, nx, ny, n, thr)
                            ? std::min<int64_t>(n, std::max<int64_t>(1, (int64_t)(((size_t)1 << 30) / ((size_t)ny * wpr * 4))))
                            : bin_chunk_pairs(cells, n);
  vhp_status result = ensure(ctx, ctx->b_out[0], (size_t)chunk * ny * wpr * 4);
  // small per-chunk arrays: row counts, row offsets, pair totals, pair offsets
  const size_t o_cnt = 0, o_off = ((size_t)chunk * ny * 2 + 255) & ~(size_t)255;
  const size_t o_tot = o_off + (((size_t)chunk * ny * 4 + 255) & ~(size_t)255);
  const size_t o_ptr = o_tot + (((size_t)chunk * 4 + 255) & ~(size_t)255);
  if (result == VHP_OK) result = ensure(ctx, ctx->b_misc, o_ptr + ((size_t)chunk + 1) * 8 + 256);
  unsigned long long total = 0;
  for (int64_t p0 = 0; p0 < n && result == VHP_OK; p0 += chunk) {
    const int64_t np = std::min(chunk, n - p0);
    char *misc = (char *)ctx->b_misc.p;
    uint16_t *d_cnt = (uint16_t *)(misc + o_cnt);
    uint32_t *d_off = (uint32_t *)(misc + o_off);
    uint32_t *d_tot = (uint32_t *)(misc + o_tot);
    result = sweep_to_bits(ctx, (const uint8_t *)ctx->b_occ.p, nmaps, nx, ny, d_xy + 2 * p0,
                           d_map ? d_map + p0 : nullptr, np, thr, (uint32_t *)ctx->b_out[0].p, d_cnt);
    if (result != VHP_OK) break;
    unsigned long long *d_ptr = (unsigned long long *)(misc + o_ptr);
    cudaError_t e = vhp_launch_runs_count(d_cnt, np, ny, d_off, d_tot, total, d_ptr, ctx->stream, &ctx->launches);
    unsigned long long new_total = 0;
    // the row counts and pair offsets are final here: they leave on the copy stream, under the write pass
    if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_done[0], ctx->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[0], 0);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(row_count + (size_t)p0 * ny, d_cnt, (size_t)np * ny * 2, cudaMemcpyDeviceToHost, ctx->copy_stream);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(pair_ptr + p0, d_ptr, ((size_t)np + 1) * 8, cudaMemcpyDeviceToHost, ctx->copy_stream);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(&new_total, d_ptr + np, 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { result = cuda_fail(ctx, e, "row runs: count"); break; }
    if (trans_used) *trans_used = (int64_t)new_total;
    if ((int64_t)new_total > trans_cap) {
      result = fail(ctx, VHP_ERR_INVALID_ARG, "vhp_visibility_batch_runs: trans_cap too small (see *trans_used)");
      break;
    }
    const size_t chunk_elems = (size_t)(new_total - total);
    if ((result = ensure(ctx, ctx->b_out[1], std::max<size_t>(chunk_elems * 2, 256))) != VHP_OK) break;
    e = vhp_launch_runs_write((const uint32_t *)ctx->b_out[0].p, np, ny, nx, d_off, d_ptr, total,
                              (uint16_t *)ctx->b_out[1].p, ctx->sm_count, ctx->stream, &ctx->launches);
    if (e == cudaSuccess && chunk_elems)
      e = cudaMemcpyAsync(trans + total, ctx->b_out[1].p, chunk_elems * 2, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream); // the device buffers are reused by the next chunk
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->copy_stream);
    if (e != cudaSuccess) { result = cuda_fail(ctx, e, "row runs: write"); break; }
    ctx->last_d2h_bytes += (int64_t)(chunk_elems * 2 + (size_t)np * ny * 2 + ((size_t)np + 1) * 8 + 8);
    total = new_total;
  }
  (void)cudaStreamSynchronize(ctx->copy_stream); // (a failed chunk may have left its copies in flight)
  ctx->planes_sticky = false;
  ctx->tile_src = nullptr;
  ctx->last_result_bytes = ctx->last_d2h_bytes;
  if (result != VHP_OK) return result;
  return check_device_error(ctx);
}

// ---- opt-in sweep variants ------------------------------------------------------------------------
vhp_status check_variant(vhp_context *ctx, const vhp_sweep_variant *v) {
  if (!v) return fail(ctx, VHP_ERR_INVALID_ARG, "sweep variant: null descriptor");
  if (v->model != VHP_VARIANT_MATLAB && v->model != VHP_VARIANT_QUEUE)
    return fail(ctx, VHP_ERR_INVALID_ARG, "sweep variant: model is VHP_VARIANT_MATLAB or VHP_VARIANT_QUEUE");
  if (v->model == VHP_VARIANT_MATLAB && !(v->fac > 0.0))
    return fail(ctx, VHP_ERR_INVALID_ARG, "sweep variant: fac must be positive");
  return VHP_OK;
}

vhp_status run_variant_dev(vhp_context *ctx, const uint8_t *d_occ, int nx, int ny, const int32_t *d_xy,
                           const int32_t *d_map, int64_t n, const vhp_sweep_variant &v, vhp_dtype dtype,
                           void *d_out) {
  const size_t cells = (size_t)nx * ny;
  const int64_t chunk = dtype == VHP_F64 ? n : bin_chunk_pairs(cells, n);
  vhp_status st;
  if (dtype == VHP_F32 && (st = ensure(ctx, ctx->b_bin, (size_t)chunk * cells * 8)) != VHP_OK) return st;
  for (int64_t p0 = 0; p0 < n; p0 += chunk) {
    const int64_t np = std::min(chunk, n - p0);
    double *buf = dtype == VHP_F64 ? (double *)d_out + (size_t)p0 * cells : (double *)ctx->b_bin.p;
    float *o32 = dtype == VHP_F32 ? (float *)d_out + (size_t)p0 * cells : nullptr;
    VHP_CUDA(ctx, vhp_launch_sweep_variant(d_occ, nx, ny, d_xy + 2 * p0, d_map ? d_map + p0 : nullptr, np,
                                           v.model, v.alpha, v.fac, v.light_strength, v.cutoff, buf, o32,
                                           ctx->d_err, ctx->stream, &ctx->launches));
  }
  return VHP_OK;
}

// ---- planner ------------------------------------------------------------------
// One LARGE problem on the whole GPU: solve() (src/visibilityBasedSolver.cpp:76-160) as the
// one-strip case of the strip engine (giant.cu).  Per iteration: grid-mode sweep of the whole
// map (many CTAs, the +y and -y quadrants as two concurrent launches), per-cell epilogue +
// arg-min, next-source selection -- all with the loop state on the device, captured as the
// body of a CUDA-graph WHILE node: no host round trip between the first sweep and the path.
// Same results as the persistent single-CTA planner kernel, which would leave all but one SM
// idle.
vhp_status planner_grid_one(vhp_context *ctx, const VhpTilePlanes &pl, int nx, int ny,
                            const int32_t se[4], double thr, int32_t max_iter, int32_t ls_cap,
                            double *vis, double *vg, double *hc, int32_t *came, int32_t *status,
                            int32_t *nb, int32_t *ls, double *plen, int32_t *pn, int32_t *path,
                            float *vg32, float *vis32) {
  vhp_status st = vhp_i_grid_planner_run(ctx, pl, nx, ny, se, thr, max_iter, ls_cap, vis, vg, hc, came, ls);
  if (st != VHP_OK) return st;
  VHP_CUDA(ctx, vhp_launch_grid_planner_finish(vhp_i_grid_planner_ctl(ctx), nx, ny, se[2], se[3], ls_cap,
                                               ls, came, vis, vg, status, nb, plen, pn, path, vg32,
                                               vis32, ctx->stream, &ctx->launches));
  return VHP_OK;
}

// device pointers in `o` (any may be null).  Fields the caller did not ask for in
// fp64 live in the context workspace, so large batches run in chunks.
vhp_status planner_dev(vhp_context *ctx, const uint8_t *d_occ, int nmaps, int nx, int ny,
                       const int32_t *d_se, const int32_t *d_pmap, int64_t nprob, double thr,
                       int32_t max_iter, int32_t ls_cap, vhp_dtype dtype, const vhp_planner_out &o) {
  if (nprob == 0) return VHP_OK;
  // maps too wide for the persistent single-CTA kernel still run on the grid route
  const bool cta_ok = vhp_planner_supported(nx, ny);
  if (!cta_ok && !(ctx->grid_sweep != 0 && vhp_sweep_grid_supported(nx, ny)))
    return fail(ctx, VHP_ERR_UNSUPPORTED, "planner: grid too large for the single-CTA kernel");
  if (ls_cap < max_iter + 2 || max_iter < 0)
    return fail(ctx, VHP_ERR_INVALID_ARG, "planner: ls_cap must be >= max_iter + 2");
  vhp_status st = ensure_rcp2(ctx, std::max(nx, ny) + 64);
  if (st != VHP_OK) return st;
  if ((st = pack_tile(ctx, d_occ, nmaps, nx, ny, false)) != VHP_OK) return st;
  const size_t cells = (size_t)nx * ny;
  const bool vis_user = dtype == VHP_F64 && o.vis, vg_user = dtype == VHP_F64 && o.vg;
  const size_t per_prob = (vis_user ? 0 : 8 * cells) + (vg_user ? 0 : 8 * cells) +
                          (o.came ? 0 : 4 * cells) + 8 * cells; // + the cached heuristic field
  const size_t small_pp = 4 + 4 + 8 + 4 + 2 * (size_t)ls_cap * 8 + 32;
  int64_t chunk = nprob;
  const size_t ws_limit = (size_t)32 << 30; // (of 180 GB: fewer chunks, fewer kernel tails)
  if (per_prob) chunk = std::max<int64_t>(1, std::min<int64_t>(nprob, (int64_t)(ws_limit / per_prob)));
  const size_t field_bytes = (size_t)chunk * per_prob;
  const size_t small_bytes = (size_t)nprob * small_pp;
  if ((st = ensure(ctx, ctx->b_planner, field_bytes + small_bytes + 256)) != VHP_OK) return st;
  char *w = (char *)ctx->b_planner.p;
  auto take = [&](size_t bytes) { char *r = w; w += (bytes + 15) & ~(size_t)15; return r; };
  double *ws_vis = vis_user ? nullptr : (double *)take(8 * cells * chunk);
  double *ws_vg = vg_user ? nullptr : (double *)take(8 * cells * chunk);
  double *ws_hc = (double *)take(8 * cells * chunk);
  int32_t *ws_came = o.came ? nullptr : (int32_t *)take(4 * cells * chunk);
  int32_t *status = o.status ? o.status : (int32_t *)take(4 * nprob);
  int32_t *nb = o.nb_sources ? o.nb_sources : (int32_t *)take(4 * nprob);
  int32_t *ls = o.light_sources ? o.light_sources : (int32_t *)take(8 * (size_t)ls_cap * nprob);
  double *plen = o.path_len ? o.path_len : (double *)take(8 * nprob);
  int32_t *pn = o.path_n ? o.path_n : (int32_t *)take(4 * nprob);
  int32_t *path = o.path ? o.path : (int32_t *)take(8 * (size_t)ls_cap * nprob);
  // A few problems on a large map: one problem at a time on the whole GPU (grid mode) instead
  // of one persistent CTA per problem.  grid_sweep 0: never, 1: by size, 2: always.
  // Measured on B200 (tools/planner_routes.py, ms per iteration): single CTA 0.6 @256^2, 3.3 @1000^2,
  // 48 @4096^2, 184 @8192^2; grid route 0.3, 0.7, 2.8, 7.5 -- but its problems run one after the
  // other, so it only wins for a handful of problems, more of them the larger the map.
  const int64_t grid_max_prob = std::max<int64_t>(1, std::min<int64_t>(24, (int64_t)(cells / 100000)));
  const bool grid_route = !cta_ok || ctx->grid_sweep == 2 ||
                          (ctx->grid_sweep == 1 && nprob <= grid_max_prob && vhp_sweep_grid_supported(nx, ny));
  std::vector<int32_t> h_se, h_pmap;
  if (grid_route) {
    h_se.resize(4 * (size_t)nprob);
    VHP_CUDA(ctx, cudaMemcpyAsync(h_se.data(), d_se, h_se.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (d_pmap) {
      h_pmap.resize((size_t)nprob);
      VHP_CUDA(ctx, cudaMemcpyAsync(h_pmap.data(), d_pmap, h_pmap.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    VHP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int64_t k = 0; k < nprob; ++k)
      if (d_pmap && (h_pmap[k] < 0 || h_pmap[k] >= nmaps))
        return fail(ctx, VHP_ERR_INVALID_ARG, "planner: map index out of range");
  }
  for (int64_t q0 = 0; q0 < nprob; q0 += chunk) {
    const int64_t n = std::min(chunk, nprob - q0);
    double *vis = vis_user ? (double *)o.vis + q0 * cells : ws_vis;
    double *vg = vg_user ? (double *)o.vg + q0 * cells : ws_vg;
    int32_t *came = o.came ? o.came + q0 * cells : ws_came;
    float *vg32 = (dtype == VHP_F32 && o.vg) ? (float *)o.vg + q0 * cells : nullptr;
    float *vis32 = (dtype == VHP_F32 && o.vis) ? (float *)o.vis + q0 * cells : nullptr;
    if (grid_route) {
      int wx, wy, nsum;
      vhp_tile_plane_geometry(nx, ny, &wx, &wy, &nsum);
      for (int64_t k = 0; k < n; ++k) {
        const int64_t q = q0 + k;
        const size_t m = d_pmap ? (size_t)h_pmap[q] : 0;
        VhpTilePlanes pl = ctx->tile; // planes of map m (the grid kernels sweep "map 0")
        pl.rowF += m * pl.row_plane; pl.rowR += m * pl.row_plane;
        pl.colF += m * pl.col_plane; pl.colR += m * pl.col_plane;
        pl.bsum += m * (size_t)nsum;
        st = planner_grid_one(ctx, pl, nx, ny, &h_se[4 * q], thr, max_iter, ls_cap, vis + k * cells,
                              vg + k * cells, ws_hc + k * cells, came + k * cells, status + q, nb + q,
                              ls + 2 * (size_t)ls_cap * q, plen + q, pn + q,
                              path + 2 * (size_t)ls_cap * q, vg32 ? vg32 + k * cells : nullptr,
                              vis32 ? vis32 + k * cells : nullptr);
        if (st != VHP_OK) return st;
      }
      continue;
    }
    void *first_ws = nullptr;
    if (ctx->planner_first == 2 || (ctx->planner_first == 1 && n >= (int64_t)8 * ctx->sm_count)) {
      if ((st = ensure(ctx, ctx->b_misc, vhp_planner_first_ws_bytes(n))) != VHP_OK) return st;
      first_ws = ctx->b_misc.p;
    }
    VHP_CUDA(ctx, vhp_launch_planner(ctx->tile, nx, ny, d_se + 4 * q0, d_pmap ? d_pmap + q0 : nullptr,
                                     n, thr, max_iter, ls_cap, ctx->rcp2_table, vis, vg, ws_hc, came,
                                     status + q0, nb + q0, ls + 2 * (size_t)ls_cap * q0, plen + q0,
                                     pn + q0, path + 2 * (size_t)ls_cap * q0, vg32, vis32,
                                     ctx->d_err, ctx->stream, &ctx->launches, first_ws, ctx->planner_first_rounds));
  }
  return VHP_OK;
}

} // namespace

vhp_status vhp_i_fail(vhp_context *ctx, vhp_status st, const std::string &msg) { return fail(ctx, st, msg); }
vhp_status vhp_i_ensure_rcp2(vhp_context *ctx, int len) { return ensure_rcp2(ctx, len); }
int vhp_i_grid_ctas(const vhp_context *ctx, int nx, int ny, int rows) {
  const bool cta_ok = vhp_sweep_tile_supported(nx, ny);
  const int mode = ctx->grid_sweep;
  if (!cta_ok || (mode && vhp_sweep_grid_supported(nx, ny) &&
                  (int64_t)nx * rows >= (mode > 1 ? 0 : (1 << 20))))
    return grid_sweep_ctas(ctx, rows);
  return 1;
}

extern "C" {

int vhp_abi_version(void) { return VHP_ABI_VERSION; }

const char *vhp_version_string(void) {
  return "vhp_b200 0.1 (sm_100a; visibility sweep / planner / ray casting)";
}

int vhp_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return n;
}

const char *vhp_last_error(const vhp_context *ctx) {
  return ctx ? ctx->last_error.c_str() : g_last_error.c_str();
}

int64_t vhp_launch_count(const vhp_context *ctx) { return ctx ? ctx->launches : 0; }

vhp_status vhp_context_create(int device, void *cuda_stream, vhp_context **out) {
  if (!out) return fail(nullptr, VHP_ERR_INVALID_ARG, "null out pointer");
  *out = nullptr;
  const int n = vhp_device_count();
  if (n == 0) return fail(nullptr, VHP_ERR_NO_DEVICE, "no CUDA device available (no CPU fallback)");
  if (device < 0 || device >= n) return fail(nullptr, VHP_ERR_INVALID_ARG, "bad device index");
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceProperties");
  if (prop.major != 10)
    return fail(nullptr, VHP_ERR_NO_DEVICE,
                "device is not sm_100 (this library ships sm_100a code only)");
#define CTX_TRY(call)                                      \
  do {                                                     \
    cudaError_t e_ = (call);                               \
    if (e_ != cudaSuccess) {                               \
      vhp_context_destroy(ctx);                            \
      return cuda_fail(nullptr, e_, #call);                \
    }                                                      \
  } while (0)
  vhp_context *ctx = new vhp_context();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  DeviceGuard device_guard_(device);
  CTX_TRY(device_guard_.err);
  if (cuda_stream) {
    ctx->stream = (cudaStream_t)cuda_stream;
  } else {
    CTX_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->owns_stream = true;
  }
  CTX_TRY(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    CTX_TRY(cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming));
    CTX_TRY(cudaEventCreateWithFlags(&ctx->ev_copied[i], cudaEventDisableTiming));
  }
  CTX_TRY(cudaMalloc(&ctx->d_err, sizeof(int)));
  CTX_TRY(cudaMemsetAsync(ctx->d_err, 0, sizeof(int), ctx->stream));
  const char *impl = std::getenv("VHP_SWEEP_IMPL");
  ctx->sweep_impl = 0;
  if (impl && std::strcmp(impl, "naive") == 0) ctx->sweep_impl = 1;
  if (const char *e = std::getenv("VHP_GRID_SWEEP")) ctx->grid_sweep = std::atoi(e);
  if (const char *e = std::getenv("VHP_PLANNER_FIRST")) ctx->planner_first = std::atoi(e);
  if (const char *e = std::getenv("VHP_BIN_DIRECT")) ctx->bin_direct = std::atoi(e) != 0;
  if (const char *e = std::getenv("VHP_PLANNER_ROUNDS")) ctx->planner_first_rounds = std::max(1, std::atoi(e));
  for (int i = 0; i < vhp_context::kPackSets; ++i) {
    CTX_TRY(cudaEventCreate(&ctx->ev_pack_meta[i]));
    CTX_TRY(cudaEventCreate(&ctx->ev_pack_t0[i]));
    CTX_TRY(cudaEventCreateWithFlags(&ctx->ev_pack_lit[i], cudaEventDisableTiming));
  }
  if (const char *e = std::getenv("VHP_RESULT_TRANSPORT")) {
    if (std::strcmp(e, "plain") == 0) ctx->result_transport = 0;
    else if (std::strcmp(e, "packed") == 0) ctx->result_transport = 2;
  }
  if (const char *e = std::getenv("VHP_RESULT_DIRECT")) ctx->result_direct = std::atoi(e) != 0;
  if (const char *e = std::getenv("VHP_RESULT_GPU_SHARE")) ctx->result_gpu_share = std::atoi(e);
  ctx->transport_trace = std::getenv("VHP_TRANSPORT_TRACE") != nullptr;
#undef CTX_TRY
  *out = ctx;
  return VHP_OK;
}

void vhp_context_destroy(vhp_context *ctx) {
  if (!ctx) return;
  DeviceGuard device_guard_(ctx->device);
  vhp_i_grid_planner_release(ctx);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
  VhpDevBuf *bufs[] = {&ctx->b_occ, &ctx->b_src, &ctx->b_map, &ctx->b_out[0], &ctx->b_out[1],
                       &ctx->b_scratch, &ctx->b_planner, &ctx->b_misc, &ctx->b_grid, &ctx->b_bin};
  for (VhpDevBuf *b : bufs)
    if (b->p) cudaFree(b->p);
  delete ctx->expand_pool;
  for (int i = 0; i < vhp_context::kPackSets; ++i) {
    if (ctx->b_pack_out[i].p) cudaFree(ctx->b_pack_out[i].p);
    if (ctx->b_pack_meta[i].p) cudaFree(ctx->b_pack_meta[i].p);
    if (ctx->b_pack_lit[i].p) cudaFree(ctx->b_pack_lit[i].p);
    if (ctx->h_pack_meta[i]) cudaFreeHost(ctx->h_pack_meta[i]);
    if (ctx->h_pack_lit[i]) cudaFreeHost(ctx->h_pack_lit[i]);
    if (ctx->ev_pack_meta[i]) cudaEventDestroy(ctx->ev_pack_meta[i]);
    if (ctx->ev_pack_t0[i]) cudaEventDestroy(ctx->ev_pack_t0[i]);
    if (ctx->ev_pack_lit[i]) cudaEventDestroy(ctx->ev_pack_lit[i]);
  }
  if (ctx->tile_buf) cudaFree(ctx->tile_buf);
  if (ctx->rcp2_table) cudaFree(ctx->rcp2_table);
  if (ctx->d_err) cudaFree(ctx->d_err);
  for (int i = 0; i < 2; ++i) {
    if (ctx->ev_done[i]) cudaEventDestroy(ctx->ev_done[i]);
    if (ctx->ev_copied[i]) cudaEventDestroy(ctx->ev_copied[i]);
  }
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->owns_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  (void)cudaGetLastError();
  delete ctx;
}

vhp_status vhp_context_synchronize(vhp_context *ctx) {
  if (!ctx) return fail(nullptr, VHP_ERR_INVALID_ARG, "null context");
  VHP_ON_DEVICE(ctx);
  return check_device_error(ctx);
}

int vhp_context_device(const vhp_context *ctx) { return ctx ? ctx->device : -1; }

vhp_status vhp_prepare_maps_dev(vhp_context *ctx, const uint8_t *d_occ, int nmaps, int nx,
                                int ny) {
  if (!ctx || !d_occ || nmaps < 1 || nx < 1 || ny < 1)
    return fail(ctx, VHP_ERR_INVALID_ARG, "vhp_prepare_maps_dev: bad argument");
  VHP_ON_DEVICE(ctx);
  if (!vhp_sweep_tile_supported(nx, ny) && !vhp_sweep_grid_supported(nx, ny))
    return VHP_OK; // the naive kernel reads the byte maps
  const vhp_status st = pack_tile(ctx, d_occ, nmaps, nx, ny, true);
  if (st == VHP_OK) ctx->planes_sticky = true;
  return st;
}

vhp_status vhp_release_maps_dev(vhp_context *ctx) {
  if (!ctx) return fail(nullptr, VHP_ERR_INVALID_ARG, "null context");
  ctx->planes_sticky = false;
  ctx->tile_src = nullptr;
  return VHP_OK;
}

vhp_status vhp_visibility_batch_dev(vhp_context *ctx, const uint8_t *d_occ, int nmaps, int nx,
                                    int ny, const int32_t *d_src_xy, const int32_t *d_src_map,
                                    int64_t npairs, vhp_dtype dtype, void *d_out) {
  vhp_status st = check_common(ctx, d_occ, nmaps, nx, ny, d_src_xy, npairs, dtype, d_out);
  if (st != VHP_OK) return st;
  VHP_ON_DEVICE(ctx);
  return run_dev(ctx, Op::Sweep, d_occ, nmaps, nx, ny, d_src_xy, d_src_map, npairs, dtype, d_out);
}

vhp_status vhp_visibility_batch(vhp_context *ctx, const uint8_t *occ, int nmaps, int nx, int ny,
                                const int32_t *src_xy, const int32_t *src_map, int64_t npairs,
                                vhp_dtype dtype, void *out) {
  return run_host(ctx, Op::Sweep, occ, nmaps, nx, ny, src_xy, src_map, npairs, dtype, out);
}

vhp_status vhp_visibility_batch_packed(vhp_context *ctx, const uint8_t *occ, int nmaps, int nx, int ny,
                                       const int32_t *src_xy, const int32_t *src_map, int64_t npairs,
                                       vhp_dtype dtype, vhp_packed **handle) {
  if (!handle) return fail(ctx, VHP_ERR_INVALID_ARG, "vhp_visibility_batch_packed: null handle pointer");
  vhp_status st = check_common(ctx, occ, nmaps, nx, ny, src_xy, npairs, dtype, handle);
  if (st != VHP_OK) return st;
  if ((st = check_points(ctx, src_xy, 2, src_map, npairs, nmaps, nx, ny, "vhp_visibility_batch_packed")) != VHP_OK)
    return st;
  ctx->last_d2h_bytes = ctx->last_result_bytes = 0;
  ctx->last_transport_packed = 0;
  VHP_ON_DEVICE(ctx);
  vhp_packed *h = *handle;
  const bool fresh = h == nullptr;
  if (fresh) {
    h = new (std::nothrow) vhp_packed();
    if (!h) return fail(ctx, VHP_ERR_CUDA, "vhp_visibility_batch_packed: out of memory");
  } else if (h->device != ctx->device && !h->blocks.empty()) {
    return fail(ctx, VHP_ERR_INVALID_ARG, "vhp_visibility_batch_packed: the handle belongs to another device");
  }
  h->device = ctx->device;
  h->npairs = 0;
  h->chunks.clear();
  auto run = [&]() -> vhp_status {
    if (npairs == 0) return VHP_OK;
    const size_t cells = (size_t)nx * ny, occ_bytes = (size_t)nmaps * cells;
    vhp_status r;
    if ((r = ensure(ctx, ctx->b_occ, occ_bytes)) != VHP_OK) return r;
    if ((r = ensure(ctx, ctx->b_src, (size_t)npairs * 8)) != VHP_OK) return r;
    if (src_map && (r = ensure(ctx, ctx->b_map, (size_t)npairs * 4)) != VHP_OK) return r;
    VHP_CUDA(ctx, cudaMemcpyAsync(ctx->b_occ.p, occ, occ_bytes, cudaMemcpyHostToDevice, ctx->stream));
    VHP_CUDA(ctx, cudaMemcpyAsync(ctx->b_src.p, src_xy, (size_t)npairs * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (src_map)
      VHP_CUDA(ctx, cudaMemcpyAsync(ctx->b_map.p, src_map, (size_t)npairs * 4, cudaMemcpyHostToDevice, ctx->stream));
    ctx->planes_sticky = false;
    ctx->tile_src = nullptr;
    if (ctx->sweep_impl == 0 && (vhp_sweep_tile_supported(nx, ny) || vhp_sweep_grid_supported(nx, ny))) {
      if ((r = pack_tile(ctx, (const uint8_t *)ctx->b_occ.p, nmaps, nx, ny, true)) != VHP_OK) return r;
      ctx->planes_sticky = true;
    }
    r = run_host_to_packed(ctx, nmaps, nx, ny, (const int32_t *)ctx->b_src.p,
                           src_map ? (const int32_t *)ctx->b_map.p : nullptr, npairs, dtype, h);
    ctx->planes_sticky = false;
    ctx->tile_src = nullptr;
    if (r != VHP_OK) return r;
    return check_device_error(ctx);
  };
  st = run();
  if (st != VHP_OK) {
    h->npairs = 0;
    h->chunks.clear();
    if (fresh) vhp_packed_destroy(h);
    return st;
  }
  *handle = h;
  return VHP_OK;
}

int64_t vhp_packed_pairs(const vhp_packed *h) { return h ? h->npairs : 0; }
int64_t vhp_packed_bytes(const vhp_packed *h) { return h ? h->bytes_held : 0; }
int64_t vhp_packed_pair_bytes(const vhp_packed *h) { return h ? (int64_t)h->pair_bytes : 0; }

vhp_status vhp_packed_expand(const vhp_packed *h, int64_t first_pair, int64_t npairs, void *out, int nthreads) {
  if (!h || first_pair < 0 || npairs < 0 || first_pair + npairs > h->npairs || (npairs > 0 && !out))
    return vhp_i_fail(nullptr, VHP_ERR_INVALID_ARG, "vhp_packed_expand: null handle / buffer or pairs out of range");
  if (npairs == 0) return VHP_OK;
  const size_t b_begin = (size_t)first_pair * h->pair_bytes, b_end = (size_t)(first_pair + npairs) * h->pair_bytes;
  const size_t chunk_bytes = (size_t)h->chunk_pairs * h->pair_bytes;
  // slices of about equal size, cut at 128-byte units of the batch's first chunk (any cut is valid)
  const size_t total = b_end - b_begin;
  int nt = nthreads > 0 ? nthreads : (int)std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()),
                                                           (total + ((size_t)64 << 20) - 1) / ((size_t)64 << 20));
  nt = std::max(1, std::min(nt, 64));
  auto work = [&](size_t lo, size_t hi) { // bytes [lo, hi) of the whole batch
    while (lo < hi) {
      const size_t ci = lo / chunk_bytes, c0 = ci * chunk_bytes;
      const size_t upto = std::min(hi, c0 + chunk_bytes);
      vhp_expand_bytes(h->view(h->chunks[ci]), lo - c0, upto - c0, (char *)out + (lo - b_begin));
      lo = upto;
    }
  };
  if (nt == 1) {
    work(b_begin, b_end);
    return VHP_OK;
  }
  const size_t per = ((total / nt) + 127) & ~(size_t)127;
  std::vector<std::thread> th;
  try {
    for (int t = 0; t < nt; ++t) {
      const size_t lo = b_begin + std::min(total, (size_t)t * per), hi = t == nt - 1 ? b_end : b_begin + std::min(total, (size_t)(t + 1) * per);
      if (lo < hi) th.emplace_back(work, lo, hi);
    }
  } catch (...) { // no more threads to be had: the caller's thread does the rest
    const size_t done = th.size();
    for (auto &t : th) t.join();
    work(b_begin + std::min(total, done * per), b_end);
    return VHP_OK;
  }
  for (auto &t : th) t.join();
  return VHP_OK;
}

void vhp_packed_destroy(vhp_packed *h) {
  if (!h) return;
  int cur = 0;
  const bool have = cudaGetDevice(&cur) == cudaSuccess;
  if (!h->blocks.empty()) cudaSetDevice(h->device);
  for (auto &b : h->blocks) cudaFreeHost(b.p);
  if (have) cudaSetDevice(cur);
  (void)cudaGetLastError();
  delete h;
}

vhp_status vhp_visibility_variant_batch_dev(vhp_context *ctx, const uint8_t *d_occ, int nmaps, int nx,
                                            int ny, const int32_t *d_src_xy, const int32_t *d_src_map,
                                            int64_t npairs, const vhp_sweep_variant *variant,
                                            vhp_dtype dtype, void *d_out) {
  vhp_status st = check_common(ctx, d_occ, nmaps, nx, ny, d_src_xy, npairs, dtype, d_out);
  if (st != VHP_OK) return st;
  if ((st = check_variant(ctx, variant)) != VHP_OK) return st;
  VHP_ON_DEVICE(ctx);
  if (npairs == 0) return VHP_OK;
  return run_variant_dev(ctx, d_occ, nx, ny, d_src_xy, d_src_map, npairs, *variant, dtype, d_out);
}

vhp_status vhp_visibility_variant_batch(vhp_context *ctx, const uint8_t *occ, int nmaps, int nx, int ny,
                                        const int32_t *src_xy, const int32_t *src_map, int64_t npairs,
                                        const vhp_sweep_variant *variant, vhp_dtype dtype, void *out) {
  vhp_status st = check_common(ctx, occ, nmaps, nx, ny, src_xy, npairs, dtype, out);
  if (st != VHP_OK) return st;
  if ((st = check_variant(ctx, variant)) != VHP_OK) return st;
  if ((st = check_points(ctx, src_xy, 2, src_map, npairs, nmaps, nx, ny, "vhp_visibility_variant_batch")) != VHP_OK)
    return st;
  if (npairs == 0) return VHP_OK;
  VHP_ON_DEVICE(ctx);
  const size_t cells = (size_t)nx * ny, esz = dtype == VHP_F32 ? 4 : 8, occ_bytes = (size_t)nmaps * cells;
  if ((st = ensure(ctx, ctx->b_occ, occ_bytes)) != VHP_OK) return st;
  if ((st = ensure(ctx, ctx->b_src, (size_t)npairs * 8)) != VHP_OK) return st;
  if (src_map && (st = ensure(ctx, ctx->b_map, (size_t)npairs * 4)) != VHP_OK) return st;
  VHP_CUDA(ctx, cudaMemcpyAsync(ctx->b_occ.p, occ, occ_bytes, cudaMemcpyHostToDevice, ctx->stream));
  VHP_CUDA(ctx, cudaMemcpyAsync(ctx->b_src.p, src_xy, (size_t)npairs * 8, cudaMemcpyHostToDevice, ctx->stream));
  if (src_map)
    VHP_CUDA(ctx, cudaMemcpyAsync(ctx->b_map.p, src_map, (size_t)npairs * 4, cudaMemcpyHostToDevice, ctx->stream));
  ctx->planes_sticky = false; // b_occ content changed
  ctx->tile_src = nullptr;
  // results leave in chunks of at most 1 GB through one device buffer
  const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(npairs, (int64_t)(((size_t)1 << 30) / (cells * esz))));
  if ((st = ensure(ctx, ctx->b_out[0], (size_t)chunk * cells * esz)) != VHP_OK) return st;
  for (int64_t p0 = 0; p0 < npairs; p0 += chunk) {
    const int64_t np = std::min(chunk, npairs - p0);
    st = run_variant_dev(ctx, (const uint8_t *)ctx->b_occ.p, nx, ny, (const int32_t *)ctx->b_src.p + 2 * p0,
                         src_map ? (const int32_t *)ctx->b_map.p + p0 : nullptr, np, *variant, dtype,
                         ctx->b_out[0].p);
    if (st != VHP_OK) return st;
    VHP_CUDA(ctx, cudaMemcpyAsync((char *)out + (size_t)p0 * cells * esz, ctx->b_out[0].p, (size_t)np * cells * esz,
                                  cudaMemcpyDeviceToHost, ctx->stream));
    VHP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return check_device_error(ctx);
}

vhp_status vhp_visibility_batch_bin(vhp_context *ctx, const uint8_t *occ, int nmaps, int nx, int ny,
                                    const int32_t *src_xy, const int32_t *src_map, int64_t npairs,
                                    double threshold, uint32_t *out_bits) {
  return run_bin_host(ctx, occ, nmaps, nx, ny, src_xy, src_map, npairs, threshold, out_bits);
}

vhp_status vhp_visibility_batch_runs(vhp_context *ctx, const uint8_t *occ, int nmaps, int nx, int ny,
                                     const int32_t *src_xy, const int32_t *src_map, int64_t npairs,
                                     double threshold, uint16_t *row_count, uint64_t *pair_ptr,
                                     uint16_t *trans, int64_t trans_cap, int64_t *trans_used) {
  return run_runs_host(ctx, occ, nmaps, nx, ny, src_xy, src_map, npairs, threshold, row_count, pair_ptr, trans,
                       trans_cap, trans_used);
}

vhp_status vhp_runs_to_bits(const uint16_t *row_count, const uint64_t *pair_ptr, const uint16_t *trans,
                            int64_t npairs, int nx, int ny, uint32_t *out_bits) {
  if (!row_count || !pair_ptr || !out_bits || npairs < 0 || nx < 1 || ny < 1)
    return fail(nullptr, VHP_ERR_INVALID_ARG, "vhp_runs_to_bits: bad argument");
  const size_t wpr = (size_t)(nx + 31) / 32;
  std::memset(out_bits, 0, (size_t)npairs * ny * wpr * 4);
  for (int64_t p = 0; p < npairs; ++p) {
    const uint16_t *t = trans + pair_ptr[p];
    for (int y = 0; y < ny; ++y) {
      const int c = row_count[(size_t)p * ny + y];
      uint32_t *row = out_bits + ((size_t)p * ny + y) * wpr;
      for (int k = 0; k + 1 < c; k += 2)
        for (int x = t[k]; x < t[k + 1] && x < nx; ++x) row[x >> 5] |= 1u << (x & 31);
      t += c;
    }
  }
  return VHP_OK;
}

vhp_status vhp_visibility_batch_bin_dev(vhp_context *ctx, const uint8_t *d_occ, int nmaps, int nx,
                                        int ny, const int32_t *d_src_xy, const int32_t *d_src_map,
                                        int64_t npairs, double threshold, uint32_t *d_out_bits) {
  vhp_status st = check_common(ctx, d_occ, nmaps, nx, ny, d_src_xy, npairs, VHP_F64, d_out_bits);
  if (st != VHP_OK) return st;
  VHP_ON_DEVICE(ctx);
  return run_bin_dev(ctx, d_occ, nmaps, nx, ny, d_src_xy, d_src_map, npairs, threshold, d_out_bits);
}

vhp_status vhp_raycast_batch_dev(vhp_context *ctx, const uint8_t *d_occ, int nmaps, int nx,
                                 int ny, const int32_t *d_src_xy, const int32_t *d_src_map,
                                 int64_t npairs, vhp_dtype dtype, void *d_out) {
  vhp_status st = check_common(ctx, d_occ, nmaps, nx, ny, d_src_xy, npairs, dtype, d_out);
  if (st != VHP_OK) return st;
  VHP_ON_DEVICE(ctx);
  return run_dev(ctx, Op::Raycast, d_occ, nmaps, nx, ny, d_src_xy, d_src_map, npairs, dtype,
                 d_out);
}

vhp_status vhp_raycast_batch(vhp_context *ctx, const uint8_t *occ, int nmaps, int nx, int ny,
                             const int32_t *src_xy, const int32_t *src_map, int64_t npairs,
                             vhp_dtype dtype, void *out) {
  return run_host(ctx, Op::Raycast, occ, nmaps, nx, ny, src_xy, src_map, npairs, dtype, out);
}

vhp_status vhp_selftest_ratio(vhp_context *ctx, int kmax, int64_t *mismatches) {
  if (!ctx || !mismatches || kmax < 1 || kmax > 16384)
    return fail(ctx, VHP_ERR_INVALID_ARG, "vhp_selftest_ratio: bad argument");
  VHP_ON_DEVICE(ctx);
  vhp_status st = ensure_rcp2(ctx, kmax + 8);
  if (st != VHP_OK) return st;
  if ((st = ensure(ctx, ctx->b_misc, 64)) != VHP_OK) return st;
  VHP_CUDA(ctx, cudaMemsetAsync(ctx->b_misc.p, 0, 8, ctx->stream));
  VHP_CUDA(ctx, vhp_launch_ratio2_selftest(ctx->rcp2_table, kmax,
                                           (unsigned long long *)ctx->b_misc.p, ctx->stream,
                                           &ctx->launches));
  unsigned long long bad = 0;
  VHP_CUDA(ctx, cudaMemcpyAsync(&bad, ctx->b_misc.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
  VHP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *mismatches = (int64_t)bad;
  return VHP_OK;
}

uint32_t vhp_environment_draw(uint64_t seed, uint64_t map, uint64_t obstacle, uint32_t d) {
  return vhp_env_draw(seed, map, obstacle, d);
}

vhp_status vhp_environment_generate_batch_dev(vhp_context *ctx, const vhp_config *cfg,
                                              uint64_t seed, int64_t first_map, int nmaps,
                                              uint8_t *d_occ) {
  if (!ctx || !cfg || !d_occ || nmaps < 1 || first_map < 0)
    return fail(ctx, VHP_ERR_INVALID_ARG, "vhp_environment_generate_batch_dev: bad argument");
  if (cfg->ncols < 1 || cfg->nrows < 1 || cfg->ncols > 16384 || cfg->nrows > 16384 ||
      cfg->nb_of_obstacles < 0 || cfg->nb_of_obstacles > 0x7fffffff || cfg->min_width < 0 ||
      cfg->min_height < 0 || cfg->max_width < cfg->min_width || cfg->max_height < cfg->min_height)
    return fail(ctx, VHP_ERR_INVALID_ARG, "vhp_environment_generate_batch_dev: bad environment settings");
  VHP_ON_DEVICE(ctx);
  VHP_CUDA(ctx, vhp_launch_env_generate(d_occ, nmaps, (int)cfg->ncols, (int)cfg->nrows, first_map,
                                        seed, cfg->nb_of_obstacles, cfg->min_width, cfg->max_width,
                                        cfg->min_height, cfg->max_height, ctx->stream,
                                        &ctx->launches));
  if (ctx->tile_src == d_occ) { // bit planes of an older content of this buffer
    ctx->planes_sticky = false;
    ctx->tile_src = nullptr;
  }
  return VHP_OK;
}

vhp_status vhp_context_set_result_transport(vhp_context *ctx, int mode) {
  if (!ctx || mode < 0 || mode > 2)
    return fail(ctx, VHP_ERR_INVALID_ARG, "vhp_context_set_result_transport: mode is 0, 1 or 2");
  ctx->result_transport = mode;
  return VHP_OK;
}

vhp_status vhp_context_set_result_gpu_share(vhp_context *ctx, int sixteenths) {
  if (!ctx || sixteenths < 0 || sixteenths > 16)
    return fail(ctx, VHP_ERR_INVALID_ARG, "vhp_context_set_result_gpu_share: 0..16");
  ctx->result_gpu_share = sixteenths;
  return VHP_OK;
}

vhp_status vhp_context_last_transport(const vhp_context *ctx, int64_t *d2h_bytes,
                                      int64_t *result_bytes, int32_t *packed) {
  if (!ctx) return fail(nullptr, VHP_ERR_INVALID_ARG, "null context");
  if (d2h_bytes) *d2h_bytes = ctx->last_d2h_bytes;
  if (result_bytes) *result_bytes = ctx->last_result_bytes;
  if (packed) *packed = ctx->last_transport_packed;
  return VHP_OK;
}

vhp_status vhp_expand_packed_chunk(const uint32_t *mask, const uint32_t *word_base,
                                   const uint32_t *vmask, int elem_bytes, const void *literals,
                                   int64_t nunits, int64_t valid_bytes, void *dst, int threads) {
  if (elem_bytes != 4 && elem_bytes != 8)
    return fail(nullptr, VHP_ERR_INVALID_ARG, "vhp_expand_packed_chunk: elem_bytes is 4 or 8");
  if (nunits < 0 || valid_bytes < 0 || valid_bytes > nunits * (int64_t)kVhpPackUnit ||
      valid_bytes <= (nunits - 1) * (int64_t)kVhpPackUnit)
    return fail(nullptr, VHP_ERR_INVALID_ARG, "vhp_expand_packed_chunk: nunits does not match valid_bytes");
  if (nunits == 0) return VHP_OK;
  if (!mask || !word_base || !vmask || !dst)
    return fail(nullptr, VHP_ERR_INVALID_ARG, "vhp_expand_packed_chunk: null argument");
  VhpPackedChunk c;
  c.mask = mask;
  c.word_base = word_base;
  c.vmask = vmask;
  c.elem_bytes = elem_bytes;
  c.literals = (const char *)literals;
  c.dst = (char *)dst;
  c.nunits = nunits;
  c.valid_bytes = (size_t)valid_bytes;
  VhpExpandPool pool(std::max(1, std::min(threads, 64)));
  pool.wait(pool.submit(c));
  return VHP_OK;
}

vhp_status vhp_expand_packed_range(const uint32_t *mask, const uint32_t *word_base,
                                   const uint32_t *vmask, int elem_bytes, const void *literals,
                                   int64_t nunits, int64_t valid_bytes, int64_t b0, int64_t b1, void *out) {
  if (elem_bytes != 4 && elem_bytes != 8)
    return fail(nullptr, VHP_ERR_INVALID_ARG, "vhp_expand_packed_range: elem_bytes is 4 or 8");
  if (nunits < 0 || valid_bytes < 0 || valid_bytes > nunits * (int64_t)kVhpPackUnit ||
      valid_bytes <= (nunits - 1) * (int64_t)kVhpPackUnit)
    return fail(nullptr, VHP_ERR_INVALID_ARG, "vhp_expand_packed_range: nunits does not match valid_bytes");
  if (b0 < 0 || b1 < b0 || b1 > valid_bytes || b0 % elem_bytes != 0)
    return fail(nullptr, VHP_ERR_INVALID_ARG, "vhp_expand_packed_range: bad byte range");
  if (b0 == b1) return VHP_OK;
  if (!mask || !word_base || !vmask || !literals || !out)
    return fail(nullptr, VHP_ERR_INVALID_ARG, "vhp_expand_packed_range: null argument");
  VhpPackedChunk c;
  c.mask = mask;
  c.word_base = word_base;
  c.vmask = vmask;
  c.elem_bytes = elem_bytes;
  c.literals = (const char *)literals;
  c.nunits = nunits;
  c.valid_bytes = (size_t)valid_bytes;
  vhp_expand_bytes(c, (size_t)b0, (size_t)b1, (char *)out);
  return VHP_OK;
}

vhp_status vhp_context_set_planner_loop(vhp_context *ctx, int mode) {
  if (!ctx || mode < 0 || mode > 3) return fail(ctx, VHP_ERR_INVALID_ARG, "vhp_context_set_planner_loop: bad argument");
  ctx->grid_loop_mode = mode;
  return VHP_OK;
}

vhp_status vhp_context_set_grid_sweep(vhp_context *ctx, int mode) {
  if (!ctx || mode < 0 || mode > 2) return fail(ctx, VHP_ERR_INVALID_ARG, "vhp_context_set_grid_sweep: bad argument");
  ctx->grid_sweep = mode;
  return VHP_OK;
}

void vhp_strip_halo_rows(int nx, int ny, int sx, int sy, int y0, int y1, int32_t rows[4]) {
  vhp_window_halo_rows(nx, ny, sx, sy, y0, y1, rows);
}

vhp_status vhp_strip_sweep_dev(vhp_context *ctx, const uint8_t *d_occ, int nx, int ny, int sx,
                               int sy, int y0, int y1, const double *const d_halo[4],
                               vhp_dtype dtype, void *d_vis_strip) {
  if (!ctx || !d_occ || !d_vis_strip || nx < 1 || ny < 1 || y0 < 0 || y1 > ny || y0 >= y1)
    return fail(ctx, VHP_ERR_INVALID_ARG, "vhp_strip_sweep_dev: bad argument");
  if (sx < 0 || sx >= nx || sy < 0 || sy >= ny)
    return fail(ctx, VHP_ERR_INVALID_ARG, "vhp_strip_sweep_dev: source outside the grid");
  if (dtype != VHP_F32 && dtype != VHP_F64) return fail(ctx, VHP_ERR_INVALID_ARG, "bad dtype");
  const bool cta_ok = vhp_sweep_tile_supported(nx, ny);
  if (!cta_ok && !(ctx->grid_sweep != 0 && vhp_sweep_grid_supported(nx, ny)))
    return fail(ctx, VHP_ERR_UNSUPPORTED, "vhp_strip_sweep_dev: grid too wide for the tile kernel");
  int32_t rows[4];
  vhp_window_halo_rows(nx, ny, sx, sy, y0, y1, rows);
  for (int q = 0; q < 4; ++q)
    if (rows[q] >= 0 && (!d_halo || !d_halo[q]))
      return fail(ctx, VHP_ERR_INVALID_ARG, "vhp_strip_sweep_dev: missing halo row");
  VHP_ON_DEVICE(ctx);
  vhp_status st = ensure_rcp2(ctx, std::max(nx, ny) + 64);
  if (st != VHP_OK) return st;
  if ((st = pack_tile(ctx, d_occ, 1, nx, ny, false)) != VHP_OK) return st;
  // Large windows are swept by many CTAs at once (grid mode): enough 8-warp CTAs for the tile
  // rows of the window, at most two per SM.
  const int grid_mode = ctx->grid_sweep;
  int grid_ctas = 1;
  if (!cta_ok || (grid_mode && vhp_sweep_grid_supported(nx, ny) &&
                  (int64_t)nx * (y1 - y0) >= (grid_mode > 1 ? 0 : (1 << 20)))) {
    grid_ctas = grid_sweep_ctas(ctx, y1 - y0);
    if ((st = ensure(ctx, ctx->b_grid, vhp_sweep_grid_ws_bytes(nx, ny))) != VHP_OK) return st;
  }
  VHP_CUDA(ctx, vhp_launch_sweep_window(ctx->tile, nx, ny, sx, sy, y0, y1, d_halo, dtype,
                                        d_vis_strip, ctx->rcp2_table, ctx->d_err,
                                        grid_ctas > 1 ? ctx->b_grid.p : nullptr, grid_ctas,
                                        ctx->stream, &ctx->launches));
  return VHP_OK;
}

vhp_status vhp_strip_epilogue_dev(vhp_context *ctx, int nx, int ny, int y0, int y1, int sx, int sy,
                                  int ex, int ey, double threshold, int32_t nb,
                                  const int32_t *d_light_sources, const double *d_vis_strip,
                                  double *d_vg_strip, double *d_h_strip, int32_t *d_came_strip,
                                  uint64_t *d_best) {
  if (!ctx || !d_light_sources || !d_vis_strip || !d_vg_strip || !d_h_strip || !d_came_strip ||
      !d_best || nx < 1 || ny < 1 || y0 < 0 || y1 > ny || y0 >= y1 || nb < 0)
    return fail(ctx, VHP_ERR_INVALID_ARG, "vhp_strip_epilogue_dev: bad argument");
  VHP_ON_DEVICE(ctx);
  const int nblocks = vhp_strip_epilogue_blocks(ctx->sm_count);
  vhp_status st = ensure(ctx, ctx->b_misc, (size_t)nblocks * 16 + 64);
  if (st != VHP_OK) return st;
  VHP_CUDA(ctx, vhp_launch_strip_epilogue(nx, ny, y0, y1, sx, sy, ex, ey, threshold, nb,
                                          d_light_sources, d_vis_strip, d_vg_strip, d_h_strip,
                                          d_came_strip, (unsigned long long *)ctx->b_misc.p,
                                          nblocks, (unsigned long long *)d_best, ctx->stream,
                                          &ctx->launches));
  return VHP_OK;
}

vhp_status vhp_planner_batch_dev(vhp_context *ctx, const uint8_t *d_occ, int nmaps, int nx,
                                 int ny, const int32_t *d_se_xy, const int32_t *d_prob_map,
                                 int64_t nprob, double threshold, int32_t max_iter,
                                 int32_t ls_cap, vhp_dtype dtype, const vhp_planner_out *d_out) {
  static const vhp_planner_out none = {};
  vhp_status st = check_common(ctx, d_occ, nmaps, nx, ny, d_se_xy, nprob, dtype, ctx);
  if (st != VHP_OK) return st;
  VHP_ON_DEVICE(ctx);
  return planner_dev(ctx, d_occ, nmaps, nx, ny, d_se_xy, d_prob_map, nprob, threshold, max_iter,
                     ls_cap, dtype, d_out ? *d_out : none);
}

vhp_status vhp_planner_batch(vhp_context *ctx, const uint8_t *occ, int nmaps, int nx, int ny,
                             const int32_t *se_xy, const int32_t *prob_map, int64_t nprob,
                             double threshold, int32_t max_iter, int32_t ls_cap, vhp_dtype dtype,
                             const vhp_planner_out *out) {
  static const vhp_planner_out none = {};
  const vhp_planner_out &h = out ? *out : none;
  vhp_status st = check_common(ctx, occ, nmaps, nx, ny, se_xy, nprob, dtype, ctx);
  if (st != VHP_OK) return st;
  if (prob_map)
    if ((st = check_points(ctx, nullptr, 0, prob_map, nprob, nmaps, nx, ny, "vhp_planner_batch")) != VHP_OK)
      return st;
  if (nprob == 0) return VHP_OK;
  VHP_ON_DEVICE(ctx);
  const size_t cells = (size_t)nx * ny, esz = dtype == VHP_F32 ? 4 : 8;
  if ((st = ensure(ctx, ctx->b_occ, (size_t)nmaps * cells)) != VHP_OK) return st;
  if ((st = ensure(ctx, ctx->b_src, (size_t)nprob * 16)) != VHP_OK) return st;
  if (prob_map && (st = ensure(ctx, ctx->b_map, (size_t)nprob * 4)) != VHP_OK) return st;
  VHP_CUDA(ctx, cudaMemcpyAsync(ctx->b_occ.p, occ, (size_t)nmaps * cells, cudaMemcpyHostToDevice, ctx->stream));
  VHP_CUDA(ctx, cudaMemcpyAsync(ctx->b_src.p, se_xy, (size_t)nprob * 16, cudaMemcpyHostToDevice, ctx->stream));
  if (prob_map)
    VHP_CUDA(ctx, cudaMemcpyAsync(ctx->b_map.p, prob_map, (size_t)nprob * 4, cudaMemcpyHostToDevice, ctx->stream));
  ctx->tile_src = nullptr;
  ctx->planes_sticky = false;
  if ((st = pack_tile(ctx, (const uint8_t *)ctx->b_occ.p, nmaps, nx, ny, true)) != VHP_OK) return st;
  ctx->planes_sticky = true;
  // device staging for the requested outputs, a chunk of problems at a time
  const size_t per_prob = (h.vis ? esz * cells : 0) + (h.vg ? esz * cells : 0) + (h.came ? 4 * cells : 0) +
                          4 + 4 + 8 + 4 + 2 * (size_t)ls_cap * 8 + 64;
  const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(nprob, (int64_t)(((size_t)2 << 30) / per_prob)));
  vhp_status result = VHP_OK;
  if ((st = ensure(ctx, ctx->b_out[0], (size_t)chunk * per_prob + 256)) != VHP_OK) result = st;
  for (int64_t q0 = 0; q0 < nprob && result == VHP_OK; q0 += chunk) {
    const int64_t n = std::min(chunk, nprob - q0);
    char *w = (char *)ctx->b_out[0].p;
    auto take = [&](size_t bytes) { char *r = w; w += (bytes + 15) & ~(size_t)15; return r; };
    vhp_planner_out d = {};
    d.vis = h.vis ? take(esz * cells * n) : nullptr;
    d.vg = h.vg ? take(esz * cells * n) : nullptr;
    d.came = h.came ? (int32_t *)take(4 * cells * n) : nullptr;
    d.status = (int32_t *)take(4 * n);
    d.nb_sources = (int32_t *)take(4 * n);
    d.light_sources = (int32_t *)take(8 * (size_t)ls_cap * n);
    d.path_len = (double *)take(8 * n);
    d.path_n = (int32_t *)take(4 * n);
    d.path = (int32_t *)take(8 * (size_t)ls_cap * n);
    result = planner_dev(ctx, (const uint8_t *)ctx->b_occ.p, nmaps, nx, ny,
                         (const int32_t *)ctx->b_src.p + 4 * q0,
                         prob_map ? (const int32_t *)ctx->b_map.p + q0 : nullptr, n, threshold,
                         max_iter, ls_cap, dtype, d);
    if (result != VHP_OK) break;
    auto back = [&](void *dst, const void *src, size_t bytes) {
      if (dst && result == VHP_OK) {
        cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream);
        if (e != cudaSuccess) result = cuda_fail(ctx, e, "planner D2H");
      }
    };
    back(h.vis ? (char *)h.vis + q0 * cells * esz : nullptr, d.vis, esz * cells * n);
    back(h.vg ? (char *)h.vg + q0 * cells * esz : nullptr, d.vg, esz * cells * n);
    back(h.came ? h.came + q0 * cells : nullptr, d.came, 4 * cells * n);
    back(h.status ? h.status + q0 : nullptr, d.status, 4 * n);
    back(h.nb_sources ? h.nb_sources + q0 : nullptr, d.nb_sources, 4 * n);
    back(h.light_sources ? h.light_sources + 2 * (size_t)ls_cap * q0 : nullptr, d.light_sources, 8 * (size_t)ls_cap * n);
    back(h.path_len ? h.path_len + q0 : nullptr, d.path_len, 8 * n);
    back(h.path_n ? h.path_n + q0 : nullptr, d.path_n, 4 * n);
    back(h.path ? h.path + 2 * (size_t)ls_cap * q0 : nullptr, d.path, 8 * (size_t)ls_cap * n);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess && result == VHP_OK) result = cuda_fail(ctx, e, "planner sync");
  }
  ctx->planes_sticky = false;
  ctx->tile_src = nullptr;
  if (result != VHP_OK) return result;
  return check_device_error(ctx);
}

void vhp_export_came_from_u64(const int32_t *came, int64_t n, uint64_t *out) {
  for (int64_t i = 0; i < n; ++i)
    out[i] = came[i] < 0 ? VHP_NO_PARENT_U64 : (uint64_t)came[i];
}

} // extern "C"
