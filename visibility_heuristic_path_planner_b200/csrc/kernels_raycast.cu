// kernels_raycast.cu -- K4: batched ray casting, the secondary method the
// reference benchmarks its sweep against.
//
// Replaces visibilityBasedSolver::raycasting (src/visibilityBasedSolver.cpp:267-290)
// called for every target cell as in benchmark() (:228-232).  The reference walks a
// Bresenham line from the source to each target; the first occupied cell met
// BEFORE reaching the target zeroes itself and the target in
// visibilityRayCasting_ (initialised to 1.0, :45).  The target cell itself is
// never tested.  All writes are stores of 0 on a field of 1, so any execution
// order gives the same field: one thread per target cell.
#include <cstdint>

#include "vhp_internal.h"

namespace {

template <typename OutT>
__global__ void fill_ones_kernel(OutT *out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    out[i] = (OutT)1;
}

// grid: x = target tiles of 256 cells (x fastest), y = pair
template <typename OutT>
__global__ void __launch_bounds__(256)
raycast_kernel(const uint8_t *__restrict__ occ, int nx, int ny,
               const int32_t *__restrict__ src_xy, const int32_t *__restrict__ src_map,
               OutT *__restrict__ out_all, int *__restrict__ err_flag) {
  const int64_t pair = blockIdx.y;
  const size_t cell = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (cell >= (size_t)nx * ny) return;
  const int x1 = (int)(cell % nx), y1 = (int)(cell / nx);
  int x0 = __ldg(src_xy + 2 * pair), y0 = __ldg(src_xy + 2 * pair + 1);
  if ((unsigned)x0 >= (unsigned)nx || (unsigned)y0 >= (unsigned)ny) {
    if (cell == 0) atomicOr(err_flag, 1);
    return;
  }
  const int map = src_map ? __ldg(src_map + pair) : 0;
  const uint8_t *__restrict__ o = occ + (size_t)map * nx * ny;
  OutT *__restrict__ out = out_all + (size_t)pair * nx * ny;

  const int dx = abs(x1 - x0), dy = abs(y1 - y0);
  const int sx = (x0 < x1) ? 1 : -1, sy = (y0 < y1) ? 1 : -1;
  int err = dx - dy;
  while (x0 != x1 || y0 != y1) {
    if (__ldg(o + (size_t)y0 * nx + x0) == 0) {
      out[(size_t)y0 * nx + x0] = (OutT)0;
      out[(size_t)y1 * nx + x1] = (OutT)0;
      return;
    }
    const int e2 = 2 * err;
    if (e2 > -dy) { err -= dy; x0 += sx; }
    if (e2 < dx)  { err += dx; y0 += sy; }
  }
}

} // namespace

cudaError_t vhp_launch_raycast(const uint8_t *d_occ, int nx, int ny, const int32_t *d_src_xy,
                               const int32_t *d_src_map, int64_t npairs, vhp_dtype dtype,
                               void *d_out, int *d_err, cudaStream_t st, int64_t *launches) {
  const size_t cells = (size_t)nx * ny;
  const size_t n = cells * (size_t)npairs;
  const unsigned fill_grid = (unsigned)std::min<size_t>((n + 255) / 256, 148u * 16u);
  const unsigned tiles = (unsigned)((cells + 255) / 256);
  // gridDim.y is limited to 65535: chunk the pair axis
  if (dtype == VHP_F32) fill_ones_kernel<float><<<fill_grid, 256, 0, st>>>((float *)d_out, n);
  else fill_ones_kernel<double><<<fill_grid, 256, 0, st>>>((double *)d_out, n);
  if (launches) *launches += 1;
  for (int64_t p0 = 0; p0 < npairs; p0 += 65535) {
    const unsigned np = (unsigned)std::min<int64_t>(65535, npairs - p0);
    dim3 grid(tiles, np);
    if (dtype == VHP_F32)
      raycast_kernel<float><<<grid, 256, 0, st>>>(d_occ, nx, ny, d_src_xy + 2 * p0,
                                                  d_src_map ? d_src_map + p0 : nullptr,
                                                  (float *)d_out + (size_t)p0 * cells, d_err);
    else
      raycast_kernel<double><<<grid, 256, 0, st>>>(d_occ, nx, ny, d_src_xy + 2 * p0,
                                                   d_src_map ? d_src_map + p0 : nullptr,
                                                   (double *)d_out + (size_t)p0 * cells, d_err);
    if (launches) *launches += 1;
  }
  return cudaGetLastError();
}
