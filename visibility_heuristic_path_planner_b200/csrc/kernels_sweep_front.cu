// kernels_sweep_front.cu -- K1, the tuned batched visibility sweep for sm_100a.
//
// Replaces visibilityBasedSolver::computeVisibility
// (reference src/visibilityBasedSolver.cpp:570-696) for batches of (map, source)
// pairs.  Same L-front dynamic program and arithmetic contract as the simple
// kernel in kernels_sweep.cu (which stays as the in-library cross-check).
//
// One CTA per (map, source) pair.  Thread `tid` owns the four ABSOLUTE coordinates
// u = 4*tid + e (e = 0..3), used both as x for the two "row fronts" and as y for
// the two "column fronts":
//     RU: cells (x, sy+k)   RD: cells (x, sy-k)     |x-sx| < k   (row parts)
//     CR: cells (sx+k, y)   CL: cells (sx-k, y)     |y-sy| < k   (column parts)
// Each of the 16 front values a thread owns stays in an fp64 register for the
// whole sweep.  The upstream neighbour b comes from the adjacent register, one
// warp shuffle at thread boundaries and a double-buffered shared slot at warp
// boundaries.  c = |u-s|/k costs one table read of RN(1/k) plus three fp64 ops
// and is shared by the two fronts of a type.  Occupancy is one 32-bit word of the
// bit planes per front per step, prefetched one step ahead.  Elements that are
// not on the front yet are held at 0 by masking their occupancy bit, so the
// update itself is branch-free; the element that joins the front in a step is
// patched in a short rare path.
//
// Stores: row fronts write 4 consecutive x per thread (one 128-bit store when the
// row pitch allows it).  Column fronts produce one x per step, so each warp
// stages S = 32 B / sizeof(OutT) steps of its 128 rows in shared memory
// (tile[kk][y], pitch 132: conflict-free both ways) and flushes sector-sized row
// segments aligned on absolute X.
//
// Diagonal cell (k,k) of a quadrant = q[k][k-1]*occ(k,k) (the reference has no
// i==j branch, so `v` keeps the value of the previous inner iteration).
// q[k][k-1] is a column-front value: the column owner of |y-sy| == k derives the
// diagonal from its own neighbour, the row owner of |x-sx| == k receives
// q[k][k-1] through a 4-entry shared slot and also stores the diagonal cell.
#include <cstdint>

#include "sweep_front_body.cuh"

namespace {


template <typename OutT, bool VEC, int MAXNT, int MINB>
__global__ void __launch_bounds__(MAXNT, MINB) sweep_front_kernel(const FrontParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int64_t pair = blockIdx.x;
  const int nx = p.nx, ny = p.ny;
  const int sx = __ldg(p.src_xy + 2 * pair), sy = __ldg(p.src_xy + 2 * pair + 1);
  if ((unsigned)sx >= (unsigned)nx || (unsigned)sy >= (unsigned)ny) { // CTA-uniform
    if (threadIdx.x == 0) atomicOr(p.err, 1);
    return;
  }
  const int map = p.src_map ? __ldg(p.src_map + pair) : 0;
  sweep_front_body<OutT, VEC>(p, sx, sy, p.rowbits + (size_t)map * p.row_plane,
                              p.colbits + (size_t)map * p.col_plane,
                              reinterpret_cast<OutT *>(p.out) + (size_t)pair * nx * ny, smem_raw);
}

template <typename OutT, bool VEC, int MAXNT, int MINB>
cudaError_t launch_front(const FrontParams &p, int64_t npairs, int nt, size_t smem,
                         cudaStream_t st) {
  auto kern = sweep_front_kernel<OutT, VEC, MAXNT, MINB>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<(unsigned)npairs, nt, smem, st>>>(p);
  return cudaGetLastError();
}

template <typename OutT>
cudaError_t launch_front_t(const FrontParams &p, int64_t npairs, int nt, size_t smem, bool vec,
                           cudaStream_t st) {
  if (nt <= 256)
    return vec ? launch_front<OutT, true, 256, 2>(p, npairs, nt, smem, st)
               : launch_front<OutT, false, 256, 2>(p, npairs, nt, smem, st);
  return vec ? launch_front<OutT, true, 1024, 1>(p, npairs, nt, smem, st)
             : launch_front<OutT, false, 1024, 1>(p, npairs, nt, smem, st);
}

} // namespace

bool vhp_sweep_front_supported(int nx, int ny) {
  int nt; uint32_t p2; size_t smem;
  front_geometry(nx, ny, VHP_F64, nt, p2, smem);
  return nt <= 1024 && smem <= 227 * 1024;
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
  int nt; size_t smem;
  front_geometry(nx, ny, dtype, nt, p.edge_p2, smem);
  if (smem > 227 * 1024 || nt > 1024) return cudaErrorInvalidConfiguration;
  // 128-bit stores need 16-byte aligned rows: nx % 4 == 0 (f32) / nx % 2 == 0 (f64)
  const bool vec = (dtype == VHP_F32) ? (nx % 4 == 0) : (nx % 2 == 0);
  cudaError_t e = dtype == VHP_F32 ? launch_front_t<float>(p, npairs, nt, smem, vec, st)
                                   : launch_front_t<double>(p, npairs, nt, smem, vec, st);
  if (launches) *launches += 1;
  return e;
}
