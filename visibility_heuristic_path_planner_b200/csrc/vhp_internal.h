// vhp_internal.h -- shared declarations between the C-ABI layer and the kernels.
#ifndef VHP_INTERNAL_H
#define VHP_INTERNAL_H

#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "vhp.h"

// Packed occupancy planes of one batch of maps (built by vhp_pack_maps):
//   rowbits[m][y][wx]  bit (x & 31) of word x>>5 = occupancy(x, y)   (row-major)
//   colbits[m][x][wy]  bit (y & 31) of word y>>5 = occupancy(x, y)   (column-major)
// Row pitch is padded (+1 word, rounded up to 4 words) so that a lookahead read
// one word past the last data word stays inside the row.
struct VhpPackedMaps {
  const uint32_t *rowbits = nullptr;
  const uint32_t *colbits = nullptr;
  int wpr = 0;            // words per row of rowbits
  int wpc = 0;            // words per column of colbits
  size_t row_plane = 0;   // words per map in rowbits (= ny * wpr)
  size_t col_plane = 0;   // words per map in colbits (= nx * wpc)
};

// Bit planes of the octant sweep kernel (built by vhp_launch_pack_oct): forward and
// mirrored, row and column major, 4*NS words per line, never-written border baked
// in as occupied (see kernels_sweep_octant.cu).
struct VhpOctPlanes {
  const uint32_t *row_f = nullptr, *row_r = nullptr, *col_f = nullptr, *col_r = nullptr;
  size_t row_plane = 0, col_plane = 0; // words per map
};

// Bit planes of the tile sweep kernel (built by vhp_launch_pack_tile): row / column
// major, forward / mirrored, wx (wy) words per row (column) line including one zero
// word of padding (see kernels_sweep_tile.cu).
struct VhpTilePlanes {
  const uint32_t *rowF = nullptr, *rowR = nullptr, *colF = nullptr, *colR = nullptr;
  const uint32_t *bsum = nullptr;      // [map][block row][words]: 1 = aligned 32x32 block is free
  int wx = 0, wy = 0;
  size_t row_plane = 0, col_plane = 0; // words per map
};

// grow-only device buffer
struct VhpDevBuf {
  void *p = nullptr;
  size_t cap = 0;
};

struct vhp_context {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr; // D2H of the host-buffer entry points
  bool owns_stream = false;
  int sm_count = 0;
  int64_t launches = 0;
  std::string last_error;
  // device-side error word (bit 0: a source / start / end outside the grid)
  int *d_err = nullptr;
  // workspace buffers (grown on demand, reused across calls)
  VhpDevBuf b_occ, b_src, b_map, b_out[2], b_scratch, b_planner, b_misc;
  cudaEvent_t ev_done[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
  // cached packed maps for _dev calls (vhp_prepare_maps_dev)
  const uint8_t *packed_src = nullptr;
  bool packed_sticky = false; // set by vhp_prepare_maps_dev: reuse planes for this pointer
  int packed_nmaps = 0, packed_nx = 0, packed_ny = 0;
  uint32_t *packed_buf = nullptr;
  size_t packed_bytes = 0;
  VhpPackedMaps packed;
  // 1/k table for the sweep kernels (exact __drcp_rn), device resident
  double *rcp_table = nullptr;
  int rcp_len = 0;
  // double-double 1/k table {rh, rl} of the octant kernel
  double *rcp2_table = nullptr;
  int rcp2_len = 0;
  // octant-kernel bit planes (cached like `packed`)
  const uint8_t *oct_src = nullptr;
  int oct_nmaps = 0, oct_nx = 0, oct_ny = 0;
  uint32_t *oct_buf = nullptr;
  size_t oct_bytes = 0;
  VhpOctPlanes oct;
  // tile-kernel bit planes (cached like `packed`)
  const uint8_t *tile_src = nullptr;
  int tile_nmaps = 0, tile_nx = 0, tile_ny = 0;
  uint32_t *tile_buf = nullptr;
  size_t tile_bytes = 0;
  VhpTilePlanes tile;
  // which K1 implementation vhp_visibility_batch* uses (env VHP_SWEEP_IMPL):
  // 0 = auto (octant kernel where it fits), 1 = naive reference kernel, 2 = front
  // kernel (every thread serves all four fronts), 3 = ring kernel (one front per
  // warp, block barrier per ring), 4 = octant kernel (one octant per warp, no barriers),
  // 5 = tile wavefront kernel (the default where it fits)
  int sweep_impl = 0;
};

// ---- kernel launchers (all enqueue on `st`, return cudaGetLastError()) --------
// Each launcher returns the number of kernel launches it made through *launches.

cudaError_t vhp_launch_pack_maps(const uint8_t *d_occ, int nmaps, int nx, int ny,
                                 uint32_t *d_rowbits, uint32_t *d_colbits, int wpr,
                                 int wpc, cudaStream_t st, int64_t *launches);

cudaError_t vhp_launch_rcp_table(double *d_table, int len, cudaStream_t st,
                                 int64_t *launches);

cudaError_t vhp_launch_ratio_selftest(const double *d_rcp, int kmax,
                                      unsigned long long *d_mismatches, cudaStream_t st,
                                      int64_t *launches);

// K1, straightforward L-front kernel (one CTA per (pair, quadrant), fronts in
// shared/global memory).  Correctness anchor for the tuned kernel.
cudaError_t vhp_launch_sweep_naive(const uint8_t *d_occ, int nx, int ny,
                                   const int32_t *d_src_xy, const int32_t *d_src_map,
                                   int64_t npairs, vhp_dtype dtype, void *d_out,
                                   double *d_scratch, int *d_err, cudaStream_t st,
                                   int64_t *launches);
size_t vhp_sweep_naive_scratch_bytes(int nx, int ny, int64_t npairs);

// K1, tuned front kernel (absolute-coordinate ownership, bit-plane occupancy,
// register-resident fp64 fronts, sector-staged column stores).
bool vhp_sweep_front_supported(int nx, int ny);
cudaError_t vhp_launch_sweep_front(const VhpPackedMaps &maps, int nx, int ny,
                                   const int32_t *d_src_xy, const int32_t *d_src_map,
                                   int64_t npairs, vhp_dtype dtype, void *d_out,
                                   const double *d_rcp, int *d_err, cudaStream_t st,
                                   int64_t *launches);

// K1, front-specialised warps (grids up to 1024 x 1024): the default sweep kernel.
bool vhp_sweep_ring_supported(int nx, int ny);
cudaError_t vhp_launch_sweep_ring(const VhpPackedMaps &maps, int nx, int ny,
                                  const int32_t *d_src_xy, const int32_t *d_src_map,
                                  int64_t npairs, vhp_dtype dtype, void *d_out,
                                  const double *d_rcp, int *d_err, cudaStream_t st,
                                  int64_t *launches);

// K1, one warp per octant, no block barriers (grids up to 1021 x 1021): the default.
bool vhp_sweep_octant_supported(int nx, int ny);
int vhp_oct_words_per_line(int nx, int ny);
cudaError_t vhp_launch_pack_oct(const uint8_t *d_occ, int nmaps, int nx, int ny, uint32_t *row_f,
                                uint32_t *row_r, uint32_t *col_f, uint32_t *col_r,
                                cudaStream_t st, int64_t *launches);
cudaError_t vhp_launch_rcp2_table(double *d_table, int len, cudaStream_t st, int64_t *launches);
cudaError_t vhp_launch_ratio2_selftest(const double *d_rcp2, int kmax,
                                       unsigned long long *d_mismatches, cudaStream_t st,
                                       int64_t *launches);
cudaError_t vhp_launch_sweep_octant(const VhpOctPlanes &pl, const uint8_t *d_occ, int nx, int ny,
                                    const int32_t *d_src_xy, const int32_t *d_src_map,
                                    int64_t npairs, vhp_dtype dtype, void *d_out,
                                    const double *d_rcp2, int *d_err, cudaStream_t st,
                                    int64_t *launches);

// K1, tile wavefront (32 x 32 tiles, uniform tiles are plain fills): the default.
bool vhp_sweep_tile_supported(int nx, int ny);
void vhp_tile_plane_geometry(int nx, int ny, int *wx, int *wy, int *sum_words_per_map);
cudaError_t vhp_launch_pack_tile(const uint8_t *d_occ, int nmaps, int nx, int ny, uint32_t *rowF,
                                 uint32_t *rowR, uint32_t *colF, uint32_t *colR, uint32_t *bsum,
                                 cudaStream_t st, int64_t *launches);
cudaError_t vhp_launch_sweep_tile(const VhpTilePlanes &pl, int nx, int ny, const int32_t *d_src_xy,
                                  const int32_t *d_src_map, int64_t npairs, vhp_dtype dtype,
                                  void *d_out, const double *d_rcp2, int *d_err, cudaStream_t st,
                                  int64_t *launches);

// K4 ray casting
cudaError_t vhp_launch_raycast(const uint8_t *d_occ, int nx, int ny,
                               const int32_t *d_src_xy, const int32_t *d_src_map,
                               int64_t npairs, vhp_dtype dtype, void *d_out,
                               int *d_err, cudaStream_t st, int64_t *launches);

// K2/K3/K5 planner: one persistent CTA per problem (sweep + fused epilogue +
// arg-min + next-source selection + path reconstruction, no host round trip).
// Working fields are fp64; vg32/vis32 are optional fp32 exports.
bool vhp_planner_supported(int nx, int ny);
cudaError_t vhp_launch_planner(const VhpPackedMaps &maps, int nx, int ny, const int32_t *d_se_xy,
                               const int32_t *d_prob_map, int64_t nprob, double threshold,
                               int32_t max_iter, int32_t ls_cap, const double *d_rcp,
                               double *d_vis, double *d_vg, int32_t *d_came, int32_t *d_status,
                               int32_t *d_nb, int32_t *d_ls, double *d_path_len, int32_t *d_path_n,
                               int32_t *d_path, float *d_vg32, float *d_vis32, int *d_err,
                               cudaStream_t st, int64_t *launches);

// helpers
static inline int vhp_words_padded(int n) { return (((n + 31) >> 5) + 1 + 3) & ~3; }

#endif
