// giant_internal.h -- shared between csrc/giant.cu (engine + C-ABI of the strip-partitioned
// planner) and csrc/kernels_giant.cu (its control kernels).
#ifndef VHP_GIANT_INTERNAL_H
#define VHP_GIANT_INTERNAL_H

#include <cuda_runtime.h>

#include <cstdint>

#include "vhp_internal.h"

// one row strip [y0, y1) of the fp64 working fields, device pointers (row y0 first)
struct GiantStripDev {
  int y0, y1;
  double *vis, *vg, *hc;
  int32_t *came;
};

constexpr int kGiantMaxLocal = 16; // strips one process may own
struct GiantLocal {
  int n;
  GiantStripDev s[kGiantMaxLocal];
};

constexpr int kGiantCtlInts = 12; // {done, sx, sy, status, nb, iterations, ex, ey, thr (double), max_iter, -}

cudaError_t vhp_launch_giant_reset(const GiantStripDev &s, int nx, cudaStream_t st, int64_t *launches);
cudaError_t vhp_launch_giant_begin(const VhpTilePlanes &pl, int nx, int ny, const int32_t se[4],
                                   double thr, int max_iter, int32_t *d_ls, const GiantLocal &loc,
                                   int *d_ctl, cudaStream_t st, int64_t *launches);
cudaError_t vhp_launch_giant_halo_gather(const int *d_ctl, int nx, int ny, const GiantStripDev &src,
                                         int cy0, int cy1, int qfirst, double *d_dst, cudaStream_t st,
                                         int64_t *launches);
cudaError_t vhp_launch_giant_pack_key(const int *d_ctl, const unsigned long long *d_bests,
                                      const GiantLocal &loc, int nx, unsigned long long *d_out,
                                      cudaStream_t st, int64_t *launches);
cudaError_t vhp_launch_giant_step(const unsigned long long *d_all, int world, int *d_ctl, int32_t *d_ls,
                                  unsigned long long cond_handle, int use_cond, cudaStream_t st,
                                  int64_t *launches);
cudaError_t vhp_launch_giant_came_at(const int *d_ctl, const int32_t *d_ls, int nx,
                                     const GiantLocal &loc, int cap, int32_t *d_came_at,
                                     cudaStream_t st, int64_t *launches);
cudaError_t vhp_launch_giant_finish(const int *d_ctl, const int32_t *d_came_at, int ls_cap,
                                    int32_t *d_ls, int32_t *d_status, int32_t *d_nb,
                                    double *d_path_len, int32_t *d_path_n, int32_t *d_path,
                                    int32_t *d_iters, cudaStream_t st, int64_t *launches);

#endif
