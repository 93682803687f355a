// kernels_env.cu -- random-rectangle environments generated on the device, a batch of maps
// at a time (the maps of BASELINE configs[3] never touch the host).
//
// Rectangle rule of environment::generateNewEnvironmentFromSettings
// (reference src/environment.cpp:57-79), per obstacle o of map m:
//   col_1 = 1 + r0 % (nx + 1);  col_2 = col_1 + minWidth  + r1 % (maxWidth  - minWidth  + 1)
//   row_1 = 1 + r2 % (ny + 1);  row_2 = row_1 + minHeight + r3 % (maxHeight - minHeight + 1)
//   all four clamped to n - 1; cells [col_1, col_2) x [row_1, row_2) become occupied.
// The reference draws r0..r3 from one sequential glibc rand() stream per map
// (vhp_environment_generate reproduces that on the host).  A batch needs independent draws,
// so here r_d = vhp_env_draw(seed, m, o, d): a counter-based generator (SplitMix64 finaliser of
// a counter built from the key), reduced to 31 bits like rand().  The parity test checks it
// against a CPU restatement of the same definition.
#include <cstdint>

#include "vhp_internal.h"

namespace {

__host__ __device__ inline uint32_t env_draw(uint64_t seed, uint64_t map, uint64_t obstacle, uint32_t d) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (map * 0x100000001B3ull + obstacle * 4ull + d + 1ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (uint32_t)(z >> 33); // 31 bits, like rand()
}

// grid (obstacles, maps); the batch was set to 1 (free) before
__global__ void env_rectangles_kernel(uint8_t *occ, int nx, int ny, long long first_map,
                                      unsigned long long seed, long long min_w, long long max_w,
                                      long long min_h, long long max_h) {
  const unsigned long long m = (unsigned long long)first_map + blockIdx.y, o = blockIdx.x;
  long long col_1 = 1 + (long long)(env_draw(seed, m, o, 0) % (unsigned long long)(nx + 1));
  long long col_2 = col_1 + min_w + (long long)(env_draw(seed, m, o, 1) % (unsigned long long)(max_w - min_w + 1));
  long long row_1 = 1 + (long long)(env_draw(seed, m, o, 2) % (unsigned long long)(ny + 1));
  long long row_2 = row_1 + min_h + (long long)(env_draw(seed, m, o, 3) % (unsigned long long)(max_h - min_h + 1));
  col_1 = min(col_1, (long long)nx - 1); col_2 = min(col_2, (long long)nx - 1);
  row_1 = min(row_1, (long long)ny - 1); row_2 = min(row_2, (long long)ny - 1);
  const int w = (int)(col_2 - col_1), h = (int)(row_2 - row_1);
  if (w <= 0 || h <= 0) return;
  uint8_t *base = occ + (size_t)blockIdx.y * nx * ny + (size_t)row_1 * nx + col_1;
  for (int c = threadIdx.x; c < w * h; c += blockDim.x) base[(size_t)(c / w) * nx + (c % w)] = 0;
}

} // namespace

uint32_t vhp_env_draw(uint64_t seed, uint64_t map, uint64_t obstacle, uint32_t d) {
  return env_draw(seed, map, obstacle, d);
}

cudaError_t vhp_launch_env_generate(uint8_t *d_occ, int nmaps, int nx, int ny, int64_t first_map,
                                    uint64_t seed, int64_t nb_of_obstacles, int64_t min_w,
                                    int64_t max_w, int64_t min_h, int64_t max_h, cudaStream_t st,
                                    int64_t *launches) {
  cudaError_t e = cudaMemsetAsync(d_occ, 1, (size_t)nmaps * nx * ny, st);
  if (e != cudaSuccess) return e;
  for (int m0 = 0; m0 < nmaps && nb_of_obstacles > 0; m0 += 65535) { // gridDim.y limit
    const int nm = nmaps - m0 < 65535 ? nmaps - m0 : 65535;
    env_rectangles_kernel<<<dim3((unsigned)nb_of_obstacles, (unsigned)nm), 256, 0, st>>>(
        d_occ + (size_t)m0 * nx * ny, nx, ny, (long long)(first_map + m0), (unsigned long long)seed,
        (long long)min_w, (long long)max_w, (long long)min_h, (long long)max_h);
    if (launches) *launches += 1;
  }
  return cudaGetLastError();
}
