// host_solver.cpp -- the drop-in for one vbs::visibilityBasedSolver instance (host
// C++ over the batched C-ABI entry points).
//
// Mirrors the reference's public surface (include/solver/visibilityBasedSolver.h:23-175):
//   solve()                 src/visibilityBasedSolver.cpp:76-160   -> vhp_solver_solve
//   standAloneVisibility()  :165-189                               -> vhp_solver_stand_alone_visibility
//   benchmark()             :194-262                               -> vhp_solver_benchmark
//   benchmarkSeries()       :295-374                               -> vhp_solver_benchmark_series
//   saveResults()           :1022-1178                             -> vhp_solver_save_results
// with the same stdout lines and the same ./output/*.txt formats, so interface.m
// style consumers keep working.  The compute runs on the GPU (K1/K2-K5/K4); this
// file only validates, prints and writes files.  The PNG renderings
//   saveStandAloneVisibility() :898-955, saveRayCastingVisibility() :960-1017,
//   saveImageWithPath()        :1218-1292
// are reproduced pixel for pixel (same base image, same drawing order) and written as 8-bit
// RGBA PNGs with zlib (the reference writes them through SFML).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include <cuda_runtime.h>
#include <zlib.h>

#include "vhp.h"

struct vhp_solver {
  vhp_context *ctx = nullptr;
  vhp_config cfg;
  int nx = 0, ny = 0;
  std::vector<uint8_t> occ;
  std::vector<double> vis, vg, ray;
  std::vector<int32_t> came;
  std::vector<int32_t> ls;   // lightSources_[0..nb]
  std::vector<int32_t> path; // reconstructed path, start -> end
  int64_t nb = 0;
  double path_len = 0;
  int ls_x = 0, ls_y = 0;    // ls_ (last stand-alone source)
};

namespace {

const char *kBanner =
    "############################## Solver output ##############################";

bool in_grid(const vhp_solver *s, int x, int y) {
  // isValid() compares as size_t (.h:101-103): negative coordinates fail
  return (size_t)x < (size_t)s->nx && (size_t)y < (size_t)s->ny;
}

bool ensure_dir(const std::string &dir);

// ---- PNG renderings --------------------------------------------------------------
struct Rgb { unsigned char r, g, b; };
constexpr Rgb kBlack{0, 0, 0}, kWhite{255, 255, 255}, kRed{255, 0, 0}, kGreen{0, 255, 0},
    kYellow{255, 255, 0}, kMagenta{255, 0, 255}, kCyan{0, 255, 255};

struct Canvas {
  int w = 0, h = 0;
  std::vector<Rgb> px;
  void set(int x, int y, Rgb c) { px[(size_t)y * w + x] = c; }
};

// the solver's environment image (constructor, :20-33): created black, then rows j = ny-1 .. 1
// white where free -- image row ny-1-j; the row of j == 0 keeps the creation colour
Canvas base_image(const vhp_solver *s) {
  Canvas im;
  im.w = s->nx; im.h = s->ny;
  im.px.assign((size_t)s->nx * s->ny, kBlack);
  for (int j = s->ny - 1; j > 0; --j)
    for (int i = 0; i < s->nx; ++i)
      im.set(i, s->ny - 1 - j, s->occ[(size_t)j * s->nx + i] < 1 ? kBlack : kWhite);
  return im;
}

void draw_ball(Canvas &im, const vhp_solver *s, int x0, int y0, int radius, Rgb c) {
  for (int j = -radius; j <= radius; ++j)
    for (int k = -radius; k <= radius; ++k)
      if (in_grid(s, x0 + j, y0 + k) && j * j + k * k <= radius * radius) im.set(x0 + j, y0 + k, c);
}

void png_chunk(std::ofstream &f, const char type[4], const std::vector<unsigned char> &data) {
  auto be32 = [&](uint32_t v) {
    const unsigned char b[4] = {(unsigned char)(v >> 24), (unsigned char)(v >> 16), (unsigned char)(v >> 8),
                                (unsigned char)v};
    f.write((const char *)b, 4);
  };
  be32((uint32_t)data.size());
  f.write(type, 4);
  if (!data.empty()) f.write((const char *)data.data(), (std::streamsize)data.size());
  uLong crc = crc32(0L, (const Bytef *)type, 4);
  if (!data.empty()) crc = crc32(crc, data.data(), (uInt)data.size());
  be32((uint32_t)crc);
}

bool write_png(const std::string &path, const Canvas &im) {
  std::vector<unsigned char> raw((size_t)im.h * (1 + 4 * (size_t)im.w));
  size_t o = 0;
  for (int y = 0; y < im.h; ++y) {
    raw[o++] = 0; // filter type None
    for (int x = 0; x < im.w; ++x) {
      const Rgb c = im.px[(size_t)y * im.w + x];
      raw[o++] = c.r; raw[o++] = c.g; raw[o++] = c.b; raw[o++] = 255;
    }
  }
  uLongf zlen = compressBound((uLong)raw.size());
  std::vector<unsigned char> z(zlen);
  if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return false;
  z.resize(zlen);
  std::ofstream f(path, std::ios::binary);
  if (!f) return false;
  const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  f.write((const char *)sig, 8);
  std::vector<unsigned char> ihdr(13);
  const uint32_t w = (uint32_t)im.w, h = (uint32_t)im.h;
  for (int b = 0; b < 4; ++b) { ihdr[b] = (unsigned char)(w >> (24 - 8 * b)); ihdr[4 + b] = (unsigned char)(h >> (24 - 8 * b)); }
  ihdr[8] = 8; ihdr[9] = 6; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0; // 8-bit RGBA, no interlace
  png_chunk(f, "IHDR", ihdr);
  png_chunk(f, "IDAT", z);
  png_chunk(f, "IEND", {});
  return (bool)f;
}

// saveStandAloneVisibility (:898-955) / saveRayCastingVisibility (:960-1017): grey levels of
// the field (rows j >= 1), the light source as a yellow ball with a black ring, occupied
// cells red
void save_visibility_image(const vhp_solver *s, const std::vector<double> &field, const char *name) {
  Canvas im = base_image(s);
  const int nx = s->nx, ny = s->ny;
  for (int j = ny - 1; j > 0; --j)
    for (int i = 0; i < nx; ++i) {
      const unsigned char v = (unsigned char)(255 * field[(size_t)j * nx + i]); // sf::Color(Uint8...)
      im.set(i, ny - 1 - j, Rgb{v, v, v});
    }
  const int r = s->cfg.ball_radius, x0 = s->ls_x, y0 = ny - 1 - s->ls_y;
  draw_ball(im, s, x0, y0, r, kYellow);
  for (int j = -r - 1; j <= r + 1; ++j)
    for (int k = -r - 1; k <= r + 1; ++k)
      if (in_grid(s, x0 + j, y0 + k) && j * j + k * k > r * r && j * j + k * k <= (r + 1) * (r + 1))
        im.set(x0 + j, y0 + k, kBlack);
  for (int j = ny - 1; j > 0; --j)
    for (int i = 0; i < nx; ++i)
      if (s->occ[(size_t)j * nx + i] == 0) im.set(i, ny - 1 - j, kRed);
  if (ensure_dir("output")) write_png(std::string("output/") + name, im);
}

// saveImageWithPath (:1218-1292): magenta Bresenham segments, a cyan ball per way point,
// the start green, the end red
void save_path_image(const vhp_solver *s) {
  const size_t n = s->path.size() / 2;
  if (n == 0) return;
  Canvas im = base_image(s);
  const int ny = s->ny;
  for (size_t i = 0; i + 1 < n; ++i) {
    int x0 = s->path[2 * i], y0 = ny - 1 - s->path[2 * i + 1];
    const int x1 = s->path[2 * i + 2], y1 = ny - 1 - s->path[2 * i + 3];
    const int dx = std::abs(x1 - x0), dy = std::abs(y1 - y0);
    const int sx = x0 < x1 ? 1 : -1, sy = y0 < y1 ? 1 : -1;
    int err = dx - dy;
    while (x0 != x1 || y0 != y1) {
      im.set(x0, y0, kMagenta);
      const int e2 = 2 * err;
      if (e2 > -dy) { err -= dy; x0 += sx; }
      if (e2 < dx) { err += dx; y0 += sy; }
    }
  }
  const int r = s->cfg.ball_radius;
  for (size_t i = 0; i < n; ++i) draw_ball(im, s, s->path[2 * i], ny - 1 - s->path[2 * i + 1], r, kCyan);
  draw_ball(im, s, s->path[0], ny - 1 - s->path[1], r, kGreen);
  draw_ball(im, s, s->path[2 * (n - 1)], ny - 1 - s->path[2 * (n - 1) + 1], r, kRed);
  if (ensure_dir("output")) write_png("output/ResultingPath.png", im);
}

template <typename T, typename F>
bool write_field(const std::string &path, const vhp_solver *s, F get) {
  std::fstream of(path, std::ios::out | std::ios::trunc);
  if (!of.is_open()) {
    std::cerr << "Failed to open output file " << path << std::endl;
    return false;
  }
  std::ostream &os = of;
  // mode 2 writes the bottom row first (:1043-1049)
  for (int r = 0; r < s->ny; ++r) {
    const int j = s->cfg.mode == 2 ? s->ny - 1 - r : r;
    for (int i = 0; i < s->nx; ++i) os << (T)get((size_t)j * s->nx + i) << " ";
    os << "\n";
  }
  return true;
}

bool ensure_dir(const std::string &dir) {
  namespace fs = std::filesystem;
  std::error_code ec;
  if (!fs::exists(dir, ec) && !fs::create_directories(dir, ec)) {
    std::cerr << "Failed to create directory " << dir << std::endl;
    return false;
  }
  return true;
}

// environment::saveEnvironment (src/environment.cpp:227-260): always bottom row first
void save_environment(const vhp_solver *s, const std::string &dir) {
  if (!ensure_dir(dir) || !s->cfg.save_visibility_field) return;
  const std::string path = dir + "/visibilityField.txt";
  std::fstream of(path, std::ios::out | std::ios::trunc);
  if (!of.is_open()) {
    std::cerr << "Failed to open output file " << path << std::endl;
    return;
  }
  for (int j = s->ny - 1; j >= 0; --j) {
    for (int i = 0; i < s->nx; ++i) of << (double)s->occ[(size_t)j * s->nx + i] << " ";
    of << "\n";
  }
  if (!s->cfg.silent) std::cout << "Saved visibility field" << std::endl;
}

double density(const vhp_solver *s) {
  size_t count = 0;
  for (uint8_t v : s->occ) count += v == 0;
  return (double)count / ((double)s->nx * s->ny) * 100;
}

// device-resident timing of one K1 sweep and one all-targets ray cast
vhp_status time_sweep_and_raycast(vhp_solver *s, const uint8_t *occ, int nx, int ny, int sx, int sy,
                                  double *vis_out, double *ray_out, long long *us_vis,
                                  long long *us_ray) {
  const size_t cells = (size_t)nx * ny;
  uint8_t *d_occ = nullptr;
  int32_t *d_src = nullptr;
  double *d_out = nullptr;
  // (the bit planes the context cached for d_occ must not outlive it: a later allocation may get
  // the same address)
  auto cleanup = [&]() { vhp_release_maps_dev(s->ctx); cudaFree(d_occ); cudaFree(d_src); cudaFree(d_out); };
  // allocate on the context's device, whatever the caller's current device is; restore it after
  int prev_dev = -1;
  cudaGetDevice(&prev_dev);
  if (cudaSetDevice(vhp_context_device(s->ctx)) != cudaSuccess) return VHP_ERR_CUDA;
  struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore_{prev_dev};
  if (cudaMalloc(&d_occ, cells) != cudaSuccess || cudaMalloc(&d_src, 8) != cudaSuccess ||
      cudaMalloc(&d_out, cells * 8) != cudaSuccess) {
    cleanup();
    return VHP_ERR_CUDA;
  }
  const int32_t src[2] = {sx, sy};
  cudaMemcpy(d_occ, occ, cells, cudaMemcpyHostToDevice);
  cudaMemcpy(d_src, src, 8, cudaMemcpyHostToDevice);
  vhp_status st = vhp_prepare_maps_dev(s->ctx, d_occ, 1, nx, ny);
  if (st == VHP_OK) st = vhp_context_synchronize(s->ctx);
  using clk = std::chrono::high_resolution_clock;
  if (st == VHP_OK) {
    const auto t0 = clk::now(); // one un-warmed call, like benchmark() :217-222
    st = vhp_visibility_batch_dev(s->ctx, d_occ, 1, nx, ny, d_src, nullptr, 1, VHP_F64, d_out);
    if (st == VHP_OK) st = vhp_context_synchronize(s->ctx);
    *us_vis = std::chrono::duration_cast<std::chrono::microseconds>(clk::now() - t0).count();
    if (st == VHP_OK && vis_out) cudaMemcpy(vis_out, d_out, cells * 8, cudaMemcpyDeviceToHost);
  }
  if (st == VHP_OK) {
    const auto t0 = clk::now();
    st = vhp_raycast_batch_dev(s->ctx, d_occ, 1, nx, ny, d_src, nullptr, 1, VHP_F64, d_out);
    if (st == VHP_OK) st = vhp_context_synchronize(s->ctx);
    *us_ray = std::chrono::duration_cast<std::chrono::microseconds>(clk::now() - t0).count();
    if (st == VHP_OK && ray_out) cudaMemcpy(ray_out, d_out, cells * 8, cudaMemcpyDeviceToHost);
  }
  cleanup();
  return st;
}

} // namespace

extern "C" {

vhp_status vhp_solver_create(vhp_context *ctx, const vhp_config *cfg, const uint8_t *occ, int nx,
                             int ny, vhp_solver **out) {
  if (!ctx || !cfg || !occ || !out || nx < 1 || ny < 1) return VHP_ERR_INVALID_ARG;
  vhp_solver *s = new vhp_solver();
  s->ctx = ctx;
  s->cfg = *cfg;
  s->nx = nx;
  s->ny = ny;
  const size_t cells = (size_t)nx * ny;
  s->occ.assign(occ, occ + cells);
  for (auto &v : s->occ) v = v ? 1 : 0;
  // reset(), :42-60
  s->vis.assign(cells, 0.0);
  s->vg.assign(cells, 0.0);
  s->ray.assign(cells, 1.0);
  s->came.assign(cells, VHP_NO_PARENT);
  // the reference's environment writes visibilityField.txt when it is built
  if (cfg->save_results == 1) save_environment(s, "./output"); // (2: written by an earlier instance)
  *out = s;
  return VHP_OK;
}

void vhp_solver_destroy(vhp_solver *s) { delete s; }

vhp_status vhp_solver_solve(vhp_solver *s) {
  if (!s) return VHP_ERR_INVALID_ARG;
  const auto t_start = std::chrono::high_resolution_clock::now();
  int sx = s->cfg.start_x, sy = s->cfg.start_y, ex = s->cfg.end_x, ey = s->cfg.end_y;
  if (s->cfg.mode == 2) { // config points use a bottom-left origin, :83-86
    sy = s->ny - 1 - sy;
    ey = s->ny - 1 - ey;
  }
  const int64_t max_iter = s->cfg.max_iter;
  const int32_t ls_cap = (int32_t)max_iter + 2;
  const int32_t se[4] = {sx, sy, ex, ey};
  int32_t status = 0, nb = 0, path_n = 0;
  s->ls.assign((size_t)ls_cap * 2, 0);
  s->path.assign((size_t)ls_cap * 2, 0);
  vhp_planner_out o = {};
  o.status = &status; o.nb_sources = &nb; o.light_sources = s->ls.data();
  o.path_len = &s->path_len; o.path_n = &path_n; o.path = s->path.data();
  o.vg = s->vg.data(); o.came = s->came.data(); o.vis = s->vis.data();
  const vhp_status rc = vhp_planner_batch(s->ctx, s->occ.data(), 1, s->nx, s->ny, se, nullptr, 1,
                                          s->cfg.visibility_threshold, (int32_t)max_iter, ls_cap,
                                          VHP_F64, &o);
  if (rc < 0) return rc;
  s->nb = nb;
  s->path.resize((size_t)path_n * 2);
  s->ls.resize((size_t)(nb + 1) * 2);
  switch (status) {
  case VHP_START_OOB: std::cout << kBanner << std::endl << "Start point is out of bounds." << std::endl; return VHP_START_OOB;
  case VHP_END_OOB: std::cout << kBanner << std::endl << "End point is out of bounds." << std::endl; return VHP_END_OOB;
  case VHP_START_OCCUPIED: std::cout << kBanner << std::endl << "Start point is not valid (occupied)" << std::endl; return VHP_START_OCCUPIED;
  case VHP_END_OCCUPIED: std::cout << kBanner << std::endl << "End point is not valid (occupied)" << std::endl; return VHP_END_OCCUPIED;
  case VHP_MAX_ITER:
    std::cout << "Max iters hit. Solution could not be found. Try lowering visibility threshold."
              << std::endl;
    return VHP_MAX_ITER; // the reference returns without saving (:134-139)
  default: break;
  }
  const auto duration = std::chrono::duration_cast<std::chrono::microseconds>(
      std::chrono::high_resolution_clock::now() - t_start);
  if (!s->cfg.silent && s->cfg.timer)
    std::cout << kBanner << "\n" << "Execution time in us: " << duration.count() << "us" << std::endl;
  vhp_solver_save_results(s, "./output");
  if (!s->cfg.silent) std::cout << "Path length: " << s->path_len << std::endl;
  if (s->cfg.save_results) save_path_image(s); // :1209-1211
  return VHP_OK;
}

vhp_status vhp_solver_stand_alone_visibility(vhp_solver *s) {
  if (!s) return VHP_ERR_INVALID_ARG;
  const int sx = s->cfg.start_x, sy = s->cfg.start_y; // no mode-2 flip here (:167)
  if (!in_grid(s, sx, sy)) {
    std::cout << kBanner << std::endl << "Start point is out of bounds." << std::endl;
    return VHP_START_OOB;
  }
  if (s->occ[(size_t)sy * s->nx + sx] == 0) {
    std::cout << kBanner << std::endl << "Start point is not valid (occupied)" << std::endl;
    return VHP_START_OCCUPIED;
  }
  s->ls_x = sx; s->ls_y = sy;
  // updateVisibility() from `start` (:187) == the first planner sweep: run the planner
  // with max_iter = 0 (one sweep, then the iteration limit stops it)
  const int32_t se[4] = {sx, sy, sx, sy};
  int32_t status = 0, nb = 0;
  int32_t ls[4] = {0, 0, 0, 0};
  vhp_planner_out o = {};
  o.status = &status; o.nb_sources = &nb; o.light_sources = ls;
  o.vg = s->vg.data(); o.came = s->came.data(); o.vis = s->vis.data();
  const vhp_status rc = vhp_planner_batch(s->ctx, s->occ.data(), 1, s->nx, s->ny, se, nullptr, 1,
                                          s->cfg.visibility_threshold, 0, 2, VHP_F64, &o);
  if (rc < 0) return rc;
  save_visibility_image(s, s->vis, "standAloneVisibility.png"); // :188
  return VHP_OK;
}

vhp_status vhp_solver_benchmark(vhp_solver *s) {
  if (!s) return VHP_ERR_INVALID_ARG;
  const int sx = s->cfg.start_x, sy = s->cfg.start_y;
  if (!in_grid(s, sx, sy)) {
    std::cout << kBanner << std::endl << "Start point is out of bounds." << std::endl;
    return VHP_START_OOB;
  }
  if (s->occ[(size_t)sy * s->nx + sx] == 0) {
    std::cout << kBanner << std::endl << "Start point is not valid (occupied)" << std::endl;
    return VHP_START_OCCUPIED;
  }
  s->ls_x = sx; s->ls_y = sy;
  long long us_vis = 0, us_ray = 0;
  const vhp_status rc = time_sweep_and_raycast(s, s->occ.data(), s->nx, s->ny, sx, sy, s->vis.data(),
                                               s->ray.data(), &us_vis, &us_ray);
  if (rc != VHP_OK) return rc;
  save_visibility_image(s, s->vis, "standAloneVisibility.png");  // :224
  save_visibility_image(s, s->ray, "rayCastingVisibility.png");  // :237
  if (!s->cfg.silent) {
    std::cout << kBanner << "\n" << "Visibility computation time in us: " << us_vis << "us" << std::endl;
    std::cout << "Raycasting computation time in us: " << us_ray << "us" << std::endl;
    std::cout << "Ratio. Proposed method is: " << (double)us_ray / us_vis
              << " faster than typical raycasting." << std::endl;
  }
  std::cout << "Density of the occupancy grid: " << density(s) << "%" << std::endl;
  return VHP_OK;
}

vhp_status vhp_solver_benchmark_series(vhp_solver *s, int n_sizes) {
  if (!s) return VHP_ERR_INVALID_ARG;
  const int num_points = 60; // :298-315
  const double start_value = 50, end_value = 5000;
  if (n_sizes <= 0 || n_sizes > num_points) n_sizes = num_points;
  std::vector<int> sizes;
  for (int i = 0; i < num_points; ++i)
    sizes.push_back((int)std::round(start_value * std::exp((std::log(end_value / start_value) * i) / (num_points - 1))));
  std::vector<double> tv, tr, ratios;
  for (int it = 0; it < n_sizes; ++it) {
    const int n = sizes[it];
    std::vector<uint8_t> occ((size_t)n * n, 1); // empty grid, source in the centre (:320-326)
    long long us_vis = 0, us_ray = 0;
    const vhp_status rc = time_sweep_and_raycast(s, occ.data(), n, n, n / 2, n / 2, nullptr, nullptr,
                                                 &us_vis, &us_ray);
    if (rc != VHP_OK) return rc;
    std::cout << "***************************" << std::endl;
    std::cout << "For grid size: " << n << "x" << n << std::endl;
    std::cout << "Visibility computation time in us: " << us_vis << "us" << std::endl;
    std::cout << "Raycasting computation time in us: " << us_ray << "us" << std::endl;
    std::cout << "Ratio. Proposed method is: " << (double)us_ray / us_vis
              << " faster than typical raycasting." << std::endl;
    tv.push_back((double)us_vis);
    tr.push_back((double)us_ray);
    ratios.push_back((double)us_ray / us_vis);
  }
  std::cout << kBanner << "\n" << "Ratios: " << std::endl;
  for (double r : ratios) std::cout << r << std::endl;
  ensure_dir("output");
  std::ofstream file("output/benchmark_results.txt", std::ios::app);
  for (size_t i = 0; i < ratios.size(); ++i)
    file << tv[i] << " " << tr[i] << " " << ratios[i] << " " << sizes[i] << "x" << sizes[i] << std::endl;
  return VHP_OK;
}

vhp_status vhp_solver_get_field(const vhp_solver *s, vhp_field which, void *dst) {
  if (!s || !dst) return VHP_ERR_INVALID_ARG;
  const size_t cells = (size_t)s->nx * s->ny;
  switch (which) {
  case VHP_FIELD_VISIBILITY: std::memcpy(dst, s->vis.data(), cells * 8); break;
  case VHP_FIELD_VISIBILITY_GLOBAL: std::memcpy(dst, s->vg.data(), cells * 8); break;
  case VHP_FIELD_CAME_FROM: std::memcpy(dst, s->came.data(), cells * 4); break;
  case VHP_FIELD_RAYCASTING: std::memcpy(dst, s->ray.data(), cells * 8); break;
  case VHP_FIELD_OCCUPANCY: std::memcpy(dst, s->occ.data(), cells); break;
  default: return VHP_ERR_INVALID_ARG;
  }
  return VHP_OK;
}

int64_t vhp_solver_nb_of_sources(const vhp_solver *s) { return s ? s->nb : 0; }

int64_t vhp_solver_light_sources(const vhp_solver *s, int32_t *xy, int64_t cap) {
  if (!s) return 0;
  const int64_t n = (int64_t)s->ls.size() / 2;
  if (xy) std::memcpy(xy, s->ls.data(), (size_t)std::min(n, cap) * 8);
  return n;
}

int64_t vhp_solver_path(const vhp_solver *s, int32_t *xy, int64_t cap, double *length) {
  if (!s) return 0;
  const int64_t n = (int64_t)s->path.size() / 2;
  if (xy) std::memcpy(xy, s->path.data(), (size_t)std::min(n, cap) * 8);
  if (length) *length = s->path_len;
  return n;
}

vhp_status vhp_solver_save_results(const vhp_solver *s, const char *dir_c) {
  if (!s || !dir_c) return VHP_ERR_INVALID_ARG;
  const std::string dir(dir_c);
  if (!ensure_dir(dir)) return VHP_ERR_IO;
  const vhp_config &c = s->cfg;
  if (c.save_came_from) {
    // Field<size_t> with 1e15 for "no parent" (:46, :1035-1063)
    if (!write_field<unsigned long long>(dir + "/cameFrom.txt", s, [&](size_t k) {
          return s->came[k] < 0 ? VHP_NO_PARENT_U64 : (unsigned long long)s->came[k];
        }))
      return VHP_ERR_IO;
    if (!c.silent) std::cout << "Saved cameFrom_" << std::endl;
  }
  if (c.save_light_sources) {
    const std::string path = dir + "/lightSources.txt";
    std::fstream of(path, std::ios::out | std::ios::trunc);
    if (!of.is_open()) {
      std::cerr << "Failed to open output file " << path << std::endl;
      return VHP_ERR_IO;
    }
    // only the nb swept sources are listed, y flipped back in mode 2 (:1073-1080)
    for (int64_t i = 0; i < s->nb; ++i) {
      const int x = s->ls[2 * i], y = s->ls[2 * i + 1];
      of << x << " " << (c.mode == 2 ? s->ny - 1 - y : y) << "\n";
    }
    if (!c.silent) std::cout << "Saved lightSources" << std::endl;
  }
  if (c.save_global_visibility) {
    if (!write_field<double>(dir + "/VisibilityMap.txt", s, [&](size_t k) { return s->vg[k]; }))
      return VHP_ERR_IO;
    if (!c.silent) std::cout << "Saved GlobalVisibility" << std::endl;
  }
  if (c.save_local_visibility) {
    if (!write_field<double>(dir + "/LocalVisibilityMap.txt", s, [&](size_t k) { return s->vis[k]; }))
      return VHP_ERR_IO;
    if (!c.silent) std::cout << "Saved LocalVisibility" << std::endl;
  }
  if (c.save_visibility_field) {
    if (!write_field<double>(dir + "/visibilityField.txt", s, [&](size_t k) { return (double)s->occ[k]; }))
      return VHP_ERR_IO;
    if (!c.silent) std::cout << "Saved OccupancyComplement" << std::endl;
  }
  return VHP_OK;
}

} // extern "C"
