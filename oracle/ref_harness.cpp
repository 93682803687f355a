// ref_harness.cpp -- white-box C entry points into the UNMODIFIED reference.
//
// TEST INFRASTRUCTURE (not product code).  oracle/Makefile compiles this file
// together with /root/reference/src/{visibilityBasedSolver,environment,parser}.cpp
// (sources read where they lie; nothing is copied into the repo) into
// oracle/_ref/libvhp_ref*.so.  The reference keeps all results in private
// members and only prints / writes files, so this harness re-declares `private`
// as `public` around the reference headers and calls the reference's own
// methods (computeVisibility, updateVisibility, solve, reconstructPath,
// raycasting, environment ctor) on state copied in/out through plain C arrays.
//
// Used for: (1) pinning oracle/vhp_oracle.c bit-for-bit, (2) generating
// tests/golden/*.npz (oracle/gen_golden.py), (3) the timed CPU baseline of
// bench.py (`cpu_baseline.kind = "reference"`, `--impl reference`).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <filesystem>
#include <iostream>
#include <limits>
#include <memory>
#include <queue>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include <unistd.h>

#include <SFML/Graphics.hpp>

#define private public
#include "solver/visibilityBasedSolver.h"
#undef private

namespace {

constexpr std::uint64_t kNoParent = 1000000000000000ULL;

vbs::Config makeConfig(int nx, int ny) {
  vbs::Config c;
  c.mode = 1;
  c.ncols = static_cast<std::size_t>(nx);
  c.nrows = static_cast<std::size_t>(ny);
  c.nb_of_obstacles = 0;
  c.randomSeed = false;
  c.seedValue = 1;
  c.timer = false;
  c.saveResults = false;
  c.saveLocalVisibility = false;
  c.saveCameFrom = false;
  c.saveLightSources = false;
  c.saveGlobalVisibility = false;
  c.saveVisibilityField = false;
  c.silent = true;
  return c;
}

void loadOccupancy(vbs::environment &env, const double *occ, int nx, int ny) {
  auto &field = env.getVisibilityField();
  for (int y = 0; y < ny; ++y)
    for (int x = 0; x < nx; ++x)
      field->set(x, y, occ[x + static_cast<std::size_t>(y) * nx]);
}

// solve() writes ./output (directory creation is unconditional in
// saveResults()); run it from a scratch directory.
struct ScratchCwd {
  std::filesystem::path old;
  ScratchCwd() {
    old = std::filesystem::current_path();
    auto dir = std::filesystem::temp_directory_path() /
               ("vhp_ref_" + std::to_string(::getpid()));
    std::filesystem::create_directories(dir);
    std::filesystem::current_path(dir);
  }
  ~ScratchCwd() { std::filesystem::current_path(old); }
};

// stateless sink, safe to share between the timing threads
struct NullBuf : std::streambuf {
  int overflow(int ch) override { return ch; }
  std::streamsize xsputn(const char *, std::streamsize n) override { return n; }
};
struct CoutMute {
  NullBuf nb;
  std::streambuf *old;
  CoutMute() : old(std::cout.rdbuf(&nb)) {}
  ~CoutMute() { std::cout.rdbuf(old); }
};

struct CoutCapture {
  std::ostringstream buf;
  std::streambuf *old;
  CoutCapture() : old(std::cout.rdbuf(buf.rdbuf())) {}
  ~CoutCapture() { std::cout.rdbuf(old); }
};

int statusFromMessage(const std::string &s) {
  if (s.find("Start point is out of bounds") != std::string::npos) return 1;
  if (s.find("End point is out of bounds") != std::string::npos) return 2;
  if (s.find("Start point is not valid") != std::string::npos) return 3;
  if (s.find("End point is not valid") != std::string::npos) return 4;
  if (s.find("Max iters hit") != std::string::npos) return 5;
  return 0;
}

} // namespace

extern "C" {

const char *ref_build_flags() {
#ifdef VHP_REF_FLAGS
  return VHP_REF_FLAGS;
#else
  return "unknown";
#endif
}

// computeVisibility() on a caller-provided (in/out) visibility_ field.
void ref_compute_visibility(const double *occ, int nx, int ny, int sx, int sy,
                            double *vis) {
  vbs::Config c = makeConfig(nx, ny);
  vbs::environment env(c);
  loadOccupancy(env, occ, nx, ny);
  vbs::visibilityBasedSolver s(env);
  const std::size_t n = static_cast<std::size_t>(nx) * ny;
  std::memcpy(s.visibility_.data_.get(), vis, n * sizeof(double));
  s.ls_ = {sx, sy};
  s.computeVisibility();
  std::memcpy(vis, s.visibility_.data_.get(), n * sizeof(double));
}

// computeVisibilityUsingQueue() (:701-893, the early-terminating BFS variant the README
// recommends for dense maps; its only call site :219 is commented out) on a caller-provided
// (in/out) visibility_ field.
void ref_compute_visibility_queue(const double *occ, int nx, int ny, int sx, int sy,
                                  double *vis) {
  vbs::Config c = makeConfig(nx, ny);
  vbs::environment env(c);
  loadOccupancy(env, occ, nx, ny);
  vbs::visibilityBasedSolver s(env);
  const std::size_t n = static_cast<std::size_t>(nx) * ny;
  std::memcpy(s.visibility_.data_.get(), vis, n * sizeof(double));
  s.ls_ = {sx, sy};
  s.computeVisibilityUsingQueue();
  std::memcpy(vis, s.visibility_.data_.get(), n * sizeof(double));
}

// resetQueue() + updateVisibility() + heap_->top() on caller-provided state.
long ref_update_visibility(const double *occ, int nx, int ny, int sx, int sy,
                           int ex, int ey, double thr, double *vis, double *vg,
                           std::uint64_t *came, const int *ls_xy,
                           std::uint64_t nb, int *top_xy, double *top_h) {
  vbs::Config c = makeConfig(nx, ny);
  c.visibilityThreshold = thr;
  vbs::environment env(c);
  loadOccupancy(env, occ, nx, ny);
  vbs::visibilityBasedSolver s(env);
  const std::size_t n = static_cast<std::size_t>(nx) * ny;
  std::memcpy(s.visibility_global_.data_.get(), vg, n * sizeof(double));
  for (std::size_t k = 0; k < n; ++k) s.cameFrom_.data_[k] = came[k];
  for (std::uint64_t k = 0; k <= nb; ++k)
    s.lightSources_[k] = {ls_xy[2 * k], ls_xy[2 * k + 1]};
  s.ls_ = {sx, sy};
  s.end_ = {ex, ey};
  s.nb_of_sources_ = nb;
  s.visibilityThreshold_ = thr;
  s.resetQueue();
  s.updateVisibility();
  std::memcpy(vis, s.visibility_.data_.get(), n * sizeof(double));
  std::memcpy(vg, s.visibility_global_.data_.get(), n * sizeof(double));
  for (std::size_t k = 0; k < n; ++k) came[k] = s.cameFrom_.data_[k];
  const long pushes = static_cast<long>(s.heap_->size());
  if (pushes > 0) {
    const vbs::Node top = s.heap_->top();
    top_xy[0] = static_cast<int>(top.x);
    top_xy[1] = static_cast<int>(top.y);
    *top_h = top.h;
  }
  return pushes;
}

// solve() + reconstructPath().  Returns the status implied by the reference's
// own stdout message.  ls_xy needs max_iter+2 points, path_xy path_cap points.
int ref_solve(const double *occ, int nx, int ny, int sx, int sy, int ex, int ey,
              double thr, long max_iter, double *vis, double *vg,
              std::uint64_t *came, int *ls_xy, long *nb_of_sources,
              int *path_xy, long path_cap, long *path_n, double *path_len,
              double *printed_len) {
  vbs::Config c = makeConfig(nx, ny);
  c.start = {sx, sy};
  c.end = {ex, ey};
  c.max_iter = static_cast<std::size_t>(max_iter);
  c.visibilityThreshold = thr;
  c.silent = false; // we want "Path length:"; everything is captured
  std::unique_ptr<vbs::environment> envp;
  {
    CoutMute mute; // "Generated new environment ..." banner
    envp = std::make_unique<vbs::environment>(c);
  }
  vbs::environment &env = *envp;
  loadOccupancy(env, occ, nx, ny);
  std::string out;
  std::unique_ptr<vbs::visibilityBasedSolver> sp;
  {
    ScratchCwd cwd;
    CoutCapture cap;
    sp = std::make_unique<vbs::visibilityBasedSolver>(env);
    sp->solve();
    out = cap.buf.str();
  }
  vbs::visibilityBasedSolver &s = *sp;
  const int status = statusFromMessage(out);
  const std::size_t n = static_cast<std::size_t>(nx) * ny;
  std::memcpy(vis, s.visibility_.data_.get(), n * sizeof(double));
  std::memcpy(vg, s.visibility_global_.data_.get(), n * sizeof(double));
  for (std::size_t k = 0; k < n; ++k) came[k] = s.cameFrom_.data_[k];
  *nb_of_sources = static_cast<long>(s.nb_of_sources_);
  if (status == 0 || status == 5) {
    const long cnt = std::min<long>(static_cast<long>(s.nb_of_sources_),
                                    max_iter + 1);
    for (long k = 0; k <= cnt; ++k) {
      ls_xy[2 * k] = s.lightSources_[k].first;
      ls_xy[2 * k + 1] = s.lightSources_[k].second;
    }
  }
  *path_n = 0;
  *path_len = 0;
  *printed_len = std::numeric_limits<double>::quiet_NaN();
  if (status == 0) {
    // The reference only prints the length (6 significant digits); take the
    // printed value as a cross-check and re-run its own reconstructPath +
    // eval_d for the full-precision number.
    const auto pos = out.find("Path length: ");
    if (pos != std::string::npos) *printed_len = std::atof(out.c_str() + pos + 13);
    std::vector<vbs::point> path;
    {
      ScratchCwd cwd;
      CoutCapture cap;
      s.reconstructPath(vbs::Node{static_cast<std::size_t>(ex),
                                  static_cast<std::size_t>(ey), 0},
                        path);
    }
    double total = 0;
    for (std::size_t k = 0; k + 1 < path.size(); ++k)
      total += s.eval_d(path[k].first, path[k].second, path[k + 1].first,
                        path[k + 1].second);
    *path_len = total;
    *path_n = static_cast<long>(path.size());
    for (long k = 0; k < std::min<long>(path_cap, *path_n); ++k) {
      path_xy[2 * k] = path[k].first;
      path_xy[2 * k + 1] = path[k].second;
    }
  }
  return status;
}

// The all-targets ray-casting loop of benchmark() (:228-232) on an in/out
// visibilityRayCasting_ field.
void ref_raycast_all(const double *occ, int nx, int ny, int sx, int sy,
                     double *ray) {
  vbs::Config c = makeConfig(nx, ny);
  vbs::environment env(c);
  loadOccupancy(env, occ, nx, ny);
  vbs::visibilityBasedSolver s(env);
  const std::size_t n = static_cast<std::size_t>(nx) * ny;
  std::memcpy(s.visibilityRayCasting_.data_.get(), ray, n * sizeof(double));
  for (std::size_t i = 0; i < s.nx_; ++i)
    for (std::size_t j = 0; j < s.ny_; ++j) s.raycasting(sx, sy, i, j);
  std::memcpy(ray, s.visibilityRayCasting_.data_.get(), n * sizeof(double));
}

// environment ctor in mode 1 with a fixed seed -> occupancy complement.
void ref_generate_environment(double *occ, int nx, int ny, long nb_of_obstacles,
                              long min_w, long max_w, long min_h, long max_h,
                              int seed) {
  vbs::Config c = makeConfig(nx, ny);
  c.nb_of_obstacles = static_cast<std::size_t>(nb_of_obstacles);
  c.minWidth = static_cast<std::size_t>(min_w);
  c.maxWidth = static_cast<std::size_t>(max_w);
  c.minHeight = static_cast<std::size_t>(min_h);
  c.maxHeight = static_cast<std::size_t>(max_h);
  c.seedValue = seed;
  vbs::environment env(c);
  const auto &field = env.getVisibilityField();
  for (int y = 0; y < ny; ++y)
    for (int x = 0; x < nx; ++x)
      occ[x + static_cast<std::size_t>(y) * nx] = field->get(x, y);
}

// ---- timed CPU baseline -----------------------------------------------------
// computeVisibility() for `nsrc` sources on one map, bracketed with
// high_resolution_clock like benchmark() (:217-222).  The reference is
// single-threaded; `nthreads` independent solver instances each take a
// contiguous slice of the sources.  Returns wall seconds for the whole batch
// (max over threads), setup excluded; *checksum defeats dead-code elimination.
double ref_time_compute_visibility(const double *occ, int nx, int ny,
                                   const int *src_xy, int nsrc, int nthreads,
                                   double *checksum) {
  if (nthreads < 1) nthreads = 1;
  vbs::Config c = makeConfig(nx, ny);
  std::vector<std::unique_ptr<vbs::environment>> envs;
  std::vector<std::unique_ptr<vbs::visibilityBasedSolver>> solvers;
  for (int t = 0; t < nthreads; ++t) {
    envs.emplace_back(std::make_unique<vbs::environment>(c));
    loadOccupancy(*envs.back(), occ, nx, ny);
    solvers.emplace_back(
        std::make_unique<vbs::visibilityBasedSolver>(*envs.back()));
  }
  std::vector<double> secs(nthreads, 0.0), sums(nthreads, 0.0);
  auto work = [&](int t) {
    const int lo = static_cast<int>(static_cast<long>(nsrc) * t / nthreads);
    const int hi = static_cast<int>(static_cast<long>(nsrc) * (t + 1) / nthreads);
    vbs::visibilityBasedSolver &s = *solvers[t];
    const auto t0 = std::chrono::high_resolution_clock::now();
    for (int k = lo; k < hi; ++k) {
      s.ls_ = {src_xy[2 * k], src_xy[2 * k + 1]};
      s.computeVisibility();
      sums[t] += s.visibility_(static_cast<std::size_t>(nx) - 1,
                               static_cast<std::size_t>(ny) - 1);
    }
    const auto t1 = std::chrono::high_resolution_clock::now();
    secs[t] = std::chrono::duration<double>(t1 - t0).count();
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < nthreads; ++t) pool.emplace_back(work, t);
  work(0);
  for (auto &th : pool) th.join();
  double total = 0;
  for (double v : sums) total += v;
  if (checksum) *checksum = total;
  return *std::max_element(secs.begin(), secs.end());
}

// solve() for `nprob` (start,end) problems on one map, same threading scheme.
// Returns wall seconds (max over threads); solver construction and reset()
// are outside the timed region, exactly what solve()'s own timer covers plus
// its saveResults()/reconstructPath() tail (all save flags off).
double ref_time_solve(const double *occ, int nx, int ny, const int *se_xy,
                      int nprob, double thr, long max_iter, int nthreads,
                      long *total_sources) {
  if (nthreads < 1) nthreads = 1;
  ScratchCwd cwd;
  CoutMute mute;
  std::vector<double> secs(nthreads, 0.0);
  std::vector<long> srcs(nthreads, 0);
  auto work = [&](int t) {
    const int lo = static_cast<int>(static_cast<long>(nprob) * t / nthreads);
    const int hi = static_cast<int>(static_cast<long>(nprob) * (t + 1) / nthreads);
    for (int k = lo; k < hi; ++k) {
      vbs::Config c = makeConfig(nx, ny);
      c.start = {se_xy[4 * k], se_xy[4 * k + 1]};
      c.end = {se_xy[4 * k + 2], se_xy[4 * k + 3]};
      c.max_iter = static_cast<std::size_t>(max_iter);
      c.visibilityThreshold = thr;
      vbs::environment env(c);
      loadOccupancy(env, occ, nx, ny);
      vbs::visibilityBasedSolver s(env);
      const auto t0 = std::chrono::high_resolution_clock::now();
      s.solve();
      const auto t1 = std::chrono::high_resolution_clock::now();
      secs[t] += std::chrono::duration<double>(t1 - t0).count();
      srcs[t] += static_cast<long>(s.nb_of_sources_);
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < nthreads; ++t) pool.emplace_back(work, t);
  work(0);
  for (auto &th : pool) th.join();
  long total = 0;
  for (long v : srcs) total += v;
  if (total_sources) *total_sources = total;
  return *std::max_element(secs.begin(), secs.end());
}

} // extern "C"
