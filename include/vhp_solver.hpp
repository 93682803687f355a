// vhp_solver.hpp -- C++20 drop-in for the reference's classes, header-only over the
// C-ABI of vhp.h.  Same namespace, class and method names as
//   include/parser/parser.h              vbs::Config, vbs::ConfigParser
//   include/environment/field.h          vbs::Field<T>
//   include/environment/environment.h    vbs::environment
//   include/solver/visibilityBasedSolver.h  vbs::visibilityBasedSolver
// so the reference's main.cpp compiles against this header unchanged and runs the
// hot path on the GPU.  Results the reference keeps private stay reachable through
// the extra accessors at the bottom of visibilityBasedSolver.
#ifndef VHP_SOLVER_HPP
#define VHP_SOLVER_HPP

#include <cstring>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "vhp.h"

namespace vbs {

using size_t = std::size_t;
using point = std::pair<int, int>;

struct Config { // include/parser/parser.h:11-37
  int mode = 1;
  size_t ncols = 100, nrows = 100, nb_of_obstacles = 10;
  size_t minWidth = 10, maxWidth = 20, minHeight = 10, maxHeight = 20;
  bool randomSeed = true;
  int seedValue = 0;
  std::string imagePath = "C:\\...";
  point start, end;
  size_t max_iter = 100;
  double visibilityThreshold = 0.5;
  float lightStrength = 1;
  bool timer = true, saveResults = true, saveLocalVisibility = true, saveCameFrom = true,
       saveLightSources = true, saveGlobalVisibility = true, saveVisibilityField = true,
       silent = false;
  int ballRadius = 5;
};

inline vhp_config toC(const Config &c) {
  vhp_config o;
  vhp_config_default(&o);
  o.mode = c.mode; o.ncols = (int64_t)c.ncols; o.nrows = (int64_t)c.nrows;
  o.nb_of_obstacles = (int64_t)c.nb_of_obstacles;
  o.min_width = (int64_t)c.minWidth; o.max_width = (int64_t)c.maxWidth;
  o.min_height = (int64_t)c.minHeight; o.max_height = (int64_t)c.maxHeight;
  o.random_seed = c.randomSeed; o.seed_value = c.seedValue;
  std::snprintf(o.image_path, sizeof o.image_path, "%s", c.imagePath.c_str());
  o.start_x = c.start.first; o.start_y = c.start.second;
  o.end_x = c.end.first; o.end_y = c.end.second;
  o.max_iter = (int64_t)c.max_iter; o.visibility_threshold = c.visibilityThreshold;
  o.light_strength = c.lightStrength; o.timer = c.timer; o.save_results = c.saveResults;
  o.save_local_visibility = c.saveLocalVisibility; o.save_came_from = c.saveCameFrom;
  o.save_light_sources = c.saveLightSources; o.save_global_visibility = c.saveGlobalVisibility;
  o.save_visibility_field = c.saveVisibilityField; o.silent = c.silent;
  o.ball_radius = c.ballRadius;
  return o;
}

inline Config fromC(const vhp_config &o) {
  Config c;
  c.mode = o.mode; c.ncols = (size_t)o.ncols; c.nrows = (size_t)o.nrows;
  c.nb_of_obstacles = (size_t)o.nb_of_obstacles;
  c.minWidth = (size_t)o.min_width; c.maxWidth = (size_t)o.max_width;
  c.minHeight = (size_t)o.min_height; c.maxHeight = (size_t)o.max_height;
  c.randomSeed = o.random_seed != 0; c.seedValue = o.seed_value; c.imagePath = o.image_path;
  c.start = {o.start_x, o.start_y}; c.end = {o.end_x, o.end_y};
  c.max_iter = (size_t)o.max_iter; c.visibilityThreshold = o.visibility_threshold;
  c.lightStrength = o.light_strength; c.timer = o.timer != 0; c.saveResults = o.save_results != 0;
  c.saveLocalVisibility = o.save_local_visibility != 0; c.saveCameFrom = o.save_came_from != 0;
  c.saveLightSources = o.save_light_sources != 0;
  c.saveGlobalVisibility = o.save_global_visibility != 0;
  c.saveVisibilityField = o.save_visibility_field != 0; c.silent = o.silent != 0;
  c.ballRadius = o.ball_radius;
  return c;
}

class ConfigParser { // src/parser.cpp
public:
  bool parse(const std::string &filename) {
    vhp_config c;
    const bool ok = vhp_config_parse(filename.c_str(), &c) == VHP_OK;
    if (ok) config_ = fromC(c);
    return ok;
  }
  const Config &getConfig() const { return config_; }

private:
  Config config_;
};

template <typename T> class Field { // include/environment/field.h: index = x + y*nx
public:
  Field() = default;
  Field(size_t nx, size_t ny, T v) : nx_(nx), ny_(ny), data_(nx * ny, v) {}
  void set(size_t x, size_t y, T v) { data_[x + y * nx_] = v; }
  T get(size_t x, size_t y) const { return data_[x + y * nx_]; }
  T &operator()(size_t x, size_t y) { return data_[x + y * nx_]; }
  const T &operator()(size_t x, size_t y) const { return data_[x + y * nx_]; }
  size_t nx() const { return nx_; }
  size_t ny() const { return ny_; }
  size_t size() const { return data_.size(); }
  void reset() { data_.assign(data_.size(), T{}); }
  void resize(size_t nx, size_t ny, T v) { nx_ = nx; ny_ = ny; data_.assign(nx * ny, v); }
  T *data() { return data_.data(); }
  const T *data() const { return data_.data(); }

private:
  size_t nx_ = 0, ny_ = 0;
  std::vector<T> data_;
};

class environment { // src/environment.cpp:15-35
public:
  explicit environment(Config &config) : sharedConfig_(std::make_shared<Config>(config)) {
    vhp_config c = toC(config);
    int nx = (int)config.ncols, ny = (int)config.nrows;
    std::vector<uint8_t> occ;
    if (config.mode == 2) {
      if (vhp_environment_load_image(c.image_path, nullptr, &nx, &ny) == VHP_OK) {
        occ.resize((size_t)nx * ny);
        vhp_environment_load_image(c.image_path, occ.data(), &nx, &ny);
      } else {
        nx = ny = 0;
      }
    } else {
      occ.resize((size_t)nx * ny);
      vhp_environment_generate(&c, occ.data(), nullptr);
    }
    sharedVisibilityField_ = std::make_shared<Field<double>>((size_t)nx, (size_t)ny, 1.0);
    for (size_t k = 0; k < occ.size(); ++k) sharedVisibilityField_->data()[k] = occ[k] ? 1.0 : 0.0;
  }
  const std::shared_ptr<Field<double>> &getVisibilityField() const { return sharedVisibilityField_; }
  const std::shared_ptr<Config> &getConfig() const { return sharedConfig_; }

private:
  std::shared_ptr<Field<double>> sharedVisibilityField_;
  std::shared_ptr<Config> sharedConfig_;
};

class visibilityBasedSolver { // include/solver/visibilityBasedSolver.h:23-175
public:
  explicit visibilityBasedSolver(environment &env, int device = 0)
      : occupancyComplement_(env.getVisibilityField()), sharedConfig_(env.getConfig()) {
    if (vhp_context_create(device, nullptr, &ctx_) != VHP_OK)
      throw std::runtime_error(std::string("vhp: ") + vhp_last_error(nullptr));
    rebuild();
  }
  ~visibilityBasedSolver() {
    if (solver_) vhp_solver_destroy(solver_);
    if (ctx_) vhp_context_destroy(ctx_);
  }
  visibilityBasedSolver(const visibilityBasedSolver &) = delete;
  visibilityBasedSolver &operator=(const visibilityBasedSolver &) = delete;

  int getGlobalIter() const { return 0; } // always 0 in the reference (.h:151)
  void solve() { rebuild(); last_ = vhp_solver_solve(solver_); }
  void standAloneVisibility() { rebuild(); last_ = vhp_solver_stand_alone_visibility(solver_); }
  void benchmark() { last_ = vhp_solver_benchmark(solver_); }
  void benchmarkSeries() { last_ = vhp_solver_benchmark_series(solver_, 0); }

  // ---- beyond the reference: results without going through ./output ----------
  vhp_status lastStatus() const { return last_; }
  size_t nbOfSources() const { return (size_t)vhp_solver_nb_of_sources(solver_); }
  std::vector<point> lightSources() const { return points(&vhp_solver_light_sources); }
  std::vector<point> path(double *length = nullptr) const {
    const int64_t n = vhp_solver_path(solver_, nullptr, 0, length);
    std::vector<int32_t> xy((size_t)n * 2);
    vhp_solver_path(solver_, xy.data(), n, length);
    std::vector<point> out;
    for (int64_t k = 0; k < n; ++k) out.emplace_back(xy[2 * k], xy[2 * k + 1]);
    return out;
  }
  // every field as doubles, like the reference's members: cameFrom_ holds the parent index or
  // 1e15 (Field<size_t> filled with 1e15, :46), occupancyComplement_ 1.0 / 0.0
  Field<double> field(vhp_field which) const {
    const size_t nx = occupancyComplement_->nx(), ny = occupancyComplement_->ny();
    Field<double> f(nx, ny, 0.0);
    if (which == VHP_FIELD_CAME_FROM) {
      std::vector<int32_t> tmp(nx * ny);
      vhp_solver_get_field(solver_, which, tmp.data());
      for (size_t k = 0; k < tmp.size(); ++k)
        f.data()[k] = tmp[k] < 0 ? (double)VHP_NO_PARENT_U64 : (double)tmp[k];
    } else if (which == VHP_FIELD_OCCUPANCY) {
      std::vector<uint8_t> tmp(nx * ny);
      vhp_solver_get_field(solver_, which, tmp.data());
      for (size_t k = 0; k < tmp.size(); ++k) f.data()[k] = tmp[k] ? 1.0 : 0.0;
    } else {
      vhp_solver_get_field(solver_, which, f.data());
    }
    return f;
  }

private:
  // the reference shares the occupancy field with the environment, so edits made
  // through environment::getVisibilityField() after construction are picked up
  void rebuild() {
    if (solver_) vhp_solver_destroy(solver_);
    solver_ = nullptr;
    const size_t nx = occupancyComplement_->nx(), ny = occupancyComplement_->ny();
    std::vector<uint8_t> occ(nx * ny);
    for (size_t k = 0; k < occ.size(); ++k) occ[k] = occupancyComplement_->data()[k] != 0.0;
    vhp_config c = toC(*sharedConfig_);
    // visibilityField.txt is written once, when the environment is built: 2 = "save results,
    // the environment file exists already" (vhp.h)
    if (built_ && c.save_results) c.save_results = 2;
    if (vhp_solver_create(ctx_, &c, occ.data(), (int)nx, (int)ny, &solver_) != VHP_OK)
      throw std::runtime_error("vhp_solver_create failed");
    built_ = true;
  }
  template <typename F> std::vector<point> points(F fn) const {
    const int64_t n = fn(solver_, nullptr, 0);
    std::vector<int32_t> xy((size_t)n * 2);
    fn(solver_, xy.data(), n);
    std::vector<point> out;
    for (int64_t k = 0; k < n; ++k) out.emplace_back(xy[2 * k], xy[2 * k + 1]);
    return out;
  }
  std::shared_ptr<Field<double>> occupancyComplement_;
  std::shared_ptr<Config> sharedConfig_;
  vhp_context *ctx_ = nullptr;
  vhp_solver *solver_ = nullptr;
  vhp_status last_ = VHP_OK;
  bool built_ = false;
};

} // namespace vbs
#endif
