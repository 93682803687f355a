// host_config.cpp -- settings.config boundary (host C++).
//
// Mirrors vbs::ConfigParser::parse (reference src/parser.cpp:12-338) and struct
// Config (include/parser/parser.h:11-37): same 25 keys, defaults, validation,
// diagnostics on stderr and the settings echo on stdout, so a config file written
// for the reference behaves the same here.  Table-driven instead of the
// reference's if/else chain.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>

#include "vhp.h"

namespace {

enum class Kind { Mode, NonNegative, AnyInt, Bool, Path, Pair, MaxIter, Unit01 };

struct KeySpec {
  const char *key;
  Kind kind;
  size_t offset; // into vhp_config
  bool is64;     // int64_t field (else int32_t)
};

#define OFF(f) offsetof(vhp_config, f)
const KeySpec kKeys[] = {
    {"mode", Kind::Mode, OFF(mode), false},
    {"ncols", Kind::NonNegative, OFF(ncols), true},
    {"nrows", Kind::NonNegative, OFF(nrows), true},
    {"nb_of_obstacles", Kind::AnyInt, OFF(nb_of_obstacles), true},
    {"minWidth", Kind::NonNegative, OFF(min_width), true},
    {"maxWidth", Kind::NonNegative, OFF(max_width), true},
    {"minHeight", Kind::NonNegative, OFF(min_height), true},
    {"maxHeight", Kind::NonNegative, OFF(max_height), true},
    {"randomSeed", Kind::Bool, OFF(random_seed), false},
    {"seedValue", Kind::AnyInt, OFF(seed_value), false},
    {"imagePath", Kind::Path, OFF(image_path), false},
    {"saveLocalVisibility", Kind::Bool, OFF(save_local_visibility), false},
    {"start", Kind::Pair, OFF(start_x), false},
    {"end", Kind::Pair, OFF(end_x), false},
    {"max_iter", Kind::MaxIter, OFF(max_iter), true},
    {"visibilityThreshold", Kind::Unit01, OFF(visibility_threshold), false},
    {"lightStrength", Kind::Unit01, OFF(light_strength), false},
    {"timer", Kind::Bool, OFF(timer), false},
    {"saveResults", Kind::Bool, OFF(save_results), false},
    {"saveCameFrom", Kind::Bool, OFF(save_came_from), false},
    {"saveLightSources", Kind::Bool, OFF(save_light_sources), false},
    {"saveGlobalVisibility", Kind::Bool, OFF(save_global_visibility), false},
    {"saveVisibilityField", Kind::Bool, OFF(save_visibility_field), false},
    {"silent", Kind::Bool, OFF(silent), false},
    {"ballRadius", Kind::AnyInt, OFF(ball_radius), false},
};
#undef OFF

void invalid(const std::string &key, const std::string &value, const char *hint) {
  std::cerr << "Invalid value for " << key << ": " << value << '\n';
  if (hint) std::cerr << hint;
}

void store_int(vhp_config *cfg, const KeySpec &k, long v) {
  char *base = reinterpret_cast<char *>(cfg) + k.offset;
  if (k.is64) *reinterpret_cast<int64_t *>(base) = v;
  else *reinterpret_cast<int32_t *>(base) = (int32_t)v;
}

// returns false where the reference's parse() returns false
bool apply(vhp_config *cfg, const KeySpec &k, const std::string &key, const std::string &value) {
  switch (k.kind) {
  case Kind::Mode:
    try {
      int m = std::stoi(value);
      if (m != 1 && m != 2) {
        std::cerr << "Invalid value for " << key << ": " << value << ", using default value 1\n";
        m = 1;
      }
      cfg->mode = m;
    } catch (...) {
      invalid(key, value, "It must be an integer 1 or 2 \n");
      return false;
    }
    return true;
  case Kind::NonNegative:
    try {
      const int v = std::stoi(value);
      if (v < 0) {
        invalid(key, value, "It must be a positive integer\n");
        return false;
      }
      store_int(cfg, k, v);
    } catch (...) {
      // the reference's minHeight branch only prints the first line and keeps parsing
      // (src/parser.cpp:117-127); every other key of this kind is fatal
      if (key == "minHeight") {
        invalid(key, value, nullptr);
        return true;
      }
      invalid(key, value, "It must be a positive integer\n");
      return false;
    }
    return true;
  case Kind::AnyInt:
    try {
      store_int(cfg, k, std::stoi(value));
    } catch (...) {
      invalid(key, value, "It must be an integer\n");
      return false;
    }
    return true;
  case Kind::Bool:
    if (value == "0" || value == "false") store_int(cfg, k, 0);
    else if (value == "1" || value == "true") store_int(cfg, k, 1);
    else {
      invalid(key, value, "It must be a boolean\n");
      return false;
    }
    return true;
  case Kind::Path:
    std::snprintf(cfg->image_path, sizeof cfg->image_path, "%s", value.c_str());
    return true;
  case Kind::Pair: {
    // parsePairString (src/parser.cpp:343-353): "{x,y}", (0,0) on a malformed string
    int32_t *xy = reinterpret_cast<int32_t *>(reinterpret_cast<char *>(cfg) + k.offset);
    int a = 0, b = 0;
    if (std::sscanf(value.c_str(), "{%d,%d}", &a, &b) != 2) {
      std::cerr << "Error: Invalid pair string: " << value << std::endl;
      a = b = 0;
    }
    xy[0] = a;
    xy[1] = b;
    return true;
  }
  case Kind::MaxIter:
    try {
      const int v = std::stoi(value);
      if (v < 0) {
        invalid(key, value, "It must be a positive integer\n");
        return false;
      }
      cfg->max_iter = v;
    } catch (...) {
      invalid(key, value, nullptr); // the reference keeps parsing here
    }
    return true;
  case Kind::Unit01:
    try {
      const double v = std::stod(value);
      if (k.offset == offsetof(vhp_config, visibility_threshold)) cfg->visibility_threshold = v;
      else cfg->light_strength = (float)v;
      const double chk = k.offset == offsetof(vhp_config, visibility_threshold)
                             ? cfg->visibility_threshold : (double)cfg->light_strength;
      if (chk > 1.0 || chk < 0.0) {
        invalid(key, value, "It must be a double between 0 and 1\n");
        return false;
      }
    } catch (...) {
      invalid(key, value, "It must be a positive double between 0 and 1\n");
      return false;
    }
    return true;
  }
  return true;
}

void echo(const vhp_config &c) {
  if (c.mode == 1) {
    std::cout << "Random environment mode" << std::endl;
    std::cout << "################### Environment settings ################## \n"
              << "nrows: " << c.nrows << "\n"
              << "ncols: " << c.ncols << "\n"
              << "Nb of obstacles: " << c.nb_of_obstacles << "\n"
              << "Min width: " << c.min_width << "\n"
              << "Max width: " << c.max_width << "\n"
              << "Min height: " << c.min_height << "\n"
              << "Max height: " << c.max_height << std::endl;
    if (c.random_seed) std::cout << "Random seed: " << (c.random_seed != 0) << std::endl;
    else std::cout << "Fixed seed value: " << c.seed_value << std::endl;
  } else if (c.mode == 2) {
    std::cout << "Import image mode" << "\n" << "Image path: " << c.image_path << std::endl;
  }
  std::cout << "#################### Solver settings ###################### \n"
            << "Start point: " << c.start_x << ", " << c.start_y << "\n"
            << "End point: " << c.end_x << ", " << c.end_y << "\n"
            << "Maximum iterations: " << c.max_iter << "\n"
            << "Solver visibility threshold: " << c.visibility_threshold << "\n"
            << "Light strength: " << c.light_strength << std::endl;
  // the reference labels saveCameFrom as "saveLightSourceEnum" etc. (:328-335)
  std::cout << "#################### Output settings ###################### \n"
            << "timer: " << (c.timer != 0) << "\n"
            << "saveLightSourceEnum: " << (c.save_came_from != 0) << "\n"
            << "saveLightSources: " << (c.save_light_sources != 0) << "\n"
            << "saveVisibilityField: " << (c.save_global_visibility != 0) << "\n"
            << "saveLocalVisibility: " << (c.save_local_visibility != 0) << "\n"
            << "saveVisibilityMapEnv: " << (c.save_visibility_field != 0) << std::endl;
}

} // namespace

extern "C" {

void vhp_config_default(vhp_config *cfg) {
  if (!cfg) return;
  std::memset(cfg, 0, sizeof *cfg);
  cfg->mode = 1;
  cfg->ncols = 100;
  cfg->nrows = 100;
  cfg->nb_of_obstacles = 10;
  cfg->min_width = 10;
  cfg->max_width = 20;
  cfg->min_height = 10;
  cfg->max_height = 20;
  cfg->random_seed = 1;
  cfg->seed_value = 0;
  std::snprintf(cfg->image_path, sizeof cfg->image_path, "%s", "C:\\...");
  cfg->max_iter = 100;
  cfg->visibility_threshold = 0.5;
  cfg->light_strength = 1.0f;
  cfg->timer = 1;
  cfg->save_results = 1;
  cfg->save_local_visibility = 1;
  cfg->save_came_from = 1;
  cfg->save_light_sources = 1;
  cfg->save_global_visibility = 1;
  cfg->save_visibility_field = 1;
  cfg->silent = 0;
  cfg->ball_radius = 5;
}

vhp_status vhp_config_parse(const char *filename, vhp_config *cfg) {
  if (!filename || !cfg) return VHP_ERR_INVALID_ARG;
  vhp_config_default(cfg);
  std::ifstream file(filename);
  if (!file) {
    std::cerr << "Failed to open " << filename << '\n';
    return VHP_ERR_IO;
  }
  std::string line;
  while (std::getline(file, line)) {
    if (line.empty() || line[0] == '#') continue; // comments start in column 0 only
    std::istringstream iss(line);
    std::string key, value;
    if (!std::getline(iss, key, '=')) continue;
    if (!std::getline(iss, value)) continue;
    auto trim = [](std::string &s) {
      s.erase(0, s.find_first_not_of(" \t"));
      s.erase(s.find_last_not_of(" \t") + 1);
    };
    trim(key);
    trim(value);
    const KeySpec *spec = nullptr;
    for (const KeySpec &k : kKeys)
      if (key == k.key) { spec = &k; break; }
    if (!spec) {
      std::cerr << "Invalid/irrelavent key: " << key << '\n';
      continue;
    }
    if (!apply(cfg, *spec, key, value)) return VHP_ERR_INVALID_ARG;
  }
  if (!cfg->silent) echo(*cfg);
  return VHP_OK;
}

} // extern "C"
