// cli_main.cpp -- drop-in `visibility_heuristic_planner` executable.
// Same flow as the reference's main() (src/main.cpp:6-29): parse
// config/settings.config relative to the working directory, build the
// environment, solve(), benchmark().  Written against the drop-in classes of
// vhp_solver.hpp, which run the hot path on the GPU through libvhp_b200.so.
#include <iostream>

#include "vhp_solver.hpp"

int main(int argc, char **argv) {
  const char *config_path = argc > 1 ? argv[1] : "config/settings.config";
  vbs::ConfigParser parser;
  if (!parser.parse(config_path)) {
    std::cout << "################## Parsing results: ##################### \n";
    std::cout << "Error parsing config file" << std::endl;
    return 1;
  }
  std::cout << "################## Parsing results: ##################### \n";
  std::cout << "Config file parsed successfully \n" << std::endl;
  auto config = parser.getConfig();
  try {
    vbs::environment env(config);
    vbs::visibilityBasedSolver solver(env);
    solver.solve();
    solver.benchmark();
  } catch (const std::exception &e) {
    std::cerr << e.what() << std::endl;
    return 2;
  }
  return 0;
}
