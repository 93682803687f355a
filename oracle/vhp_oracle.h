/*
 * vhp_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the reference's visibility / planner hot path
 * (IbrahimSquared/visibility-heuristic-path-planner, src/visibilityBasedSolver.cpp).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product (libvhp_b200.so) never
 * links, loads or calls anything in oracle/.
 *
 * Parity status: PINNED -- checked bit-for-bit against the unmodified reference
 * compiled from /root/reference (oracle/_ref, see oracle/Makefile and
 * oracle/ref_harness.cpp) and against the golden fixtures in tests/golden/
 * generated from that build by oracle/gen_golden.py.
 *
 * Build flags that define the floating-point contract: -O2 -ffp-contract=off
 * (strict IEEE-754 binary64, one rounding per operation, no FMA contraction).
 *
 * All 2-D fields use the reference's Field<T> layout: index = x + y*nx
 * (include/environment/field.h:26-29).  Occupancy "complement" convention:
 * 1.0 = free, 0.0 = occupied (src/environment.cpp:198-209).
 */
#ifndef VHP_ORACLE_H
#define VHP_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cameFrom_ sentinel: Field<size_t> filled with 1e15 (visibilityBasedSolver.cpp:46) */
#define VHP_ORACLE_NO_PARENT 1000000000000000ULL

/* status codes of solve(), in the order the reference tests them (:89-116,:134-139) */
enum {
  VHP_ORACLE_OK = 0,
  VHP_ORACLE_START_OOB = 1,
  VHP_ORACLE_END_OOB = 2,
  VHP_ORACLE_START_OCCUPIED = 3,
  VHP_ORACLE_END_OCCUPIED = 4,
  VHP_ORACLE_MAX_ITER = 5
};

/* computeVisibility(), visibilityBasedSolver.cpp:570-696.  `vis` is read-modify-
 * write exactly like the member visibility_ (no reset; cells the sweep never
 * visits keep their previous content). */
void vhp_oracle_compute_visibility(const double *occ, int nx, int ny, int sx,
                                   int sy, double *vis);

/* One planner sweep = resetQueue() + updateVisibility() + heap_->top(),
 * visibilityBasedSolver.cpp:65-71, 379-565, 130.
 *   vis      : local visibility, zeroed first (visibility_.reset(), :386)
 *   vg       : global visibility, max-merged (:417-418)
 *   came     : parents, first-writer-wins (:419-423)
 *   ls_xy    : light sources so far, (x,y) pairs, at least nb+1 entries valid
 *   nb       : nb_of_sources_ (label written into `came`)
 *   top_xyh  : out, heap top {x, y} and *top_h; returns number of pushes
 *              (0 => heap empty, top undefined in the reference). */
long vhp_oracle_update_visibility(const double *occ, int nx, int ny, int sx,
                                  int sy, int ex, int ey, double thr,
                                  double *vis, double *vg, uint64_t *came,
                                  const int *ls_xy, uint64_t nb, int *top_xy,
                                  double *top_h);

/* solve(), visibilityBasedSolver.cpp:76-160 (start/end already in the internal
 * frame, i.e. after the mode-2 flip of :83-86).  Buffers are (re)initialised
 * like reset() (:42-60).  ls_xy needs room for max_iter+2 points.
 * On OK, ls_xy[nb] = end (:141).  Returns a status code above. */
int vhp_oracle_solve(const double *occ, int nx, int ny, int sx, int sy, int ex,
                     int ey, double thr, long max_iter, double *vis, double *vg,
                     uint64_t *came, int *ls_xy, long *nb_of_sources);

/* reconstructPath(), visibilityBasedSolver.cpp:1183-1213.  Writes the path
 * start->end into path_xy (capacity path_cap points), returns the number of
 * points; *length = sum of eval_d segment lengths. */
long vhp_oracle_reconstruct_path(const uint64_t *came, int nx, const int *ls_xy,
                                 int ex, int ey, int *path_xy, long path_cap,
                                 double *length);

/* raycasting() for every target cell, visibilityBasedSolver.cpp:267-290 called
 * as in benchmark() :228-232 (i outer over x, j inner over y).  `ray` is
 * read-modify-write like visibilityRayCasting_ (initialised to 1.0 in reset()). */
void vhp_oracle_raycast_all(const double *occ, int nx, int ny, int sx, int sy,
                            double *ray);

/* generateNewEnvironmentFromSettings(), src/environment.cpp:40-88: glibc
 * srand(seed) + 4 rand() per obstacle.  occ is filled with 1.0 then rectangles
 * of 0.0. */
void vhp_oracle_generate_environment(double *occ, int nx, int ny,
                                     long nb_of_obstacles, long min_w,
                                     long max_w, long min_h, long max_h,
                                     int seed);

/* The same rectangle rule (src/environment.cpp:57-79) with the counter-based draws of the
 * device-side batch generator (include/vhp.h, vhp_environment_generate_batch_dev): draw d of
 * obstacle o of map `map` = top 31 bits of SplitMix64's finaliser over
 * seed + 0x9E3779B97F4A7C15 * (map * 0x100000001B3 + 4 * o + d + 1).  Not reference code: the
 * reference draws from one sequential rand() stream; this restates the generator this
 * repository defines for batches so that the CUDA kernel can be checked. */
unsigned vhp_oracle_env_draw(unsigned long long seed, unsigned long long map,
                             unsigned long long obstacle, unsigned d);
void vhp_oracle_generate_environment_counter(double *occ, int nx, int ny,
                                             long nb_of_obstacles, long min_w,
                                             long max_w, long min_h, long max_h,
                                             unsigned long long seed,
                                             unsigned long long map);

/* ---- variants of the sweep --------------------------------------
 *
 * getAccessibilityMap.m (MATLAB_code/visibility/getAccessibilityMap.m:1-118), the paper's
 * Algorithm 1: every cell of the grid is computed (no never-written border), the diagonal
 * branch `i == j*fac` takes the value of (i-1, j-1), `alpha` multiplies every cell (decay) and
 * `fac` bends the octant boundary (c = (j*fac)/i for i > j*fac, i/(j*fac) for j*fac > i).
 * 0-based restatement; `vis` must hold nx*ny doubles and is completely overwritten.
 * Parity status of THIS function: UNPINNED -- MATLAB / Octave cannot run in the build
 * container, so it is checked only against its own reading of the .m file (and, for
 * alpha = fac = 1, against the reference's C++ sweep away from the diagonal and the border). */
void vhp_oracle_accessibility_map(const double *occ, int nx, int ny, int sx, int sy,
                                  double alpha, double fac, double light_strength,
                                  double *vis);

/* computeVisibilityUsingQueue(), visibilityBasedSolver.cpp:701-893, as an ORDER-FREE rule:
 *   - the source holds lightStrength_ = 1 whatever its occupancy (:707);
 *   - a free cell is computed iff a cell that pushes it holds a value > cutoff (0.001 in the
 *     reference) -- its neighbour (i-1, j) unless that lies on the i == 0 axis, (i, j-1) unless
 *     that lies on the j == 0 axis, (i-1, j-1) for diagonal cells -- or it is one of the eight
 *     neighbours of the source (:709-716); every other cell keeps 0;
 *   - the value is the DP value of computeVisibility with the diagonal taken from (i-1, j-1)
 *     (:750-751) and no never-written border, upstream cells that were not computed reading 0.
 * The reference's function is a FIFO breadth-first search: a cell is computed when it is first
 * popped, from whatever its upstream neighbours hold AT THAT MOMENT.  Where an upstream
 * neighbour is itself reached only by a detour it may still be unvisited (0) then, so the
 * reference's output depends on the queue order; this rule is what the search computes when
 * every upstream cell is visited first, and equals it bit for bit on maps where the order does
 * not matter (tests/test_oracle_golden.py states how often, and pins the rule to the compiled
 * reference there).  `vis` is completely overwritten. */
void vhp_oracle_visibility_cutoff(const double *occ, int nx, int ny, int sx, int sy,
                                  double cutoff, double *vis);

/* eval_d(), include/solver/visibilityBasedSolver.h:112-115 */
double vhp_oracle_eval_d(int sx, int sy, int tx, int ty);

#ifdef __cplusplus
}
#endif
#endif
