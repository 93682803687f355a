/*
 * vhp_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 * See vhp_oracle.h for the contract.  Compile with:
 *     gcc -std=c11 -O2 -ffp-contract=off -fPIC -shared
 *
 * The reference spells the four quadrants out as four copies of the same loop
 * nest (src/visibilityBasedSolver.cpp:388-564 and :575-695).  This restatement
 * folds them into one routine driven by a sign pair (dirx, diry) and an extent
 * pair, keeping the reference's iteration order (quadrants Q1,Q2,Q3,Q4; i outer,
 * j inner), its operation order and its quirks:
 *   - no i==j branch: `v` keeps the value of the previous inner iteration
 *     (:396-414 has no final else), i.e. vis(k,k) = vis(k,k-1)*occ(k,k);
 *   - Q2/Q3 stop at X=1, Q3/Q4 stop at Y=1 (loop bounds :434-438,:478-483,:522-527);
 *   - heap top == first pushed element attaining the minimum h (Node::operator<
 *     in include/solver/visibilityBasedSolver.h:16-21 is strict, so
 *     std::push_heap never lifts an element over an equal one).
 */
#include "vhp_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define IDX(x, y) ((size_t)(x) + (size_t)(y) * (size_t)nx)

double vhp_oracle_eval_d(int sx, int sy, int tx, int ty) {
  /* include/solver/visibilityBasedSolver.h:112-115: dx squared in double, dy
   * squared in int then converted. */
  return sqrt((double)(sx - tx) * (sx - tx) + (sy - ty) * (sy - ty));
}

/* per-cell planner epilogue state (updateVisibility only) */
typedef struct {
  int enabled;
  double thr, scale;
  double *vg;
  uint64_t *came;
  const int *ls_xy;
  uint64_t nb;
  int ex, ey;
  /* running arg-min in push order == heap top */
  long pushes;
  int top_x, top_y;
  double top_h;
} epilogue_t;

/* One quadrant.  Local cell (i,j) maps to (X,Y) = (sx + dirx*i, sy + diry*j);
 * its upstream neighbours are one step back towards the source. */
static void sweep_quadrant(const double *occ, int nx, double *vis, int sx,
                           int sy, int dirx, int diry, size_t max_x,
                           size_t max_y, double *v_carry, epilogue_t *ep) {
  double v = *v_carry; /* `v` is function-scope in the reference (:381,:572) */
  for (size_t i = 0; i < max_x; ++i) {
    const int X = sx + dirx * (int)i;
    for (size_t j = 0; j < max_y; ++j) {
      const int Y = sy + diry * (int)j;
      if (i == 0 && j == 0) {
        v = 1.0; /* lightStrength_ (.h:144) */
      } else if (i == 0) {
        v = vis[IDX(X, Y - diry)];
      } else if (j == 0) {
        v = vis[IDX(X - dirx, Y)];
      } else if (i > j) {
        const double c = (double)((double)j + 0.0) / ((double)i + 0.0);
        const double a = vis[IDX(X - dirx, Y)];
        const double b = vis[IDX(X - dirx, Y - diry)];
        v = a - c * (a - b);
      } else if (j > i) {
        const double c = (double)((double)i + 0.0) / ((double)j + 0.0);
        const double a = vis[IDX(X, Y - diry)];
        const double b = vis[IDX(X - dirx, Y - diry)];
        v = a - c * (a - b);
      } /* i == j > 0: v is stale on purpose */
      v = v * occ[IDX(X, Y)];
      vis[IDX(X, Y)] = v;
      if (ep && ep->enabled) {
        double *g = &ep->vg[IDX(X, Y)];
        *g = (v > *g) ? v : *g; /* std::max(v, vg): returns vg when equal */
        if (v >= ep->thr) {
          if (ep->came[IDX(X, Y)] == VHP_ORACLE_NO_PARENT)
            ep->came[IDX(X, Y)] = ep->nb;
        }
        if (*g >= ep->thr) {
          const uint64_t p = ep->came[IDX(X, Y)];
          const int px = ep->ls_xy[2 * p], py = ep->ls_xy[2 * p + 1];
          const double h =
              (ep->scale * *g) + (vhp_oracle_eval_d(X, Y, ep->ex, ep->ey) +
                                  vhp_oracle_eval_d(X, Y, px, py));
          if (ep->pushes == 0 || h < ep->top_h) {
            ep->top_h = h;
            ep->top_x = X;
            ep->top_y = Y;
          }
          ++ep->pushes;
        }
      }
    }
  }
  *v_carry = v;
}

static void sweep_all(const double *occ, int nx, int ny, int sx, int sy,
                      double *vis, epilogue_t *ep) {
  double v = 0.0;
  /* Q1 (+,+)  :388-432 / :575-605 */
  sweep_quadrant(occ, nx, vis, sx, sy, +1, +1, (size_t)(nx - sx),
                 (size_t)(ny - sy), &v, ep);
  /* Q2 (-,+)  :433-476 / :606-635 */
  sweep_quadrant(occ, nx, vis, sx, sy, -1, +1, (size_t)sx, (size_t)(ny - sy),
                 &v, ep);
  /* Q3 (-,-)  :477-520 / :636-665 */
  sweep_quadrant(occ, nx, vis, sx, sy, -1, -1, (size_t)sx, (size_t)sy, &v, ep);
  /* Q4 (+,-)  :521-564 / :666-695 */
  sweep_quadrant(occ, nx, vis, sx, sy, +1, -1, (size_t)(nx - sx), (size_t)sy,
                 &v, ep);
}

void vhp_oracle_compute_visibility(const double *occ, int nx, int ny, int sx,
                                   int sy, double *vis) {
  sweep_all(occ, nx, ny, sx, sy, vis, NULL);
}

long vhp_oracle_update_visibility(const double *occ, int nx, int ny, int sx,
                                  int sy, int ex, int ey, double thr,
                                  double *vis, double *vg, uint64_t *came,
                                  const int *ls_xy, uint64_t nb, int *top_xy,
                                  double *top_h) {
  epilogue_t ep;
  memset(&ep, 0, sizeof ep);
  ep.enabled = 1;
  ep.thr = thr;
  /* scale_ = sqrt(ny_*ny_ + nx_*nx_) on size_t (:49) */
  ep.scale = sqrt((double)((size_t)ny * (size_t)ny + (size_t)nx * (size_t)nx));
  ep.vg = vg;
  ep.came = came;
  ep.ls_xy = ls_xy;
  ep.nb = nb;
  ep.ex = ex;
  ep.ey = ey;
  memset(vis, 0, sizeof(double) * (size_t)nx * (size_t)ny); /* :386 */
  sweep_all(occ, nx, ny, sx, sy, vis, &ep);
  if (top_xy) {
    top_xy[0] = ep.top_x;
    top_xy[1] = ep.top_y;
  }
  if (top_h) *top_h = ep.top_h;
  return ep.pushes;
}

int vhp_oracle_solve(const double *occ, int nx, int ny, int sx, int sy, int ex,
                     int ey, double thr, long max_iter, double *vis, double *vg,
                     uint64_t *came, int *ls_xy, long *nb_of_sources) {
  const size_t n = (size_t)nx * (size_t)ny;
  /* reset(), :42-60 */
  for (size_t k = 0; k < n; ++k) {
    vg[k] = 0.0;
    vis[k] = 0.0;
    came[k] = VHP_ORACLE_NO_PARENT;
  }
  long nb = 0;
  *nb_of_sources = 0;
  /* isValid() compares as size_t (.h:101-103): negatives wrap and fail */
  if (!((size_t)sx < (size_t)nx && (size_t)sy < (size_t)ny))
    return VHP_ORACLE_START_OOB;
  if (!((size_t)ex < (size_t)nx && (size_t)ey < (size_t)ny))
    return VHP_ORACLE_END_OOB;
  if (occ[IDX(sx, sy)] == 0) return VHP_ORACLE_START_OCCUPIED;
  if (occ[IDX(ex, ey)] == 0) return VHP_ORACLE_END_OCCUPIED;

  int lx = sx, ly = sy;
  ls_xy[0] = sx;
  ls_xy[1] = sy;
  came[IDX(sx, sy)] = 0;
  vg[IDX(ex, ey)] = 0;
  while (vg[IDX(ex, ey)] <= thr) { /* :127 */
    int top[2] = {0, 0};
    double h = 0;
    vhp_oracle_update_visibility(occ, nx, ny, lx, ly, ex, ey, thr, vis, vg,
                                 came, ls_xy, (uint64_t)nb, top, &h);
    lx = top[0];
    ly = top[1];
    ++nb;
    ls_xy[2 * nb] = lx;
    ls_xy[2 * nb + 1] = ly;
    *nb_of_sources = nb;
    if (nb > max_iter) return VHP_ORACLE_MAX_ITER; /* :134-139 */
  }
  ls_xy[2 * nb] = ex; /* :141 */
  ls_xy[2 * nb + 1] = ey;
  *nb_of_sources = nb;
  return VHP_ORACLE_OK;
}

long vhp_oracle_reconstruct_path(const uint64_t *came, int nx, const int *ls_xy,
                                 int ex, int ey, int *path_xy, long path_cap,
                                 double *length) {
  /* :1183-1213.  The walk runs end->start, then the list is reversed. */
  long n = 0;
  int x = ex, y = ey;
  double t = (double)came[IDX(x, y)];
  double t_old = 1.7976931348623157e308; /* numeric_limits<double>::max() */
  while (t != t_old) {
    if (n < path_cap) {
      path_xy[2 * n] = x;
      path_xy[2 * n + 1] = y;
    }
    ++n;
    t_old = t;
    const uint64_t ti = (uint64_t)t;
    x = ls_xy[2 * ti];
    y = ls_xy[2 * ti + 1];
    t = (double)came[IDX(x, y)];
  }
  if (n < path_cap) {
    path_xy[2 * n] = x;
    path_xy[2 * n + 1] = y;
  }
  ++n;
  const long m = n < path_cap ? n : path_cap;
  for (long a = 0, b = m - 1; a < b; ++a, --b) {
    int tx = path_xy[2 * a], ty = path_xy[2 * a + 1];
    path_xy[2 * a] = path_xy[2 * b];
    path_xy[2 * a + 1] = path_xy[2 * b + 1];
    path_xy[2 * b] = tx;
    path_xy[2 * b + 1] = ty;
  }
  double total = 0;
  for (long k = 0; k + 1 < m; ++k)
    total += vhp_oracle_eval_d(path_xy[2 * k], path_xy[2 * k + 1],
                               path_xy[2 * k + 2], path_xy[2 * k + 3]);
  if (length) *length = total;
  return n;
}

static void raycast_one(const double *occ, int nx, double *ray, int x0, int y0,
                        int x1, int y1) {
  /* :267-290 */
  const int dx = abs(x1 - x0), dy = abs(y1 - y0);
  const int sx = (x0 < x1) ? 1 : -1, sy = (y0 < y1) ? 1 : -1;
  int err = dx - dy;
  while (x0 != x1 || y0 != y1) {
    if (occ[IDX(x0, y0)] == 0) {
      ray[IDX(x0, y0)] = 0;
      ray[IDX(x1, y1)] = 0;
      return;
    }
    const int e2 = 2 * err;
    if (e2 > -dy) {
      err -= dy;
      x0 += sx;
    }
    if (e2 < dx) {
      err += dx;
      y0 += sy;
    }
  }
}

void vhp_oracle_raycast_all(const double *occ, int nx, int ny, int sx, int sy,
                            double *ray) {
  for (int i = 0; i < nx; ++i)   /* :228-232 */
    for (int j = 0; j < ny; ++j) raycast_one(occ, nx, ray, sx, sy, i, j);
}

void vhp_oracle_generate_environment(double *occ, int nx, int ny,
                                     long nb_of_obstacles, long min_w,
                                     long max_w, long min_h, long max_h,
                                     int seed) {
  /* src/environment.cpp:40-88.  All modulo arithmetic is done in size_t there
   * (nx_, minWidth... are size_t); rand() is non-negative so unsigned long
   * reproduces it. */
  const size_t n = (size_t)nx * (size_t)ny;
  for (size_t k = 0; k < n; ++k) occ[k] = 1.0;
  srand((unsigned)seed);
  for (long o = 0; o < nb_of_obstacles; ++o) {
    int col_1 = (int)(1 + ((unsigned long)rand() % ((unsigned long)nx + 1)));
    int col_2 = (int)((unsigned long)col_1 + (unsigned long)min_w +
                      ((unsigned long)rand() %
                       ((unsigned long)max_w - (unsigned long)min_w + 1)));
    if (col_1 > nx - 1) col_1 = nx - 1;
    if (col_2 > nx - 1) col_2 = nx - 1;
    int row_1 = (int)(1 + ((unsigned long)rand() % ((unsigned long)ny + 1)));
    int row_2 = (int)((unsigned long)row_1 + (unsigned long)min_h +
                      ((unsigned long)rand() %
                       ((unsigned long)max_h - (unsigned long)min_h + 1)));
    if (row_1 > ny - 1) row_1 = ny - 1;
    if (row_2 > ny - 1) row_2 = ny - 1;
    for (int x = col_1; x < col_2; ++x)
      for (int y = row_1; y < row_2; ++y) occ[IDX(x, y)] = 0.0;
  }
}

unsigned vhp_oracle_env_draw(unsigned long long seed, unsigned long long map,
                             unsigned long long obstacle, unsigned d) {
  unsigned long long z =
      seed + 0x9E3779B97F4A7C15ull * (map * 0x100000001B3ull + obstacle * 4ull + d + 1ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (unsigned)(z >> 33);
}

void vhp_oracle_generate_environment_counter(double *occ, int nx, int ny,
                                             long nb_of_obstacles, long min_w,
                                             long max_w, long min_h, long max_h,
                                             unsigned long long seed,
                                             unsigned long long map) {
  /* rectangle rule of src/environment.cpp:57-79, draws as documented in the header */
  const size_t n = (size_t)nx * (size_t)ny;
  for (size_t k = 0; k < n; ++k) occ[k] = 1.0;
  for (long o = 0; o < nb_of_obstacles; ++o) {
    long col_1 = 1 + (long)(vhp_oracle_env_draw(seed, map, (unsigned long long)o, 0) %
                            ((unsigned long)nx + 1));
    long col_2 = col_1 + min_w + (long)(vhp_oracle_env_draw(seed, map, (unsigned long long)o, 1) %
                                        ((unsigned long)max_w - (unsigned long)min_w + 1));
    long row_1 = 1 + (long)(vhp_oracle_env_draw(seed, map, (unsigned long long)o, 2) %
                            ((unsigned long)ny + 1));
    long row_2 = row_1 + min_h + (long)(vhp_oracle_env_draw(seed, map, (unsigned long long)o, 3) %
                                        ((unsigned long)max_h - (unsigned long)min_h + 1));
    if (col_1 > nx - 1) col_1 = nx - 1;
    if (col_2 > nx - 1) col_2 = nx - 1;
    if (row_1 > ny - 1) row_1 = ny - 1;
    if (row_2 > ny - 1) row_2 = ny - 1;
    for (long x = col_1; x < col_2; ++x)
      for (long y = row_1; y < row_2; ++y) occ[IDX(x, y)] = 0.0;
  }
}


/* ---- variants ------------------------------------------------------------------------------ */

/* getAccessibilityMap.m:10-117, one quadrant (dx, dy = +-1); i outer, j inner as in the .m file */
static void accessibility_quadrant(const double *occ, int nx, int ny, int sx, int sy, int dx,
                                   int dy, double alpha, double fac, double ls, double *vis) {
  const int Ex = dx > 0 ? nx - 1 - sx : sx, Ey = dy > 0 ? ny - 1 - sy : sy;
  for (int i = 0; i <= Ex; ++i) {
    const int X = sx + dx * i;
    for (int j = 0; j <= Ey; ++j) {
      const int Y = sy + dy * j;
      const size_t c0 = (size_t)X + (size_t)Y * nx;
      double v;
      const double jf = (double)j * fac;
      if (i == 0 && j == 0) {
        v = ls;
      } else if (i == 0) {
        v = alpha * vis[(size_t)X + (size_t)(Y - dy) * nx];
      } else if (j == 0) {
        v = alpha * vis[(size_t)(X - dx) + (size_t)Y * nx];
      } else if ((double)i == jf) {
        v = alpha * vis[(size_t)(X - dx) + (size_t)(Y - dy) * nx];
      } else if ((double)i > jf) {
        const double c = jf / (double)i;
        const double a = vis[(size_t)(X - dx) + (size_t)Y * nx];
        const double b = vis[(size_t)(X - dx) + (size_t)(Y - dy) * nx];
        const double f = a - c * (a - b);
        v = alpha * f;
      } else {
        const double c = (double)i / jf;
        const double a = vis[(size_t)X + (size_t)(Y - dy) * nx];
        const double b = vis[(size_t)(X - dx) + (size_t)(Y - dy) * nx];
        const double f = a - c * (a - b);
        v = alpha * f;
      }
      vis[c0] = v * occ[c0];
    }
  }
}

void vhp_oracle_accessibility_map(const double *occ, int nx, int ny, int sx, int sy,
                                  double alpha, double fac, double light_strength,
                                  double *vis) {
  for (size_t c = 0; c < (size_t)nx * ny; ++c) vis[c] = 0.0;
  accessibility_quadrant(occ, nx, ny, sx, sy, 1, 1, alpha, fac, light_strength, vis);   /* %% 1 */
  accessibility_quadrant(occ, nx, ny, sx, sy, -1, 1, alpha, fac, light_strength, vis);  /* %% 2 */
  accessibility_quadrant(occ, nx, ny, sx, sy, -1, -1, alpha, fac, light_strength, vis); /* %% 3 */
  accessibility_quadrant(occ, nx, ny, sx, sy, 1, -1, alpha, fac, light_strength, vis);  /* %% 4 */
}

/* computeVisibilityUsingQueue() as an order-free rule, see vhp_oracle.h.  Anti-diagonal order
 * d = i + j: every cell a cell depends on (values and pushers) has a smaller d; the quadrants
 * advance together because the axis cells belong to one quadrant each (dx >= 0 / dy >= 0 in the
 * reference's tests :738, :777, :816, :855) and are read by its neighbour. */
void vhp_oracle_visibility_cutoff(const double *occ, int nx, int ny, int sx, int sy,
                                  double cutoff, double *vis) {
  static const int qdx[4] = {1, -1, -1, 1}, qdy[4] = {1, 1, -1, -1};
  for (size_t c = 0; c < (size_t)nx * ny; ++c) vis[c] = 0.0;
  vis[(size_t)sx + (size_t)sy * nx] = 1.0; /* lightStrength_, :707 */
  for (int d = 1; d <= nx + ny; ++d) {
    for (int q = 0; q < 4; ++q) {
      const int dx = qdx[q], dy = qdy[q];
      const int Ex = dx > 0 ? nx - 1 - sx : sx, Ey = dy > 0 ? ny - 1 - sy : sy;
      for (int i = (d > Ey ? d - Ey : 0); i <= Ex && i <= d; ++i) {
        const int j = d - i;
        if ((i == 0 && dx < 0) || (j == 0 && dy < 0)) continue; /* the axis belongs to the + side */
        const int X = sx + dx * i, Y = sy + dy * j;
        const size_t c0 = (size_t)X + (size_t)Y * nx;
        if (occ[c0] == 0.0) continue; /* :732-734 */
#define VAT(a, b) vis[(size_t)(sx + dx * (a)) + (size_t)(sy + dy * (b)) * nx]
        int pushed = (i <= 1 && j <= 1); /* the eight neighbours of the source, :709-716 */
        if (!pushed && i - 1 >= 1 && VAT(i - 1, j) > cutoff) pushed = 1;
        if (!pushed && j - 1 >= 1 && VAT(i, j - 1) > cutoff) pushed = 1;
        if (!pushed && i == j && VAT(i - 1, j - 1) > cutoff) pushed = 1;
        if (!pushed) continue;
        double v;
        if (i == 0) v = VAT(0, j - 1);
        else if (j == 0) v = VAT(i - 1, 0);
        else if (i == j) v = VAT(i - 1, j - 1);
        else if (i > j) {
          const double c = (double)j / (double)i;
          const double a = VAT(i - 1, j), b = VAT(i - 1, j - 1);
          v = a - c * (a - b);
        } else {
          const double c = (double)i / (double)j;
          const double a = VAT(i, j - 1), b = VAT(i - 1, j - 1);
          v = a - c * (a - b);
        }
#undef VAT
        vis[c0] = v * occ[c0];
      }
    }
  }
}
