// tests/emu/emu_coarse.cpp -- host emulator of the single-CTA coarse V-cycle program (TEST INFRASTRUCTURE): runs
// csrc/pmg_coarse_cycle.h thread by thread (1024 emulated threads, barriers = phase boundaries) on host arrays, with the
// tables of host/pmg_fe.c and the Chebyshev parameters the caller passes.  Built only by tests/.
#include <cstring>
#include <vector>
#include "pmg_coarse_cycle.h"

extern "C" {
void pmg_fe_pencil(int p, double *M, double *K);
void pmg_fe_dinv_table(int p, const double h[3], int dim, double *tab);
void pmg_fe_prolongation_h(int p, double *P);
}

struct HostExecC {
  template <class F> void for_each_thread(F f) { for (int t = 0; t < 1024; ++t) f(t); }
  void sync() {}
};

// levels coarse -> fine: level l has n0 * 2^l cells per direction (n0 = {nx0, ny0, nz0}); cheb[l] = {degree, theta, delta};
// dst = cycle(src) on the finest level
extern "C" int emu_coarse_cycle(int p, int n_levels, const int *n0, unsigned faces, int pre, int post, const int *cheb_degree,
                                const double *theta, const double *delta, double *dst, const double *src)
{
  if (p + 1 > PMG_CC_MAX_N1 || n_levels > PMG_CC_MAX_LEVELS) return -3;
  PmgCoarseParams q;
  std::memset(&q, 0, sizeof(q));
  q.n_levels = n_levels; q.pre = pre; q.post = post; q.p = p; q.faces = faces;
  pmg_fe_pencil(p, q.M, q.K);
  pmg_fe_prolongation_h(p, q.P1d);
  const int T = p + 2;
  std::vector<std::vector<double>> tabs(n_levels), bufs(4 * n_levels);
  for (int l = 0; l < n_levels; ++l) {
    PmgCoarseLevel &c = q.lv[l];
    c.nx = n0[0] << l; c.ny = n0[1] << l; c.nz = n0[2] << l;
    c.Nx = c.nx * p + 1; c.Ny = c.ny * p + 1; c.Nz = c.nz * p + 1;
    const double h[3] = {1.0 / c.nx, 1.0 / c.ny, 1.0 / c.nz};
    c.cx = h[1] * h[2] / h[0]; c.cy = h[0] * h[2] / h[1]; c.cz = h[0] * h[1] / h[2];
    c.degree = cheb_degree[l]; c.theta = theta[l]; c.delta = delta[l];
    tabs[l].resize((size_t)T * T * T);
    pmg_fe_dinv_table(p, h, 3, tabs[l].data());
    c.dinv_tab = tabs[l].data();
    const size_t n = (size_t)c.Nx * c.Ny * c.Nz;
    for (int k = 0; k < 4; ++k) bufs[4 * l + k].assign(n, 1e300); // poisoned: every vector is written before it is read
    c.sol = bufs[4 * l].data(); c.rhs = bufs[4 * l + 1].data(); c.tmp = bufs[4 * l + 2].data(); c.res = bufs[4 * l + 3].data();
  }
  q.lv[n_levels - 1].sol = dst;
  q.lv[n_levels - 1].rhs = const_cast<double *>(src);
  HostExecC ex;
  switch (p) {
    case 1: PmgCoarseCycle<1024, 1>::run(q, ex); break;
    case 2: PmgCoarseCycle<1024, 2>::run(q, ex); break;
    case 3: PmgCoarseCycle<1024, 3>::run(q, ex); break;
    case 4: PmgCoarseCycle<1024, 4>::run(q, ex); break;
    default: return -3;
  }
  return 0;
}
