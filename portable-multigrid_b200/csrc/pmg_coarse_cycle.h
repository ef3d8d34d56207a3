// pmg_coarse_cycle.h -- the coarse part of the V-cycle as ONE single-CTA program.
//
// VCycleMultigrid::v_cycle (reference include/multigrid/portable_v_cycle_multigrid.h:128-190) recurses down to a mesh of
// one cell; every level runs (pre + post) * k + 1 operator applies, a restriction and a prolongation whatever its size
// (k = Chebyshev degree, configured at source/geometric_multigrid/program.cc:267-279).  On the small levels (8 .. ~40 000
// DoFs) each of these is a kernel whose run time is its launch-to-result latency (5-7 us inside a CUDA graph on B200,
// tools/small_levels.py): ~23 launches per level, ~0.13 ms per level and cycle, 20 % of the C2 cycle and 60 % of the C1 cycle.
// Here the levels 0 .. L of one degree (h-hierarchy, constant coefficient, held by one GPU) are processed by a single CTA:
// the same sequence of operations as host/pmg_vcycle.c + host/pmg_smoother.c issue (fused Chebyshev steps with the same
// coefficients, residual, transfers), separated by block barriers instead of kernel boundaries; vectors stay in global
// memory (L1 / L2 resident at these sizes).  Every operation is a loop over the output DoFs of a level, each DoF
// gathering its row (as csrc/pmg_dim2.h does in 2-D): complete results, no atomics, deterministic.
//   A u (row gx, gy, gz) = sum over the <= 8 cells containing the DoF of the tensor-product cell matrix row
//       cx K (x) M (x) M + cy M (x) K (x) M + cz M (x) M (x) K,  Dirichlet values read as 0, Dirichlet rows = identity;
//   prolongation: a fine DoF takes its value from one coarse cell that contains it; restriction: transposed gather.
// STATUS: opt-in (PMG_COARSE_KERNEL=1).  Measured on B200 it is slower than the per-level kernels inside the CUDA graph at every
// size bound (profiles/r01_coarse_cycle_kernel_sweep.txt): a block barrier plus an L2 round trip per operation and one SM's
// worth of gather throughput cost more than the 5.4 us a per-level launch costs.  Kept, tested, as the starting point for a
// shared-memory / thread-block-cluster version.
// Written against the executor interface of the other tile programs (for_each_thread / sync): the same source is the
// CUDA kernel (csrc/pmg_coarse_cycle.cu) and runs under the host emulator (tests/emu/emu_coarse.cpp).
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define PMG_HD __host__ __device__ __forceinline__
#else
#define PMG_HD inline
#endif

#define PMG_CC_MAX_LEVELS 8
#define PMG_CC_MAX_N1 5 /* degrees 1 .. 4 */

struct PmgCoarseLevel {
  int nx, ny, nz, Nx, Ny, Nz;
  double cx, cy, cz;       // hy hz / hx, hx hz / hy, hx hy / hz
  int degree;              // Chebyshev degree of the level's smoother
  double theta, delta;     // its parameters (host/pmg_smoother.c)
  const double *dinv_tab;  // (p+2)^3 inverse-diagonal table by position type
  double *sol, *rhs, *tmp, *res; // the level's vectors; sol / rhs of the top level are the cycle's output / input
};

struct PmgCoarseParams {
  int n_levels, pre, post;
  int p;                   // the levels' common degree
  unsigned faces;
  double M[PMG_CC_MAX_N1 * PMG_CC_MAX_N1], K[PMG_CC_MAX_N1 * PMG_CC_MAX_N1]; // 1-D cell matrices on [0,1]
  double P1d[PMG_CC_MAX_N1 * (2 * PMG_CC_MAX_N1 - 1)];                        // h-prolongation, (p+1) x (2p+1), row = coarse node
  PmgCoarseLevel lv[PMG_CC_MAX_LEVELS];
};

// NT: threads of the CTA; P: the levels' degree (compile time: the gather loops unroll)
template <int NT, int P>
struct PmgCoarseCycle {
  static PMG_HD bool dirichlet(const PmgCoarseParams &q, const PmgCoarseLevel &l, int gx, int gy, int gz)
  {
    const unsigned f = q.faces;
    return (gx == 0 && (f & 1u)) || (gx == l.Nx - 1 && (f >> 1 & 1u)) || (gy == 0 && (f >> 2 & 1u)) ||
           (gy == l.Ny - 1 && (f >> 3 & 1u)) || (gz == 0 && (f >> 4 & 1u)) || (gz == l.Nz - 1 && (f >> 5 & 1u));
  }
  static PMG_HD int pos_type(int g, int N, int p) { return (g == 0) ? p : (g == N - 1) ? p + 1 : g % p; }

  // (A u)(gx, gy, gz) without the Dirichlet identity
  static PMG_HD double row(const PmgCoarseParams &q, const PmgCoarseLevel &l, const double *u, int gx, int gy, int gz)
  {
    constexpr int p = P, n1 = p + 1;
    const unsigned f = q.faces;
    double acc = 0.0;
    for (int ez = 0; ez < 2; ++ez) {
      const int cz = gz / p - ez, k = gz - cz * p;
      if (cz < 0 || cz >= l.nz || k > p) continue;
      for (int ey = 0; ey < 2; ++ey) {
        const int cy = gy / p - ey, j = gy - cy * p;
        if (cy < 0 || cy >= l.ny || j > p) continue;
        for (int ex = 0; ex < 2; ++ex) {
          const int cx = gx / p - ex, i = gx - cx * p;
          if (cx < 0 || cx >= l.nx || i > p) continue;
#pragma unroll
          for (int kk = 0; kk < n1; ++kk) {
            const int z = cz * p + kk;
            if ((z == 0 && (f >> 4 & 1u)) || (z == l.Nz - 1 && (f >> 5 & 1u))) continue;
            const double mz = q.M[k * n1 + kk], kz = q.K[k * n1 + kk];
#pragma unroll
            for (int jj = 0; jj < n1; ++jj) {
              const int y = cy * p + jj;
              if ((y == 0 && (f >> 2 & 1u)) || (y == l.Ny - 1 && (f >> 3 & 1u))) continue;
              const double my = q.M[j * n1 + jj], ky = q.K[j * n1 + jj];
              const double a = l.cx * my * mz, b = l.cy * ky * mz + l.cz * my * kz;
              const double *r = u + ((int64_t)z * l.Ny + y) * l.Nx + cx * p;
              double s = 0.0;
#pragma unroll
              for (int ii = 0; ii < n1; ++ii) {
                const int x = cx * p + ii;
                if ((x == 0 && (f & 1u)) || (x == l.Nx - 1 && (f >> 1 & 1u))) continue;
                s += r[ii] * (q.K[i * n1 + ii] * a + q.M[i * n1 + ii] * b);
              }
              acc += s;
            }
          }
        }
      }
    }
    return acc;
  }

  // fused apply over the level, modes of PmgApplyMode: 1 residual, 2 first Chebyshev step, 3 Chebyshev step; out may alias xold
  static PMG_HD void apply(const PmgCoarseParams &q, const PmgCoarseLevel &l, int tid, int mode, const double *u, const double *b,
                           const double *xold, double *out, double f1, double f2)
  {
    constexpr int T = P + 2;
    const int64_t n = (int64_t)l.Nx * l.Ny * l.Nz;
    for (int64_t g = tid; g < n; g += NT) {
      const int gx = (int)(g % l.Nx), gy = (int)((g / l.Nx) % l.Ny), gz = (int)(g / ((int64_t)l.Nx * l.Ny));
      const bool dir = dirichlet(q, l, gx, gy, gz);
      const double uc = u[g];
      const double Au = dir ? uc : row(q, l, u, gx, gy, gz);
      const double r = b[g] - Au;
      double v;
      if (mode == 1) v = r;
      else {
        const double dinv = dir ? 1.0 : l.dinv_tab[pos_type(gx, l.Nx, P) + T * (pos_type(gy, l.Ny, P) + T * pos_type(gz, l.Nz, P))];
        const double corr = f2 * dinv * r;
        v = (mode == 2) ? uc + corr : uc + f1 * (uc - (xold ? xold[g] : 0.0)) + corr;
      }
      out[g] = v;
    }
  }

  // out = f Dinv b (first Chebyshev step from a zero guess)
  static PMG_HD void scale_dinv(const PmgCoarseParams &q, const PmgCoarseLevel &l, int tid, double f, const double *b, double *out)
  {
    constexpr int T = P + 2;
    const int64_t n = (int64_t)l.Nx * l.Ny * l.Nz;
    for (int64_t g = tid; g < n; g += NT) {
      const int gx = (int)(g % l.Nx), gy = (int)((g / l.Nx) % l.Ny), gz = (int)(g / ((int64_t)l.Nx * l.Ny));
      const double dinv = dirichlet(q, l, gx, gy, gz)
                              ? 1.0
                              : l.dinv_tab[pos_type(gx, l.Nx, P) + T * (pos_type(gy, l.Ny, P) + T * pos_type(gz, l.Nz, P))];
      out[g] = f * dinv * b[g];
    }
  }

  static PMG_HD void copy(const PmgCoarseLevel &l, int tid, double *dst, const double *src)
  {
    const int64_t n = (int64_t)l.Nx * l.Ny * l.Nz;
    for (int64_t g = tid; g < n; g += NT) dst[g] = src[g];
  }

  // coarse rhs = P^T (fine residual with constrained fine DoFs dropped); constrained coarse DoFs get 0
  static PMG_HD void restrict_to(const PmgCoarseParams &q, const PmgCoarseLevel &c, const PmgCoarseLevel &f, int tid, double *dst,
                                 const double *src)
  {
    constexpr int p = P, NF = 2 * p + 1, fstep = 2 * p;
    const int64_t n = (int64_t)c.Nx * c.Ny * c.Nz;
    for (int64_t g = tid; g < n; g += NT) {
      const int X = (int)(g % c.Nx), Y = (int)((g / c.Nx) % c.Ny), Z = (int)(g / ((int64_t)c.Nx * c.Ny));
      double acc = 0.0;
      if (!dirichlet(q, c, X, Y, Z)) {
        for (int ez = 0; ez < 2; ++ez) {
          const int cz = Z / p - ez, iz = Z - cz * p;
          if (cz < 0 || cz >= c.nz || iz > p) continue;
          for (int lz = (ez == 0 && iz == 0 && cz > 0) ? 1 : 0; lz < NF; ++lz) { // a fine plane shared by two cells: the lower one's
            const double pz = q.P1d[iz * NF + lz];
            if (pz == 0.0) continue;
            const int zf = cz * fstep + lz;
            for (int ey = 0; ey < 2; ++ey) {
              const int cy = Y / p - ey, iy = Y - cy * p;
              if (cy < 0 || cy >= c.ny || iy > p) continue;
              for (int ly = (ey == 0 && iy == 0 && cy > 0) ? 1 : 0; ly < NF; ++ly) {
                const double pyz = pz * q.P1d[iy * NF + ly];
                if (pyz == 0.0) continue;
                const int yf = cy * fstep + ly;
                double s = 0.0;
                for (int ex = 0; ex < 2; ++ex) {
                  const int cx = X / p - ex, ix = X - cx * p;
                  if (cx < 0 || cx >= c.nx || ix > p) continue;
                  for (int lx = (ex == 0 && ix == 0 && cx > 0) ? 1 : 0; lx < NF; ++lx) {
                    const int xf = cx * fstep + lx;
                    if (dirichlet(q, f, xf, yf, zf)) continue; // weights vanish on constrained fine DoFs
                    s += q.P1d[ix * NF + lx] * src[((int64_t)zf * f.Ny + yf) * f.Nx + xf];
                  }
                }
                acc += pyz * s;
              }
            }
          }
        }
      }
      dst[g] = acc;
    }
  }

  // fine dst (+)= P coarse src on unconstrained fine DoFs; add == false: dst = P src (constrained: 0)
  static PMG_HD void prolongate(const PmgCoarseParams &q, const PmgCoarseLevel &c, const PmgCoarseLevel &f, int tid, double *dst,
                                const double *src, bool add)
  {
    constexpr int p = P, NC = p + 1, NF = 2 * p + 1, fstep = 2 * p;
    const int64_t n = (int64_t)f.Nx * f.Ny * f.Nz;
    for (int64_t g = tid; g < n; g += NT) {
      const int xf = (int)(g % f.Nx), yf = (int)((g / f.Nx) % f.Ny), zf = (int)(g / ((int64_t)f.Nx * f.Ny));
      double acc = 0.0;
      if (!dirichlet(q, f, xf, yf, zf)) {
        int cx = xf / fstep; if (cx > c.nx - 1) cx = c.nx - 1;
        int cy = yf / fstep; if (cy > c.ny - 1) cy = c.ny - 1;
        int cz = zf / fstep; if (cz > c.nz - 1) cz = c.nz - 1;
        const int lx = xf - cx * fstep, ly = yf - cy * fstep, lz = zf - cz * fstep;
        for (int iz = 0; iz < NC; ++iz) {
          const double pz = q.P1d[iz * NF + lz];
          if (pz == 0.0) continue;
          for (int iy = 0; iy < NC; ++iy) {
            const double pyz = pz * q.P1d[iy * NF + ly];
            if (pyz == 0.0) continue;
            double s = 0.0;
            for (int ix = 0; ix < NC; ++ix) {
              const int X = cx * p + ix, Y = cy * p + iy, Z = cz * p + iz;
              if (dirichlet(q, c, X, Y, Z)) continue; // constrained coarse DoFs read as 0 (:170-173)
              s += q.P1d[ix * NF + lx] * src[((int64_t)Z * c.Ny + Y) * c.Nx + X];
            }
            acc += pyz * s;
          }
        }
      } else if (add) {
        continue;
      }
      dst[g] = add ? dst[g] + acc : acc;
    }
  }

  // smooth(): host/pmg_smoother.c pmg_chebyshev_smooth, operation by operation.  cur / other ping-pong; returns (in cur)
  // the buffer that holds the result
  template <class Exec>
  static PMG_HD void smooth(const PmgCoarseParams &q, const PmgCoarseLevel &l, Exec &ex, double *&cur, double *&other, const double *rhs,
                            bool zero_guess)
  {
    const double theta = l.theta, delta = l.delta;
    bool other_is_xold = false;
    if (zero_guess) {
      double *c = cur;
      ex.for_each_thread([&](int tid) { scale_dinv(q, l, tid, 1.0 / theta, rhs, c); });
      ex.sync();
    } else {
      double *c = cur, *o = other;
      ex.for_each_thread([&](int tid) { apply(q, l, tid, 2, c, rhs, nullptr, o, 0.0, 1.0 / theta); });
      ex.sync();
      cur = o; other = c;
      other_is_xold = true;
    }
    if (l.degree >= 2 && fabs(delta) >= 1e-40) {
      double rhok = delta / theta;
      const double sigma = theta / delta;
      for (int k = 0; k < l.degree - 1; ++k) {
        const double rhokp = 1.0 / (2.0 * sigma - rhok);
        const double f1 = rhokp * rhok, f2 = 2.0 * rhokp / delta;
        rhok = rhokp;
        double *c = cur, *o = other;
        const double *xo = other_is_xold ? o : nullptr;
        ex.for_each_thread([&](int tid) { apply(q, l, tid, 3, c, rhs, xo, o, f1, f2); });
        ex.sync();
        cur = o; other = c;
        other_is_xold = true;
      }
    }
  }

  // v_cycle(top level, sol = 0 on entry): host/pmg_vcycle.c v_cycle unrolled into a down sweep and an up sweep
  template <class Exec>
  static PMG_HD void run(const PmgCoarseParams &q, Exec &ex)
  {
    const int top = q.n_levels - 1;
    double *cur[PMG_CC_MAX_LEVELS], *other[PMG_CC_MAX_LEVELS];
    bool zg[PMG_CC_MAX_LEVELS];
    for (int lev = top; lev >= 1; --lev) {
      const PmgCoarseLevel &l = q.lv[lev];
      cur[lev] = l.sol; other[lev] = l.tmp; zg[lev] = true;
      for (int s = 0; s < q.pre; ++s) { smooth(q, l, ex, cur[lev], other[lev], l.rhs, zg[lev]); zg[lev] = false; }
      // residual (no pre-smoothing and a zero iterate: the residual is the right-hand side), restricted to the next level
      const double *res = l.rhs;
      if (!zg[lev]) {
        const double *c = cur[lev];
        ex.for_each_thread([&](int tid) { apply(q, l, tid, 1, c, l.rhs, nullptr, l.res, 0.0, 0.0); });
        ex.sync();
        res = l.res;
      }
      const PmgCoarseLevel &c = q.lv[lev - 1];
      ex.for_each_thread([&](int tid) { restrict_to(q, c, l, tid, c.rhs, res); });
      ex.sync();
    }
    { // coarsest level: one smooth() from a zero guess
      const PmgCoarseLevel &l = q.lv[0];
      cur[0] = l.sol; other[0] = l.tmp;
      smooth(q, l, ex, cur[0], other[0], l.rhs, true);
      if (cur[0] != l.sol) {
        const double *c = cur[0];
        ex.for_each_thread([&](int tid) { copy(l, tid, l.sol, c); });
        ex.sync();
      }
    }
    for (int lev = 1; lev <= top; ++lev) {
      const PmgCoarseLevel &l = q.lv[lev];
      const PmgCoarseLevel &c = q.lv[lev - 1];
      {
        double *dst = cur[lev];
        const bool add = !zg[lev];
        ex.for_each_thread([&](int tid) { prolongate(q, c, l, tid, dst, c.sol, add); });
        ex.sync();
      }
      for (int s = 0; s < q.post; ++s) smooth(q, l, ex, cur[lev], other[lev], l.rhs, false);
      if (cur[lev] != l.sol) {
        const double *cc = cur[lev];
        ex.for_each_thread([&](int tid) { copy(l, tid, l.sol, cc); });
        ex.sync();
      }
    }
  }
};
