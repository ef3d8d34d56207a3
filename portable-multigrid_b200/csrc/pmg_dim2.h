// pmg_dim2.h -- the 2-D operator and transfers (dim = 2), one thread per DoF.
//
// The reference's polynomial_multigrid driver is 2-D (source/polynomial_multigrid/program.cc:439-441: dim = 2,
// fe_degree = 7, seven p-levels, at most (7 * 64 + 1)^2 = 201 601 DoFs), so the library has to run
// LaplaceOperator<2>::vmult (include/operators/portable_laplace_operator.h:557-719), PolynomialTransfer<2> and
// GeometricTransfer<2> (include/multigrid/portable_polynomial_tranfer.h:674-901,
// include/multigrid/portable_geometric_transfer.h:760-888) too.  Problems of that size live in L2 and are bound by
// launch latency, not by HBM or FP64: instead of the tile programs of the 3-D path each output DoF is one thread that
// gathers its row -- complete results, no atomics, no shared memory, the fused epilogues of the 3-D kernels unchanged:
//   A = cx Kx (x) My + cy Mx (x) Ky  (cx = hy / hx, cy = hx / hy; M, K the 1-D cell matrices of csrc/pmg_apply_sweep.h),
//   row (gx, gy) = sum over the <= 2 x 2 cells that contain the DoF of the cell matrix row, Dirichlet values read as 0
//   (:250-254), Dirichlet rows = identity (:718).
// Transfers: a fine DoF takes its value from one coarse cell that contains it (conforming: every containing cell gives
// the same value, which is the reference's sum of weight-1/multiplicity contributions, portable_geometric_transfer.h:
// 1336-1349); restriction is the transposed gather per coarse DoF, every fine DoF counted once.
// PMG_HD functions: the kernels of csrc/pmg_dim2.cu call them per DoF, tests/emu calls them in a loop.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PMG_HD __host__ __device__ __forceinline__
#else
#define PMG_HD inline
#endif

#define PMG2_MAX_N1 10

struct Pmg2Level {
  int p, nx, ny, Nx, Ny;
  unsigned faces;                   // Dirichlet faces bitmask, bits 0..3
  double M[PMG2_MAX_N1 * PMG2_MAX_N1], K[PMG2_MAX_N1 * PMG2_MAX_N1]; // 1-D cell matrices on [0,1], row-major (p+1)^2
  double cx, cy;                    // hy / hx, hx / hy
};

PMG_HD bool pmg2_dirichlet(int gx, int gy, int Nx, int Ny, unsigned faces)
{
  return (gx == 0 && (faces & 1u)) || (gx == Nx - 1 && (faces >> 1 & 1u)) || (gy == 0 && (faces >> 2 & 1u)) ||
         (gy == Ny - 1 && (faces >> 3 & 1u));
}

PMG_HD int pmg2_pos_type(int g, int N, int p) { return (g == 0) ? p : (g == N - 1) ? p + 1 : g % p; }

// (A u)(gx, gy) without the Dirichlet identity
PMG_HD double pmg2_row(const Pmg2Level &l, const double *u, int gx, int gy)
{
  const int p = l.p, n1 = p + 1;
  double acc = 0.0;
  for (int ey = 0; ey < 2; ++ey) {
    const int cy = gy / p - ey, j = gy - cy * p;
    if (cy < 0 || cy >= l.ny || j > p) continue;
    for (int ex = 0; ex < 2; ++ex) {
      const int cx = gx / p - ex, i = gx - cx * p;
      if (cx < 0 || cx >= l.nx || i > p) continue;
      for (int jj = 0; jj < n1; ++jj) {
        const int y = cy * p + jj;
        if ((y == 0 && (l.faces >> 2 & 1u)) || (y == l.Ny - 1 && (l.faces >> 3 & 1u))) continue;
        const double my = l.cx * l.M[j * n1 + jj], ky = l.cy * l.K[j * n1 + jj];
        const double *row = u + (int64_t)y * l.Nx + cx * p;
        double s = 0.0;
        for (int ii = 0; ii < n1; ++ii) {
          const int x = cx * p + ii;
          if ((x == 0 && (l.faces & 1u)) || (x == l.Nx - 1 && (l.faces >> 1 & 1u))) continue;
          s += row[ii] * (l.K[i * n1 + ii] * my + l.M[i * n1 + ii] * ky);
        }
        acc += s;
      }
    }
  }
  return acc;
}

// fused epilogue, the modes of PmgApplyMode (csrc/pmg_apply_sweep.h); dinv_tab: (p+2)^2 table by position type
PMG_HD double pmg2_apply_dof(const Pmg2Level &l, int mode, const double *u, const double *b, const double *xold, double f1,
                             double f2, const double *dinv_vec, const double *dinv_tab, int gx, int gy)
{
  const int64_t g = (int64_t)gy * l.Nx + gx;
  const bool dir = pmg2_dirichlet(gx, gy, l.Nx, l.Ny, l.faces);
  const double uc = u[g];
  const double Au = dir ? uc : pmg2_row(l, u, gx, gy);
  if (mode == 0) return Au;
  const double r = b[g] - Au;
  if (mode == 1) return r;
  const double dinv = dir ? 1.0 : (dinv_vec ? dinv_vec[g] : dinv_tab[pmg2_pos_type(gx, l.Nx, l.p) + (l.p + 2) * pmg2_pos_type(gy, l.Ny, l.p)]);
  const double corr = f2 * dinv * r;
  if (mode == 2) return uc + corr;
  return uc + f1 * (uc - (xold ? xold[g] : 0.0)) + corr;
}

struct Pmg2Xfer {
  int kind;            // 0 = h (fine = coarse refined once, same degree), 1 = p (same mesh, pc < pf)
  int pc, NC, NF;      // coarse degree, 1-D sizes of the coarse cell / fine patch
  int fstep;           // fine DoF offset per coarse cell: 2 p (h) or pf (p)
  int ncx, ncy;        // coarse cells
  int Ncx, Ncy, Nfx, Nfy;
  unsigned faces;
};

// value of the prolongated coarse vector at fine DoF (xf, yf); P1d: NC x NF, row = coarse node
PMG_HD double pmg2_prolongate_dof(const Pmg2Xfer &t, const double *P1d, const double *src, int xf, int yf)
{
  int cx = xf / t.fstep; if (cx > t.ncx - 1) cx = t.ncx - 1;
  int cy = yf / t.fstep; if (cy > t.ncy - 1) cy = t.ncy - 1;
  const int lx = xf - cx * t.fstep, ly = yf - cy * t.fstep;
  double acc = 0.0;
  for (int iy = 0; iy < t.NC; ++iy) {
    const int gy = cy * t.pc + iy;
    const double py = P1d[iy * t.NF + ly];
    if (py == 0.0) continue;
    double s = 0.0;
    for (int ix = 0; ix < t.NC; ++ix) {
      const int gx = cx * t.pc + ix;
      // h: constrained coarse DoFs read as 0 (dof_indices_coarse == invalid, portable_geometric_transfer.h:170-173);
      // p: read unmasked (portable_polynomial_tranfer.h:115-121)
      if (t.kind == 0 && pmg2_dirichlet(gx, gy, t.Ncx, t.Ncy, t.faces)) continue;
      s += P1d[ix * t.NF + lx] * src[(int64_t)gy * t.Ncx + gx];
    }
    acc += py * s;
  }
  return acc;
}

// 1-D: the fine DoFs that coarse DoF X reaches: cells c0..c1, and per cell the local coarse index; the fine DoF shared by
// two cells is taken from the left one
PMG_HD double pmg2_restrict_dof(const Pmg2Xfer &t, const double *P1d, const double *src, int X, int Y)
{
  double acc = 0.0;
  for (int ey = 0; ey < 2; ++ey) {
    const int cy = Y / t.pc - ey, iy = Y - cy * t.pc;
    if (cy < 0 || cy >= t.ncy || iy > t.pc) continue;
    const bool lower_y_exists = (ey == 0 && iy == 0 && cy > 0); // the cell below also contains Y: it takes the shared row
    for (int ly = lower_y_exists ? 1 : 0; ly < t.NF; ++ly) {
      const double py = P1d[iy * t.NF + ly];
      if (py == 0.0) continue;
      const int yf = cy * t.fstep + ly;
      double s = 0.0;
      for (int ex = 0; ex < 2; ++ex) {
        const int cx = X / t.pc - ex, ix = X - cx * t.pc;
        if (cx < 0 || cx >= t.ncx || ix > t.pc) continue;
        const bool left_exists = (ex == 0 && ix == 0 && cx > 0);
        for (int lx = left_exists ? 1 : 0; lx < t.NF; ++lx) {
          const int xf = cx * t.fstep + lx;
          if (pmg2_dirichlet(xf, yf, t.Nfx, t.Nfy, t.faces)) continue; // weights vanish on constrained fine DoFs
          s += P1d[ix * t.NF + lx] * src[(int64_t)yf * t.Nfx + xf];
        }
      }
      acc += py * s;
    }
  }
  return acc;
}
