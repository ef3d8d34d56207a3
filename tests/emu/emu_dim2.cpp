// tests/emu/emu_dim2.cpp -- host emulation of the 2-D kernels (TEST INFRASTRUCTURE): the per-DoF functions of
// csrc/pmg_dim2.h, which the kernels of csrc/pmg_dim2.cu call with one thread per DoF, run here in a loop over the DoFs
// with the tables of host/pmg_fe.c.  Built only by tests/; the product has no CPU path.
#include <cstring>
#include <vector>
#include "pmg_dim2.h"

extern "C" {
void pmg_fe_pencil(int p, double *M, double *K);
void pmg_fe_dinv_table(int p, const double h[3], int dim, double *tab);
void pmg_fe_prolongation_h(int p, double *P);
void pmg_fe_prolongation_p(int pc, int pf, double *P);

int emu2_apply(int p, int nx, int ny, unsigned faces, int mode, const double *u, const double *b, const double *xold, double *out,
               double f1, double f2, const double *dinv_vec)
{
  if (p + 1 > PMG2_MAX_N1) return -3;
  Pmg2Level l;
  l.p = p; l.nx = nx; l.ny = ny; l.Nx = nx * p + 1; l.Ny = ny * p + 1; l.faces = faces & 0xFu;
  pmg_fe_pencil(p, l.M, l.K);
  const double h[3] = {1.0 / nx, 1.0 / ny, 1.0};
  l.cx = h[1] / h[0]; l.cy = h[0] / h[1];
  const int T = p + 2;
  std::vector<double> tab((size_t)T * T * T);
  pmg_fe_dinv_table(p, h, 2, tab.data());
  std::vector<double> res((size_t)l.Nx * l.Ny); // out may alias xold: every "thread" reads before any writes, as on the device
  for (int gy = 0; gy < l.Ny; ++gy)
    for (int gx = 0; gx < l.Nx; ++gx)
      res[(size_t)gy * l.Nx + gx] = pmg2_apply_dof(l, mode, u, b, xold, f1, f2, dinv_vec, tab.data(), gx, gy);
  std::memcpy(out, res.data(), res.size() * sizeof(double));
  return 0;
}

static int make_xfer(int kind, int pc, int pf, int ncx, int ncy, unsigned faces, Pmg2Xfer *t, std::vector<double> *P)
{
  t->kind = kind; t->pc = pc; t->NC = pc + 1; t->NF = (kind == 0) ? 2 * pc + 1 : pf + 1; t->fstep = t->NF - 1;
  t->ncx = ncx; t->ncy = ncy; t->Ncx = ncx * pc + 1; t->Ncy = ncy * pc + 1;
  t->Nfx = (kind == 0) ? 2 * ncx * pc + 1 : ncx * pf + 1; t->Nfy = (kind == 0) ? 2 * ncy * pc + 1 : ncy * pf + 1;
  t->faces = faces & 0xFu;
  P->resize((size_t)t->NC * t->NF);
  if (kind == 0) pmg_fe_prolongation_h(pc, P->data()); else pmg_fe_prolongation_p(pc, pf, P->data());
  return 0;
}

int emu2_prolongate_and_add(int kind, int pc, int pf, int ncx, int ncy, unsigned faces, double *dst_fine, const double *src_coarse)
{
  Pmg2Xfer t; std::vector<double> P;
  make_xfer(kind, pc, pf, ncx, ncy, faces, &t, &P);
  for (int yf = 0; yf < t.Nfy; ++yf)
    for (int xf = 0; xf < t.Nfx; ++xf)
      if (!pmg2_dirichlet(xf, yf, t.Nfx, t.Nfy, t.faces)) dst_fine[(size_t)yf * t.Nfx + xf] += pmg2_prolongate_dof(t, P.data(), src_coarse, xf, yf);
  return 0;
}

int emu2_restrict_and_add(int kind, int pc, int pf, int ncx, int ncy, unsigned faces, double *dst_coarse, const double *src_fine)
{
  Pmg2Xfer t; std::vector<double> P;
  make_xfer(kind, pc, pf, ncx, ncy, faces, &t, &P);
  for (int Y = 0; Y < t.Ncy; ++Y)
    for (int X = 0; X < t.Ncx; ++X)
      if (!pmg2_dirichlet(X, Y, t.Ncx, t.Ncy, t.faces)) dst_coarse[(size_t)Y * t.Ncx + X] += pmg2_restrict_dof(t, P.data(), src_fine, X, Y);
  return 0;
}
}
