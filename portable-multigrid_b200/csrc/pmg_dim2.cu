// pmg_dim2.cu -- kernels of the 2-D path (csrc/pmg_dim2.h): one thread per DoF, rows along x.  sm_100a only.
#include "pmg_dim2.h"
#include "pmg_cuda_common.h"
#include "pmg_kernels.h"

namespace {

__global__ void __launch_bounds__(128) k2_apply(const __grid_constant__ Pmg2Level l, int mode, const double *__restrict__ u,
                                               const double *__restrict__ b, const double *xold, double *out, double f1, double f2,
                                               const double *__restrict__ dinv_vec, const double *__restrict__ dinv_tab)
{
  const int gx = blockIdx.x * blockDim.x + threadIdx.x, gy = blockIdx.y;
  if (gx >= l.Nx) return;
  out[(int64_t)gy * l.Nx + gx] = pmg2_apply_dof(l, mode, u, b, xold, f1, f2, dinv_vec, dinv_tab, gx, gy);
}

__global__ void __launch_bounds__(128) k2_prolongate(const __grid_constant__ Pmg2Xfer t, const double *__restrict__ P1d, double *dst,
                                                    const double *__restrict__ src)
{
  const int xf = blockIdx.x * blockDim.x + threadIdx.x, yf = blockIdx.y;
  if (xf >= t.Nfx || pmg2_dirichlet(xf, yf, t.Nfx, t.Nfy, t.faces)) return;
  dst[(int64_t)yf * t.Nfx + xf] += pmg2_prolongate_dof(t, P1d, src, xf, yf);
}

__global__ void __launch_bounds__(128) k2_restrict(const __grid_constant__ Pmg2Xfer t, const double *__restrict__ P1d, double *dst,
                                                  const double *__restrict__ src)
{
  const int X = blockIdx.x * blockDim.x + threadIdx.x, Y = blockIdx.y;
  if (X >= t.Ncx || pmg2_dirichlet(X, Y, t.Ncx, t.Ncy, t.faces)) return; // constrained coarse DoFs are skipped (:674-682)
  dst[(int64_t)Y * t.Ncx + X] += pmg2_restrict_dof(t, P1d, src, X, Y);
}

int fill_xfer(int kind, const pmgk_level *c, const pmgk_level *f, Pmg2Xfer *t)
{
  t->kind = kind; t->pc = c->degree; t->NC = c->degree + 1;
  t->NF = (kind == 0) ? 2 * c->degree + 1 : f->degree + 1;
  t->fstep = t->NF - 1;
  t->ncx = c->nx; t->ncy = c->ny;
  t->Ncx = c->Nx; t->Ncy = c->Ny; t->Nfx = f->Nx; t->Nfy = f->Ny;
  t->faces = c->faces & 0xFu;
  if (t->NC > PMG2_MAX_N1) return PMG_ERR_UNSUPPORTED;
  return 0;
}

} // namespace

// called by pmgk_apply (pmg_apply.cu) for levels with dim == 2
int pmg_dim2_apply(const pmgk_level *lv, int mode, const double *u, const double *b, const double *xold, double *out, double f1,
                   double f2, cudaStream_t s, int *geom)
{
  const int n1 = lv->degree + 1;
  if (n1 > PMG2_MAX_N1) return PMG_ERR_UNSUPPORTED;
  Pmg2Level l;
  l.p = lv->degree; l.nx = lv->nx; l.ny = lv->ny; l.Nx = lv->Nx; l.Ny = lv->Ny; l.faces = lv->faces & 0xFu;
  for (int i = 0; i < n1 * n1; ++i) { l.M[i] = lv->Mref[i]; l.K[i] = lv->Kref[i]; }
  l.cx = lv->h[1] / lv->h[0]; l.cy = lv->h[0] / lv->h[1];
  const dim3 grid((unsigned)((lv->Nx + 127) / 128), (unsigned)lv->Ny);
  if (geom) { geom[0] = (int)(grid.x * grid.y); geom[1] = 128; geom[2] = 0; geom[3] = 1; return 0; }
  /* the (p+2)^3 table of a 2-D level repeats its (p+2)^2 values for every z type (host/pmg_fe.c) */
  k2_apply<<<grid, 128, 0, s>>>(l, mode, u, b, xold, out, f1, f2, lv->dinv_vec, lv->dinv_tab);
  PMG_CUDA_CHECK(cudaGetLastError());
  pmg_count_launch(1);
  return 0;
}

int pmg_dim2_prolongate(int kind, const pmgk_level *c, const pmgk_level *f, const double *P1d, double *dst, const double *src, cudaStream_t s)
{
  Pmg2Xfer t;
  const int rc = fill_xfer(kind, c, f, &t);
  if (rc) return rc;
  k2_prolongate<<<dim3((unsigned)((t.Nfx + 127) / 128), (unsigned)t.Nfy), 128, 0, s>>>(t, P1d, dst, src);
  PMG_CUDA_CHECK(cudaGetLastError());
  pmg_count_launch(1);
  return 0;
}

int pmg_dim2_restrict(int kind, const pmgk_level *c, const pmgk_level *f, const double *P1d, double *dst, const double *src, cudaStream_t s)
{
  Pmg2Xfer t;
  const int rc = fill_xfer(kind, c, f, &t);
  if (rc) return rc;
  k2_restrict<<<dim3((unsigned)((t.Ncx + 127) / 128), (unsigned)t.Ncy), 128, 0, s>>>(t, P1d, dst, src);
  PMG_CUDA_CHECK(cudaGetLastError());
  pmg_count_launch(1);
  return 0;
}
