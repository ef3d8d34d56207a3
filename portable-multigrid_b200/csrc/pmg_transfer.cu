// pmg_transfer.cu -- h- and p-multigrid transfer kernels (K4-K7) and inverse-diagonal kernels (K2/K3).
//
// Replaces, on the structured box mesh,
//   h_mg_transfer::CellProlongationKernel / CellRestrictionKernel
//     (reference include/multigrid/portable_geometric_transfer.h:150-387, 450-684) and
//   p_mg_transfer::CellProlongationKernel / CellRestrictionKernel
//     (reference include/multigrid/portable_polynomial_tranfer.h:103-326, 390-615).
// Same tensor algebra (three 1-D contractions with the (pc+1) x nf1 matrix), different data
// movement: indices, weights and masks are computed from the cell position instead of being
// streamed (the reference reads a u32 index and an f64 weight per patch dof), and every fine
// dof is handled by exactly one patch (low-side ownership) so neither atomics nor the weight
// array are needed:  sum_patches w_f P (.) with w_f = 1/multiplicity  ==  P (.) taken once.
// Restriction is two deterministic passes (cell-local contraction, then a gather per coarse dof).
#include "pmg_cuda_common.h"
#include "pmg_kernels.h"

namespace {

struct XferGeom {
  int kind;            // 0 = h, 1 = p
  int pc, pf;          // degrees
  int NC, NF;          // 1-D sizes of the coarse cell / fine patch
  int fstep;           // fine dof offset per coarse cell: 2p (h) or pf (p)
  int ncx, ncy, ncz;   // coarse cells (global)
  int Ncx, Ncy, Ncz;   // coarse dofs (global)
  int Nfx, Nfy, Nfz;   // fine dofs (global)
  unsigned faces;
  // slabs
  int c_z0, f_z0;      // first stored plane of the coarse / fine local vectors
  int ccz_lo, ccz_hi;  // coarse cell layers handled by this rank
  int c_zown_lo, c_zown_hi, f_zown_lo, f_zown_hi;
  int c_nzl, f_nzl;
};

__device__ __forceinline__ bool on_dirichlet(int gx, int gy, int gz, int Nx, int Ny, int Nz, unsigned faces)
{
  return (gx == 0 && (faces & 1u)) || (gx == Nx - 1 && (faces >> 1 & 1u)) || (gy == 0 && (faces >> 2 & 1u)) ||
         (gy == Ny - 1 && (faces >> 3 & 1u)) || (gz == 0 && (faces >> 4 & 1u)) || (gz == Nz - 1 && (faces >> 5 & 1u));
}

// cell coordinates of the CTA's cells, decoded once per CTA (64-bit divisions) into shared memory
__device__ __forceinline__ void decode_cells(const XferGeom &g, int64_t cell0, int64_t ncells, int cpb, int *cc)
{
  for (int c = threadIdx.x; c < cpb; c += blockDim.x) {
    const int64_t cell = cell0 + c;
    if (cell < ncells) {
      cc[3 * c] = (int)(cell % g.ncx);
      cc[3 * c + 1] = (int)((cell / g.ncx) % g.ncy);
      cc[3 * c + 2] = g.ccz_lo + (int)(cell / ((int64_t)g.ncx * g.ncy));
    } else {
      cc[3 * c] = -1; cc[3 * c + 1] = 0; cc[3 * c + 2] = 0;
    }
  }
}

// one CTA handles CPB consecutive coarse cells (x fastest).  TNC, TNF: compile-time 1-D sizes (0 = read them from g):
// the contractions unroll and the index arithmetic becomes multiply-shift
template <int CPB, int TNC, int TNF>
__global__ void __launch_bounds__(256) k_prolongate(const XferGeom g, const double *__restrict__ P1d, double *dst, const double *__restrict__ src)
{
  extern __shared__ double sm[];
  const int NC = TNC ? TNC : g.NC, NF = TNF ? TNF : g.NF;
  const int nvc = NC * NC * NC, nt1 = NC * NC * NF, nt2 = NC * NF * NF;
  double *sP = sm;                       // NC*NF
  double *vc = sP + NC * NF;             // CPB * nvc
  double *t1 = vc + CPB * nvc;           // CPB * nt1
  double *t2 = t1 + CPB * nt1;           // CPB * nt2
  int *cc = (int *)(t2 + CPB * nt2);     // 3 * CPB cell coordinates
  const int tid = threadIdx.x, nth = blockDim.x;
  const int64_t ncells = (int64_t)g.ncx * g.ncy * (g.ccz_hi - g.ccz_lo);
  const int64_t cell0 = (int64_t)blockIdx.x * CPB;

  for (int i = tid; i < NC * NF; i += nth) sP[i] = P1d[i];
  decode_cells(g, cell0, ncells, CPB, cc);
  __syncthreads();
  // gather coarse cell values
  for (int w = tid; w < CPB * nvc; w += nth) {
    const int c = w / nvc, i = w % nvc;
    double v = 0.0;
    const int cx = cc[3 * c], cy = cc[3 * c + 1], cz = cc[3 * c + 2];
    if (cx >= 0) {
      const int ix = i % NC, iy = (i / NC) % NC, iz = i / (NC * NC);
      const int gx = cx * g.pc + ix, gy = cy * g.pc + iy, gz = cz * g.pc + iz;
      // h: constrained coarse dofs read as 0 (dof_indices_coarse == invalid, :170-173);
      // p: read unmasked (:115-121)
      if (g.kind == 1 || !on_dirichlet(gx, gy, gz, g.Ncx, g.Ncy, g.Ncz, g.faces))
        v = src[((int64_t)(gz - g.c_z0) * g.Ncy + gy) * g.Ncx + gx];
    }
    vc[w] = v;
  }
  __syncthreads();
  // x: t1[z][y][xf]
  for (int w = tid; w < CPB * nt1; w += nth) {
    const int c = w / nt1, r = w % nt1;
    const int xf = r % NF, zy = r / NF;
    const double *in = vc + c * nvc + zy * NC;
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < NC; ++k) s += sP[k * NF + xf] * in[k];
    t1[w] = s;
  }
  __syncthreads();
  // y: t2[z][yf][xf]
  for (int w = tid; w < CPB * nt2; w += nth) {
    const int c = w / nt2, r = w % nt2;
    const int xf = r % NF, yf = (r / NF) % NF, z = r / (NF * NF);
    const double *in = t1 + c * nt1 + z * NC * NF + xf;
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < NC; ++k) s += sP[k * NF + yf] * in[k * NF];
    t2[w] = s;
  }
  __syncthreads();
  // z + owner write: dst += value on owned, unconstrained fine dofs
  const int nf3 = NF * NF * NF;
  for (int w = tid; w < CPB * nf3; w += nth) {
    const int c = w / nf3, r = w % nf3;
    const int cx = cc[3 * c], cy = cc[3 * c + 1], cz = cc[3 * c + 2];
    if (cx < 0) continue;
    const int xf = r % NF, yf = (r / NF) % NF, zf = r / (NF * NF);
    if ((xf == NF - 1 && cx != g.ncx - 1) || (yf == NF - 1 && cy != g.ncy - 1) || (zf == NF - 1 && cz != g.ncz - 1)) continue;
    const int gx = cx * g.fstep + xf, gy = cy * g.fstep + yf, gz = cz * g.fstep + zf;
    if (gz < g.f_zown_lo || gz >= g.f_zown_hi) continue;
    if (on_dirichlet(gx, gy, gz, g.Nfx, g.Nfy, g.Nfz, g.faces)) continue; // weight 0 / masked (:1346-1349, p :306-324)
    const double *in = t2 + c * nt2 + yf * NF + xf;
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < NC; ++k) s += sP[k * NF + zf] * in[k * NF * NF];
    dst[((int64_t)(gz - g.f_z0) * g.Nfy + gy) * g.Nfx + gx] += s;
  }
}

// pass 1 of the restriction: cell-local (P^T x P^T x P^T) applied to the patch's owned fine dofs
template <int CPB, int TNC, int TNF>
__global__ void __launch_bounds__(256) k_restrict_cells(const XferGeom g, const double *__restrict__ P1d, double *scratch, const double *__restrict__ src)
{
  extern __shared__ double sm[];
  const int NC = TNC ? TNC : g.NC, NF = TNF ? TNF : g.NF;
  const int nf3 = NF * NF * NF, nt1 = NC * NF * NF, nt2 = NC * NC * NF, nvc = NC * NC * NC;
  double *sP = sm;                 // NC*NF
  double *vf = sP + NC * NF;       // CPB*nf3
  double *t1 = vf + CPB * nf3;     // CPB*nt1   [zc][yf][xf]
  double *t2 = t1 + CPB * nt1;     // CPB*nt2   [zc][yc][xf]
  int *cc = (int *)(t2 + CPB * nt2);
  const int tid = threadIdx.x, nth = blockDim.x;
  const int64_t ncells = (int64_t)g.ncx * g.ncy * (g.ccz_hi - g.ccz_lo);
  const int64_t cell0 = (int64_t)blockIdx.x * CPB;
  for (int i = tid; i < NC * NF; i += nth) sP[i] = P1d[i];
  decode_cells(g, cell0, ncells, CPB, cc);
  __syncthreads();
  for (int w = tid; w < CPB * nf3; w += nth) {
    const int c = w / nf3, r = w % nf3;
    double v = 0.0;
    const int cx = cc[3 * c], cy = cc[3 * c + 1], cz = cc[3 * c + 2];
    if (cx >= 0) {
      const int xf = r % NF, yf = (r / NF) % NF, zf = r / (NF * NF);
      const bool owned = !((xf == NF - 1 && cx != g.ncx - 1) || (yf == NF - 1 && cy != g.ncy - 1) || (zf == NF - 1 && cz != g.ncz - 1));
      const int gx = cx * g.fstep + xf, gy = cy * g.fstep + yf, gz = cz * g.fstep + zf;
      if (owned && !on_dirichlet(gx, gy, gz, g.Nfx, g.Nfy, g.Nfz, g.faces))
        v = src[((int64_t)(gz - g.f_z0) * g.Nfy + gy) * g.Nfx + gx];
    }
    vf[w] = v;
  }
  __syncthreads();
  // z: t1[zc][yf][xf] = sum_zf P[zc][zf] vf[zf][yf][xf]
  for (int w = tid; w < CPB * nt1; w += nth) {
    const int c = w / nt1, r = w % nt1;
    const int xy = r % (NF * NF), zc = r / (NF * NF);
    const double *in = vf + c * nf3 + xy;
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < NF; ++k) s += sP[zc * NF + k] * in[k * NF * NF];
    t1[w] = s;
  }
  __syncthreads();
  // y: t2[zc][yc][xf]
  for (int w = tid; w < CPB * nt2; w += nth) {
    const int c = w / nt2, r = w % nt2;
    const int xf = r % NF, yc = (r / NF) % NC, zc = r / (NF * NC);
    const double *in = t1 + c * nt1 + zc * NF * NF + xf;
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < NF; ++k) s += sP[yc * NF + k] * in[k * NF];
    t2[w] = s;
  }
  __syncthreads();
  // x: scratch[cell][zc][yc][xc]
  for (int w = tid; w < CPB * nvc; w += nth) {
    const int c = w / nvc, r = w % nvc;
    if (cc[3 * c] < 0) continue;
    const int xc = r % NC, zy = r / NC;
    const double *in = t2 + c * nt2 + zy * NF;
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < NF; ++k) s += sP[xc * NF + k] * in[k];
    scratch[(cell0 + c) * nvc + r] = s;
  }
}

// ---- column kernels for compile-time sizes ---------------------------------------------------------------
// Thread = one (xf, yf) column of a coarse cell's fine patch; CPB cells, consecutive in x, per CTA; grid = (x groups, cy,
// local cz): no integer division on the data path and ~20 instructions per fine dof (the generic kernels above: ~400,
// issue bound).  Lane order xf fastest, then cell: a warp touches consecutive fine dofs of one row.

// fine dof (xf, yf, zf) of cell (cx, cy, cz) belongs to this patch (low-side ownership) and is not constrained
__device__ __forceinline__ bool xfer_col_owned(const XferGeom &g, int NF, int cx, int cy, int xf, int yf, int gx, int gy)
{
  if ((xf == NF - 1 && cx != g.ncx - 1) || (yf == NF - 1 && cy != g.ncy - 1)) return false;
  return !((gx == 0 && (g.faces & 1u)) || (gx == g.Nfx - 1 && (g.faces >> 1 & 1u)) || (gy == 0 && (g.faces >> 2 & 1u)) ||
           (gy == g.Nfy - 1 && (g.faces >> 3 & 1u)));
}

template <int CPB, int NC, int NF>
__global__ void __launch_bounds__(CPB *NF *NF) k_prolongate_col(const XferGeom g, const double *__restrict__ P1d, double *dst,
                                                               const double *__restrict__ src)
{
  __shared__ double sP[NC * NF];
  __shared__ double vc[CPB][NC * NC * NC];
  const int tid = threadIdx.x;
  const int xf = tid % NF, c = (tid / NF) % CPB, yf = tid / (NF * CPB);
  const int cx0 = blockIdx.x * CPB, cy = blockIdx.y, cz = g.ccz_lo + blockIdx.z;
  for (int i = tid; i < NC * NF; i += CPB * NF * NF) sP[i] = P1d[i];
  // coarse values of the CTA's cells (h: constrained coarse dofs read as 0, :170-173; p: read unmasked, :115-121)
  for (int w = tid; w < CPB * NC * NC * NC; w += CPB * NF * NF) {
    const int cc = w / (NC * NC * NC), i = w % (NC * NC * NC);
    const int ix = i % NC, iy = (i / NC) % NC, iz = i / (NC * NC);
    const int cx = cx0 + cc;
    double v = 0.0;
    if (cx < g.ncx) {
      const int gx = cx * g.pc + ix, gy = cy * g.pc + iy, gz = cz * g.pc + iz;
      if (g.kind == 1 || !on_dirichlet(gx, gy, gz, g.Ncx, g.Ncy, g.Ncz, g.faces))
        v = src[((int64_t)(gz - g.c_z0) * g.Ncy + gy) * g.Ncx + gx];
    }
    vc[cc][i] = v;
  }
  __syncthreads();
  const int cx = cx0 + c;
  if (cx >= g.ncx) return;
  const int gx = cx * g.fstep + xf, gy = cy * g.fstep + yf;
  if (!xfer_col_owned(g, NF, cx, cy, xf, yf, gx, gy)) return;
  // x and y contractions for this column: t[iz] = sum_iy P[iy][yf] sum_ix P[ix][xf] c[iz][iy][ix]
  double t[NC];
#pragma unroll
  for (int iz = 0; iz < NC; ++iz) {
    double s = 0.0;
#pragma unroll
    for (int iy = 0; iy < NC; ++iy) {
      double r = 0.0;
#pragma unroll
      for (int ix = 0; ix < NC; ++ix) r = fma(sP[ix * NF + xf], vc[c][(iz * NC + iy) * NC + ix], r);
      s = fma(sP[iy * NF + yf], r, s);
    }
    t[iz] = s;
  }
  // z contraction + owner update: dst += value on owned, unconstrained fine dofs (weight 0 / masked otherwise, :1346-1349)
  double *col = dst + ((int64_t)(cz * g.fstep - g.f_z0) * g.Nfy + gy) * g.Nfx + gx;
  const int64_t plane = (int64_t)g.Nfx * g.Nfy;
#pragma unroll
  for (int zf = 0; zf < NF; ++zf) {
    const int gz = cz * g.fstep + zf;
    if (zf == NF - 1 && cz != g.ncz - 1) continue;
    if (gz < g.f_zown_lo || gz >= g.f_zown_hi) continue;
    if ((gz == 0 && (g.faces >> 4 & 1u)) || (gz == g.Nfz - 1 && (g.faces >> 5 & 1u))) continue;
    double s = 0.0;
#pragma unroll
    for (int iz = 0; iz < NC; ++iz) s = fma(sP[iz * NF + zf], t[iz], s);
    col[zf * plane] += s;
  }
}

// pass 1 of the restriction, column form: scratch[cell][zc][yc][xc] = (P^T x P^T x P^T) applied to the patch's owned fine dofs
template <int CPB, int NC, int NF>
__global__ void __launch_bounds__(CPB *NF *NF) k_restrict_cells_col(const XferGeom g, const double *__restrict__ P1d, double *scratch,
                                                                   const double *__restrict__ src)
{
  __shared__ double sP[NC * NF];
  __shared__ double t1[CPB][NC][NF * NF]; // [zc][yf][xf]
  __shared__ double t2[CPB][NC * NC][NF]; // [zc][yc][xf]
  constexpr int NTH = CPB * NF * NF;
  const int tid = threadIdx.x;
  const int xf = tid % NF, c = (tid / NF) % CPB, yf = tid / (NF * CPB);
  const int cx0 = blockIdx.x * CPB, cy = blockIdx.y, cz = g.ccz_lo + blockIdx.z;
  for (int i = tid; i < NC * NF; i += NTH) sP[i] = P1d[i];
  __syncthreads();
  const int cx = cx0 + c;
  {
    // z contraction of this column's owned, unconstrained fine values
    double v[NF];
    const int gx = cx * g.fstep + xf, gy = cy * g.fstep + yf;
    const bool live = cx < g.ncx && xfer_col_owned(g, NF, cx, cy, xf, yf, gx, gy);
    const double *col = src + ((int64_t)(cz * g.fstep - g.f_z0) * g.Nfy + gy) * g.Nfx + gx;
    const int64_t plane = (int64_t)g.Nfx * g.Nfy;
#pragma unroll
    for (int zf = 0; zf < NF; ++zf) {
      const int gz = cz * g.fstep + zf;
      const bool ok = live && !(zf == NF - 1 && cz != g.ncz - 1) &&
                      !((gz == 0 && (g.faces >> 4 & 1u)) || (gz == g.Nfz - 1 && (g.faces >> 5 & 1u)));
      v[zf] = ok ? col[zf * plane] : 0.0;
    }
#pragma unroll
    for (int zc = 0; zc < NC; ++zc) {
      double s = 0.0;
#pragma unroll
      for (int zf = 0; zf < NF; ++zf) s = fma(sP[zc * NF + zf], v[zf], s);
      t1[c][zc][yf * NF + xf] = s;
    }
  }
  __syncthreads();
  // y contraction: t2[zc][yc][xf]
  for (int w = tid; w < CPB * NC * NC * NF; w += NTH) {
    const int x2 = w % NF, cc = (w / NF) % CPB, zy = w / (NF * CPB);
    const int yc = zy % NC, zc = zy / NC;
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < NF; ++k) s = fma(sP[yc * NF + k], t1[cc][zc][k * NF + x2], s);
    t2[cc][zc * NC + yc][x2] = s;
  }
  __syncthreads();
  // x contraction -> scratch[cell][zc][yc][xc], cell numbered x fastest over the rank's coarse cells
  for (int w = tid; w < CPB * NC * NC * NC; w += NTH) {
    const int cc = w / (NC * NC * NC), r = w % (NC * NC * NC);
    if (cx0 + cc >= g.ncx) continue;
    const int xc = r % NC, zy = r / NC;
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < NF; ++k) s = fma(sP[xc * NF + k], t2[cc][zy][k], s);
    const int64_t cell = ((int64_t)blockIdx.z * g.ncy + cy) * g.ncx + cx0 + cc;
    scratch[cell * (NC * NC * NC) + r] = s;
  }
}

// pass 2: every coarse dof of the handled layers gathers its <= 8 cell-local contributions
__global__ void k_restrict_gather(const XferGeom g, double *dst, const double *__restrict__ scratch)
{
  const int NC = g.NC, p = g.pc;
  const int nvc = NC * NC * NC;
  const int z_lo = g.ccz_lo * p, z_hi = g.ccz_hi * p; // planes z_lo..z_hi inclusive receive contributions
  const int64_t plane = (int64_t)g.Ncx * g.Ncy;
  const int64_t total = plane * (z_hi - z_lo + 1);
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int gx = (int)(idx % g.Ncx), gy = (int)((idx / g.Ncx) % g.Ncy), gz = z_lo + (int)(idx / plane);
    if (gz - g.c_z0 < 0 || gz - g.c_z0 >= g.c_nzl) continue;
    if (on_dirichlet(gx, gy, gz, g.Ncx, g.Ncy, g.Ncz, g.faces)) continue; // constrained coarse dofs are skipped (:674-682)
    double s = 0.0;
    for (int sz = 0; sz < 2; ++sz) {
      int cz = gz / p, iz = gz % p;
      if (sz == 1) { if (iz != 0) continue; cz -= 1; iz = p; }
      if (cz < g.ccz_lo || cz >= g.ccz_hi) continue;
      for (int sy = 0; sy < 2; ++sy) {
        int cy = gy / p, iy = gy % p;
        if (sy == 1) { if (iy != 0) continue; cy -= 1; iy = p; }
        if (cy < 0 || cy >= g.ncy) continue;
        for (int sx = 0; sx < 2; ++sx) {
          int cx = gx / p, ix = gx % p;
          if (sx == 1) { if (ix != 0) continue; cx -= 1; ix = p; }
          if (cx < 0 || cx >= g.ncx) continue;
          const int64_t cell = ((int64_t)(cz - g.ccz_lo) * g.ncy + cy) * g.ncx + cx;
          s += scratch[cell * nvc + (iz * NC + iy) * NC + ix];
        }
      }
    }
    dst[(int64_t)(gz - g.c_z0) * plane + (int64_t)gy * g.Ncx + gx] += s;
  }
}

int make_geom(int kind, const pmgk_level *c, const pmgk_level *f, XferGeom *g)
{
  g->kind = kind; g->pc = c->degree; g->pf = f->degree;
  if (kind == 0) {
    if (c->degree != f->degree || f->nx != 2 * c->nx || f->ny != 2 * c->ny || f->nz != 2 * c->nz) return PMG_ERR_ARG;
    g->NC = c->degree + 1; g->NF = 2 * c->degree + 1; g->fstep = 2 * c->degree;
    if (f->cz_lo != 2 * c->cz_lo || f->cz_hi != 2 * c->cz_hi) return PMG_ERR_ARG;
  } else {
    if (c->degree >= f->degree || f->nx != c->nx || f->ny != c->ny || f->nz != c->nz) return PMG_ERR_ARG;
    g->NC = c->degree + 1; g->NF = f->degree + 1; g->fstep = f->degree;
    if (f->cz_lo != c->cz_lo || f->cz_hi != c->cz_hi) return PMG_ERR_ARG;
  }
  if (c->faces != f->faces) return PMG_ERR_ARG;
  g->ncx = c->nx; g->ncy = c->ny; g->ncz = c->nz;
  g->Ncx = c->Nx; g->Ncy = c->Ny; g->Ncz = c->Nz;
  g->Nfx = f->Nx; g->Nfy = f->Ny; g->Nfz = f->Nz;
  g->faces = c->faces;
  g->c_z0 = c->z0; g->f_z0 = f->z0; g->c_nzl = c->nzl; g->f_nzl = f->nzl;
  g->ccz_lo = c->cz_lo; g->ccz_hi = c->cz_hi;
  g->c_zown_lo = c->z_own_lo; g->c_zown_hi = c->z_own_hi;
  g->f_zown_lo = f->z_own_lo; g->f_zown_hi = f->z_own_hi;
  return 0;
}

int pick_cpb(int NC, int NF)
{
  const int per_cell = NF * NF * NF + NC * NF * NF + NC * NC * NF; // doubles of shared memory per cell (restriction)
  int cpb = 1;
  while (cpb < 8 && (cpb * 2) * per_cell * 8 <= 40 * 1024 && (cpb * NF * NF * NF) < 1024) cpb *= 2;
  return cpb;
}

} // namespace

// 2-D levels: csrc/pmg_dim2.cu
int pmg_dim2_prolongate(int kind, const pmgk_level *c, const pmgk_level *f, const double *P1d, double *dst, const double *src, cudaStream_t s);
int pmg_dim2_restrict(int kind, const pmgk_level *c, const pmgk_level *f, const double *P1d, double *dst, const double *src, cudaStream_t s);

extern "C" int64_t pmgk_restrict_scratch_doubles(int kind, const pmgk_level *coarse, const pmgk_level *fine)
{
  (void)kind; (void)fine;
  if (coarse->dim == 2) return 0; /* gather per coarse dof, no cell-local scratch */
  const int NC = coarse->degree + 1;
  return (int64_t)coarse->nx * coarse->ny * (coarse->cz_hi - coarse->cz_lo) * NC * NC * NC;
}

template <int CPB, int TNC, int TNF>
static int launch_prolongate(const XferGeom &g, const double *P1d, double *dst, const double *src, cudaStream_t s)
{
  const int NC = g.NC, NF = g.NF;
  const size_t smem = sizeof(double) * (NC * NF + CPB * (NC * NC * NC + NC * NC * NF + NC * NF * NF)) + sizeof(int) * 3 * CPB;
  const int64_t ncells = (int64_t)g.ncx * g.ncy * (g.ccz_hi - g.ccz_lo);
  if (smem > 48 * 1024) PMG_CUDA_CHECK(cudaFuncSetAttribute(k_prolongate<CPB, TNC, TNF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_prolongate<CPB, TNC, TNF><<<(unsigned)((ncells + CPB - 1) / CPB), 256, smem, s>>>(g, P1d, dst, src);
  PMG_CUDA_CHECK(cudaGetLastError());
  pmg_count_launch(1);
  return 0;
}

template <int CPB, int TNC, int TNF>
static int launch_restrict(const XferGeom &g, const double *P1d, double *scratch, const double *src, cudaStream_t s)
{
  const int NC = g.NC, NF = g.NF;
  const size_t smem = sizeof(double) * (NC * NF + CPB * (NF * NF * NF + NC * NF * NF + NC * NC * NF)) + sizeof(int) * 3 * CPB;
  const int64_t ncells = (int64_t)g.ncx * g.ncy * (g.ccz_hi - g.ccz_lo);
  if (smem > 48 * 1024) PMG_CUDA_CHECK(cudaFuncSetAttribute(k_restrict_cells<CPB, TNC, TNF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_restrict_cells<CPB, TNC, TNF><<<(unsigned)((ncells + CPB - 1) / CPB), 256, smem, s>>>(g, P1d, scratch, src);
  PMG_CUDA_CHECK(cudaGetLastError());
  pmg_count_launch(1);
  return 0;
}

template <int CPB, int NC, int NF>
static int launch_prolongate_col(const XferGeom &g, const double *P1d, double *dst, const double *src, cudaStream_t s)
{
  const dim3 grid((unsigned)((g.ncx + CPB - 1) / CPB), (unsigned)g.ncy, (unsigned)(g.ccz_hi - g.ccz_lo));
  k_prolongate_col<CPB, NC, NF><<<grid, CPB * NF * NF, 0, s>>>(g, P1d, dst, src);
  PMG_CUDA_CHECK(cudaGetLastError());
  pmg_count_launch(1);
  return 0;
}

template <int CPB, int NC, int NF>
static int launch_restrict_col(const XferGeom &g, const double *P1d, double *scratch, const double *src, cudaStream_t s)
{
  const dim3 grid((unsigned)((g.ncx + CPB - 1) / CPB), (unsigned)g.ncy, (unsigned)(g.ccz_hi - g.ccz_lo));
  k_restrict_cells_col<CPB, NC, NF><<<grid, CPB * NF * NF, 0, s>>>(g, P1d, scratch, src);
  PMG_CUDA_CHECK(cudaGetLastError());
  pmg_count_launch(1);
  return 0;
}

// the (NC, NF) pairs of the drivers' hierarchies get compile-time sizes: h-transfer of Q1..Q4 and the p-transfers
// p -> p/2 (Q2->Q1 and Q4->Q2 share their sizes with the h-transfers of Q1 and Q2); everything else runs the generic code
#define PMG_XFER_DISPATCH(LAUNCH, ...)                                                        \
  do {                                                                                        \
    const int cpb = pick_cpb(g.NC, g.NF);                                                     \
    if (g.NC == 2 && g.NF == 3 && cpb == 8) return LAUNCH<8, 2, 3>(__VA_ARGS__);              \
    if (g.NC == 3 && g.NF == 5 && cpb == 8) return LAUNCH<8, 3, 5>(__VA_ARGS__);              \
    if (g.NC == 2 && g.NF == 4 && cpb == 8) return LAUNCH<8, 2, 4>(__VA_ARGS__);              \
    if (g.NC == 4 && g.NF == 7 && cpb == 4) return LAUNCH<4, 4, 7>(__VA_ARGS__);              \
    if (g.NC == 5 && g.NF == 9 && cpb == 2) return LAUNCH<2, 5, 9>(__VA_ARGS__);              \
    switch (cpb) {                                                                            \
      case 8: return LAUNCH<8, 0, 0>(__VA_ARGS__);                                            \
      case 4: return LAUNCH<4, 0, 0>(__VA_ARGS__);                                            \
      case 2: return LAUNCH<2, 0, 0>(__VA_ARGS__);                                            \
      default: return LAUNCH<1, 0, 0>(__VA_ARGS__);                                           \
    }                                                                                         \
  } while (0)

// column kernels (grid.y, grid.z <= 65535) for the sizes of the drivers' hierarchies; generic kernels otherwise
#define PMG_XFER_COL(LAUNCH, ...)                                                            \
  do {                                                                                        \
    if (g.ncy <= 65535 && g.ccz_hi - g.ccz_lo <= 65535) {                                     \
      if (g.NC == 2 && g.NF == 3) return LAUNCH<16, 2, 3>(__VA_ARGS__);                       \
      if (g.NC == 3 && g.NF == 5) return LAUNCH<8, 3, 5>(__VA_ARGS__);                        \
      if (g.NC == 2 && g.NF == 4) return LAUNCH<8, 2, 4>(__VA_ARGS__);                        \
      if (g.NC == 4 && g.NF == 7) return LAUNCH<4, 4, 7>(__VA_ARGS__);                        \
      if (g.NC == 5 && g.NF == 9) return LAUNCH<2, 5, 9>(__VA_ARGS__);                        \
    }                                                                                         \
  } while (0)

static int dispatch_prolongate(const XferGeom &g, const double *P1d, double *dst, const double *src, cudaStream_t s)
{
  PMG_XFER_COL(launch_prolongate_col, g, P1d, dst, src, s);
  PMG_XFER_DISPATCH(launch_prolongate, g, P1d, dst, src, s);
}

static int dispatch_restrict(const XferGeom &g, const double *P1d, double *scratch, const double *src, cudaStream_t s)
{
  PMG_XFER_COL(launch_restrict_col, g, P1d, scratch, src, s);
  PMG_XFER_DISPATCH(launch_restrict, g, P1d, scratch, src, s);
}

extern "C" int pmgk_prolongate_and_add(int kind, const pmgk_level *coarse, const pmgk_level *fine, const double *P1d,
                                       double *dst_fine, const double *src_coarse, void *stream)
{
  if (coarse->dim == 2) return pmg_dim2_prolongate(kind, coarse, fine, P1d, dst_fine, src_coarse, (cudaStream_t)stream);
  XferGeom g;
  const int rc = make_geom(kind, coarse, fine, &g);
  if (rc) return rc;
  if (g.ccz_hi <= g.ccz_lo) return 0;
  return dispatch_prolongate(g, P1d, dst_fine, src_coarse, (cudaStream_t)stream);
}

extern "C" int pmgk_restrict_and_add(int kind, const pmgk_level *coarse, const pmgk_level *fine, const double *P1d,
                                     double *dst_coarse, const double *src_fine, double *scratch, void *stream)
{
  if (coarse->dim == 2) return pmg_dim2_restrict(kind, coarse, fine, P1d, dst_coarse, src_fine, (cudaStream_t)stream);
  XferGeom g;
  int rc = make_geom(kind, coarse, fine, &g);
  if (rc) return rc;
  if (g.ccz_hi <= g.ccz_lo) return 0;
  rc = dispatch_restrict(g, P1d, scratch, src_fine, (cudaStream_t)stream);
  if (rc) return rc;
  const int64_t total = (int64_t)g.Ncx * g.Ncy * ((g.ccz_hi - g.ccz_lo) * g.pc + 1);
  int64_t nb = (total + 255) / 256;
  const int cap = pmgk_device_sm_count() * 8;
  if (nb > cap) nb = cap;
  k_restrict_gather<<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(g, dst_coarse, scratch);
  PMG_CUDA_CHECK(cudaGetLastError());
  pmg_count_launch(1);
  return 0;
}

// ---- inverse diagonal ------------------------------------------------------
namespace {

struct DiagGeom {
  int P, Nx, Ny, Nz, z0, z_lo, z_hi;
  unsigned faces;
};

__device__ __forceinline__ int pos_type(int g, int N, int P) { return (g == 0) ? P : (g == N - 1) ? P + 1 : g % P; }

// one CTA row per (plane, y-row) pair: no integer division per element
__global__ void k_dinv(const DiagGeom g, const double *__restrict__ tab, double f, const double *b, double *out)
{
  const int T = g.P + 2;
  const int nrows = g.Ny * (g.z_hi - g.z_lo);
  for (int row = blockIdx.x; row < nrows; row += gridDim.x) {
    const int gy = row % g.Ny, gz = g.z_lo + row / g.Ny;
    const bool dyz = (gy == 0 && (g.faces >> 2 & 1u)) || (gy == g.Ny - 1 && (g.faces >> 3 & 1u)) ||
                     (gz == 0 && (g.faces >> 4 & 1u)) || (gz == g.Nz - 1 && (g.faces >> 5 & 1u));
    const int tyz = T * (pos_type(gy, g.Ny, g.P) + T * pos_type(gz, g.Nz, g.P));
    const int64_t l0 = ((int64_t)(gz - g.z0) * g.Ny + gy) * g.Nx;
    for (int gx = threadIdx.x; gx < g.Nx; gx += blockDim.x) {
      const bool dir = dyz || (gx == 0 && (g.faces & 1u)) || (gx == g.Nx - 1 && (g.faces >> 1 & 1u));
      const int tx = (gx == 0) ? g.P : (gx == g.Nx - 1) ? g.P + 1 : gx % g.P;
      const double d = dir ? 1.0 : tab[tx + tyz];
      out[l0 + gx] = b ? f * d * b[l0 + gx] : d;
    }
  }
}

__global__ void k_dinv_vec(const double *__restrict__ dinv, double f, const double *__restrict__ b, double *out, int64_t n)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = f * dinv[i] * b[i];
}

int launch_dinv(const pmgk_level *lv, double f, const double *b, double *out, cudaStream_t s)
{
  DiagGeom g;
  g.P = lv->degree; g.Nx = lv->Nx; g.Ny = lv->Ny; g.Nz = lv->Nz; g.z0 = lv->z0;
  g.z_lo = lv->z0; g.z_hi = lv->z0 + lv->nzl; // all stored planes (ghost planes get consistent values)
  g.faces = lv->faces;
  const int64_t total = (int64_t)lv->Nx * lv->Ny * lv->nzl;
  int64_t nb = (total + 255) / 256;
  const int cap = pmgk_device_sm_count() * 8;
  if (nb > cap) nb = cap;
  if (nb < 1) return 0;
  if (b && lv->dinv_vec) k_dinv_vec<<<(unsigned)nb, 256, 0, s>>>(lv->dinv_vec, f, b, out, total);
  else {
    int64_t rows = (int64_t)lv->Ny * lv->nzl;
    if (rows > (int64_t)cap * 4) rows = (int64_t)cap * 4;
    k_dinv<<<(unsigned)rows, lv->Nx >= 192 ? 256 : (lv->Nx >= 96 ? 128 : (lv->Nx >= 48 ? 64 : 32)), 0, s>>>(g, lv->dinv_tab, f, b, out);
  }
  PMG_CUDA_CHECK(cudaGetLastError());
  pmg_count_launch(1);
  return 0;
}

} // namespace

extern "C" int pmgk_fill_dinv(const pmgk_level *lv, double *dinv, void *stream)
{
  return launch_dinv(lv, 1.0, nullptr, dinv, (cudaStream_t)stream);
}

extern "C" int pmgk_scale_dinv(const pmgk_level *lv, double f, const double *b, double *out, void *stream)
{
  return launch_dinv(lv, f, b, out, (cudaStream_t)stream);
}
