// tests/emu/emu_var.cpp -- host emulator of the variable-coefficient tile program (TEST INFRASTRUCTURE).
//
// Compiles the product's kernel source (csrc/pmg_apply_var.h) for the CPU and runs every CTA's phases thread by
// thread; the coefficient grid and the inverse diagonal are produced by the same PMG_HD functions the set-up kernels of
// csrc/pmg_apply_var.cu call per element.  Built only by tests/; the product has no CPU path.
#include <cstring>
#include <vector>
#include "pmg_apply_var.h"

extern "C" void pmg_fe_shape_tables(int p, double *Sq, double *Dco, double *G, double *gq, double *gw);

template <class Tile>
struct HostExecV {
  std::vector<typename Tile::ThreadState> st;
  template <class F> void for_each_thread(F f) { for (int t = 0; t < Tile::NT; ++t) f(t, st[t]); }
  void sync() {}
};

struct VarArgs {
  int nx, ny, nz; unsigned faces; int z0, nzl, cz_lo, cz_hi, z_own_lo, z_own_hi, n_chunks;
  const double *h; int mode; const double *u, *b, *xold; double *out; double f1, f2; const double *dinv_vec; double *dinv_out;
};

template <int P, int BX, int BY>
static void go(const VarArgs &a)
{
  using Tile = PmgVarTile<P, BX, BY>;
  constexpr int N1 = P + 1;
  PmgVarParams<P> p;
  std::memset(&p, 0, sizeof(p));
  p.nx = a.nx; p.ny = a.ny; p.nz = a.nz;
  p.Nx = a.nx * P + 1; p.Ny = a.ny * P + 1; p.Nz = a.nz * P + 1;
  p.faces = a.faces; p.z0 = a.z0; p.nzl = a.nzl; p.cz_lo = a.cz_lo; p.cz_hi = a.cz_hi;
  p.z_own_lo = a.z_own_lo; p.z_own_hi = a.z_own_hi;
  double G[N1 * N1], gq[N1], gw[N1], S2[N1 * N1], G2[N1 * N1];
  pmg_fe_shape_tables(P, p.S, p.D, G, gq, gw);
  for (int i = 0; i < N1 * N1; ++i) { S2[i] = p.S[i] * p.S[i]; G2[i] = G[i] * G[i]; }
  p.c[0] = a.h[1] * a.h[2] / a.h[0]; p.c[1] = a.h[0] * a.h[2] / a.h[1]; p.c[2] = a.h[0] * a.h[1] / a.h[2];
  // coefficient grid of the locally stored cell layers
  const int coef_cz0 = a.z0 / P;
  const int Qx = a.nx * N1, Qy = a.ny * N1, nqz = (a.cz_hi - coef_cz0) * N1;
  std::vector<double> coef((size_t)Qx * Qy * nqz);
  for (int qz = 0; qz < nqz; ++qz)
    for (int qy = 0; qy < Qy; ++qy)
      for (int qx = 0; qx < Qx; ++qx)
        coef[((size_t)qz * Qy + qy) * Qx + qx] = pmg_var_coef_c5<P>(qx, qy, coef_cz0 * N1 + qz, gq, gw, a.h);
  p.coef = coef.data(); p.coef_cz0 = coef_cz0;
  // inverse diagonal on all stored planes, as k_var_dinv computes it
  std::vector<double> dinv;
  if (a.dinv_out || (a.mode >= 2 && !a.dinv_vec)) {
    dinv.resize((size_t)p.Nx * p.Ny * a.nzl);
    for (int l = 0; l < a.nzl; ++l)
      for (int gy = 0; gy < p.Ny; ++gy)
        for (int gx = 0; gx < p.Nx; ++gx) {
          const int gz = a.z0 + l;
          const bool dir = (gx == 0 && (a.faces & 1u)) || (gx == p.Nx - 1 && (a.faces >> 1 & 1u)) ||
                           (gy == 0 && (a.faces >> 2 & 1u)) || (gy == p.Ny - 1 && (a.faces >> 3 & 1u)) ||
                           (gz == 0 && (a.faces >> 4 & 1u)) || (gz == p.Nz - 1 && (a.faces >> 5 & 1u));
          const double d = dir ? 1.0 : pmg_var_diag_entry<P>(gx, gy, gz, a.nx, a.ny, coef_cz0, a.cz_hi, S2, G2, p.c, coef.data(), coef_cz0);
          dinv[((size_t)l * p.Ny + gy) * p.Nx + gx] = 1.0 / d;
        }
    if (a.dinv_out) { std::memcpy(a.dinv_out, dinv.data(), dinv.size() * sizeof(double)); }
  }
  if (!a.out) return;
  p.mode = a.mode; p.u = a.u; p.b = a.b; p.xold = a.xold; p.out = a.out; p.f1 = a.f1; p.f2 = a.f2;
  p.dinv_vec = a.dinv_vec ? a.dinv_vec : (dinv.empty() ? nullptr : dinv.data());
  p.dinv_tab = nullptr;
  p.tiles_x = (p.nx + BX - 1) / BX;
  p.tiles_y = (p.ny + BY - 1) / BY;
  const int layers = p.cz_hi - p.cz_lo;
  int n_chunks = a.n_chunks < 1 ? 1 : a.n_chunks;
  if (n_chunks > layers) n_chunks = layers;
  p.layers_per_chunk = (layers + n_chunks - 1) / n_chunks;
  p.n_chunks = (layers + p.layers_per_chunk - 1) / p.layers_per_chunk;
  std::vector<double> smem(Tile::SMEM_DOUBLES);
  for (int chunk = 0; chunk < p.n_chunks; ++chunk)
    for (int ty = 0; ty < p.tiles_y; ++ty)
      for (int tx = 0; tx < p.tiles_x; ++tx) {
        HostExecV<Tile> ex;
        ex.st.resize(Tile::NT);
        for (auto &v : smem) v = 1e300; // poison shared memory so stale reads show up
        Tile::run(p, ex, smem.data(), tx, ty, chunk);
      }
}

// small_tiles != 0 selects deliberately tiny tiles so that small meshes exercise many tiles; otherwise the tiles
// pmg_apply_var.cu launches.  out == NULL: only the inverse diagonal (dinv_out, all stored planes) is produced.
extern "C" int emu_var(int degree, int small_tiles, int nx, int ny, int nz, unsigned faces, int z0, int nzl, int cz_lo,
                       int cz_hi, int z_own_lo, int z_own_hi, int n_chunks, const double *h, int mode, const double *u,
                       const double *b, const double *xold, double *out, double f1, double f2, const double *dinv_vec,
                       double *dinv_out)
{
  const VarArgs a = {nx, ny, nz, faces, z0, nzl, cz_lo, cz_hi, z_own_lo, z_own_hi, n_chunks, h, mode, u, b, xold, out, f1, f2, dinv_vec, dinv_out};
  if (small_tiles) {
    switch (degree) {
      case 1: go<1, 3, 2>(a); return 0;
      case 2: go<2, 2, 3>(a); return 0;
      case 3: go<3, 2, 2>(a); return 0;
      case 4: go<4, 3, 2>(a); return 0;
      case 5: go<5, 2, 2>(a); return 0;
      case 6: go<6, 2, 1>(a); return 0;
      case 7: go<7, 1, 2>(a); return 0;
      case 8: go<8, 2, 2>(a); return 0;
      case 9: go<9, 1, 2>(a); return 0;
    }
    return -3;
  }
  switch (degree) {
#define PMG_VAR_CASE(P, BX, BY, MINB) case P: go<P, BX, BY>(a); return 0;
#include "pmg_apply_var_tiles.inc"
#undef PMG_VAR_CASE
  }
  return -3;
}
