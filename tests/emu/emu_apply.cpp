// tests/emu/emu_apply.cpp -- host emulator of the CUDA tile programs (TEST INFRASTRUCTURE).
//
// Compiles the product's kernel source (csrc/pmg_apply_tile.h) for the CPU and runs every
// CTA's phases thread by thread, so the CPU test-suite can check the kernels' index logic,
// halo/ownership rules and fused epilogues against the oracle without a GPU.  It is built only
// by tests/ and is not part of libpmg.so: the product has no CPU path.
#include <vector>
#include <cstring>
#include "pmg_apply_tile.h"

template <class Tile>
struct HostExec {
  std::vector<typename Tile::ThreadState> st;
  template <class F> void for_each_thread(F f) { for (int t = 0; t < Tile::NT; ++t) f(t, st[t]); }
  void sync() {}
};

template <int P, int BX, int BY>
static void run_all(PmgApplyParams<P> p, int n_chunks)
{
  using Tile = PmgApplyTile<P, BX, BY>;
  p.tiles_x = (p.nx + BX - 1) / BX;
  p.tiles_y = (p.ny + BY - 1) / BY;
  const int layers = p.cz_hi - p.cz_lo;
  if (n_chunks < 1) n_chunks = 1;
  if (n_chunks > layers) n_chunks = layers;
  p.layers_per_chunk = (layers + n_chunks - 1) / n_chunks;
  p.n_chunks = (layers + p.layers_per_chunk - 1) / p.layers_per_chunk;
  std::vector<double> smem(Tile::SMEM_DOUBLES);
  for (int chunk = 0; chunk < p.n_chunks; ++chunk)
    for (int ty = 0; ty < p.tiles_y; ++ty)
      for (int tx = 0; tx < p.tiles_x; ++tx) {
        HostExec<Tile> ex;
        ex.st.resize(Tile::NT);
        // poison shared memory so stale reads show up
        for (auto &v : smem) v = 1e300;
        Tile::run(p, ex, smem.data(), tx, ty, chunk);
      }
}

template <int P, int BX, int BY>
static void go(int nx, int ny, int nz, unsigned faces, int z0, int nzl, int cz_lo, int cz_hi, int z_own_lo,
               int z_own_hi, int n_chunks, const double *S, const double *lam, const double *h, int mode,
               const double *u, const double *b, const double *xold, double *out, double f1, double f2,
               const double *dinv_vec, const double *dinv_tab)
{
  PmgApplyParams<P> p;
  std::memset(&p, 0, sizeof(p));
  p.nx = nx; p.ny = ny; p.nz = nz;
  p.Nx = nx * P + 1; p.Ny = ny * P + 1; p.Nz = nz * P + 1;
  p.faces = faces; p.z0 = z0; p.nzl = nzl; p.cz_lo = cz_lo; p.cz_hi = cz_hi;
  p.z_own_lo = z_own_lo; p.z_own_hi = z_own_hi;
  for (int i = 0; i < (P + 1) * (P + 1); ++i) p.S[i] = S[i];
  for (int i = 0; i < P + 1; ++i) p.lam[i] = lam[i];
  p.c[0] = h[1] * h[2] / h[0]; p.c[1] = h[0] * h[2] / h[1]; p.c[2] = h[0] * h[1] / h[2];
  p.mode = mode; p.u = u; p.b = b; p.xold = xold; p.out = out; p.f1 = f1; p.f2 = f2;
  p.dinv_vec = dinv_vec; p.dinv_tab = dinv_tab;
  run_all<P, BX, BY>(p, n_chunks);
}

#define ARGS nx, ny, nz, faces, z0, nzl, cz_lo, cz_hi, z_own_lo, z_own_hi, n_chunks, S, lam, h, mode, u, b, xold, out, f1, f2, dinv_vec, dinv_tab

// small_tiles != 0 selects deliberately tiny tiles so that small meshes exercise many tiles
extern "C" int emu_apply(int degree, int small_tiles, int nx, int ny, int nz, unsigned faces, int z0, int nzl,
                         int cz_lo, int cz_hi, int z_own_lo, int z_own_hi, int n_chunks, const double *S,
                         const double *lam, const double *h, int mode, const double *u, const double *b,
                         const double *xold, double *out, double f1, double f2, const double *dinv_vec,
                         const double *dinv_tab)
{
  if (small_tiles) {
    switch (degree) {
      case 1: go<1, 3, 2>(ARGS); return 0;
      case 2: go<2, 2, 3>(ARGS); return 0;
      case 3: go<3, 2, 2>(ARGS); return 0;
      case 4: go<4, 3, 2>(ARGS); return 0;
      case 5: go<5, 2, 2>(ARGS); return 0;
      case 6: go<6, 2, 1>(ARGS); return 0;
      case 7: go<7, 1, 2>(ARGS); return 0;
      case 8: go<8, 2, 2>(ARGS); return 0;
      case 9: go<9, 2, 1>(ARGS); return 0;
    }
    return -3;
  }
  switch (degree) { // the tiles pmg_apply.cu launches by default
    case 1: go<1, 10, 10>(ARGS); return 0;
    case 2: go<2, 8, 8>(ARGS); return 0;
    case 3: go<3, 7, 7>(ARGS); return 0;
    case 4: go<4, 6, 6>(ARGS); return 0;
    case 5: go<5, 5, 4>(ARGS); return 0;
    case 6: go<6, 5, 5>(ARGS); return 0;
    case 7: go<7, 4, 5>(ARGS); return 0;
    case 8: go<8, 4, 4>(ARGS); return 0;
    case 9: go<9, 3, 3>(ARGS); return 0;
  }
  return -3;
}
