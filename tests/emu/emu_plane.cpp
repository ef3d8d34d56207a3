// tests/emu/emu_plane.cpp -- host emulator of the plane-per-step apply kernel (TEST INFRASTRUCTURE).
//
// Compiles the product's kernel source (csrc/pmg_apply_plane.h) for the CPU and runs every CTA's phases thread by thread,
// so the CPU test-suite can check the kernel's index logic, ownership rules, z-chunking and fused epilogues against the
// oracle without a GPU.  Not part of libpmg.so: the product has no CPU path.
#include <vector>
#include <cstring>
#include "pmg_apply_plane.h"

template <class Tile>
struct PlaneHostExec {
  std::vector<typename Tile::ThreadState> st;
  bool reverse = false; // run the threads of a phase in descending order: exposes hazards between threads of one phase
  template <class F> void for_each_thread(F f)
  {
    if (reverse) for (int t = Tile::NT - 1; t >= 0; --t) f(t, st[t]);
    else for (int t = 0; t < Tile::NT; ++t) f(t, st[t]);
  }
  void sync() {}
  void wait_flag(unsigned long long *, int, int) {} // fused ghost exchange: no neighbour GPUs here
};

static bool g_plane_reverse = false;
// fused ghost push under the emulator: the neighbours' copies of `out` and their first stored planes (emu_plane_set_push)
static double *g_push_lo = nullptr, *g_push_hi = nullptr;
static int g_push_lo_z0 = 0, g_push_hi_z0 = 0;
extern "C" void emu_plane_set_push(double *lower, int lower_z0, double *upper, int upper_z0)
{
  g_push_lo = lower; g_push_lo_z0 = lower_z0; g_push_hi = upper; g_push_hi_z0 = upper_z0;
}

template <int P, int BX, int BY, int NT, int UZ = 1, int PR = 1, int XS = 1>
static void plane_go(int nx, int ny, int nz, unsigned faces, int z0, int nzl, int cz_lo, int cz_hi, int z_own_lo,
                     int z_own_hi, int n_chunks, const double *M, const double *K, const double *h, int mode,
                     const double *u, const double *b, const double *xold, double *out, double f1, double f2,
                     const double *dinv_vec, const double *dinv_tab)
{
  using Tile = PmgPlaneTile<P, BX, BY, NT, -1, UZ, 0, 3, 0, PR, XS, 1>; // PUSH = 1: the emulator also covers the fused ghost push
  PmgSweepParams<P> p;
  std::memset(&p, 0, sizeof(p));
  p.nx = nx; p.ny = ny; p.nz = nz;
  p.Nx = nx * P + 1; p.Ny = ny * P + 1; p.Nz = nz * P + 1;
  p.faces = faces; p.z0 = z0; p.nzl = nzl; p.cz_lo = cz_lo; p.cz_hi = cz_hi;
  p.z_own_lo = z_own_lo; p.z_own_hi = z_own_hi;
  pmg_sweep_fill_matrices<P>(p, M, K, h);
  p.mode = mode; p.u = u; p.b = b; p.xold = xold; p.out = out; p.f1 = f1; p.f2 = f2;
  p.dinv_vec = dinv_vec; p.dinv_tab = dinv_tab;
  { // as the launcher does (csrc/pmg_apply_plane_launch.h): pointers indexed like `out`
    const int64_t plane = (int64_t)p.Nx * p.Ny;
    if (g_push_lo) p.push_lo = g_push_lo + plane * (z0 - g_push_lo_z0);
    if (g_push_hi) p.push_hi = g_push_hi + plane * (z0 - g_push_hi_z0);
  }
  p.tiles_x = Tile::tiles_of(nx, faces >> 1 & 1u, BX);
  p.tiles_y = Tile::tiles_of(ny, faces >> 3 & 1u, BY);
  const int layers = cz_hi - cz_lo;
  if (n_chunks < 1) n_chunks = 1;
  if (n_chunks > layers) n_chunks = layers;
  p.layers_per_chunk = (layers + n_chunks - 1) / n_chunks;
  p.n_chunks = (layers + p.layers_per_chunk - 1) / p.layers_per_chunk;
  std::vector<double> smem(Tile::SMEM_DOUBLES + 16);
  for (int chunk = 0; chunk < p.n_chunks; ++chunk)
    for (int ty = 0; ty < p.tiles_y; ++ty)
      for (int tx = 0; tx < p.tiles_x; ++tx) {
        PlaneHostExec<Tile> ex;
        ex.st.resize(Tile::NT);
        ex.reverse = g_plane_reverse;
        for (auto &v : smem) v = 1e300; // poison shared memory so stale reads show up
        Tile::run(p, ex, smem.data(), tx, ty, chunk);
        // fused ghost push (csrc/pmg_apply_plane_launch.h): the bottom / top chunk's CTAs copy the slab's boundary planes
        if (p.push_lo || p.push_hi) Tile::push_boundary(p, ex, tx, ty, chunk == 0, chunk == p.n_chunks - 1);
      }
}

#define ARGS nx, ny, nz, faces, z0, nzl, cz_lo, cz_hi, z_own_lo, z_own_hi, n_chunks, M, K, h, mode, u, b, xold, out, f1, f2, dinv_vec, dinv_tab

// small_tiles: 0 = the tiles csrc/pmg_apply_plane_tiles.inc launches; 1 = tiny tiles and thread counts (several tiles,
// several items per thread); 2 = tiny tiles, threads run in descending order; 3 = tiny tiles, the layer's steps rolled (UZ = 0);
// 4 = tiny tiles, two x items per cell row (XS = 2)
extern "C" int emu_plane(int degree, int small_tiles, int nx, int ny, int nz, unsigned faces, int z0, int nzl,
                         int cz_lo, int cz_hi, int z_own_lo, int z_own_hi, int n_chunks, const double *M,
                         const double *K, const double *h, int mode, const double *u, const double *b,
                         const double *xold, double *out, double f1, double f2, const double *dinv_vec,
                         const double *dinv_tab)
{
  g_plane_reverse = (small_tiles == 2);
  if (small_tiles == 4) { // two x items per cell row (XS = 2), rolled and unrolled, with and without 16-byte pairs
    switch (degree) {
      case 2: plane_go<2, 2, 3, 32, 1, 1, 2>(ARGS); return 0;
      case 3: plane_go<3, 2, 2, 32, 0, 0, 2>(ARGS); return 0;
      case 4: plane_go<4, 3, 2, 64, 1, 0, 2>(ARGS); return 0;
      case 5: plane_go<5, 2, 2, 32, 0, 1, 2>(ARGS); return 0;
      case 6: plane_go<6, 2, 1, 32, 0, 0, 2>(ARGS); return 0;
      case 7: plane_go<7, 1, 2, 32, 0, 0, 2>(ARGS); return 0;
      case 8: plane_go<8, 2, 2, 64, 0, 0, 2>(ARGS); return 0;
      case 9: plane_go<9, 1, 2, 64, 0, 0, 2>(ARGS); return 0;
    }
    return -3;
  }
  if (small_tiles == 3) {
    switch (degree) {
      case 1: plane_go<1, 3, 2, 32, 0, 0>(ARGS); return 0;
      case 2: plane_go<2, 2, 3, 32, 0, 0>(ARGS); return 0;
      case 3: plane_go<3, 2, 2, 32, 0>(ARGS); return 0;
      case 4: plane_go<4, 3, 2, 32, 0, 0>(ARGS); return 0;
      case 5: plane_go<5, 2, 2, 32, 0>(ARGS); return 0;
      case 6: plane_go<6, 2, 1, 32, 0>(ARGS); return 0;
      case 7: plane_go<7, 1, 2, 32, 0, 0>(ARGS); return 0;
      case 8: plane_go<8, 2, 2, 64, 0>(ARGS); return 0;
    }
    return -3;
  }
  if (small_tiles) {
    switch (degree) {
      case 1: plane_go<1, 3, 2, 32>(ARGS); return 0;
      case 2: plane_go<2, 2, 3, 32>(ARGS); return 0;
      case 3: plane_go<3, 2, 2, 32>(ARGS); return 0;
      case 4: plane_go<4, 3, 2, 32>(ARGS); return 0;
      case 5: plane_go<5, 2, 2, 32>(ARGS); return 0;
      case 6: plane_go<6, 2, 1, 32>(ARGS); return 0;
      case 7: plane_go<7, 1, 2, 32>(ARGS); return 0;
      case 8: plane_go<8, 2, 2, 64>(ARGS); return 0;
      case 9: plane_go<9, 1, 2, 64>(ARGS); return 0;
    }
    return -3;
  }
  // the shipped tiles: the plain apply's for mode 0, the fused modes' otherwise
  if (mode == 0) {
    switch (degree) {
#define PMG_PLANE_CASE(P, BX, BY, NT, MINB, UZ, PR, XS) case P: plane_go<P, BX, BY, NT, UZ, PR, XS>(ARGS); return 0;
#define PMG_PLANE_CASE_F(P, BX, BY, NT, MINB, UZ, PR, XS)
#include "pmg_apply_plane_tiles.inc"
#undef PMG_PLANE_CASE
#undef PMG_PLANE_CASE_F
    }
  } else {
    switch (degree) {
#define PMG_PLANE_CASE(P, BX, BY, NT, MINB, UZ, PR, XS)
#define PMG_PLANE_CASE_F(P, BX, BY, NT, MINB, UZ, PR, XS) case P: plane_go<P, BX, BY, NT, UZ, PR, XS>(ARGS); return 0;
#include "pmg_apply_plane_tiles.inc"
#undef PMG_PLANE_CASE
#undef PMG_PLANE_CASE_F
    }
  }
  return -3;
}
