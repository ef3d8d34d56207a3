// tests/emu/emu_sweep.cpp -- host emulator of the line-marching apply kernel (TEST INFRASTRUCTURE).
//
// Compiles the product's kernel source (csrc/pmg_apply_sweep.h) for the CPU and runs every CTA's phases
// thread by thread, so the CPU test-suite can check the kernel's index logic, halo/ownership rules and
// fused epilogues against the oracle without a GPU.  Not part of libpmg.so: the product has no CPU path.
#include <vector>
#include <cstring>
#include <type_traits>
#include "pmg_apply_sweep.h"

template <class Tile>
struct SweepHostExec {
  std::vector<typename Tile::ThreadState> st;
  bool reverse = false; // run the threads of a phase in descending order: exposes in-place hazards between threads of one phase
  template <class F> void for_each_thread(F f)
  {
    if (reverse) for (int t = Tile::NT - 1; t >= 0; --t) f(t, st[t]);
    else for (int t = 0; t < Tile::NT; ++t) f(t, st[t]);
  }
  void sync() {}
  void sync_some(int) {}
};

static bool g_reverse = false;

template <int P, int BX, int BY, int LZ, int NT, int US = 1, int SG = 1, int RL = 0, int A2 = 0, int PL = 0, int EG = 0>
static void sweep_go(int nx, int ny, int nz, unsigned faces, int z0, int nzl, int cz_lo, int cz_hi, int z_own_lo,
                     int z_own_hi, int n_chunks, const double *M, const double *K, const double *h, int mode,
                     const double *u, const double *b, const double *xold, double *out, double f1, double f2,
                     const double *dinv_vec, const double *dinv_tab)
{
  // PL: the pipelined variant (csrc/pmg_apply_sweep_pipe.h): two groups of NT threads each
  static_assert(PL == 0, "the pipelined variant was removed in round 2 (superseded by csrc/pmg_apply_plane.h)");
  using Tile = PmgSweepTile<P, BX, BY, LZ, NT, US, -1, SG, RL, A2, EG>;
  PmgSweepParams<P> p;
  std::memset(&p, 0, sizeof(p));
  p.nx = nx; p.ny = ny; p.nz = nz;
  p.Nx = nx * P + 1; p.Ny = ny * P + 1; p.Nz = nz * P + 1;
  p.faces = faces; p.z0 = z0; p.nzl = nzl; p.cz_lo = cz_lo; p.cz_hi = cz_hi;
  p.z_own_lo = z_own_lo; p.z_own_hi = z_own_hi;
  pmg_sweep_fill_matrices<P>(p, M, K, h);
  // the kernel's bulk row copies may read 16 bytes past the last element: run on padded copies (x_old may alias out)
  const size_t n_local = (size_t)p.Nx * p.Ny * nzl;
  std::vector<double> up(u, u + n_local), bp, xp;
  const size_t pad = (size_t)p.Nx + 2;
  up.resize(n_local + pad, 0.0);
  if (b) { bp.assign(b, b + n_local); bp.resize(n_local + pad, 0.0); }
  const bool alias = (xold != nullptr && xold == out);
  std::vector<double> outp;
  if (alias) { outp.assign(xold, xold + n_local); outp.resize(n_local + pad, 0.0); }
  else if (xold) { xp.assign(xold, xold + n_local); xp.resize(n_local + pad, 0.0); }
  p.mode = mode; p.u = up.data(); p.b = b ? bp.data() : nullptr;
  p.xold = alias ? outp.data() : (xold ? xp.data() : nullptr);
  p.out = alias ? outp.data() : out; p.f1 = f1; p.f2 = f2;
  p.dinv_vec = dinv_vec; p.dinv_tab = dinv_tab;
  p.tiles_x = (nx + BX - 1) / BX;
  p.tiles_y = (ny + BY - 1) / BY;
  const int layers = cz_hi - cz_lo;
  if (n_chunks < 1) n_chunks = 1;
  if (n_chunks > layers) n_chunks = layers;
  p.layers_per_chunk = (layers + n_chunks - 1) / n_chunks;
  p.n_chunks = (layers + p.layers_per_chunk - 1) / p.layers_per_chunk;
  std::vector<double> smem(Tile::SMEM_DOUBLES + 16);
  for (int chunk = 0; chunk < p.n_chunks; ++chunk)
    for (int ty = 0; ty < p.tiles_y; ++ty)
      for (int tx = 0; tx < p.tiles_x; ++tx) {
        SweepHostExec<Tile> ex;
        ex.st.resize(Tile::NT);
        ex.reverse = g_reverse;
        for (auto &v : smem) v = 1e300; // poison shared memory so stale reads show up
        Tile::run(p, ex, smem.data(), tx, ty, chunk);
      }
  if (alias) std::memcpy(out, outp.data(), n_local * sizeof(double));
}

#define ARGS nx, ny, nz, faces, z0, nzl, cz_lo, cz_hi, z_own_lo, z_own_hi, n_chunks, M, K, h, mode, u, b, xold, out, f1, f2, dinv_vec, dinv_tab

// small_tiles != 0 selects deliberately tiny tiles / thread counts so that small meshes exercise many tiles,
// several items per thread and several columns per thread; 0 = the tiles pmg_apply.cu launches;
// 2, 3 = small tiles with two segments per line (SG = 2), 3 with the threads of every phase run in descending order;
// 4 = small tiles with the cell loops of phases 1 and 2 rolled (RL = 1); 5 = small tiles, phase-2 items alternating between
// the two halves of the CTA from step to step (A2 = 1), rolled loops for the odd degrees; 6, 7 = the pipelined variant
// (pmg_apply_sweep_pipe.h) on small tiles, 7 with the threads run in descending order (the z-sweep group before the y/x group);
// 8 = the shipped tiles with rolled loops; 9 = small tiles, b / x_old read from global memory in the z sweep (EG = 1);
// 10 = the pipelined variant with EG = 1
extern "C" int emu_sweep(int degree, int small_tiles, int nx, int ny, int nz, unsigned faces, int z0, int nzl,
                         int cz_lo, int cz_hi, int z_own_lo, int z_own_hi, int n_chunks, const double *M,
                         const double *K, const double *h, int mode, const double *u, const double *b,
                         const double *xold, double *out, double f1, double f2, const double *dinv_vec,
                         const double *dinv_tab)
{
  g_reverse = (small_tiles == 3);
  if (small_tiles == 9) {
#define PMG_EG_CASE(P, BX, BY, LZ, NT, US, RL) \
  case P: sweep_go<P, BX, BY, LZ, NT, US, 1, RL, 0, 0, 1>(ARGS); return 0;
    switch (degree) {
      PMG_EG_CASE(1, 3, 2, 3, 32, 1, 0)
      PMG_EG_CASE(2, 2, 3, 2, 32, 1, 0)
      PMG_EG_CASE(3, 2, 3, 1, 32, 1, 1)
      PMG_EG_CASE(4, 3, 2, 1, 64, 1, 1)
      PMG_EG_CASE(5, 2, 3, 1, 32, 0, 0)
      PMG_EG_CASE(6, 2, 1, 1, 32, 1, 0)
      PMG_EG_CASE(7, 1, 2, 1, 32, 1, 1)
      PMG_EG_CASE(8, 1, 2, 1, 64, 0, 0)
    }
#undef PMG_EG_CASE
    return -3;
  }
  if (small_tiles == 5) {
    switch (degree) {
      case 1: sweep_go<1, 3, 2, 3, 32, 1, 1, 1, 1>(ARGS); return 0;
      case 2: sweep_go<2, 2, 3, 2, 32, 0, 1, 0, 1>(ARGS); return 0;
      case 3: sweep_go<3, 2, 3, 1, 64, 1, 1, 1, 1>(ARGS); return 0;
      case 4: sweep_go<4, 3, 2, 2, 64, 1, 1, 0, 1>(ARGS); return 0;
      case 5: sweep_go<5, 2, 3, 1, 32, 0, 1, 1, 1>(ARGS); return 0;
      case 6: sweep_go<6, 2, 1, 1, 32, 1, 1, 0, 1>(ARGS); return 0;
      case 7: sweep_go<7, 1, 2, 2, 32, 1, 1, 1, 1>(ARGS); return 0;
      case 8: sweep_go<8, 2, 2, 1, 64, 0, 1, 0, 1>(ARGS); return 0;
    }
    return -3;
  }
  if (small_tiles == 4) {
    switch (degree) {
      case 1: sweep_go<1, 3, 2, 3, 32, 1, 1, 1>(ARGS); return 0;
      case 2: sweep_go<2, 2, 3, 2, 32, 0, 1, 1>(ARGS); return 0;
      case 3: sweep_go<3, 2, 3, 1, 32, 1, 1, 1>(ARGS); return 0;
      case 4: sweep_go<4, 3, 2, 2, 64, 1, 1, 1>(ARGS); return 0;
      case 5: sweep_go<5, 2, 3, 1, 32, 0, 1, 1>(ARGS); return 0;
      case 6: sweep_go<6, 2, 1, 1, 32, 1, 1, 1>(ARGS); return 0;
      case 7: sweep_go<7, 1, 2, 2, 32, 1, 1, 1>(ARGS); return 0;
      case 8: sweep_go<8, 2, 2, 1, 64, 0, 1, 1>(ARGS); return 0;
    }
    return -3;
  }
  if (small_tiles >= 2) {
    switch (degree) {
      case 1: sweep_go<1, 4, 3, 3, 32, 1, 2>(ARGS); return 0;
      case 2: sweep_go<2, 3, 4, 2, 32, 0, 2>(ARGS); return 0;
      case 3: sweep_go<3, 2, 3, 1, 64, 1, 2>(ARGS); return 0;
      case 4: sweep_go<4, 4, 2, 1, 64, 1, 2>(ARGS); return 0;
      case 5: sweep_go<5, 2, 2, 1, 32, 0, 2>(ARGS); return 0;
      case 6: sweep_go<6, 2, 2, 1, 32, 1, 2>(ARGS); return 0;
      case 7: sweep_go<7, 2, 2, 1, 32, 1, 2>(ARGS); return 0;
      case 8: sweep_go<8, 2, 2, 1, 64, 0, 2>(ARGS); return 0;
    }
    return -3;
  }
  if (small_tiles) {
    switch (degree) {
      case 1: sweep_go<1, 3, 2, 3, 32>(ARGS); return 0;
      case 2: sweep_go<2, 2, 3, 2, 32>(ARGS); return 0;
      case 3: sweep_go<3, 2, 2, 1, 32>(ARGS); return 0;
      case 4: sweep_go<4, 3, 2, 2, 64>(ARGS); return 0;
      case 5: sweep_go<5, 2, 2, 1, 32>(ARGS); return 0;
      case 6: sweep_go<6, 2, 1, 1, 32>(ARGS); return 0;
      case 7: sweep_go<7, 1, 2, 2, 32>(ARGS); return 0;
      case 8: sweep_go<8, 2, 2, 1, 64>(ARGS); return 0;
      case 9: sweep_go<9, 1, 2, 1, 64>(ARGS); return 0;
    }
    return -3;
  }
  if (small_tiles == 8) { // the tiles pmg_apply.cu launches, cell loops rolled (what the Q4 plain apply ships with)
    switch (degree) {
#define PMG_SWEEP_CASE(P, BX, BY, LZ, NT, MINB, US) case P: sweep_go<P, BX, BY, LZ, NT, US, 1, 1>(ARGS); return 0;
#include "pmg_apply_sweep_tiles.inc"
#undef PMG_SWEEP_CASE
    }
    return -3;
  }
  switch (degree) {
#define PMG_SWEEP_CASE(P, BX, BY, LZ, NT, MINB, US) case P: sweep_go<P, BX, BY, LZ, NT, US>(ARGS); return 0;
#include "pmg_apply_sweep_tiles.inc"
#undef PMG_SWEEP_CASE
  }
  return -3;
}
