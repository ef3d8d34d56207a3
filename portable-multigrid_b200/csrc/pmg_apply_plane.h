// pmg_apply_plane.h -- the Laplace cell-loop apply, one dof plane per step (kernel K1, v7), fused with the smoother update.
//
// Replaces LaplaceOperator::vmult + LocalLaplaceOperator::operator()
// (reference include/operators/portable_laplace_operator.h:227-381, 557-719) and the PreconditionChebyshev vector update
// that follows it (include/multigrid/portable_v_cycle_multigrid.h:116-125) by ONE pass over HBM.
//
// Same mathematics as the line-marching kernel (csrc/pmg_apply_sweep.h): on the affine box mesh
//     A u = Mz [ Kx My u + Mx Ky u ] + Kz Mx My u,        c = My u, d = Ky u;  g = Kx c + Mx d, m = Mx c;  A u = Mz g + Kz m
// with the 1-D cell matrices M, K.  What differs is how the sweeps are cut into work (round-1 ablation: the marching kernel
// was bound by barriers, long serial chains per thread and 65 KB of unrolled code, not by a hardware unit):
//  * A 1-D sweep item is ONE CELL of ONE LINE: it produces the cell's P owned values (low vertex + interior nodes) from the
//    2P+1 inputs of the cell and the cell below it.  That is exactly P(P+2) FMAs per matrix -- the row length of the
//    assembled 1-D band matrix, no partial-cell recomputation -- every item runs the same straight-line code, nothing is
//    carried from item to item, and a plane of a BX x BY tile has (BX P + P + 1) BY y-items and BY P BX x-items: enough
//    for every warp of the CTA in every phase with ONE dof plane in flight.
//  * One plane per step, one block barrier per plane: u planes are double-buffered in shared memory (loaded a plane ahead
//    through registers, coalesced; Dirichlet / out-of-mesh positions are never written and stay zero), the c, d planes too.
//  * The thread that finishes the x sweep of a cell row also owns its z sweep: P (P+1) accumulators in registers (the sums of
//    the layer's P+1 planes for the cell row's P nodes), updated with every plane (output-stationary), so g and m never
//    touch shared memory.  A finished cell layer leaves through a shared-memory box that turns the (row, cell) items into
//    rows of consecutive dofs: the epilogue (Dirichlet identity / residual / Chebyshev update) reads b, x_old, u and writes
//    the result fully coalesced, spread over the steps of the next layer.
//  * Every per-thread address is one offset plus a compile-time multiple of a constant (loader rows, epilogue rows), so the
//    steady-state loop carries no index arithmetic beyond one wide multiply-add per global access.
//
// Written against an executor (for_each_thread / sync) like the other tile programs: the same source is the CUDA kernel and
// runs thread by thread under the host emulator of the CPU test-suite (tests/emu/emu_plane.cpp; test infrastructure only).
#pragma once
#include "pmg_apply_sweep.h"

// Keeps a pointer that was advanced to the current plane as ONE value in registers: without it the compiler folds the
// (64-bit) plane offset back into every access's index arithmetic -- 5 instructions per global access instead of 1 wide
// multiply-add of a 32-bit element index.
#if defined(__CUDA_ARCH__)
#define PMG_OPAQUE_PTR(ptr) asm volatile("" : "+l"(ptr))
#else
#define PMG_OPAQUE_PTR(ptr) ((void)0)
#endif
// accesses through such a pointer name their address space themselves (the compiler no longer knows it)
PMG_HD double pmg_plane_ldg(const double *ptr)
{
#if defined(__CUDA_ARCH__)
  double v;
  asm volatile("ld.global.f64 %0, [%1];\n" : "=d"(v) : "l"(ptr) : "memory");
  return v;
#else
  return *ptr;
#endif
}
// L2 / L1 prefetch of the line that holds *ptr (no register, no result)
PMG_HD void pmg_plane_prefetch(const double *ptr, int level)
{
#if defined(__CUDA_ARCH__)
  if (level == 1) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(ptr) : "memory");
  else asm volatile("prefetch.global.L1 [%0];\n" ::"l"(ptr) : "memory");
#else
  (void)ptr; (void)level;
#endif
}
// 8-byte asynchronous copy global -> shared; the destination is a byte offset from the CTA's shared-memory base (base32: its
// 32-bit shared-window address, formed once per kernel -- a generic destination pointer costs three instructions per copy)
PMG_HD void pmg_plane_cp_async8(double *smem, unsigned base32, unsigned byte_off, const double *src_global)
{
#if defined(__CUDA_ARCH__)
  (void)smem;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(base32 + byte_off), "l"(src_global) : "memory");
#else
  (void)base32;
  *(double *)((char *)smem + byte_off) = *src_global;
#endif
}
// (c, d) of one dof, stored and loaded as one 16-byte shared-memory access
struct alignas(16) PmgPlanePair { double c, d; };

// keeps a per-thread index in its register: without it the compiler, short of registers, recomputes such values from the
// thread index at every use (ncu: 15 instructions per asynchronous copy)
#if defined(__CUDA_ARCH__)
#define PMG_KEEP(v) asm volatile("" : "+r"(v))
#else
#define PMG_KEEP(v) ((void)0)
#endif
PMG_HD void pmg_plane_stg(double *ptr, double v)
{
#if defined(__CUDA_ARCH__)
  asm volatile("st.global.f64 [%0], %1;\n" ::"l"(ptr), "d"(v) : "memory");
#else
  *ptr = v;
#endif
}

// Flag words of the fused ghost exchange in a rank's mailbox (csrc/pmg_apply_plane_launch.h; words 0..5 belong to the stand-alone
// push kernel, csrc/pmg_halo.cu).  FROM_LO / FROM_HI are written by the lower / upper neighbour: the number of fused launches
// whose chunks at the slab end that faces me they have completed; EPOCH_LO / EPOCH_HI are my own counts for my two slab ends.
enum { PMG_FUSED_FROM_LO = 8, PMG_FUSED_FROM_HI = 9, PMG_FUSED_EPOCH_LO = 10, PMG_FUSED_EPOCH_HI = 11, PMG_FUSED_TICKET_LO = 12, PMG_FUSED_TICKET_HI = 13,
       PMG_FUSED_ERROR = 14 /* set when a wait for a neighbour timed out */ };

// P: degree; BX x BY: cells per CTA; NT_: threads; FM: epilogue mode the kernel is compiled for (-1: p.mode);
// UZ: 1 = the P steps of a cell layer are unrolled (z matrices as immediate constants), 0 = rolled (z matrices indexed)
// LW_: lanes per u row in the loader (>= the row length BX P + P + 1, divides NT_; 0 = smallest such number): a thread loads
// column tid % LW of rows tid / LW + k NT / LW
// NU_: u planes in the shared-memory ring (>= 2): a plane is fetched NU - 1 steps ahead with cp.async (LDGSTS, no registers
// held; the first version prefetched one plane through registers and ncu showed 42 % of the stall samples waiting for it)
// EPF: the epilogue's b / x_old / u of the plane two planes up are prefetched (1: into L2, 2: into L1) while a plane's epilogue
// runs -- a step or two before they are read; 0: not (the fused step then waits for HBM in every epilogue)
// PR: 1 = c and d of a dof are stored side by side and move as one 16-byte shared-memory access (half the LDS / STS
// instructions of the two sweeps: +11 % at Q3, +2..4 % at Q1 / Q2; at Q4 the aligned register quadruples of the wide loads cost
// more than they save under the 128-register cap: 118 against 131 GDoF/s, so PR = 0 there: two separate planes)
// XS: the P nodes of a cell row are shared by XS (1 or 2) x items -- nodes [0, H) and [H, P), H = ceil(P / 2) -- so that the
// z-sweep accumulators of a thread, H (P+1) doubles, still fit the register file at degrees 7..9 (at the price of the second
// item re-reading the cell's c, d columns)
// PUSH: 1 = the epilogue also stores the slab's boundary planes into the neighbours' ghost planes (p.push_lo / p.push_hi: fused
// ghost exchange, csrc/pmg_apply_plane_launch.h); a separate instance, so that launches without it carry none of its code
template <int P, int BX, int BY, int NT_, int FM = -1, int UZ = 1, int LW_ = 0, int NU_ = 3, int EPF = 0, int PR = 1, int XS = 1, int PUSH = 0>
struct PmgPlaneTile {
  static constexpr int N1 = P + 1;
  static constexpr int NT = NT_;
  // FM = 4: PMG_MODE_CHEB_STEP compiled for x_old == NULL (the step after a zero initial guess); FM = 3 then requires x_old
  static constexpr int MODE_STEP0 = 4;
  static PMG_HD int mode_of(const PmgSweepParams<P> &p)
  {
    return FM >= 0 ? FM : (p.mode == PMG_MODE_CHEB_STEP && p.xold == nullptr) ? MODE_STEP0 : p.mode;
  }
  static constexpr int OX = BX * P, OY = BY * P;             // dofs per plane whose A u the CTA completes
  static constexpr int XW = OX + P + 1, YW = OY + P + 1;     // u footprint: global x = (cx0 - 1) P + xl
  static constexpr int UP = XW;                              // pitch of a u row
  static constexpr int UPLANE = YW * UP;
  static constexpr int CP = OY | 1;                          // c, d are stored transposed [xl][oy]: odd pitch
  static constexpr int CPLANE = XW * CP;
  static constexpr int OP = OX | 1;                          // pitch of a row of the output box
  static constexpr int ERW = OY + 1;                         // epilogue rows: +1 = the mesh's last vertex line (Dirichlet face)
  static constexpr int OPLANE = ERW * OP;
  // output box: planes 0 .. P-2 of a closed layer, plane P-1 in two alternating copies (it leaves at the next vertex step,
  // while the box is rewritten), the mesh's top plane; P = 1: two alternating copies + the top plane
  static constexpr int NOB = P + 2;
  static constexpr int T = P + 2;                            // position types per direction (inverse-diagonal table)
  static constexpr int NU = NU_, ND = NU_ - 1;               // ring depth, fetch distance
  static_assert(NU_ >= 2 && NU_ <= 8, "ring depth");
  // c and d of a dof sit side by side (one 16-byte access): the c, d buffers start at an even offset
  static constexpr int U_OFFSET = 0, C_OFFSET = (NU * UPLANE + 1) & ~1, O_OFFSET = C_OFFSET + 4 * CPLANE, T_OFFSET = O_OFFSET + NOB * OPLANE;
  static constexpr int SMEM_DOUBLES = T_OFFSET + T * T * T;
  static constexpr int pick_lw() { int w = XW; while (NT_ % w) ++w; return w; }
  static constexpr int LW = LW_ ? LW_ : pick_lw();           // loader: lanes per u row
  static constexpr int LR = NT / LW;                         // rows per pass
  static constexpr int NLD = (YW + LR - 1) / LR;             // passes = u elements a thread loads per plane
  static_assert(NT % LW == 0 && LW >= XW, "loader: LW lanes per row, whole rows per pass");
  static constexpr int NYI = XW * BY, IY = (NYI + NT - 1) / NT; // y items (xl fastest), per thread
  static_assert(XS == 1 || (XS == 2 && P >= 2), "one or two x items per cell row");
  static constexpr int HN = (P + XS - 1) / XS;               // nodes of a cell row per x item (the first item's; the second takes the rest)
  static constexpr int NXI = OY * BX * XS, IX = (NXI + NT - 1) / NT; // x items (oy fastest, then cell, then half), per thread
  // epilogue: thread = column tid % OX of rows tid / OX + k ER
  static constexpr int ER = NT / OX;
  static constexpr int NE = (ERW + (ER > 0 ? ER : 1) - 1) / (ER > 0 ? ER : 1);
  static_assert(ER >= 1 && NE <= 8 && NLD <= 16, "epilogue rows / loader rows per thread are packed into bit masks");
  // epilogue inputs per mode: b | b, u | b, u, x_old
  static PMG_HD constexpr int n_arrays(int mode) { return mode == PMG_MODE_APPLY ? 0 : mode == PMG_MODE_RESIDUAL ? 1 : mode == PMG_MODE_CHEB_STEP ? 3 : 2; }
  static_assert(NT % 32 == 0, "whole warps");
  static_assert(NT >= ERW, "the extra epilogue column takes one thread per row");
  static_assert(SMEM_DOUBLES * 8 <= 227 * 1024, "tile does not fit the shared memory of a CTA");

#define PMG_M(i, j) p.M[pmg_sweep_canon<P>(i, j)]
#define PMG_KX(i, j) p.Kx[pmg_sweep_canon<P>(i, j)]
#define PMG_KY(i, j) p.Ky[pmg_sweep_canon<P>(i, j)]
#define PMG_MZ(i, j) (UZ ? p.Mz[pmg_sweep_canon<P>(i, j)] : p.Mz[(i) * N1 + (j)])
#define PMG_KZ(i, j) (UZ ? p.Kz[pmg_sweep_canon<P>(i, j)] : p.Kz[(i) * N1 + (j)])

  struct ThreadState {
    double acc[IX][HN][N1]; // z sweep: sums of the current layer's planes r = 0..P for the item's nodes of its cell row
    const double *ld_src;  // loader: the thread's first element (row tid / LW, column tid % LW) of the next plane to fetch
    int ld_off;            // its element offset inside a dof plane
    int ld_dst;            // its byte offset in shared memory (ring slot 0)
    unsigned ld_mask;      // bit k: element of row tid / LW + k LR is inside the mesh and not Dirichlet (else it stays 0)
    int yi[IY];            // y item: offset of its first u row in a u plane | -1
    int yo[IY];            // its first output in a c plane
    double ycnt[IY];       // cells that hold the item's vertex (1 at the mesh boundary, else 2)
    int xi[IX];            // x item: offset of its first input in a c plane | -1; XS = 2: bit 30 set for the second half's item
    int xo[IX];            // its first output in a plane of the output box
    double xcnt[IX];
    int e_off;             // epilogue: element offset of (row tid / OX, column tid % OX) from the tile's first owned dof of a plane
    int e_src;             // its index in a plane of the output box
    double ein[3][NE];     // the step's epilogue inputs b, u, x_old, loaded before the y sweep and used after the barrier
    unsigned e_flags;      // bit k: row tid / OX + k ER is written; bit 8 + k: it is a Dirichlet dof (in xy); bits 24..27: x position type
    unsigned e_ty;         // 4 bits per k: y position type
  };

  struct TileGeom {
    int cx0, cy0;       // first owned cell column
    int virt_x, virt_y; // the last vertex line of the mesh is computed as a virtual cell (face not Dirichlet)
    int64_t plane;      // dofs per plane
    int64_t tile0;      // element offset of the tile's first owned dof inside a plane
    bool xextra;        // the tile also writes the mesh's last vertex line in x (Dirichlet face: identity rows, no x item)
    bool dirxy;         // the tile holds Dirichlet dofs of the x / y faces
    int zlo, zhi;       // dof planes that are Dirichlet faces (-1: none)
  };

  static PMG_HD TileGeom geom(const PmgSweepParams<P> &p, int tile_x, int tile_y)
  {
    TileGeom t;
    t.cx0 = tile_x * BX; t.cy0 = tile_y * BY;
    t.virt_x = !(p.faces >> 1 & 1u); t.virt_y = !(p.faces >> 3 & 1u);
    t.plane = (int64_t)p.Nx * p.Ny;
    t.tile0 = (int64_t)(t.cy0 * P) * p.Nx + t.cx0 * P;
    t.xextra = !t.virt_x && (t.cx0 * P + OX == p.Nx - 1);
    const int gx1 = t.cx0 * P + OX, gy1 = t.cy0 * P + OY; // one past the tile's own dofs (the extra line included: >=)
    t.dirxy = ((p.faces & 1u) && t.cx0 == 0) || ((p.faces >> 1 & 1u) && gx1 >= p.Nx - 1) ||
              ((p.faces >> 2 & 1u) && t.cy0 == 0) || ((p.faces >> 3 & 1u) && gy1 >= p.Ny - 1);
    t.zlo = (p.faces >> 4 & 1u) ? 0 : -1;
    t.zhi = (p.faces >> 5 & 1u) ? p.Nz - 1 : -1;
    return t;
  }
  // host side: tiles per direction (cells + the virtual cell of a non-Dirichlet high face)
  static inline int tiles_of(int ncell, bool high_dirichlet, int B) { return (ncell + (high_dirichlet ? 0 : 1) + B - 1) / B; }

  static PMG_HD bool dir_xy(const PmgSweepParams<P> &p, int gx, int gy)
  {
    return (gx == 0 && (p.faces & 1u)) || (gx == p.Nx - 1 && (p.faces >> 1 & 1u)) ||
           (gy == 0 && (p.faces >> 2 & 1u)) || (gy == p.Ny - 1 && (p.faces >> 3 & 1u));
  }

  static PMG_HD void decode(const PmgSweepParams<P> &p, const TileGeom &t, int tid, ThreadState &st, double *smem)
  {
    // the u planes start as zeros (positions outside the mesh / Dirichlet are never written); inverse-diagonal table
    for (int e = tid; e < NU * UPLANE; e += NT) smem[U_OFFSET + e] = 0.0;
    if (mode_of(p) >= PMG_MODE_CHEB_FIRST && !p.dinv_vec)
      for (int e = tid; e < T * T * T; e += NT) smem[T_OFFSET + e] = p.dinv_tab[e];
    {
      const int lr0 = tid / LW, lx = tid - lr0 * LW;
      const int gx = (t.cx0 - 1) * P + lx, gy0 = (t.cy0 - 1) * P + lr0;
      st.ld_off = gy0 * p.Nx + gx;
      st.ld_dst = (U_OFFSET + lr0 * UP + lx) * 8;
      st.ld_mask = 0;
#pragma unroll
      for (int k = 0; k < NLD; ++k) {
        const int yl = lr0 + k * LR, gy = gy0 + k * LR;
        const bool in = lx < XW && yl < YW && gx >= 0 && gx < p.Nx && gy >= 0 && gy < p.Ny;
        if (in && !dir_xy(p, gx, gy)) st.ld_mask |= 1u << k;
      }
    }
#pragma unroll
    for (int r = 0; r < IY; ++r) {
      const int item = tid + r * NT;
      st.yi[r] = -1; st.yo[r] = 0; st.ycnt[r] = 1.0;
      if (item < NYI) {
        const int yc = item / XW, xl = item - yc * XW;
        const int gcy = t.cy0 + yc;
        if (gcy < p.ny || (gcy == p.ny && t.virt_y)) {
          st.yi[r] = yc * P * UP + xl;
          st.yo[r] = xl * CP + yc * P;
          st.ycnt[r] = (gcy > 0 && gcy < p.ny) ? 2.0 : 1.0;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < IX; ++r) {
      const int item = tid + r * NT;
      st.xi[r] = -1; st.xo[r] = 0; st.xcnt[r] = 1.0;
#pragma unroll
      for (int i = 0; i < HN; ++i)
#pragma unroll
        for (int q = 0; q < N1; ++q) st.acc[r][i][q] = 0.0;
      if (item < NXI) {
        const int half = item / (OY * BX), rem = item - half * (OY * BX);
        const int xc = rem / OY, oy = rem - xc * OY;
        const int gcx = t.cx0 + xc, gy = t.cy0 * P + oy;
        const bool row_ok = gy < p.Ny - 1 || (gy == p.Ny - 1 && t.virt_y);
        if (row_ok && (gcx < p.nx || (gcx == p.nx && t.virt_x))) {
          st.xi[r] = (xc * P * CP + oy) | (half << 30);
          st.xo[r] = oy * OP + xc * P;
          st.xcnt[r] = (gcx > 0 && gcx < p.nx) ? 2.0 : 1.0;
        }
      }
    }
    {
      const int er0 = tid / OX, ox = tid - er0 * OX;
      const int gx = t.cx0 * P + ox;
      st.e_off = er0 * p.Nx + ox;
      st.e_src = er0 * OP + ox;
      st.e_flags = (unsigned)pmg_sweep_pos_type<P>(gx < p.Nx ? gx : 0, p.Nx) << 24;
      st.e_ty = 0;
#pragma unroll
      for (int k = 0; k < NE; ++k) {
        const int oy = er0 + k * ER, gy = t.cy0 * P + oy;
        // the box's extra row is the mesh's last vertex line when that face is Dirichlet (identity rows, no x item)
        const bool ok = er0 < ER && gx < p.Nx && gy < p.Ny && (oy < OY || (oy == OY && !t.virt_y && gy == p.Ny - 1));
        if (ok) {
          st.e_flags |= 1u << k;
          if (dir_xy(p, gx, gy)) st.e_flags |= 1u << (8 + k);
          st.e_ty |= (unsigned)pmg_sweep_pos_type<P>(gy, p.Ny) << (4 * k);
        }
      }
    }
    PMG_KEEP(st.ld_off); PMG_KEEP(st.ld_dst); PMG_KEEP(st.ld_mask);
    PMG_KEEP(st.e_off); PMG_KEEP(st.e_src); PMG_KEEP(st.e_flags); PMG_KEEP(st.e_ty);
#pragma unroll
    for (int r = 0; r < IY; ++r) { PMG_KEEP(st.yi[r]); PMG_KEEP(st.yo[r]); }
#pragma unroll
    for (int r = 0; r < IX; ++r) { PMG_KEEP(st.xi[r]); PMG_KEEP(st.xo[r]); }
  }

  // ---- loader: the u plane at `up` -> ring slot U, asynchronously (one commit group per plane, empty ones included) --------
  static PMG_HD void load_plane(const PmgSweepParams<P> &p, ThreadState &st, int64_t plane, double *smem, unsigned base32, int slot, bool doit)
  {
    if (doit) {
      // one destination address and one source pointer per thread; row k is a constant / a multiple of the row step away
      unsigned dst = base32 + (unsigned)st.ld_dst + (unsigned)(slot * UPLANE * 8);
      PMG_KEEP(dst);
      const double *src = st.ld_src;
      const int rstep = LR * p.Nx; // 32-bit element offsets inside a plane (checked by the launcher)
#pragma unroll
      for (int k = 0; k < NLD; ++k)
        if (st.ld_mask >> k & 1u) pmg_plane_cp_async8(smem, base32, dst - base32 + k * LR * UP * 8, src + k * rstep);
    }
    st.ld_src += plane; // the plane after it, fetched or not
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.commit_group;\n" ::: "memory");
#endif
  }
  // wait until at most ND - 1 of this thread's plane groups are pending: the next step's plane has arrived
  static PMG_HD void load_wait()
  {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.wait_group %0;\n" ::"n"(ND - 1) : "memory");
#endif
  }

  // ---- y sweep: item = (xl, cell yc): u rows yc P .. yc P + 2P  ->  c, d of rows yc P .. yc P + P - 1 -------------------
  static PMG_HD void ysweep(const PmgSweepParams<P> &p, const ThreadState &st, const double *U, double *Cb)
  {
#pragma unroll
    for (int r = 0; r < IY; ++r) {
      if (st.yi[r] < 0) continue;
      const double *Uc = U + st.yi[r];
      double c[P], d[P];
      double v = Uc[0];
      c[0] = PMG_M(P, 0) * v; d[0] = PMG_KY(P, 0) * v;
#pragma unroll
      for (int j = 1; j < P; ++j) {
        v = Uc[j * UP];
        c[0] = fma(PMG_M(P, j), v, c[0]); d[0] = fma(PMG_KY(P, j), v, d[0]);
      }
      v = Uc[P * UP];
      {
        const double vv = v * st.ycnt[r];
        c[0] = fma(PMG_M(0, 0), vv, c[0]); d[0] = fma(PMG_KY(0, 0), vv, d[0]);
#pragma unroll
        for (int i = 1; i < P; ++i) { c[i] = PMG_M(i, 0) * v; d[i] = PMG_KY(i, 0) * v; }
      }
#pragma unroll
      for (int j = 1; j < N1; ++j) {
        v = Uc[(P + j) * UP];
#pragma unroll
        for (int i = 0; i < P; ++i) { c[i] = fma(PMG_M(i, j), v, c[i]); d[i] = fma(PMG_KY(i, j), v, d[i]); }
      }
      if (PR) {
        PmgPlanePair *Co = reinterpret_cast<PmgPlanePair *>(Cb) + st.yo[r];
#pragma unroll
        for (int i = 0; i < P; ++i) { PmgPlanePair v; v.c = c[i]; v.d = d[i]; Co[i] = v; }
      } else {
        double *Co = Cb + st.yo[r];
#pragma unroll
        for (int i = 0; i < P; ++i) { Co[i] = c[i]; Co[CPLANE + i] = d[i]; }
      }
    }
  }

  // ---- x sweep of item (oy, cell xc): c, d columns xc P .. xc P + 2P of row oy -> g, m of the cell row's P nodes ---------
  // (c, d) of column j of the item's row: one 16-byte load (PR) or two loads from the c and the d plane
  static PMG_HD PmgPlanePair cd_at(const double *Cc, int j)
  {
    if (PR) return reinterpret_cast<const PmgPlanePair *>(Cc)[j * CP];
    PmgPlanePair v; v.c = Cc[j * CP]; v.d = Cc[CPLANE + j * CP];
    return v;
  }
  // nodes [I0, I1) of the cell row: g[i - I0], m[i - I0].  I0 == 0 includes the cell's low vertex, which also takes the cell below
  template <int I0, int I1>
  static PMG_HD void xsweep_item(const PmgSweepParams<P> &p, const double *Cc, double cnt, double *g, double *m)
  {
    PmgPlanePair v;
    double c, d;
    if (I0 == 0) {
      v = cd_at(Cc, 0); c = v.c; d = v.d;
      g[0] = fma(PMG_KX(P, 0), c, PMG_M(P, 0) * d); m[0] = PMG_M(P, 0) * c;
#pragma unroll
      for (int j = 1; j < P; ++j) {
        v = cd_at(Cc, j); c = v.c; d = v.d;
        g[0] = fma(PMG_KX(P, j), c, fma(PMG_M(P, j), d, g[0])); m[0] = fma(PMG_M(P, j), c, m[0]);
      }
    }
    v = cd_at(Cc, P); c = v.c; d = v.d;
    {
      if (I0 == 0) {
        const double cc = c * cnt, dd = d * cnt;
        g[0] = fma(PMG_KX(0, 0), cc, fma(PMG_M(0, 0), dd, g[0])); m[0] = fma(PMG_M(0, 0), cc, m[0]);
      }
#pragma unroll
      for (int i = (I0 == 0 ? 1 : I0); i < I1; ++i) { g[i - I0] = fma(PMG_KX(i, 0), c, PMG_M(i, 0) * d); m[i - I0] = PMG_M(i, 0) * c; }
    }
#pragma unroll
    for (int j = 1; j < N1; ++j) {
      v = cd_at(Cc, P + j); c = v.c; d = v.d;
#pragma unroll
      for (int i = I0; i < I1; ++i) {
        g[i - I0] = fma(PMG_KX(i, j), c, fma(PMG_M(i, j), d, g[i - I0])); m[i - I0] = fma(PMG_M(i, j), c, m[i - I0]);
      }
    }
  }

  // ---- x sweep + z sweep of the plane with index jz inside its cell layer -------------------------------------------------
  // zero: the plane is a Dirichlet face (u reads as 0: g = m = 0).  Vertex planes (jz == 0) close the layer below
  // (has_prev; emit: its P planes go to the output box, top: so does the closed sum of the vertex plane itself, the mesh's
  // top plane) and open the layer above (has_next).
  template <int I0, int I1>
  static PMG_HD void xz_item(const PmgSweepParams<P> &p, ThreadState &st, int r, const double *Cb, double *Ob, int jz, bool zero,
                             bool has_prev, bool has_next, bool emit, bool top, int oslot)
  {
    constexpr int NI = I1 - I0;
    double g[NI], m[NI];
    if (!zero) xsweep_item<I0, I1>(p, Cb + (PR ? 2 : 1) * (st.xi[r] & 0x3FFFFFFF), st.xcnt[r], g, m);
    else {
#pragma unroll
      for (int i = 0; i < NI; ++i) { g[i] = 0.0; m[i] = 0.0; }
    }
    if (jz != 0) {
#pragma unroll
      for (int i = 0; i < NI; ++i)
#pragma unroll
        for (int q = 0; q < N1; ++q) st.acc[r][i][q] = fma(PMG_MZ(q, jz), g[i], fma(PMG_KZ(q, jz), m[i], st.acc[r][i][q]));
    } else {
      double *Oo = Ob + st.xo[r] + I0;
      double carry[NI];
#pragma unroll
      for (int i = 0; i < NI; ++i) carry[i] = 0.0;
      if (has_prev) {
        if (emit) {
#pragma unroll
          for (int i = 0; i < NI; ++i)
#pragma unroll
            for (int q = 0; q < P; ++q)
              Oo[(q == P - 1 ? P - 1 + oslot : q) * OPLANE + i] = fma(PMG_MZ(q, P), g[i], fma(PMG_KZ(q, P), m[i], st.acc[r][i][q]));
        }
#pragma unroll
        for (int i = 0; i < NI; ++i) carry[i] = fma(PMG_MZ(P, P), g[i], fma(PMG_KZ(P, P), m[i], st.acc[r][i][P]));
        if (top) {
#pragma unroll
          for (int i = 0; i < NI; ++i) Oo[(NOB - 1) * OPLANE + i] = carry[i];
        }
      }
      if (has_next) {
#pragma unroll
        for (int i = 0; i < NI; ++i) {
#pragma unroll
          for (int q = 0; q < N1; ++q) st.acc[r][i][q] = fma(PMG_MZ(q, 0), g[i], PMG_KZ(q, 0) * m[i]);
          st.acc[r][i][0] += carry[i];
        }
      }
    }
  }
  static PMG_HD void xzsweep(const PmgSweepParams<P> &p, ThreadState &st, const double *Cb, double *Ob, int jz, bool zero,
                             bool has_prev, bool has_next, bool emit, bool top, int oslot)
  {
#pragma unroll
    for (int r = 0; r < IX; ++r) {
      if (st.xi[r] < 0) continue;
      if (XS == 1) xz_item<0, P>(p, st, r, Cb, Ob, jz, zero, has_prev, has_next, emit, top, oslot);
      else if (!(st.xi[r] >> 30 & 1)) xz_item<0, HN>(p, st, r, Cb, Ob, jz, zero, has_prev, has_next, emit, top, oslot);
      else xz_item<(XS == 2 ? HN : 0), P>(p, st, r, Cb, Ob, jz, zero, has_prev, has_next, emit, top, oslot);
    }
  }

  template <int MODE>
  static PMG_HD double epi_value(const PmgSweepParams<P> &p, double y, double uc, double bb, double xo, bool dir, double dinv)
  {
    const double Au = dir ? uc : y; // Dirichlet rows are the identity (:718)
    if (MODE == PMG_MODE_APPLY) return Au;
    if (MODE == PMG_MODE_RESIDUAL) return bb - Au;
    const double corr = p.f2 * (dir ? 1.0 : dinv) * (bb - Au);
    if (MODE == PMG_MODE_CHEB_FIRST) return uc + corr;
    if (MODE == MODE_STEP0) return uc + p.f1 * uc + corr;
    return uc + p.f1 * (uc - xo) + corr;
  }

  // ---- epilogue of dof plane gz, whose A u sits in plane `slot` of the output box; eoff: element offset of the tile's first
  // owned dof in that plane.  DIR: the tile or the plane holds Dirichlet dofs (else no flag is looked at) ------------------
  // ---- epilogue of dof plane gz, whose A u sits in plane `slot` of the output box; eoff: element offset of the tile's first
  // owned dof in that plane.  It comes in two parts: issue() loads b, u, x_old of the thread's dofs into registers before the
  // step's y sweep (and prefetches those of the plane two steps on into L2), finish() uses them after the barrier -- loaded
  // where they are used, the fused step waited for HBM in every plane (ncu: 48 % of the stall samples on these loads).
  // DIR: the tile or the plane holds Dirichlet dofs (else no flag is looked at).
  template <int MODE>
  static PMG_HD void epi_issue_t(const PmgSweepParams<P> &p, const TileGeom &t, ThreadState &st, int gz, int64_t eoff)
  {
    constexpr int NA = n_arrays(MODE);
    if (NA == 0) return;
    constexpr bool has_xo = (MODE == PMG_MODE_CHEB_STEP);
    // per-thread pointers to the thread's first dof, pinned; row k is k row steps away
    const double *pb = p.b + eoff + st.e_off, *pu = p.u + eoff + st.e_off, *px = has_xo ? p.xold + eoff + st.e_off : nullptr;
    PMG_OPAQUE_PTR(pb); PMG_OPAQUE_PTR(pu); PMG_OPAQUE_PTR(px);
    const bool ahead = EPF && (gz + 2 - p.z0 < p.nzl);
    const int64_t two = 2 * t.plane;
    const int rstep = ER * p.Nx;
#pragma unroll
    for (int k = 0; k < NE; ++k) {
      const int g = k * rstep;
      if (st.e_flags >> k & 1u) {
        st.ein[0][k] = pmg_plane_ldg(pb + g);
        if (NA >= 2) st.ein[1][k] = pmg_plane_ldg(pu + g);
        if (NA >= 3 && has_xo) st.ein[2][k] = pmg_plane_ldg(px + g);
        if (ahead) {
          pmg_plane_prefetch(pb + g + two, EPF);
          if (NA >= 2) pmg_plane_prefetch(pu + g + two, EPF);
          if (NA >= 3 && has_xo) pmg_plane_prefetch(px + g + two, EPF);
        }
      }
    }
  }
  static PMG_HD void epi_issue(const PmgSweepParams<P> &p, const TileGeom &t, ThreadState &st, int gz, int64_t eoff)
  {
    switch (mode_of(p)) {
      case PMG_MODE_APPLY: break;
      case PMG_MODE_RESIDUAL: epi_issue_t<PMG_MODE_RESIDUAL>(p, t, st, gz, eoff); break;
      case PMG_MODE_CHEB_FIRST: epi_issue_t<PMG_MODE_CHEB_FIRST>(p, t, st, gz, eoff); break;
      case PMG_MODE_CHEB_STEP: epi_issue_t<PMG_MODE_CHEB_STEP>(p, t, st, gz, eoff); break;
      default: epi_issue_t<MODE_STEP0>(p, t, st, gz, eoff); break;
    }
  }

  template <int MODE, bool DIR>
  static PMG_HD void epilogue_t(const PmgSweepParams<P> &p, const TileGeom &t, const ThreadState &st, const double *smem, int slot,
                                int gz, int64_t eoff, int tid)
  {
    constexpr int NA = n_arrays(MODE);
    const bool dirz = DIR && (gz == t.zlo || gz == t.zhi);
    const int tz = pmg_sweep_pos_type<P>(gz, p.Nz);
    constexpr bool has_xo = (MODE == PMG_MODE_CHEB_STEP);
    const double *Os = smem + O_OFFSET + slot * OPLANE + st.e_src;
    const double *Tz = smem + T_OFFSET + T * T * tz + (st.e_flags >> 24 & 0xFu);
    // plane bases (uniform) + 32-bit element indices (a local vector holds < 2^31 dofs)
    const double *pu = p.u + eoff + st.e_off, *pd = (MODE >= PMG_MODE_CHEB_FIRST && p.dinv_vec) ? p.dinv_vec + eoff + st.e_off : nullptr;
    double *po = p.out + eoff + st.e_off;
    PMG_OPAQUE_PTR(pu); PMG_OPAQUE_PTR(pd); PMG_OPAQUE_PTR(po);
    const int rstep = ER * p.Nx;
#pragma unroll
    for (int k = 0; k < NE; ++k) {
      const int g = k * rstep;
      if (st.e_flags >> k & 1u) {
        const bool dir = DIR && (dirz || (st.e_flags >> (8 + k) & 1u));
        double uc = 0.0, bb = 0.0, xo = 0.0, dinv = 1.0;
        if (NA >= 2) uc = st.ein[1][k];
        else if (dir) uc = pmg_plane_ldg(pu + g);
        if (NA >= 1) bb = st.ein[0][k];
        if (NA >= 3 && has_xo) xo = st.ein[2][k];
        if (MODE >= PMG_MODE_CHEB_FIRST) dinv = pd ? pmg_plane_ldg(pd + g) : Tz[T * (st.e_ty >> (4 * k) & 0xFu)];
        const double y = Os[k * ER * OP];
        pmg_plane_stg(po + g, epi_value<MODE>(p, y, uc, bb, xo, dir, dinv));
      }
    }
    if (DIR && t.xextra && tid < ERW) { // the mesh's last vertex line in x: Dirichlet dofs (identity rows), read directly
      const int gy = t.cy0 * P + tid;
      if (gy < p.Ny && (tid < OY || (!t.virt_y && gy == p.Ny - 1))) {
        const int64_t g = eoff + (int64_t)tid * p.Nx + OX;
        const double uc = pmg_plane_ldg(p.u + g);
        const double bb = (MODE != PMG_MODE_APPLY) ? pmg_plane_ldg(p.b + g) : 0.0;
        const double xo = has_xo ? pmg_plane_ldg(p.xold + g) : 0.0;
        pmg_plane_stg(p.out + g, epi_value<MODE>(p, 0.0, uc, bb, xo, true, 1.0));
      }
    }
  }
  template <int MODE>
  static PMG_HD void epilogue_m(const PmgSweepParams<P> &p, const TileGeom &t, const ThreadState &st, const double *smem, int slot,
                                int gz, int64_t eoff, int tid)
  {
    if (t.dirxy || gz == t.zlo || gz == t.zhi) epilogue_t<MODE, true>(p, t, st, smem, slot, gz, eoff, tid);
    else epilogue_t<MODE, false>(p, t, st, smem, slot, gz, eoff, tid);
  }
  static PMG_HD void epilogue(const PmgSweepParams<P> &p, const TileGeom &t, const ThreadState &st, const double *smem, int slot,
                              int gz, int64_t eoff, int tid)
  {
    switch (mode_of(p)) {
      case PMG_MODE_APPLY: epilogue_m<PMG_MODE_APPLY>(p, t, st, smem, slot, gz, eoff, tid); break;
      case PMG_MODE_RESIDUAL: epilogue_m<PMG_MODE_RESIDUAL>(p, t, st, smem, slot, gz, eoff, tid); break;
      case PMG_MODE_CHEB_FIRST: epilogue_m<PMG_MODE_CHEB_FIRST>(p, t, st, smem, slot, gz, eoff, tid); break;
      case PMG_MODE_CHEB_STEP: epilogue_m<PMG_MODE_CHEB_STEP>(p, t, st, smem, slot, gz, eoff, tid); break;
      default: epilogue_m<MODE_STEP0>(p, t, st, smem, slot, gz, eoff, tid); break;
    }
  }

  // ---- the march of a chunk: steps gz = cz_first P, ... one dof plane each; after the chunk's top vertex plane cz_end P a
  // few steps more without a plane (drain: only the epilogues that are still due) --------------------------------------
  struct March {
    int cz_first, cz_begin, cz_end;
    bool top;
    int cur;          // c, d buffers alternate from plane to plane
    int us;           // ring slot of the current plane
    int gz_last;      // last plane of the march (the chunk's top vertex plane)
    int gz_stop;      // last step of the march: the last output plane leaves there
    int fetch_max;    // last plane that is fetched (the march's last plane, or the one below a Dirichlet top face)
    int q_lo, q_hi;   // output planes of the chunk that this rank owns: [q_lo, q_hi)
    int64_t eq;       // element offset of the tile's first owned dof in plane gz - P - 1 (the plane whose epilogue the step runs)
    unsigned base32;  // 32-bit shared-window address of the CTA's shared memory
  };
  // The output plane whose epilogue runs in the step of plane index gz, and its place in the output box: plane q leaves P + 1
  // steps after its own step (a cell layer closes at the vertex step of the layer above, its planes leave one per step; the
  // last one at the next vertex step, from the copy of box plane P - 1 that is not being rewritten).  The mesh's top plane
  // leaves in the march's last step.  P = 1: two steps after, from the copy that is not being written.
  struct Plan { int n; int q[2]; int os[2]; };
  static PMG_HD Plan plan_of(const March &mc, int gz)
  {
    Plan pl; pl.n = 0;
    const int q = gz - P - 1;
    if (q >= mc.q_lo && q < mc.q_hi) {
      const int lay = q / P, k = q - lay * P;
      pl.q[0] = q; pl.os[0] = (k == P - 1) ? P - 1 + ((lay + 1) & 1) : k; pl.n = 1;
    }
    if (mc.top && gz == mc.gz_stop) { pl.q[pl.n] = mc.cz_end * P; pl.os[pl.n] = NOB - 1; ++pl.n; }
    return pl;
  }

  // ---- one cell layer: its P steps.  GEN = false: an interior layer of the chunk -- the layer below closes into the output
  // box and is written, no plane is a Dirichlet face, every fetched plane exists: the step's flags are compile-time constants.
  template <bool GEN, class Exec>
  static PMG_HD void layer(const PmgSweepParams<P> &p, const TileGeom &t, Exec &ex, double *smem, March &mc, int L)
  {
    double *Ub = smem + U_OFFSET, *Cb = smem + C_OFFSET, *Ob = smem + O_OFFSET;
    // the layer below (L - 1) closes at this layer's vertex plane; it is written iff it belongs to the chunk
    const bool prev_layer = L > mc.cz_first, emit_layer = L - 1 >= mc.cz_begin;
#pragma unroll(UZ ? P : 1)
    for (int jz = 0; jz < P; ++jz) {
      const int gz = L * P + jz;
      if (GEN && gz > mc.gz_stop) break;
      const bool last = GEN && (gz == mc.gz_last);
      const bool drain = GEN && (gz > mc.gz_last);
      const bool zero = GEN && (gz == t.zlo || gz == t.zhi);
      // fetch plane gz + ND into the slot the previous step's y sweep read, unless it is beyond the march or a Dirichlet face
      const bool fetch = gz + ND <= mc.fetch_max;
      int fs = mc.us + ND; if (fs >= NU) fs -= NU;
      double *Ucur = Ub + mc.us * UPLANE;
      double *Ccur = Cb + mc.cur * 2 * CPLANE;
      Plan pl;
      if (GEN) pl = plan_of(mc, gz);
      else { // the plane P + 1 steps back, if the chunk writes it: from the box plane it was put in when its layer closed
        const int q = gz - P - 1;
        pl.n = (q >= mc.q_lo && q < mc.q_hi) ? 1 : 0; pl.q[0] = q; pl.q[1] = 0; pl.os[1] = 0;
        pl.os[0] = (P == 1) ? ((L - 1) & 1) : (jz == 0) ? P - 1 + ((L - 1) & 1) : jz - 1;
      }
      ex.for_each_thread([&](int, ThreadState &st) {
        load_plane(p, st, t.plane, smem, mc.base32, fs, fetch);
        if (pl.n > 0 && !drain) epi_issue(p, t, st, pl.q[0], mc.eq);
        if (!zero && !drain) ysweep(p, st, Ucur, Ccur);
        load_wait();
      });
      ex.sync();
      ex.for_each_thread([&](int tid, ThreadState &st) {
        const bool has_prev = (jz == 0) && prev_layer;
        if (!drain) xzsweep(p, st, Ccur, Ob, jz, zero, has_prev, !last, has_prev && emit_layer, last && mc.top, L & 1);
#pragma unroll
        for (int i = 0; i < 2; ++i)
          if (i < pl.n) {
            const int64_t eo = mc.eq + (int64_t)(pl.q[i] - (gz - P - 1)) * t.plane;
            if (drain || i > 0) epi_issue(p, t, st, pl.q[i], eo); // no sweep to hide behind (or the second plane of the last step)
            epilogue(p, t, st, smem, pl.os[i], pl.q[i], eo, tid);
          }
      });
      mc.cur ^= 1;
      mc.us = (mc.us + 1 == NU) ? 0 : mc.us + 1;
      mc.eq += t.plane;
    }
  }

  // ---- the tile program -------------------------------------------------------------------------------------------
  // Exec provides: template<F> void for_each_thread(F f)  with f(int tid, ThreadState&);  void sync()
  template <class Exec>
  static PMG_HD void run(const PmgSweepParams<P> &p, Exec &ex, double *smem, int tile_x, int tile_y, int chunk)
  {
    const TileGeom t = geom(p, tile_x, tile_y);
    March mc;
    mc.cz_begin = p.cz_lo + chunk * p.layers_per_chunk;
    mc.cz_end = mc.cz_begin + p.layers_per_chunk;
    if (mc.cz_end > p.cz_hi) mc.cz_end = p.cz_hi;
    if (mc.cz_begin >= mc.cz_end) return;
    // the layer below the chunk is recomputed (its sums close the chunk's first plane) when its planes are stored locally
    const bool halo = (mc.cz_begin > 0) && ((mc.cz_begin - 1) * P >= p.z0);
    mc.cz_first = halo ? mc.cz_begin - 1 : mc.cz_begin;
    mc.top = (mc.cz_end == p.cz_hi) && (mc.cz_end * P < p.z_own_hi); // the mesh's top plane belongs to this chunk
    const int gz_first = mc.cz_first * P;
    mc.cur = 0; mc.us = 0;
#if defined(__CUDA_ARCH__)
    mc.base32 = (unsigned)__cvta_generic_to_shared(smem);
    PMG_KEEP(mc.base32);
#else
    mc.base32 = 0;
#endif
    mc.gz_last = mc.cz_end * P;
    mc.gz_stop = mc.gz_last + P; // the last layer's plane P - 1 leaves P + 1 steps after its own
    mc.fetch_max = (t.zhi >= 0 && t.zhi <= mc.gz_last) ? t.zhi - 1 : mc.gz_last;
    mc.q_lo = mc.cz_begin * P > p.z_own_lo ? mc.cz_begin * P : p.z_own_lo;
    mc.q_hi = mc.cz_end * P < p.z_own_hi ? mc.cz_end * P : p.z_own_hi;
    mc.eq = (int64_t)(gz_first - P - 1 - p.z0) * t.plane + t.tile0;
    const double *up = p.u + (int64_t)(gz_first - p.z0) * t.plane;

    // Fused ghost exchange: the chunk that reaches the slab's upper ghost plane needs the upper neighbour's word that it has
    // pushed it (and has stopped reading the ghost planes this CTA pushes into after its march) only before the cell layer in
    // whose steps that plane is fetched -- the last or last but one of the march.  The wait sits between two layers, never
    // inside the step loop (a possible barrier there keeps the compiler from scheduling loads across it).
    int l_wait = -1; // the layer before which to wait; -1: no wait
    if constexpr (PUSH != 0)
      if (mc.gz_last >= p.z_own_hi && p.z_own_hi < p.Nz && (p.consume & 1) && p.mb_hi) {
        l_wait = (p.z_own_hi - ND) / P;
        if (l_wait <= mc.cz_first) { ex.wait_flag(p.mb, PMG_FUSED_FROM_HI, PMG_FUSED_EPOCH_HI); l_wait = -1; }
      }
    ex.for_each_thread([&](int tid, ThreadState &st) { decode(p, t, tid, st, smem); st.ld_src = up + st.ld_off; });
    ex.sync();
    ex.for_each_thread([&](int, ThreadState &st) {
      // planes gz_first .. gz_first + ND - 1 -> slots 0 .. ND - 1, a group each; the first plane has arrived after the wait
#pragma unroll
      for (int d = 0; d < ND; ++d) {
        const int gz = gz_first + d;
        load_plane(p, st, t.plane, smem, mc.base32, d, gz <= mc.fetch_max && gz != t.zlo);
      }
      load_wait();
    });
    ex.sync();
    // Layers without a Dirichlet plane, before the chunk's top vertex plane: the steps look at a handful of uniform flags only
    // (GEN = false).  The others -- the mesh's bottom layer under a Dirichlet face, the top vertex plane and the drain steps
    // after it -- take the general path.
    for (int L = mc.cz_first; L * P <= mc.gz_stop; ++L) {
      if constexpr (PUSH != 0)
        if (L == l_wait) ex.wait_flag(p.mb, PMG_FUSED_FROM_HI, PMG_FUSED_EPOCH_HI);
      const bool fast = L < mc.cz_end && L * P > t.zlo && (t.zhi < 0 || L * P + P - 1 < t.zhi);
      if (fast) layer<false>(p, t, ex, smem, mc, L);
      else layer<true>(p, t, ex, smem, mc, L);
    }
  }
  // ---- fused ghost push: the tile's part of the slab's boundary planes, as just written to p.out, goes to the neighbours ------
  // (called by the CTAs of the bottom / top chunk after their march, csrc/pmg_apply_plane_launch.h: the march itself carries no
  // push code -- a first version stored from the epilogue and cost the fused modes 5-10 % in registers and instruction cache)
  template <class Exec>
  static PMG_HD void push_boundary(const PmgSweepParams<P> &p, Exec &ex, int tile_x, int tile_y, bool at_lo, bool at_hi)
  {
    const int x0 = tile_x * OX, y0 = tile_y * OY;
    const int x1 = (tile_x == p.tiles_x - 1 || x0 + OX > p.Nx) ? p.Nx : x0 + OX; // the last tile also holds the mesh's last lines
    const int y1 = (tile_y == p.tiles_y - 1 || y0 + OY > p.Ny) ? p.Ny : y0 + OY;
    if (x0 >= x1 || y0 >= y1) return;
    const int w = x1 - x0, n = w * (y1 - y0);
    const int64_t plane = (int64_t)p.Nx * p.Ny;
    ex.sync(); // the CTA's stores to p.out are visible to all of its threads
    ex.for_each_thread([&](int tid, ThreadState &) {
      if (at_lo && p.push_lo) { // plane z_own_lo: the lower neighbour's upper ghost plane
        const int64_t base = (int64_t)(p.z_own_lo - p.z0) * plane;
        for (int e = tid; e < n; e += NT) {
          const int r = e / w, c = e - r * w;
          const int64_t g = base + (int64_t)(y0 + r) * p.Nx + x0 + c;
          p.push_lo[g] = p.out[g];
        }
      }
      if (at_hi && p.push_hi) { // planes [z_own_hi - P, z_own_hi): the upper neighbour's lower ghost planes
        for (int q = p.z_own_hi - P; q < p.z_own_hi; ++q) {
          const int64_t base = (int64_t)(q - p.z0) * plane;
          for (int e = tid; e < n; e += NT) {
            const int r = e / w, c = e - r * w;
            const int64_t g = base + (int64_t)(y0 + r) * p.Nx + x0 + c;
            p.push_hi[g] = p.out[g];
          }
        }
      }
    });
  }
#undef PMG_M
#undef PMG_KX
#undef PMG_KY
#undef PMG_MZ
#undef PMG_KZ
};
