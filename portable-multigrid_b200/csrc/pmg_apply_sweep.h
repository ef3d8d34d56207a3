// pmg_apply_sweep.h -- the Laplace cell-loop apply as three line-marching sweeps (kernel K1, v3),
// fused with the smoother update.
//
// Replaces LaplaceOperator::vmult + LocalLaplaceOperator::operator()
// (reference include/operators/portable_laplace_operator.h:227-381, 557-719) and the
// PreconditionChebyshev vector update that follows it in the reference
// (include/multigrid/portable_v_cycle_multigrid.h:116-125) by ONE pass over HBM.
//
// Design (B200-first, not a translation; DESIGN.md "Kernel K1"):
//  * On the affine box mesh the reference's cell integral factorises exactly (QGauss(p+1) integrates the
//    degree-2p mass and degree-2p-2 stiffness integrands of FE_Q(p) exactly), so the sum over cells is
//        A = Kx (x) My (x) Mz + Mx (x) Ky (x) Mz + Mx (x) My (x) Kz
//    with the 1-D *cell* matrices M, K (p+1 x p+1) applied cell by cell along lines -- the same 1-D
//    sum-factorisation sweeps as the reference, but only 7 of them and on p^2 (not (p+1)^2) lines per cell:
//        c = My u, d = Ky u;   g = Kx c + Mx d, m = Mx c;   A u = Mz g + Kz m.
//    FP64 is the scarce resource on B200 (34 TFLOP/s measured): this needs 7 (p+1)^2 / p FMAs per DoF
//    (Q4: 44) against 12 (p+1)^4 / p^3 (Q4: 117) for the reference's sweeps.
//  * A 1-D sweep is done by a thread that MARCHES along its line, cell after cell, carrying the partial sum
//    of the vertex DoF shared with the next cell in a register: complete results, no cross-thread sums, no
//    atomics, and the halo a CTA recomputes for its neighbours shrinks to one partial cell per line.
//  * CTA = BX x BY cell columns, marching through cell layers in z, LZ layers (LZ P dof planes) per step:
//      stage    u reaches shared memory through the bulk async copy engine (cp.async.bulk, SASS UBLKCP, completion on an
//               mbarrier): one instruction per tile row, issued a full step ahead into one of two staging buffers; the
//               epilogue's b and x_old rows arrive by cp.async (LDGSTS) issued by the warp that phase 1 leaves idle; no
//               thread spends registers on loads.  The engine wants 16-byte aligned rows while dof rows start at any 8-byte address (odd row
//               lengths): a row is fetched from the aligned address below it and read back through a one-element shift
//               that depends on the row.  (2-D TMA tensor tiles -- one instruction per plane -- would be the better fit,
//               but UTMALDG raises "illegal instruction" on this pool's B200s even for the CUDA programming guide's
//               reference example: tools/exp/tma_test2.cu.)
//      phase 1  thread = (x, plane): marches in y: u (staging buffer, left intact) -> c, d
//      phase 2  thread = (y, plane): marches in x, in place: (c, d) -> (g, m)
//      phase 3  thread = dof column (x, y): z sweep over the layer's P+1 planes -- the values of the plane shared
//               with the layer below and its partial sum stay in registers (3 doubles per column) -- and the
//               epilogue (Dirichlet identity / residual / Chebyshev update; u, b, x_old from shared memory) before
//               the single coalesced store.
//    Every shared-memory access is conflict-free (odd row pitch of the work buffers, lanes along x or along y).
//  * Vectors must be 16-byte aligned with 16 readable bytes after their last element (the library's allocator
//    provides both).
//  * The 1-D matrices are symmetric and centro-symmetric: the kernel addresses them through the canonical
//    representative of each entry, so a phase needs only ~(p+1)^2/2 distinct constants, which stay in uniform
//    registers instead of being re-fetched for every FMA.
//  * Owner-computes: every owned DoF's A u is complete inside the CTA.
//
// Written against an executor (for_each_thread / sync) so the same source is the CUDA kernel
// (csrc/pmg_apply.cu) and runs thread by thread under the host emulator of the CPU test-suite
// (tests/emu/emu_apply.cpp).  The emulator is test infrastructure, not a fallback.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define PMG_HD __host__ __device__ __forceinline__
#else
#define PMG_HD inline
#endif

#ifndef PMG_APPLY_MODES_DEFINED
#define PMG_APPLY_MODES_DEFINED
enum PmgApplyMode {
  PMG_MODE_APPLY = 0,     // out = A u
  PMG_MODE_RESIDUAL = 1,  // out = b - A u
  PMG_MODE_CHEB_FIRST = 2,// out = u + f2 * Dinv (b - A u)                       (first step of smooth())
  PMG_MODE_CHEB_STEP = 3, // out = u + f1 (u - xold) + f2 Dinv (b - A u); xold may alias out; xold==NULL => 0
};
#endif

// ---- bulk async copy engine + mbarrier (sm_90+ PTX; plain copies / no-ops under the host emulator, which runs the
// issuing loops of all threads to completion before any consumer) -----------------------------------------------------
PMG_HD void pmg_mbar_init(uint64_t *bar, int count)
{
#if defined(__CUDA_ARCH__)
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
#else
  (void)bar; (void)count;
#endif
}
PMG_HD void pmg_mbar_init_fence()
{
#if defined(__CUDA_ARCH__)
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
#endif
}
// one arrival that also announces `bytes` of bulk copies completing on the barrier
PMG_HD void pmg_mbar_arrive_expect(uint64_t *bar, int bytes)
{
#if defined(__CUDA_ARCH__)
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
#else
  (void)bar; (void)bytes;
#endif
}
PMG_HD void pmg_mbar_wait(uint64_t *bar, int parity)
{
#if defined(__CUDA_ARCH__)
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  unsigned ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
  } while (!ok);
#else
  (void)bar; (void)parity;
#endif
}
// 8-byte asynchronous global -> shared copy (LDGSTS) and the wait for all of the thread's copies
PMG_HD void pmg_sweep_cp_async8(double *dst_smem, const double *src_global)
{
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src_global) : "memory");
#else
  *dst_smem = *src_global;
#endif
}
PMG_HD void pmg_sweep_cp_async_wait_all()
{
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.wait_all;\n" ::: "memory");
#endif
}

// copy `bytes` (multiple of 16) from 16-byte aligned global memory to 16-byte aligned shared memory
PMG_HD void pmg_bulk_copy(double *dst_smem, const double *src_global, int bytes, uint64_t *bar)
{
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src_global), "r"(bytes),
                 "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
#else
  (void)bar;
  for (int i = 0; i < bytes / 8; ++i) dst_smem[i] = src_global[i];
#endif
}

template <int P>
struct PmgSweepParams {
  // level geometry (global)
  int nx, ny, nz;      // cells per direction
  int Nx, Ny, Nz;      // dofs per direction
  unsigned faces;      // Dirichlet faces bitmask (bit 2d low, 2d+1 high)
  // this rank's slab of the vector: local plane l holds global plane z0 + l
  int z0, nzl;
  int cz_lo, cz_hi;    // owned cell layers [cz_lo, cz_hi)
  int z_own_lo, z_own_hi; // owned dof planes [z_own_lo, z_own_hi)
  // decomposition of the launch
  int tiles_x, tiles_y, n_chunks, layers_per_chunk;
  // 1-D cell matrices, row-major (P+1)^2, symmetric and centro-symmetric (read through canon()).  With
  // cx = hy hz / hx, cy = hx hz / hy, cz = hx hy / hz:  A u = Mz [ Kx My u + Mx Ky u ] + Kz Mx My u  where
  // M = Mref (y and x sweeps), Kx = Kref, Ky = (cy / cx) Kref, Mz = cx Mref, Kz = cz Kref.
  double M[(P + 1) * (P + 1)], Kx[(P + 1) * (P + 1)], Ky[(P + 1) * (P + 1)];
  double Mz[(P + 1) * (P + 1)], Kz[(P + 1) * (P + 1)];
  // epilogue
  int mode;
  const double *u;
  const double *b;
  const double *xold;
  double *out;
  double f1, f2;
  const double *dinv_vec; // explicit inverse diagonal, or NULL => table
  const double *dinv_tab; // (P+2)^3 table indexed by 1-D position types
  // Fused ghost push (plane-per-step kernel; NULL = none): the slab neighbours' copies of `out`, mapped over NVLink.  The
  // epilogue stores plane z_own_lo also into push_lo (the lower neighbour holds it as its upper ghost plane) and planes
  // [z_own_hi - P, z_own_hi) into push_hi (the upper neighbour's lower ghost planes), both indexed like `out`: the launcher
  // folds the difference of the slabs' first stored planes into the pointers.  The neighbours' next apply then finds its
  // ghost planes filled and needs no exchange (reference: update_ghost_values before every vmult, :661).
  double *push_lo, *push_hi;
  // flag words that order the fused pushes (csrc/pmg_apply_plane_launch.h): this rank's mailbox and the neighbours'
  unsigned long long *mb, *mb_lo, *mb_hi;
  int consume; // u's ghost planes were pushed by the neighbours' previous fused launch: the boundary chunks wait for its flag
};

// fill the five 1-D matrices from the reference-cell pencil (Mref, Kref) and the cell sizes h; entries related by the
// matrices' symmetries are averaged so that every representative carries the same value
template <int P>
inline void pmg_sweep_fill_matrices(PmgSweepParams<P> &p, const double *Mref, const double *Kref, const double *h)
{
  constexpr int N1 = P + 1;
  const double cx = h[1] * h[2] / h[0], cy = h[0] * h[2] / h[1], cz = h[0] * h[1] / h[2];
  for (int i = 0; i < N1; ++i)
    for (int j = 0; j < N1; ++j) {
      const int a = i * N1 + j, b = j * N1 + i, c = (P - i) * N1 + (P - j), d = (P - j) * N1 + (P - i);
      const double m = 0.25 * (Mref[a] + Mref[b] + Mref[c] + Mref[d]);
      const double k = 0.25 * (Kref[a] + Kref[b] + Kref[c] + Kref[d]);
      p.M[a] = m; p.Kx[a] = k; p.Ky[a] = (cy / cx) * k; p.Mz[a] = cx * m; p.Kz[a] = cz * k;
    }
}

template <int P>
PMG_HD int pmg_sweep_pos_type(int g, int N)
{
  return (g == 0) ? P : (g == N - 1) ? P + 1 : g % P;
}

// canonical representative of entry (i, j) under (i,j) ~ (j,i) ~ (P-i,P-j) ~ (P-j,P-i)
template <int P>
PMG_HD constexpr int pmg_sweep_canon(int i, int j)
{
  int m = i * (P + 1) + j;
  const int b = j * (P + 1) + i, c = (P - i) * (P + 1) + (P - j), d = (P - j) * (P + 1) + (P - i);
  if (b < m) m = b;
  if (c < m) m = c;
  if (d < m) m = d;
  return m;
}

// P: degree; BX x BY: cell columns per CTA; LZ: cell layers per step; NT_: threads; US: 1 = in APPLY mode every second
// row of the u tile is fetched by the loader warp's cp.async instead of the copy engine (measured: +5 % at Q4, +2 % at Q2,
// -4 % at Q3; in the fused modes the loader warp is busy with the b / x_old rows and the split costs 15 %)
// FM: -1 = the epilogue mode is the launch parameter p.mode; 0..3 = the kernel is compiled for that one mode and drops the
// other three epilogue variants from its code (measured on B200, gpurun_out/exp17: apply +4 % at Q4, +10 % at Q2 / Q3,
// +30 % at Q5 where it frees 48 registers; fused step +1..5 %)
// SG: the y lines of phase 1 and the x lines of phase 2 are cut into SG segments of cells, each marched by its own thread (a
// segment starts from the partial cell before it, as the tile's first segment starts from the halo cell): SG x the work items
// in the two phases that fill only half of the CTA, for one partial cell more per line and segment
// RL: 1 = the cell loops of phases 1 and 2 stay rolled (one copy of the cell body instead of BY resp. BX copies): the
// steady-state loop of the Q4 apply kernel is 67 KB of code unrolled (tools/sass_regions.py), more than the instruction
// caches hold, and ncu attributes 10-13 % of the stall samples to instruction fetch (profiles/r01_v5_apply_q4_ncu.md)
// A2: 1 = the phase-2 items sit on the lower half of the CTA's threads in even steps and on the upper half in odd steps.  With
// 4 warps per CTA warp w runs on SM sub-partition w; phases 1 and 2 otherwise both load warps 0 and 1, whose FP64 pipes then
// carry 2.5 x the work of sub-partition 2 and 5.7 x that of sub-partition 3 (DESIGN.md "sub-partition balance")
// EG: 1 = the fused modes read b and x_old straight from global memory in the z sweep (the loads of a dof column's planes are
// issued before its z-sweep arithmetic) instead of staging them in shared memory: 18.5 KB less per CTA at Q4, and the loader warp
// is free for u rows.  Round 1 measured direct epilogue loads slower at equal occupancy; the option exists for the configurations
// where it buys a resident CTA (4 per SM for the fused step, 2 per SM for the pipelined variant).  Not yet measured.
template <int P, int BX, int BY, int LZ, int NT_, int US = 0, int FM = -1, int SG = 1, int RL = 0, int A2 = 0, int EG = 0>
struct PmgSweepTile {
  static constexpr int N1 = P + 1;
  static PMG_HD int mode_of(const PmgSweepParams<P> &p) { return FM >= 0 ? FM : p.mode; }
  static constexpr int NT = NT_;
  static constexpr int NPS = LZ * P;          // dof planes per step
  static constexpr int XW = (BX + 1) * P + 1; // x points of the tile: global x = (cx0-1) P + xl
  static constexpr int YW = (BY + 1) * P + 1; // rows of the u tile:   global y = (cy0-1) P + yl
  static constexpr int RW = BY * P + 1;       // owned rows:           global y = cy0 P + r
  static constexpr int CW = BX * P + 1;       // owned dof columns in x (the +1 only at the mesh end)
  // staging buffers (bulk-copy targets): a row is fetched from the 16-byte aligned address at or below its first
  // element, so it holds one extra element and its pitch is even; element x of row (k, y) is at [x + shift(k, y)]
  static constexpr int XPA = (XW | 1) + 1;    // elements per copied u row = pitch of A
  static constexpr int XPE = CW | 1;          // pitch of E: b / x_old rows of the owned columns (filled with cp.async)
  static constexpr int APLANE = YW * XPA, EPLANE = RW * XPE;
  static constexpr int ABUF = NPS * APLANE;   // one u staging buffer
  static constexpr int EBUF = NPS * EPLANE;   // one epilogue array (b or x_old)
  // work buffers C (c, then g) and D (d, then m): odd pitch, the column walks of phase 2 are conflict-free
  static constexpr int XP = XW | 1;
  static constexpr int CPLANE = RW * XP;
  static constexpr int CBUF = (NPS * CPLANE + 15) & ~15;
  static constexpr int C_OFFSET = 2 * ABUF, D_OFFSET = C_OFFSET + CBUF;
  static constexpr int E_OFFSET = D_OFFSET + CBUF; // b boxes, then x_old boxes (not allocated for APPLY)
  static constexpr int BAR_OFFSET_APPLY = E_OFFSET, BAR_OFFSET_EPI = E_OFFSET + 2 * EBUF;
  static constexpr int NBAR = 2;              // mbarriers: u buffer 0, u buffer 1
  // E loader: the CTA's last warp (idle in phase 1: NITEM1 <= NT - 32 for the shipped tiles), LD_LPR lanes per row
  static constexpr int LD_LPR = (CW - 1 <= 16) ? 16 : 32;
  static constexpr int LD_RPI = 32 / LD_LPR;  // rows per warp instruction
  static PMG_HD constexpr int smem_doubles(bool epilogue_inputs) { return ((epilogue_inputs && !EG) ? BAR_OFFSET_EPI : BAR_OFFSET_APPLY) + NBAR; }
  static constexpr int SMEM_DOUBLES = (EG ? BAR_OFFSET_APPLY : BAR_OFFSET_EPI) + NBAR;
  static constexpr int BXS = (BX + SG - 1) / SG, BYS = (BY + SG - 1) / SG; // cells per segment
  static constexpr int NITEM1 = XW * NPS * SG;
  static constexpr int IT2 = (RW * NPS * SG + NT - 1) / NT; // phase-2 items per thread
  static constexpr int NT2 = ((RW * NPS * SG + 31) / 32 * 32 < NT) ? (RW * NPS * SG + 31) / 32 * 32 : NT; // threads that can hold a phase-2 item in round 0
  static_assert(SG >= 1 && SG <= 2 && SG <= BX && SG <= BY, "one or two segments per line (seed2 holds one hand-over)");
  static_assert(!(A2 && SG > 1), "A2 moves the phase-2 items off the first NT2 threads, which is where the SG > 1 barrier expects them");
  static constexpr int NCOL = (CW * RW + NT - 1) / NT; // dof columns per thread in phase 3
  static_assert(NT % 32 == 0, "whole warps: the last one is the cp.async loader");
  static_assert(CW <= 255 && RW <= 255, "column coordinates are packed into 8 bits each (ThreadState::info)");
  static_assert(P + 2 <= 15, "position types are packed into 4 bits each");
  static_assert(SMEM_DOUBLES * 8 <= 227 * 1024, "tile does not fit the 227 KB of shared memory of a CTA");

#define PMG_M(i, j) p.M[pmg_sweep_canon<P>(i, j)]
#define PMG_KX(i, j) p.Kx[pmg_sweep_canon<P>(i, j)]
#define PMG_KY(i, j) p.Ky[pmg_sweep_canon<P>(i, j)]
#define PMG_MZ(i, j) p.Mz[pmg_sweep_canon<P>(i, j)]
#define PMG_KZ(i, j) p.Kz[pmg_sweep_canon<P>(i, j)]

  struct ThreadState {
    double carry[NCOL]; // z-sweep sum of the plane shared with the next cell layer, per owned dof column
    double gP[NCOL], mP[NCOL]; // (g, m) of that plane
    double uP[NCOL];    // u of that plane (epilogue input of the next layer's plane 0)
    double dinv[NCOL][P]; // inverse diagonal of the column's planes k = 0..P-1 of an interior layer (CHEB modes, table)
    int info[NCOL];     // ox | oy << 8 | Dirichlet-in-xy << 16 | x position type << 20 | y position type << 24; -1 = none
    int item2[A2 ? 2 : 1][IT2]; // phase-2 item: row | plane << 16 | segment << 24; -1 = none ([1]: the odd steps' item when A2)
    double seed2[SG > 1 ? IT2 : 1][4]; // SG > 1: what a phase-2 item reads before the in-place sweep starts (phase2_seed)
  };

  struct TileGeom {
    int cx0, cy0;     // first owned cell column
    int ncx, ncy;     // valid owned cells
    int x_end, y_end; // tile touches the high end of the mesh (owns the last vertex line)
    int cw, rows;     // owned dof columns / rows
    int nxodd, plodd; // parities of Nx and of Nx * Ny: how the row shift changes from row to row / plane to plane
    int64_t eA;       // element index (local vector) of tile point (xl=0, yl=0) in local plane 0
    int64_t n_local;  // elements of the local vector
    bool has_e, has_xo; // the mode reads b / x_old
  };

  static PMG_HD TileGeom geom(const PmgSweepParams<P> &p, int tile_x, int tile_y)
  {
    TileGeom t;
    t.cx0 = tile_x * BX; t.cy0 = tile_y * BY;
    t.ncx = (p.nx - t.cx0 < BX) ? p.nx - t.cx0 : BX;
    t.ncy = (p.ny - t.cy0 < BY) ? p.ny - t.cy0 : BY;
    t.x_end = (t.cx0 + t.ncx == p.nx);
    t.y_end = (t.cy0 + t.ncy == p.ny);
    t.cw = t.ncx * P + t.x_end;
    t.rows = t.ncy * P + t.y_end;
    t.nxodd = p.Nx & 1; t.plodd = (p.Nx & 1) & (p.Ny & 1);
    t.eA = (int64_t)((t.cy0 - 1) * P) * p.Nx + (t.cx0 - 1) * P;
    t.n_local = (int64_t)p.Nx * p.Ny * p.nzl;
    t.has_e = (mode_of(p) != PMG_MODE_APPLY) && !EG; // b / x_old staged in shared memory
    t.has_xo = (mode_of(p) == PMG_MODE_CHEB_STEP) && (p.xold != nullptr);
    return t;
  }

  // shift of the tile row that starts at element e of the vector (e may be negative: two's complement keeps parity)
  static PMG_HD int shift_of(int64_t e) { return (int)(e & 1); }

  static PMG_HD void decode(const PmgSweepParams<P> &p, const TileGeom &t, int tid, ThreadState &st)
  {
    constexpr int T = P + 2;
    const int ncols = t.cw * t.rows;
#pragma unroll
    for (int ci = 0; ci < NCOL; ++ci) {
      const int col = tid + ci * NT;
      st.carry[ci] = 0.0; st.gP[ci] = 0.0; st.mP[ci] = 0.0; st.uP[ci] = 0.0;
#pragma unroll
      for (int k = 0; k < P; ++k) st.dinv[ci][k] = 1.0;
      if (col < ncols) {
        const int oy = col / t.cw, ox = col - oy * t.cw;
        const int gx = t.cx0 * P + ox, gy = t.cy0 * P + oy;
        const int dirxy = (gx == 0 && (p.faces & 1u)) || (gx == p.Nx - 1 && (p.faces >> 1 & 1u)) ||
                          (gy == 0 && (p.faces >> 2 & 1u)) || (gy == p.Ny - 1 && (p.faces >> 3 & 1u));
        const int tx = pmg_sweep_pos_type<P>(gx, p.Nx), ty = pmg_sweep_pos_type<P>(gy, p.Ny);
        st.info[ci] = ox | (oy << 8) | (dirxy << 16) | (tx << 20) | (ty << 24);
        if (mode_of(p) >= PMG_MODE_CHEB_FIRST && !p.dinv_vec && !dirxy) {
#pragma unroll
          for (int k = 0; k < P; ++k) st.dinv[ci][k] = p.dinv_tab[tx + T * ty + T * T * k];
        }
      } else {
        st.info[ci] = -1;
      }
    }
#pragma unroll
    for (int a = 0; a < (A2 ? 2 : 1); ++a) {
      const int tid2 = (a == 0) ? tid : (tid + NT / 2) % NT; // odd steps: the other half of the CTA's warps goes first
#pragma unroll
      for (int i = 0; i < IT2; ++i) { // rows fastest: the lanes of a warp walk down a column of the buffer
        const int item = tid2 + i * NT;
        const int kk = item / t.rows;
        const int seg = kk / NPS, k = kk - seg * NPS;
        st.item2[a][i] = (seg < SG) ? ((item - kk * t.rows) | (k << 16) | (seg << 24)) : -1;
      }
    }
  }

  // ---- stage: one bulk copy per tile row.  Every thread issues its share and arrives once on the barrier ---------
  // copy `nelem` (even) elements around vector elements [e, e + nelem - 1) to the 16-byte aligned row `dst`
  static PMG_HD int issue_row(const double *vec, int64_t e, int nelem, int64_t n_local, double *dst, uint64_t *bar)
  {
    int64_t lo = e - shift_of(e);          // 16-byte aligned start (even element index)
    if (lo < 0) {                          // below the vector: only points x < 0 of the first row, never read
      const int sk = (int)(-lo);
      lo += sk; dst += sk; nelem -= sk;
    }
    if (lo + nelem > n_local + 2)          // beyond the vector's 16 bytes of padding: only x > Nx-1 of the last row
      nelem = (int)((n_local + 2 - lo) & ~(int64_t)1);
    if (nelem <= 0) return 0;
    pmg_bulk_copy(dst, vec + lo, nelem * 8, bar);
    return nelem * 8;
  }

  // which rows of the u tile the loader warp fetches with cp.async instead of the copy engine
  static PMG_HD bool row_by_loader(const TileGeom &t, int yl) { return US == 1 && !t.has_e && (yl & 1); }

  // the loader warp's share of the u planes gz0 .. gz0+npl-1: lane = x point, same row layout (shift) as the bulk copies
  static PMG_HD void load_u_async(const PmgSweepParams<P> &p, const TileGeom &t, int lane, double *A, int gz0, int npl)
  {
    const int64_t plane = (int64_t)p.Nx * p.Ny;
    const int yl_lo = (t.cy0 == 0) ? P : 0, yl_hi = (t.ncy + 1) * P;
    const int y0 = yl_lo | 1; // first odd row
    if (!row_by_loader(t, y0)) return;
#pragma unroll
    for (int xl = lane; xl < XW; xl += 32) {
      const int gx = (t.cx0 - 1) * P + xl;
      if (gx < 0 || gx >= p.Nx) continue;
#pragma unroll
      for (int k = 0; k < NPS; ++k) {
        if (k >= npl) break;
        int64_t e = t.eA + (int64_t)(gz0 + k - p.z0) * plane + (int64_t)y0 * p.Nx; // element index of the row's xl = 0
        double *dst = A + k * APLANE + y0 * XPA + xl;
#if defined(PMG_SWEEP_LEAN_LOADER)
        // experiment switch (tools/exp): the shift of a row equals that of the row two below it (2 Nx is even), so it and the
        // two pointers leave the row loop
        const double *src = p.u + e + xl;
        const int64_t sstep = 2 * (int64_t)p.Nx;
        dst += shift_of(e);
        for (int yl = y0; yl <= yl_hi; yl += 2) {
          pmg_sweep_cp_async8(dst, src);
          src += sstep; dst += 2 * XPA;
        }
#else
        for (int yl = y0; yl <= yl_hi; yl += 2) {
          pmg_sweep_cp_async8(dst + shift_of(e), p.u + e + xl);
          e += 2 * (int64_t)p.Nx; dst += 2 * XPA;
        }
#endif
      }
    }
  }

  // u planes gz0 .. gz0+npl-1 (tile footprint XW x YW) -> A
  static PMG_HD void stage_u(const PmgSweepParams<P> &p, const TileGeom &t, int tid, double *A, uint64_t *bar, int gz0, int npl)
  {
    const int64_t plane = (int64_t)p.Nx * p.Ny;
    const int yl_lo = (t.cy0 == 0) ? P : 0;   // rows below the mesh do not exist
    const int nrow = (t.ncy + 1) * P + 1 - yl_lo;
    int bytes = 0;
    for (int i = tid; i < npl * nrow; i += NT) {
      const int k = i / nrow, yl = yl_lo + (i - k * nrow);
      if (row_by_loader(t, yl)) continue;
      bytes += issue_row(p.u, t.eA + (int64_t)(gz0 + k - p.z0) * plane + (int64_t)yl * p.Nx, XPA, t.n_local,
                         A + k * APLANE + yl * XPA, bar);
    }
#if defined(__CUDA_ARCH__) && defined(PMG_SWEEP_WARP_ARRIVE)
    // experiment switch (tools/exp): one arrival per warp with the warp's byte count instead of one per thread
    bytes = __reduce_add_sync(0xffffffffu, bytes);
    if ((tid & 31) == 0) pmg_mbar_arrive_expect(bar, bytes);
#else
    pmg_mbar_arrive_expect(bar, bytes);
#endif
  }
  // arrivals a u-staging barrier expects per use: every staging thread, or every staging warp (PMG_SWEEP_WARP_ARRIVE)
  static PMG_HD constexpr int stage_arrivals(int staging_threads)
  {
#if defined(PMG_SWEEP_WARP_ARRIVE)
    return staging_threads / 32;
#else
    return staging_threads;
#endif
  }

  // b (and x_old) rows of the owned columns of output planes gz0 .. gz0+npl-1 -> E with cp.async (LDGSTS), issued by the
  // CTA's last warp while the others do the y sweep; the warp waits for them at the end of phase 2.  (These rows as bulk
  // copies as well saturate the copy engine -- ~1 small row per 25-35 cycles per SM, measured -- and cp.async stalls its
  // issuer on the address registers, which an otherwise idle warp can afford.)
  static PMG_HD void load_e_async(const PmgSweepParams<P> &p, const TileGeom &t, int lane, double *E, int gz0, int npl)
  {
    const int64_t plane = (int64_t)p.Nx * p.Ny;
    const int x = lane % LD_LPR, rsub = lane / LD_LPR;
    const int64_t g0 = (int64_t)(gz0 - p.z0) * plane + (int64_t)(t.cy0 * P + rsub) * p.Nx + t.cx0 * P + x;
    const int64_t rstep = (int64_t)LD_RPI * p.Nx;
    const bool xok = (x < t.cw);
#pragma unroll
    for (int arr = 0; arr < 2; ++arr) {
      if (arr == 1 && !t.has_xo) break;
      const double *base = (arr == 0 ? p.b : p.xold);
#pragma unroll
      for (int k = 0; k < NPS; ++k) {
        if (k >= npl) break;
        const double *src = base + g0 + k * plane;
        double *dst = E + arr * EBUF + k * EPLANE + rsub * XPE + x;
#pragma unroll 4
        for (int r = rsub; r < t.rows; r += LD_RPI) {
          if (xok) pmg_sweep_cp_async8(dst, src);
          src += rstep; dst += LD_RPI * XPE;
        }
      }
      // columns beyond the LD_LPR lanes (the mesh's last vertex line): one element per (plane, row)
      for (int xx = LD_LPR; xx < t.cw; ++xx)
        for (int i = lane; i < npl * t.rows; i += 32) {
          const int k = i / t.rows, r = i - k * t.rows;
          pmg_sweep_cp_async8(E + arr * EBUF + k * EPLANE + r * XPE + xx,
                        base + (int64_t)(gz0 + k - p.z0) * plane + (int64_t)(t.cy0 * P + r) * p.Nx + t.cx0 * P + xx);
        }
    }
  }

  // ---- phase 1: y sweep.  item = (xl, k): column xl of plane k of A (u, read only) -> c in C, d in D -------------
  static PMG_HD void phase1(const PmgSweepParams<P> &p, const TileGeom &t, int tid, const double *A, double *Cb, double *Db,
                            int gz0, int npl)
  {
    const bool dir_lo = (p.faces >> 2 & 1u), dir_hi = (p.faces >> 3 & 1u);
    for (int item = tid; item < NITEM1; item += NT) {
      const int seg = (SG == 1) ? 0 : item / (XW * NPS);
      const int rem = item - seg * (XW * NPS);
      const int xl = rem % XW, k = rem / XW;
      if (k >= npl) continue;
      const int gx = (t.cx0 - 1) * P + xl;
      if (gx < 0 || gx >= p.Nx) continue;
      const int c_lo = seg * BYS;             // first cell of the segment
      if (c_lo >= t.ncy) continue;
      const bool last_seg = (c_lo + BYS >= t.ncy); // the segment holds the tile's last valid cell
      const int gz = gz0 + k;
      double *Cc = Cb + k * CPLANE + xl;       // row r: Cc[r * XP]
      double *Dc = Db + k * CPLANE + xl;
      const bool zero_line = (gx == 0 && (p.faces & 1u)) || (gx == p.Nx - 1 && (p.faces >> 1 & 1u)) ||
                             (gz == 0 && (p.faces >> 4 & 1u)) || (gz == p.Nz - 1 && (p.faces >> 5 & 1u));
      if (zero_line) { // Dirichlet values read as 0 (:250-254): the whole line of c, d vanishes
        const int r_hi = last_seg ? t.rows : (c_lo + BYS) * P;
        for (int r = c_lo * P; r < r_hi; ++r) { Cc[r * XP] = 0.0; Dc[r * XP] = 0.0; }
        continue;
      }
      // row yl of the column is at Ae[yl * XPA] for even yl and Ao[yl * XPA] for odd yl (the row shift alternates
      // from row to row when Nx is odd)
      const int s0 = shift_of(t.eA + (int64_t)(gz - p.z0) * p.Nx * p.Ny);
      const double *Ae = A + k * APLANE + xl + s0;
      const double *Ao = A + k * APLANE + xl + (s0 ^ t.nxodd);
#define PMG_AROW(yl) ((((yl) & 1) ? Ao : Ae)[(yl) * XPA])
      double cc = 0.0, cd = 0.0, v0;
      if (t.cy0 + c_lo > 0) { // partial cell below the segment (the halo cell for the first one): only its contribution to the segment's first row
        double v[N1];
#pragma unroll
        for (int j = 0; j < N1; ++j) v[j] = PMG_AROW(c_lo * P + j);
        if (t.cy0 + c_lo == 1 && dir_lo) v[0] = 0.0;
#pragma unroll
        for (int j = 0; j < N1; ++j) { cc = fma(PMG_M(P, j), v[j], cc); cd = fma(PMG_KY(P, j), v[j], cd); }
        v0 = v[P];
      } else {
        v0 = dir_lo ? 0.0 : PMG_AROW(P);
      }
#pragma unroll (RL ? 1 : BYS)
      for (int ci = 0; ci < BYS; ++ci) {
        const int c = c_lo + ci;
        if (c < t.ncy) {
          double vj = v0;
          double sc[N1], sd[N1];
#pragma unroll
          for (int kk = 0; kk < N1; ++kk) { sc[kk] = 0.0; sd[kk] = 0.0; }
          sc[0] = cc; sd[0] = cd;
          // rolled: the parity of the cell's first row is a run-time value when P is odd; pick the two row bases once per cell
          const double *Ac0 = (((P + c * P) & 1) ? Ao : Ae) + (P + c * P) * XPA;
          const double *Ac1 = (((P + c * P) & 1) ? Ae : Ao) + (P + c * P) * XPA;
#pragma unroll
          for (int j = 0; j < N1; ++j) {
            if (j > 0) vj = RL ? ((j & 1) ? Ac1 : Ac0)[j * XPA] : PMG_AROW(P + c * P + j);
            if (j == P && dir_hi && t.cy0 + c == p.ny - 1) vj = 0.0;
#pragma unroll
            for (int kk = 0; kk < N1; ++kk) {
              sc[kk] = fma(PMG_M(kk, j), vj, sc[kk]);
              sd[kk] = fma(PMG_KY(kk, j), vj, sd[kk]);
            }
          }
#pragma unroll
          for (int kk = 0; kk < P; ++kk) { Cc[(c * P + kk) * XP] = sc[kk]; Dc[(c * P + kk) * XP] = sd[kk]; }
          cc = sc[P]; cd = sd[P]; v0 = vj;
        }
      }
      if (t.y_end && last_seg) { Cc[t.ncy * P * XP] = cc; Dc[t.ncy * P * XP] = cd; }
#undef PMG_AROW
    }
  }

  // ---- phase 2: x sweep in place.  item = (row r, plane k, segment): (c, d) -> (g, m) ----------------------------
  // SG > 1: the segments of a row run concurrently and in place, so whatever a segment needs of values another segment
  // overwrites is read first (phase2_seed; the phase-2 threads then meet at a barrier of their own): a later segment's start --
  // the partial sums of the cell before it and the (c, d) of the vertex it starts at -- and an earlier segment's last input, the
  // (c, d) of the vertex where the next segment starts.  seed2 = {cg, cm, c0, d0} resp. {c_end, d_end, -, -}.
  static PMG_HD void phase2_seed(const PmgSweepParams<P> &p, const TileGeom &t, ThreadState &st, const double *Cb, const double *Db, int npl,
                                 int alt = 0)
  {
#pragma unroll
    for (int it = 0; it < (SG > 1 ? IT2 : 0); ++it) {
      const int item = (A2 && alt) ? st.item2[A2 ? 1 : 0][it] : st.item2[0][it];
      if (item < 0) continue;
      const int r = item & 0xFFFF, k = (item >> 16) & 0xFF, seg = item >> 24;
      if (k >= npl) continue;
      const double *Cr = Cb + k * CPLANE + r * XP;
      const double *Dr = Db + k * CPLANE + r * XP;
      const int i_lo = seg * BXS;
      if (i_lo >= t.ncx) continue;
      if (seg > 0) {
        double cg = 0.0, cm = 0.0, c[N1], d[N1];
#pragma unroll
        for (int j = 0; j < N1; ++j) { c[j] = Cr[i_lo * P + j]; d[j] = Dr[i_lo * P + j]; }
#pragma unroll
        for (int j = 0; j < N1; ++j) {
          cg = fma(PMG_KX(P, j), c[j], fma(PMG_M(P, j), d[j], cg));
          cm = fma(PMG_M(P, j), c[j], cm);
        }
        st.seed2[it][0] = cg; st.seed2[it][1] = cm; st.seed2[it][2] = c[P]; st.seed2[it][3] = d[P];
      } else if (i_lo + BXS < t.ncx) { // first segment with a successor: its last input is the successor's first output
        st.seed2[it][0] = Cr[P + (i_lo + BXS) * P]; st.seed2[it][1] = Dr[P + (i_lo + BXS) * P];
      }
    }
  }

  static PMG_HD void phase2(const PmgSweepParams<P> &p, const TileGeom &t, const ThreadState &st, double *Cb, double *Db, int npl,
                            int alt = 0)
  {
#pragma unroll
    for (int it = 0; it < IT2; ++it) {
      const int item = (A2 && alt) ? st.item2[A2 ? 1 : 0][it] : st.item2[0][it];
      if (item < 0) continue;
      const int r = item & 0xFFFF, k = (item >> 16) & 0xFF, seg = (SG == 1) ? 0 : item >> 24;
      if (k >= npl) continue;
      const int i_lo = seg * BXS;
      if (i_lo >= t.ncx) continue;
      const bool last_seg = (i_lo + BXS >= t.ncx);
      double *Cr = Cb + k * CPLANE + r * XP;
      double *Dr = Db + k * CPLANE + r * XP;
      double cg = 0.0, cm = 0.0, c0, d0;
      if (SG > 1 && seg > 0) {
        cg = st.seed2[it][0]; cm = st.seed2[it][1]; c0 = st.seed2[it][2]; d0 = st.seed2[it][3];
      } else if (t.cx0 > 0) { // partial cell left of the tile: only its contribution to the first owned column
        double c[N1], d[N1];
#pragma unroll
        for (int j = 0; j < N1; ++j) { c[j] = Cr[j]; d[j] = Dr[j]; }
#pragma unroll
        for (int j = 0; j < N1; ++j) {
          cg = fma(PMG_KX(P, j), c[j], fma(PMG_M(P, j), d[j], cg));
          cm = fma(PMG_M(P, j), c[j], cm);
        }
        c0 = c[P]; d0 = d[P];
      } else {
        c0 = Cr[P]; d0 = Dr[P];
      }
#pragma unroll (RL ? 1 : BXS)
      for (int ii = 0; ii < BXS; ++ii) {
        const int i = i_lo + ii;
        if (i < t.ncx) {
          double *Cc = Cr + P + i * P, *Dc = Dr + P + i * P;
          double sg[N1], sm[N1];
#pragma unroll
          for (int kk = 0; kk < N1; ++kk) { sg[kk] = 0.0; sm[kk] = 0.0; }
          sg[0] = cg; sm[0] = cm;
          double cj = c0, dj = d0;
#pragma unroll
          for (int j = 0; j < N1; ++j) {
            if (j > 0) { cj = Cc[j]; dj = Dc[j]; }
            if (SG > 1 && j == P && ii == BXS - 1 && !last_seg) { cj = st.seed2[it][0]; dj = st.seed2[it][1]; } // overwritten by the next segment
#pragma unroll
            for (int kk = 0; kk < N1; ++kk) {
              sg[kk] = fma(PMG_KX(kk, j), cj, fma(PMG_M(kk, j), dj, sg[kk]));
              sm[kk] = fma(PMG_M(kk, j), cj, sm[kk]);
            }
          }
#pragma unroll
          for (int kk = 0; kk < P; ++kk) { Cc[kk] = sg[kk]; Dc[kk] = sm[kk]; }
          cg = sg[P]; cm = sm[P]; c0 = cj; d0 = dj;
        }
      }
      if (t.x_end && last_seg) { Cr[P + t.ncx * P] = cg; Dr[P + t.ncx * P] = cm; }
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
    return uc + p.f1 * (uc - xo) + corr;
  }

  // ---- phase 3: z sweep + epilogue of cell layer cz = layer l of the step; thread = NCOL dof columns --------------
  // All inputs are in shared memory: g, m in C, D; u in A (planes q = l P + k of the step: plane q - 1 of A holds dof plane
  // cz0 P + q, q >= 1; dof plane cz0 P itself is kept in a register from the previous step); b, x_old in E (plane q).
  // FULL: every plane of the layer is owned, written, inside the mesh's first/last planes: straight-line code.
  template <int MODE, bool FULL>
  static PMG_HD void phase3_layer(const PmgSweepParams<P> &p, const TileGeom &t, ThreadState &st, const double *A, const double *Cb,
                                  const double *Db, const double *E, int l, int cz0, bool write)
  {
    constexpr int T = P + 2;
    const int cz = cz0 + l;
    const int64_t plane = (int64_t)p.Nx * p.Ny;
    const bool has_xo = (MODE == PMG_MODE_CHEB_STEP) && (p.xold != nullptr);
    const int64_t glayer = (int64_t)(t.cy0 * P) * p.Nx + t.cx0 * P + ((int64_t)cz * P - p.z0) * plane;
    const int64_t zoff = (int64_t)(cz0 * P - p.z0) * plane;
#pragma unroll
    for (int ci = 0; ci < NCOL; ++ci) {
      const int info = st.info[ci];
      if (info < 0) continue;
      const int ox = info & 0xFF, oy = (info >> 8) & 0xFF;
      const double *G = Cb + oy * XP + P + ox;
      const double *Mm = Db + oy * XP + P + ox;
      double bbv[P], xov[P];
      if (EG && FULL && MODE != PMG_MODE_APPLY) { // direct epilogue inputs: issue the loads before the z-sweep arithmetic
        const int64_t goff = glayer + (int64_t)oy * p.Nx + ox; // the column's dof in plane 0 of the layer
#pragma unroll
        for (int k = 0; k < P; ++k) { bbv[k] = p.b[goff + k * plane]; xov[k] = has_xo ? p.xold[goff + k * plane] : 0.0; }
      }
      double g[N1], m[N1];
      g[0] = st.gP[ci]; m[0] = st.mP[ci];
#pragma unroll
      for (int kp = 1; kp < N1; ++kp) { g[kp] = G[(l * P + kp - 1) * CPLANE]; m[kp] = Mm[(l * P + kp - 1) * CPLANE]; }
      double top = 0.0; // row P: the partial sum of the plane shared with the next layer
#pragma unroll
      for (int kp = 0; kp < N1; ++kp) top = fma(PMG_MZ(P, kp), g[kp], fma(PMG_KZ(P, kp), m[kp], top));
      // shifts of this column's rows in dof plane cz0 P; plane cz0 P + q has them flipped for odd q when Nx Ny is odd
      const int sa = shift_of(t.eA + zoff + (int64_t)(oy + P) * p.Nx);
      const double *Uc = A + (oy + P) * XPA + P + ox;
      const double *Ec = E + oy * XPE + ox;
      const int qt = l * P + P;          // the layer's top plane
      const double utop = Uc[(qt - 1) * APLANE + (sa ^ ((qt & 1) & t.plodd))];
      if (write) {
        const bool dirxy = (info >> 16) & 1;
        const int tbase = ((info >> 20) & 0xF) + T * ((info >> 24) & 0xF);
        double *po = p.out + glayer + (int64_t)oy * p.Nx + ox;
#pragma unroll
        for (int k = 0; k < P; ++k) {
          const int gz = cz * P + k;
          if (FULL || (gz >= p.z_own_lo && gz < p.z_own_hi)) {
            double y = (k == 0) ? st.carry[ci] : 0.0;
#pragma unroll
            for (int kp = 0; kp < N1; ++kp) y = fma(PMG_MZ(k, kp), g[kp], fma(PMG_KZ(k, kp), m[kp], y));
            const bool dir = dirxy || (!FULL && ((gz == 0 && (p.faces >> 4 & 1u)) || (gz == p.Nz - 1 && (p.faces >> 5 & 1u))));
            const int q = l * P + k;
            const int fl = (q & 1) & t.plodd;
            double uc = 0.0, bb = 0.0, xo = 0.0, dinv = 1.0;
            if (MODE != PMG_MODE_APPLY || dir) uc = (q == 0) ? st.uP[ci] : Uc[(q - 1) * APLANE + (sa ^ fl)];
            if (MODE != PMG_MODE_APPLY) bb = !EG ? Ec[q * EPLANE] : FULL ? bbv[k] : p.b[(po - p.out) + k * plane];
            if (has_xo) xo = !EG ? Ec[EBUF + q * EPLANE] : FULL ? xov[k] : p.xold[(po - p.out) + k * plane];
            if (MODE >= PMG_MODE_CHEB_FIRST) {
              if (p.dinv_vec) dinv = p.dinv_vec[glayer + (int64_t)oy * p.Nx + ox + k * plane];
              else if (FULL) dinv = st.dinv[ci][k];
              else dinv = p.dinv_tab[tbase + T * T * pmg_sweep_pos_type<P>(gz, p.Nz)];
            }
            po[k * plane] = epi_value<MODE>(p, y, uc, bb, xo, dir, dinv);
          }
        }
      }
      st.carry[ci] = top; st.gP[ci] = g[P]; st.mP[ci] = m[P]; st.uP[ci] = utop;
    }
  }

  template <int MODE>
  static PMG_HD void phase3_t(const PmgSweepParams<P> &p, const TileGeom &t, ThreadState &st, const double *A, const double *Cb,
                              const double *Db, const double *E, int cz0, int nlay, int cz_write)
  {
#pragma unroll
    for (int l = 0; l < LZ; ++l) {
      if (l < nlay) {
        const int cz = cz0 + l;
        const bool write = (cz >= cz_write);
        // plane 0 of the mesh (Dirichlet or not, its inverse-diagonal type differs) takes the general path
        const bool full = write && cz > 0 && cz * P >= p.z_own_lo && cz * P + P <= p.z_own_hi;
        if (full) phase3_layer<MODE, true>(p, t, st, A, Cb, Db, E, l, cz0, true);
        else phase3_layer<MODE, false>(p, t, st, A, Cb, Db, E, l, cz0, write);
      }
    }
  }

  static PMG_HD void phase3(const PmgSweepParams<P> &p, const TileGeom &t, ThreadState &st, const double *A, const double *Cb,
                            const double *Db, const double *E, int cz0, int nlay, int cz_write)
  {
    switch (mode_of(p)) {
      case PMG_MODE_APPLY: phase3_t<PMG_MODE_APPLY>(p, t, st, A, Cb, Db, E, cz0, nlay, cz_write); break;
      case PMG_MODE_RESIDUAL: phase3_t<PMG_MODE_RESIDUAL>(p, t, st, A, Cb, Db, E, cz0, nlay, cz_write); break;
      case PMG_MODE_CHEB_FIRST: phase3_t<PMG_MODE_CHEB_FIRST>(p, t, st, A, Cb, Db, E, cz0, nlay, cz_write); break;
      default: phase3_t<PMG_MODE_CHEB_STEP>(p, t, st, A, Cb, Db, E, cz0, nlay, cz_write); break;
    }
  }

  // after the prologue: (g, m, u) of the first layer's plane 0 (dof plane gz, in plane slot 0 of A0) into the registers
  static PMG_HD void phase3_init(const PmgSweepParams<P> &p, const TileGeom &t, ThreadState &st, const double *A0, const double *Cb,
                                 const double *Db, int gz)
  {
    const int64_t zoff = (int64_t)(gz - p.z0) * p.Nx * p.Ny;
#pragma unroll
    for (int ci = 0; ci < NCOL; ++ci) {
      const int info = st.info[ci];
      if (info < 0) continue;
      const int ox = info & 0xFF, oy = (info >> 8) & 0xFF;
      st.gP[ci] = Cb[oy * XP + P + ox];
      st.mP[ci] = Db[oy * XP + P + ox];
      st.uP[ci] = A0[(oy + P) * XPA + P + ox + shift_of(t.eA + zoff + (int64_t)(oy + P) * p.Nx)];
    }
  }

  // the mesh's top plane: the carried sums are complete (no cell layer above); a single plane, inputs read from global
  template <int MODE>
  static PMG_HD void flush_t(const PmgSweepParams<P> &p, const TileGeom &t, ThreadState &st, int gz)
  {
    constexpr int T = P + 2;
    const int64_t plane = (int64_t)p.Nx * p.Ny;
    const bool dirz = (gz == 0 && (p.faces >> 4 & 1u)) || (gz == p.Nz - 1 && (p.faces >> 5 & 1u));
#pragma unroll
    for (int ci = 0; ci < NCOL; ++ci) {
      const int info = st.info[ci];
      if (info < 0) continue;
      const int ox = info & 0xFF, oy = (info >> 8) & 0xFF;
      const int64_t g = (int64_t)(gz - p.z0) * plane + (int64_t)(t.cy0 * P + oy) * p.Nx + t.cx0 * P + ox;
      const bool dir = dirz || ((info >> 16) & 1);
      double uc = 0.0, bb = 0.0, xo = 0.0, dinv = 1.0;
      if (MODE != PMG_MODE_APPLY || dir) uc = p.u[g];
      if (MODE != PMG_MODE_APPLY) bb = p.b[g];
      if (MODE == PMG_MODE_CHEB_STEP && p.xold) xo = p.xold[g];
      if (MODE >= PMG_MODE_CHEB_FIRST)
        dinv = p.dinv_vec ? p.dinv_vec[g]
                          : p.dinv_tab[((info >> 20) & 0xF) + T * ((info >> 24) & 0xF) + T * T * pmg_sweep_pos_type<P>(gz, p.Nz)];
      p.out[g] = epi_value<MODE>(p, st.carry[ci], uc, bb, xo, dir, dinv);
    }
  }

  static PMG_HD void flush(const PmgSweepParams<P> &p, const TileGeom &t, ThreadState &st, int gz)
  {
    switch (mode_of(p)) {
      case PMG_MODE_APPLY: flush_t<PMG_MODE_APPLY>(p, t, st, gz); break;
      case PMG_MODE_RESIDUAL: flush_t<PMG_MODE_RESIDUAL>(p, t, st, gz); break;
      case PMG_MODE_CHEB_FIRST: flush_t<PMG_MODE_CHEB_FIRST>(p, t, st, gz); break;
      default: flush_t<PMG_MODE_CHEB_STEP>(p, t, st, gz); break;
    }
  }

  // ---- the tile program -----------------------------------------------------
  // Exec provides: template<F> void for_each_thread(F f)  with f(int tid, ThreadState&);  void sync()
  // smem must be 16-byte aligned.
  template <class Exec>
  static PMG_HD void run(const PmgSweepParams<P> &p, Exec &ex, double *smem, int tile_x, int tile_y, int chunk)
  {
    const TileGeom t = geom(p, tile_x, tile_y);
    const int cz_begin = p.cz_lo + chunk * p.layers_per_chunk;
    int cz_end = cz_begin + p.layers_per_chunk;
    if (cz_end > p.cz_hi) cz_end = p.cz_hi;
    if (cz_begin >= cz_end) return;
    // the layer below the chunk is recomputed when its planes are stored locally (they are unless it is outside the mesh)
    const bool halo = (cz_begin > 0) && ((cz_begin - 1) * P >= p.z0);
    const int cz_first = halo ? cz_begin - 1 : cz_begin;
    double *Cb = smem + C_OFFSET, *Db = smem + D_OFFSET, *E = smem + E_OFFSET;
    uint64_t *bars = (uint64_t *)(smem + (t.has_e ? BAR_OFFSET_EPI : BAR_OFFSET_APPLY));
    int par_u[2] = {0, 0}; // phase parity each barrier completes next

    // prologue: plane 0 of the first layer -> buffer 0; the first step's planes -> buffer 1
    int n0 = cz_end - cz_first; if (n0 > LZ) n0 = LZ;
    ex.for_each_thread([&](int tid, ThreadState &) {
      if (tid == 0) {
        for (int i = 0; i < NBAR; ++i) pmg_mbar_init(bars + i, stage_arrivals(NT));
        pmg_mbar_init_fence();
      }
    });
    ex.sync();
    ex.for_each_thread([&](int tid, ThreadState &) {
      stage_u(p, t, tid, smem, bars + 0, cz_first * P, 1);
      stage_u(p, t, tid, smem + ABUF, bars + 1, cz_first * P + 1, n0 * P);
      if (tid >= NT - 32) {
        load_u_async(p, t, tid - (NT - 32), smem, cz_first * P, 1);
        load_u_async(p, t, tid - (NT - 32), smem + ABUF, cz_first * P + 1, n0 * P);
        pmg_sweep_cp_async_wait_all();
      }
    });
    ex.sync();
    ex.for_each_thread([&](int tid, ThreadState &st) {
      decode(p, t, tid, st);
      pmg_mbar_wait(bars + 0, par_u[0]);
      phase1(p, t, tid, smem, Cb, Db, cz_first * P, 1);
    });
    par_u[0] ^= 1;
    ex.sync();
    if (SG > 1) {
      ex.for_each_thread([&](int, ThreadState &st) { phase2_seed(p, t, st, Cb, Db, 1); });
      ex.sync_some(NT2);
    }
    ex.for_each_thread([&](int, ThreadState &st) { phase2(p, t, st, Cb, Db, 1); });
    ex.sync();
    ex.for_each_thread([&](int, ThreadState &st) { phase3_init(p, t, st, smem, Cb, Db, cz_first * P); });
    ex.sync();

    int cur = 1;
    int alt = 0; // A2: which half of the CTA's warps holds the phase-2 items in this step
    for (int cz = cz_first; cz < cz_end; cz += LZ, alt ^= 1) {
      double *A = smem + cur * ABUF;
      int nlay = cz_end - cz; if (nlay > LZ) nlay = LZ;
      int nnext = cz_end - (cz + LZ); if (nnext > LZ) nnext = LZ;
      ex.for_each_thread([&](int tid, ThreadState &) {
        // the other staging buffer and E were last read before the barrier that ended the previous step: fetch the next
        // step's u planes (a full step ahead) and, by the last warp, this step's b / x_old rows (two phases ahead)
        if (nnext > 0) stage_u(p, t, tid, smem + (cur ^ 1) * ABUF, bars + (cur ^ 1), (cz + LZ) * P + 1, nnext * P);
        if (tid >= NT - 32) {
          if (t.has_e) load_e_async(p, t, tid - (NT - 32), E, cz * P, nlay * P);
          if (nnext > 0) load_u_async(p, t, tid - (NT - 32), smem + (cur ^ 1) * ABUF, (cz + LZ) * P + 1, nnext * P);
        }
        pmg_mbar_wait(bars + cur, par_u[cur]);
        phase1(p, t, tid, A, Cb, Db, cz * P + 1, nlay * P);
      });
      par_u[cur] ^= 1;
      ex.sync();
      if (SG > 1) { // the phase-2 threads read what another segment of their row overwrites, then meet at their own barrier
        ex.for_each_thread([&](int, ThreadState &st) { phase2_seed(p, t, st, Cb, Db, nlay * P, alt); });
        ex.sync_some(NT2);
      }
      ex.for_each_thread([&](int tid, ThreadState &st) {
        phase2(p, t, st, Cb, Db, nlay * P, alt);
        if (tid >= NT - 32) pmg_sweep_cp_async_wait_all(); // the loader warp's copies (E for phase 3, u rows for the next step)
      });
      ex.sync();
      ex.for_each_thread([&](int, ThreadState &st) { phase3(p, t, st, A, Cb, Db, E, cz, nlay, cz_begin); });
      ex.sync();
      cur ^= 1;
    }
    // top plane of the mesh (owned by the chunk that ends there)
    if (cz_end == p.cz_hi && cz_end * P < p.z_own_hi)
      ex.for_each_thread([&](int, ThreadState &st) { flush(p, t, st, cz_end * P); });
  }
#undef PMG_M
#undef PMG_KX
#undef PMG_KY
#undef PMG_MZ
#undef PMG_KZ
};
