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
//      stage    the u planes of the NEXT step are copied global -> shared with cp.async (LDGSTS) while this step
//               computes (two staging buffers), so no sweep ever waits on HBM latency
//      phase 1  thread = (x, plane): marches in y, in place in the staging buffer: u -> c, and d -> second buffer
//      phase 2  thread = (y, plane): marches in x, in place: (c, d) -> (g, m)
//      phase 3  thread = dof column (x, y): z sweep over the layer's P+1 planes -- the values of the plane shared
//               with the layer below and its partial sum stay in registers (3 doubles per column) -- and the
//               epilogue (Dirichlet identity / residual / Chebyshev update) before the single coalesced store.
//    Every shared-memory access is conflict-free (odd row pitch, lanes along x or along y).
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

// 8-byte asynchronous global -> shared copy (LDGSTS); a plain copy under the host emulator
PMG_HD void pmg_sweep_cp_async8(double *dst_smem, const double *src_global)
{
#if defined(__CUDA_ARCH__)
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(src_global) : "memory");
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

// hint: bring the 128-byte lines covering [ptr, ptr + bytes) into L1; nothing under the host emulator
PMG_HD void pmg_sweep_prefetch_l1(const double *ptr, int bytes)
{
#if defined(__CUDA_ARCH__)
  const unsigned long long a = (unsigned long long)ptr;
  for (unsigned long long l = a & ~127ull; l < a + (unsigned)bytes; l += 128)
    asm volatile("prefetch.global.L1 [%0];\n" ::"l"(l) : "memory");
#else
  (void)ptr; (void)bytes;
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

// P: degree; BX x BY: cell columns per CTA; LZ: cell layers per step; NT_: threads
template <int P, int BX, int BY, int LZ, int NT_>
struct PmgSweepTile {
  static constexpr int N1 = P + 1;
  static constexpr int NT = NT_;
  static constexpr int NPS = LZ * P;          // dof planes per step
  static constexpr int XW = (BX + 1) * P + 1; // x points of the tile buffers: global x = (cx0-1) P + xl
  static constexpr int XP = XW | 1;           // odd row pitch: the column walks of phase 2 are conflict-free
  static constexpr int YW = (BY + 1) * P + 1; // rows of the staging buffer A: global y = (cy0-1) P + yl
  static constexpr int RW = BY * P + 1;       // rows of buffer B:            global y = cy0 P + r
  static constexpr int CW = BX * P + 1;       // owned dof columns in x (the +1 only at the mesh end)
  static constexpr int APLANE = YW * XP, BPLANE = RW * XP;
  static constexpr int ABUF = NPS * APLANE;   // one staging buffer: u, then c, then g (in place)
  static constexpr int B_OFFSET = 2 * ABUF;   // buffer B: d, then m
  static constexpr int SMEM_DOUBLES = 2 * ABUF + NPS * BPLANE;
  static constexpr int NITEM1 = XW * NPS;
  static constexpr int IT2 = (RW * NPS + NT - 1) / NT; // phase-2 items per thread
  static constexpr int NCOL = (CW * RW + NT - 1) / NT; // dof columns per thread in phase 3

#define PMG_M(i, j) p.M[pmg_sweep_canon<P>(i, j)]
#define PMG_KX(i, j) p.Kx[pmg_sweep_canon<P>(i, j)]
#define PMG_KY(i, j) p.Ky[pmg_sweep_canon<P>(i, j)]
#define PMG_MZ(i, j) p.Mz[pmg_sweep_canon<P>(i, j)]
#define PMG_KZ(i, j) p.Kz[pmg_sweep_canon<P>(i, j)]

  struct ThreadState {
    double carry[NCOL]; // z-sweep sum of the plane shared with the next cell layer, per owned dof column
    double gP[NCOL], mP[NCOL]; // (g, m) of that plane
    int info[NCOL];     // ox | oy << 8 | Dirichlet-in-xy << 16 | x position type << 20 | y position type << 24; -1 = none
    int item2[IT2];     // phase-2 item: row | plane << 16; -1 = none
  };

  struct TileGeom {
    int cx0, cy0;     // first owned cell column
    int ncx, ncy;     // valid owned cells
    int x_end, y_end; // tile touches the high end of the mesh (owns the last vertex line)
    int cw, rows;     // owned dof columns / rows
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
    return t;
  }

  static PMG_HD void decode(const PmgSweepParams<P> &p, const TileGeom &t, int tid, ThreadState &st)
  {
    const int ncols = t.cw * t.rows;
#pragma unroll
    for (int ci = 0; ci < NCOL; ++ci) {
      const int col = tid + ci * NT;
      st.carry[ci] = 0.0; st.gP[ci] = 0.0; st.mP[ci] = 0.0;
      if (col < ncols) {
        const int oy = col / t.cw, ox = col - oy * t.cw;
        const int gx = t.cx0 * P + ox, gy = t.cy0 * P + oy;
        const int dirxy = (gx == 0 && (p.faces & 1u)) || (gx == p.Nx - 1 && (p.faces >> 1 & 1u)) ||
                          (gy == 0 && (p.faces >> 2 & 1u)) || (gy == p.Ny - 1 && (p.faces >> 3 & 1u));
        st.info[ci] = ox | (oy << 8) | (dirxy << 16) | (pmg_sweep_pos_type<P>(gx, p.Nx) << 20) |
                      (pmg_sweep_pos_type<P>(gy, p.Ny) << 24);
      } else {
        st.info[ci] = -1;
      }
    }
#pragma unroll
    for (int i = 0; i < IT2; ++i) { // rows fastest: the lanes of a warp walk down a column of the buffer
      const int item = tid + i * NT;
      const int k = item / t.rows;
      st.item2[i] = (k < NPS) ? ((item - k * t.rows) | (k << 16)) : -1;
    }
  }

  // ---- stage: copy `npl` dof planes gz0 .. of the tile's (XW x YW) footprint into staging buffer A ------------
  // lane = x point (coalesced), warp w takes rows w, w + NW, ...: two pointer increments per copy
  static PMG_HD void stage(const PmgSweepParams<P> &p, const TileGeom &t, int tid, double *A, int gz0, int npl)
  {
#ifdef PMG_EXP_NOSTAGE
    return;
#endif
    constexpr int NW = NT / 32;
    const int64_t plane = (int64_t)p.Nx * p.Ny;
    const int gx0 = (t.cx0 - 1) * P, gy0 = (t.cy0 - 1) * P;
    const int yl_lo = (t.cy0 == 0) ? P : 0; // rows below the mesh do not exist
    const int yl_hi = (t.ncy + 1) * P;      // last row the sweeps read
    const int64_t row_step = (int64_t)NW * p.Nx;
#pragma unroll
    for (int xl = tid % 32; xl < XW; xl += 32) {
      const int gx = gx0 + xl;
      if (gx < 0 || gx >= p.Nx) continue;
      const int yl0 = yl_lo + tid / 32;
      const double *src0 = p.u + (int64_t)(gz0 - p.z0) * plane + (int64_t)(gy0 + yl0) * p.Nx + gx;
      double *dst0 = A + yl0 * XP + xl;
#pragma unroll
      for (int k = 0; k < NPS; ++k) {
        if (k < npl) {
          const double *src = src0 + k * plane;
          double *dst = dst0 + k * APLANE;
#pragma unroll 4
          for (int yl = yl0; yl <= yl_hi; yl += NW) {
            pmg_sweep_cp_async8(dst, src);
            src += row_step; dst += NW * XP;
          }
        }
      }
    }
  }

  // ---- phase 1: y sweep in place.  item = (xl, k): column xl of plane k; u -> c in A, d -> B ------------------
  static PMG_HD void phase1(const PmgSweepParams<P> &p, const TileGeom &t, int tid, double *A, double *B, int gz0, int npl)
  {
    const bool dir_lo = (p.faces >> 2 & 1u), dir_hi = (p.faces >> 3 & 1u);
    for (int item = tid; item < NITEM1; item += NT) {
      const int xl = item % XW, k = item / XW;
      if (k >= npl) continue;
      const int gx = (t.cx0 - 1) * P + xl;
      if (gx < 0 || gx >= p.Nx) continue;
      const int gz = gz0 + k;
      double *Ac = A + k * APLANE + xl;       // row yl of the column: Ac[yl * XP]; owned row r is yl = r + P
      double *Bc = B + k * BPLANE + xl;       // row r: Bc[r * XP]
      const bool zero_line = (gx == 0 && (p.faces & 1u)) || (gx == p.Nx - 1 && (p.faces >> 1 & 1u)) ||
                             (gz == 0 && (p.faces >> 4 & 1u)) || (gz == p.Nz - 1 && (p.faces >> 5 & 1u));
      if (zero_line) { // Dirichlet values read as 0 (:250-254): the whole line of c, d vanishes
        for (int r = 0; r < t.rows; ++r) { Ac[(r + P) * XP] = 0.0; Bc[r * XP] = 0.0; }
        continue;
      }
      double cc = 0.0, cd = 0.0, v0;
      if (t.cy0 > 0) { // partial cell below the tile: only its contribution to the first owned row
        double v[N1];
#pragma unroll
        for (int j = 0; j < N1; ++j) v[j] = Ac[j * XP];
        if (t.cy0 == 1 && dir_lo) v[0] = 0.0;
#pragma unroll
        for (int j = 0; j < N1; ++j) { cc = fma(PMG_M(P, j), v[j], cc); cd = fma(PMG_KY(P, j), v[j], cd); }
        v0 = v[P];
      } else {
        v0 = dir_lo ? 0.0 : Ac[P * XP];
      }
#pragma unroll
      for (int c = 0; c < BY; ++c) {
        if (c < t.ncy) {
          double *Acc = Ac + (P + c * P) * XP;
          double vj = v0;
          double sc[N1], sd[N1];
#pragma unroll
          for (int kk = 0; kk < N1; ++kk) { sc[kk] = 0.0; sd[kk] = 0.0; }
          sc[0] = cc; sd[0] = cd;
#pragma unroll
          for (int j = 0; j < N1; ++j) {
            if (j > 0) vj = Acc[j * XP];
            if (j == P && dir_hi && t.cy0 + c == p.ny - 1) vj = 0.0;
#pragma unroll
            for (int kk = 0; kk < N1; ++kk) {
              sc[kk] = fma(PMG_M(kk, j), vj, sc[kk]);
              sd[kk] = fma(PMG_KY(kk, j), vj, sd[kk]);
            }
          }
#pragma unroll
          for (int kk = 0; kk < P; ++kk) { Acc[kk * XP] = sc[kk]; Bc[(c * P + kk) * XP] = sd[kk]; }
          cc = sc[P]; cd = sd[P]; v0 = vj;
        }
      }
      if (t.y_end) { Ac[(P + t.ncy * P) * XP] = cc; Bc[t.ncy * P * XP] = cd; }
    }
  }

  // ---- phase 2: x sweep in place.  item = (row r, plane k): (c, d) -> (g, m) ---------------------------------
  // gz_pf >= 0: the item also asks L1 for its row of the epilogue's inputs (b, xold) in plane gz_pf + k, i.e. of the NEXT
  // step's output planes, so that phase 3 finds them in L1
  static PMG_HD void phase2(const PmgSweepParams<P> &p, const TileGeom &t, const ThreadState &st, double *A, double *B, int npl,
                            int gz_pf)
  {
#pragma unroll
    for (int it = 0; it < IT2; ++it) {
      const int item = st.item2[it];
      if (item < 0) continue;
      const int r = item & 0xFFFF, k = item >> 16;
      if (gz_pf >= 0 && p.mode != PMG_MODE_APPLY && gz_pf + k < p.z_own_hi) {
        const int64_t g = ((int64_t)(gz_pf + k - p.z0) * p.Ny + (t.cy0 * P + r)) * p.Nx + t.cx0 * P;
        // (measured on B200, Q4: L1 line prefetch 2.6 -> 1.9 ms per fused step; a bulk L2 prefetch gave nothing)
        pmg_sweep_prefetch_l1(p.b + g, t.cw * 8);
        pmg_sweep_prefetch_l1(p.u + g, t.cw * 8);
        if (p.mode == PMG_MODE_CHEB_STEP && p.xold) pmg_sweep_prefetch_l1(p.xold + g, t.cw * 8);
      }
      if (k >= npl) continue;
      double *Cr = A + k * APLANE + (r + P) * XP;
      double *Dr = B + k * BPLANE + r * XP;
      double cg = 0.0, cm = 0.0, c0, d0;
      if (t.cx0 > 0) { // partial cell left of the tile: only its contribution to the first owned column
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
#pragma unroll
      for (int i = 0; i < BX; ++i) {
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
      if (t.x_end) { Cr[P + t.ncx * P] = cg; Dr[P + t.ncx * P] = cm; }
    }
  }

  // inputs of the epilogue for the P planes of one dof column of one layer
  struct EpiIn { double u[P], b[P], xo[P]; };

  // issue the global loads of the epilogue of planes gz0 .. gz0+P-1 of one column (nothing in APPLY mode: Dirichlet rows
  // in x/y are written by fixup_dirichlet_xy, Dirichlet planes in z by the rare branch of epi_store)
  template <int MODE>
  static PMG_HD void epi_load(const PmgSweepParams<P> &p, int64_t g0, int64_t plane, int gz0, EpiIn &in)
  {
    if (MODE != PMG_MODE_APPLY) {
#pragma unroll
      for (int k = 0; k < P; ++k) {
        const int gz = gz0 + k;
        const bool act = (gz >= p.z_own_lo && gz < p.z_own_hi);
        const int64_t g = act ? g0 + k * plane : g0; // any valid address: the value is not used
        in.u[k] = p.u[g];
        in.b[k] = p.b[g];
        if (MODE == PMG_MODE_CHEB_STEP) in.xo[k] = p.xold ? p.xold[g] : 0.0;
      }
    }
  }

  template <int MODE>
  static PMG_HD void epi_store(const PmgSweepParams<P> &p, int64_t g, double y, double uc, double bb, double xo, bool dirxy,
                               bool dirz, int tab_index)
  {
    if (MODE == PMG_MODE_APPLY) {
      if (dirz) p.out[g] = p.u[g];       // Dirichlet rows are the identity (:718); first and last plane only
      else if (!dirxy) p.out[g] = y;     // (Dirichlet rows in x/y: fixup_dirichlet_xy)
      return;
    }
    const bool dir = dirxy || dirz;
    const double Au = dir ? uc : y;
    double r;
    if (MODE == PMG_MODE_RESIDUAL) {
      r = bb - Au;
    } else {
      double dinv;
      if (dir) dinv = 1.0;
      else if (p.dinv_vec) dinv = p.dinv_vec[g];
      else dinv = p.dinv_tab[tab_index];
      const double corr = p.f2 * dinv * (bb - Au);
      if (MODE == PMG_MODE_CHEB_FIRST) r = uc + corr;
      else r = uc + p.f1 * (uc - xo) + corr;
    }
    p.out[g] = r;
  }

  // ---- phase 3: z sweep + epilogue for the `nlay` cell layers cz0 .. of the step; thread = NCOL dof columns -----
  // FULL: every plane of the layer is owned, written and not a Dirichlet plane (all layers but the first and last of the
  // mesh / slab / halo): straight-line code.  The epilogue's global loads of column i+1 are issued before column i is
  // computed.
  template <int MODE, bool FULL>
  static PMG_HD void phase3_layer(const PmgSweepParams<P> &p, const TileGeom &t, ThreadState &st, const double *A, const double *B,
                                  int l, int cz, bool write)
  {
    constexpr int T = P + 2;
    const int64_t plane = (int64_t)p.Nx * p.Ny;
    const int64_t glayer = (int64_t)(t.cy0 * P) * p.Nx + t.cx0 * P + ((int64_t)cz * P - p.z0) * plane;
    EpiIn in[2];
    auto col_g0 = [&](int ci) { const int info = st.info[ci]; return glayer + (int64_t)((info >> 8) & 0xFF) * p.Nx + (info & 0xFF); };
    if (write && st.info[0] >= 0) epi_load<MODE>(p, col_g0(0), plane, cz * P, in[0]);
#pragma unroll
    for (int ci = 0; ci < NCOL; ++ci) {
      if (write && ci + 1 < NCOL && st.info[ci + 1] >= 0) epi_load<MODE>(p, col_g0(ci + 1), plane, cz * P, in[(ci + 1) & 1]);
      const int info = st.info[ci];
      if (info < 0) continue;
      const int ox = info & 0xFF, oy = (info >> 8) & 0xFF;
      const double *G = A + (oy + P) * XP + P + ox;
      const double *Mm = B + oy * XP + P + ox;
      double g[N1], m[N1];
      g[0] = st.gP[ci]; m[0] = st.mP[ci];
#pragma unroll
      for (int kp = 1; kp < N1; ++kp) { g[kp] = G[(l * P + kp - 1) * APLANE]; m[kp] = Mm[(l * P + kp - 1) * BPLANE]; }
      double top = 0.0; // row P: the partial sum of the plane shared with the next layer
#pragma unroll
      for (int kp = 0; kp < N1; ++kp) top = fma(PMG_MZ(P, kp), g[kp], fma(PMG_KZ(P, kp), m[kp], top));
      if (write) {
        const bool dirxy = (info >> 16) & 1;
        const int tbase = ((info >> 20) & 0xF) + T * ((info >> 24) & 0xF);
        const int64_t g0 = col_g0(ci);
        const EpiIn &e = in[ci & 1];
#pragma unroll
        for (int k = 0; k < P; ++k) {
          const int gz = cz * P + k;
          if (FULL || (gz >= p.z_own_lo && gz < p.z_own_hi)) {
            double y = (k == 0) ? st.carry[ci] : 0.0;
#pragma unroll
            for (int kp = 0; kp < N1; ++kp) y = fma(PMG_MZ(k, kp), g[kp], fma(PMG_KZ(k, kp), m[kp], y));
            const bool dirz = FULL ? false : ((gz == 0 && (p.faces >> 4 & 1u)) || (gz == p.Nz - 1 && (p.faces >> 5 & 1u)));
            const int ztype = FULL ? k : pmg_sweep_pos_type<P>(gz, p.Nz);
            epi_store<MODE>(p, g0 + k * plane, y, e.u[k], e.b[k], e.xo[k], dirxy, dirz, tbase + T * T * ztype);
          }
        }
      }
      st.carry[ci] = top; st.gP[ci] = g[P]; st.mP[ci] = m[P];
    }
  }

  template <int MODE>
  static PMG_HD void phase3_t(const PmgSweepParams<P> &p, const TileGeom &t, ThreadState &st, const double *A, const double *B,
                              int cz0, int nlay, int cz_write)
  {
#pragma unroll
    for (int l = 0; l < LZ; ++l) {
      if (l < nlay) {
        const int cz = cz0 + l;
        const bool write = (cz >= cz_write);
        // plane 0 of the mesh (Dirichlet or not, its inverse-diagonal type differs) takes the general path
        const bool full = write && cz > 0 && cz * P >= p.z_own_lo && cz * P + P <= p.z_own_hi;
        if (full) phase3_layer<MODE, true>(p, t, st, A, B, l, cz, true);
        else phase3_layer<MODE, false>(p, t, st, A, B, l, cz, write);
      }
    }
  }

  static PMG_HD void phase3(const PmgSweepParams<P> &p, const TileGeom &t, ThreadState &st, const double *A, const double *B,
                            int cz0, int nlay, int cz_write)
  {
    switch (p.mode) {
      case PMG_MODE_APPLY: phase3_t<PMG_MODE_APPLY>(p, t, st, A, B, cz0, nlay, cz_write); break;
      case PMG_MODE_RESIDUAL: phase3_t<PMG_MODE_RESIDUAL>(p, t, st, A, B, cz0, nlay, cz_write); break;
      case PMG_MODE_CHEB_FIRST: phase3_t<PMG_MODE_CHEB_FIRST>(p, t, st, A, B, cz0, nlay, cz_write); break;
      default: phase3_t<PMG_MODE_CHEB_STEP>(p, t, st, A, B, cz0, nlay, cz_write); break;
    }
  }

  // after the prologue: (g, m) of the first layer's plane 0 into the registers
  static PMG_HD void phase3_init(ThreadState &st, const double *A, const double *B)
  {
#pragma unroll
    for (int ci = 0; ci < NCOL; ++ci) {
      const int info = st.info[ci];
      if (info < 0) continue;
      const int ox = info & 0xFF, oy = (info >> 8) & 0xFF;
      st.gP[ci] = A[(oy + P) * XP + P + ox];
      st.mP[ci] = B[oy * XP + P + ox];
    }
  }

  // the mesh's top plane: the carried sums are complete (no cell layer above)
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
      double uc = 0.0, bb = 0.0, xo = 0.0;
      if (MODE != PMG_MODE_APPLY) { uc = p.u[g]; bb = p.b[g]; }
      if (MODE == PMG_MODE_CHEB_STEP && p.xold) xo = p.xold[g];
      epi_store<MODE>(p, g, st.carry[ci], uc, bb, xo, (info >> 16) & 1, dirz,
                      ((info >> 20) & 0xF) + T * ((info >> 24) & 0xF) + T * T * pmg_sweep_pos_type<P>(gz, p.Nz));
    }
  }

  // APPLY mode: Dirichlet rows in x / y are the identity (:718).  Their lines are copied u -> out here, after the march,
  // for planes [pz_lo, pz_hi): many independent loads in flight instead of one dependent load per store in the epilogue.
  static PMG_HD void fixup_dirichlet_xy(const PmgSweepParams<P> &p, const TileGeom &t, int tid, int pz_lo, int pz_hi)
  {
    const int npz = pz_hi - pz_lo;
    if (npz <= 0) return;
    const int64_t plane = (int64_t)p.Nx * p.Ny;
    const int64_t base = (int64_t)(pz_lo - p.z0) * plane + (int64_t)(t.cy0 * P) * p.Nx + t.cx0 * P;
    for (int side = 0; side < 4; ++side) {
      int len; int64_t first, step;
      if (side == 0) { if (!(t.cy0 == 0 && (p.faces >> 2 & 1u))) continue; len = t.cw; first = 0; step = 1; }
      else if (side == 1) { if (!(t.y_end && (p.faces >> 3 & 1u))) continue; len = t.cw; first = (int64_t)(t.rows - 1) * p.Nx; step = 1; }
      else if (side == 2) { if (!(t.cx0 == 0 && (p.faces & 1u))) continue; len = t.rows; first = 0; step = p.Nx; }
      else { if (!(t.x_end && (p.faces >> 1 & 1u))) continue; len = t.rows; first = t.cw - 1; step = p.Nx; }
      for (int idx = tid; idx < len * npz; idx += NT) {
        const int pz = idx / len, i = idx - pz * len;
        const int64_t g = base + pz * plane + first + i * step;
        p.out[g] = p.u[g];
      }
    }
  }

  static PMG_HD void flush(const PmgSweepParams<P> &p, const TileGeom &t, ThreadState &st, int gz)
  {
    switch (p.mode) {
      case PMG_MODE_APPLY: flush_t<PMG_MODE_APPLY>(p, t, st, gz); break;
      case PMG_MODE_RESIDUAL: flush_t<PMG_MODE_RESIDUAL>(p, t, st, gz); break;
      case PMG_MODE_CHEB_FIRST: flush_t<PMG_MODE_CHEB_FIRST>(p, t, st, gz); break;
      default: flush_t<PMG_MODE_CHEB_STEP>(p, t, st, gz); break;
    }
  }

  // ---- the tile program -----------------------------------------------------
  // Exec provides: template<F> void for_each_thread(F f)  with f(int tid, ThreadState&);  void sync()
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
    double *B = smem + B_OFFSET;

    // prologue: plane 0 of the first layer through the y and x sweeps; meanwhile the first step's planes arrive
    int cur = 1;
    ex.for_each_thread([&](int tid, ThreadState &st) {
      int n1 = cz_end - cz_first; if (n1 > LZ) n1 = LZ;
      stage(p, t, tid, smem, cz_first * P, 1);
      stage(p, t, tid, smem + ABUF, cz_first * P + 1, n1 * P);
      decode(p, t, tid, st);
      pmg_sweep_cp_async_wait_all();
    });
    ex.sync();
    ex.for_each_thread([&](int tid, ThreadState &) { phase1(p, t, tid, smem, B, cz_first * P, 1); });
    ex.sync();
    ex.for_each_thread([&](int, ThreadState &st) { phase2(p, t, st, smem, B, 1, -1); });
    ex.sync();
    ex.for_each_thread([&](int, ThreadState &st) { phase3_init(st, smem, B); });
    ex.sync();

    for (int cz = cz_first; cz < cz_end; cz += LZ) {
      double *A = smem + cur * ABUF;
      int nlay = cz_end - cz; if (nlay > LZ) nlay = LZ;
      int nnext = cz_end - (cz + LZ); if (nnext > LZ) nnext = LZ;
      // the next step's u planes land in the other staging buffer while this step computes
      if (nnext > 0)
        ex.for_each_thread([&](int tid, ThreadState &) { stage(p, t, tid, smem + (cur ^ 1) * ABUF, (cz + LZ) * P + 1, nnext * P); });
#ifndef PMG_EXP_NOP1
      ex.for_each_thread([&](int tid, ThreadState &) { phase1(p, t, tid, A, B, cz * P + 1, nlay * P); });
#endif
      ex.sync();
#ifndef PMG_EXP_NOP2
      ex.for_each_thread([&](int, ThreadState &st) { phase2(p, t, st, A, B, nlay * P, (nnext > 0) ? (cz + LZ) * P : -1); });
#endif
      ex.sync();
      ex.for_each_thread([&](int, ThreadState &st) {
#ifndef PMG_EXP_NOP3
        phase3(p, t, st, A, B, cz, nlay, cz_begin);
#endif
        pmg_sweep_cp_async_wait_all();
      });
      ex.sync();
      cur ^= 1;
    }
    // top plane of the mesh (owned by the chunk that ends there)
    const bool top = (cz_end == p.cz_hi && cz_end * P < p.z_own_hi);
    if (top) ex.for_each_thread([&](int, ThreadState &st) { flush(p, t, st, cz_end * P); });
    if (p.mode == PMG_MODE_APPLY) {
      int pz_lo = cz_begin * P, pz_hi = cz_end * P + (top ? 1 : 0);
      if (pz_lo < p.z_own_lo) pz_lo = p.z_own_lo;
      if (pz_hi > p.z_own_hi) pz_hi = p.z_own_hi;
      ex.for_each_thread([&](int tid, ThreadState &) { fixup_dirichlet_xy(p, t, tid, pz_lo, pz_hi); });
    }
  }
#undef PMG_M
#undef PMG_KX
#undef PMG_KY
#undef PMG_MZ
#undef PMG_KZ
};
