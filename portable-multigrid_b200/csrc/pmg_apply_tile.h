// pmg_apply_tile.h -- the cell-loop Laplace apply, fused with the smoother update: cell-tile kernel.
// (The library launches it on the small coarse levels, where its direct loads give the lowest launch-to-result latency;
// large levels run the line-marching kernel of pmg_apply_sweep.h.  PMG_TILE_VARIANT=2|3 forces it everywhere.)
//
// Replaces LaplaceOperator::vmult + LocalLaplaceOperator::operator()
// (reference include/operators/portable_laplace_operator.h:227-381, 557-719) and the
// PreconditionChebyshev vector update that follows it in the reference
// (include/multigrid/portable_v_cycle_multigrid.h:116-125) by ONE pass over HBM.
//
// Design (B200-first, not a translation; see DESIGN.md "Kernel K1"):
//  * Structured box mesh, lexicographic DoF vector: indices, Dirichlet masks and geometry
//    are computed, not streamed (the reference streams ~112 B per local DoF of them).
//  * Owner-computes, no atomics, no zeroing pass: a CTA owns BX x BY cell columns, marches
//    through cell layers in z and recomputes one halo cell row/column on its low sides, so every
//    owned DoF's A*u is complete inside the CTA and the epilogue (residual / Chebyshev step /
//    Dirichlet identity) can be applied before the single store.
//  * Per cell the operator is applied in the simultaneous-diagonalisation basis of the 1-D
//    mass/stiffness pencil: A_cell = (S^T x S^T x S^T) diag(cx l_a + cy l_b + cz l_c) (S x S x S)
//    -- 6 one-dimensional sweeps instead of the reference's 12 + q-point op, same result to
//    round-off on affine cells (tests bound the difference by 1e-12 rel. l2).
//  * A work item is (cell, line j): it keeps an (x,z) slab of the cell in registers, does the x and
//    z sweeps there, carries the shared z-plane to the next layer in registers (both the
//    transformed input and the partial output sums), and exchanges only for the y sweep through
//    shared memory.
//
// The algorithm is written against an executor (for_each_thread / sync) so the very same source
// runs as the CUDA kernel (csrc/pmg_apply.cu) and under the host emulator used by the CPU test
// suite (tests/emu/emu_apply.cpp).  The emulator is test infrastructure, not a fallback.
#pragma once
#include <stdint.h>

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

template <int P>
struct PmgApplyParams {
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
  // 1-D tables: S maps nodal -> pencil eigenbasis (row a, column i), lam eigenvalues,
  // c[d] = h_e h_f / h_d
  double S[(P + 1) * (P + 1)];
  double lam[P + 1];
  double c[3];
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

template <int P>
PMG_HD int pmg_pos_type(int g, int N)
{
  return (g == 0) ? P : (g == N - 1) ? P + 1 : g % P;
}

template <int P, int BX, int BY>
struct PmgApplyTile {
  static constexpr int N1 = P + 1;
  static constexpr int CXC = BX + 1, CYC = BY + 1; // computed cells per layer (incl. low halo)
  static constexpr int NCELL = CXC * CYC;
  static constexpr int NITEM = NCELL * N1;
  static constexpr int NT = ((NITEM + 31) / 32) * 32;
  // exchange tile T1[m][item][a]: the a-block of an item is padded to an odd length so that lanes
  // (= consecutive items) hit different banks; tools/bank_conflicts.py models the access patterns
  static constexpr int AP = N1 | 1;
  static constexpr int T1_SIZE = N1 * NITEM * AP;
  static constexpr bool Y_A_FASTEST = (N1 % 2) == 1; // phase-Y lane mapping (a fastest for odd N1)
  static constexpr int OXPAD = (P == 1 || P == 3 || P == 7) ? 1 : 0;
  static constexpr int OX = CXC * N1 + OXPAD, OY = CYC * N1; // cell-local output plane
  static constexpr int O_SIZE = P * OX * OY;
  // exchange tile and output planes are separate buffers (3 barriers per layer) when both fit,
  // otherwise the output planes alias the exchange tile (5 barriers per layer)
  static constexpr bool ALIAS = (T1_SIZE + O_SIZE) * 8 > 200 * 1024;
  static constexpr int O_OFFSET = ALIAS ? 0 : T1_SIZE;
  static constexpr int SMEM_DOUBLES = ALIAS ? ((T1_SIZE > O_SIZE) ? T1_SIZE : O_SIZE) : T1_SIZE + O_SIZE;
  // (staging the next layer's planes with cp.async was measured 5-15 % slower than these direct loads, which hit L1/L2)
  // epilogue iteration space: owned dof columns of the tile incl. the optional high face
  static constexpr int EW = BX * P + 1, EH = BY * P + 1;
  static constexpr int ECOLS = EW * EH;
  static constexpr int EITER = (ECOLS + NT - 1) / NT;

  struct ThreadState {
    double X[N1][N1]; // [k or m][a]
    double cin[N1];   // transformed input of the plane shared with the next layer
    double cout[N1];  // partial output sums of that plane
    int tcx, tcy, j;  // item decode
    int cx, cy;       // global cell
    int valid;        // cell inside the mesh and item < NITEM
    int zero_row, zero_i0, zero_iP;
  };

  // ---- helpers ------------------------------------------------------------
  static PMG_HD void decode(int tid, int cx0, int cy0, const PmgApplyParams<P> &p, ThreadState &st)
  {
    st.tcx = tid % CXC;
    st.j = (tid / CXC) % N1;
    st.tcy = tid / (CXC * N1);
    st.cx = cx0 - 1 + st.tcx;
    st.cy = cy0 - 1 + st.tcy;
    st.valid = (tid < NITEM) && st.cx >= 0 && st.cx < p.nx && st.cy >= 0 && st.cy < p.ny;
    const int gy = st.cy * P + st.j;
    st.zero_row = (gy == 0 && (p.faces >> 2 & 1u)) || (gy == p.Ny - 1 && (p.faces >> 3 & 1u));
    st.zero_i0 = (st.cx == 0 && (p.faces & 1u));
    st.zero_iP = (st.cx == p.nx - 1 && (p.faces >> 1 & 1u));
#pragma unroll
    for (int a = 0; a < N1; ++a) { st.cin[a] = 0.0; st.cout[a] = 0.0; }
  }

  static PMG_HD int item_index(int tcx, int j, int tcy) { return tcx + CXC * (j + N1 * tcy); }
  static PMG_HD int t1_index(int item, int m, int a) { return (m * NITEM + item) * AP + a; }

  // load one x-line of plane gz and transform it along x: X[k][a] = sum_i S[a][i] u[i]
  static PMG_HD void load_xfwd(const PmgApplyParams<P> &p, const ThreadState &st, const double *row, int gz, double *Xk)
  {
    double v[N1];
#pragma unroll
    for (int i = 0; i < N1; ++i) v[i] = row[i];
    if (st.zero_row | st.zero_i0 | st.zero_iP | (gz == 0) | (gz == p.Nz - 1)) { // rare: Dirichlet values read as 0
      const bool zero_plane = (gz == 0 && (p.faces >> 4 & 1u)) || (gz == p.Nz - 1 && (p.faces >> 5 & 1u));
      if (zero_plane || st.zero_row) {
#pragma unroll
        for (int i = 0; i < N1; ++i) v[i] = 0.0;
      }
      if (st.zero_i0) v[0] = 0.0;
      if (st.zero_iP) v[P] = 0.0;
    }
#pragma unroll
    for (int a = 0; a < N1; ++a) {
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < N1; ++i) s += p.S[a * N1 + i] * v[i];
      Xk[a] = s;
    }
  }

  // ---- phases (one call per thread; separated by sync) ----------------------
  // F: load the layer's new planes, x-forward each, accumulate the z-forward sweep, publish to T1
  static PMG_HD void phase_forward(const PmgApplyParams<P> &p, ThreadState &st, double *smem, int cz, bool first_layer)
  {
    if (!st.valid) return;
#pragma unroll
    for (int m = 0; m < N1; ++m)
#pragma unroll
      for (int a = 0; a < N1; ++a) st.X[m][a] = 0.0;
    const int64_t plane = (int64_t)p.Nx * p.Ny;
    const double *row = p.u + (int64_t)(cz * P - p.z0) * plane + (int64_t)(st.cy * P + st.j) * p.Nx + st.cx * P;
#pragma unroll
    for (int k = 0; k < N1; ++k) {
      double xk[N1];
      if (k == 0 && !first_layer) {
#pragma unroll
        for (int a = 0; a < N1; ++a) xk[a] = st.cin[a];
      } else {
        load_xfwd(p, st, row + k * plane, cz * P + k, xk);
      }
      if (k == P) {
#pragma unroll
        for (int a = 0; a < N1; ++a) st.cin[a] = xk[a];
      }
#pragma unroll
      for (int m = 0; m < N1; ++m)
#pragma unroll
        for (int a = 0; a < N1; ++a) st.X[m][a] += p.S[m * N1 + k] * xk[a];
    }
    double *dst = smem + t1_index(item_index(st.tcx, st.j, st.tcy), 0, 0);
#pragma unroll
    for (int m = 0; m < N1; ++m)
#pragma unroll
      for (int a = 0; a < N1; ++a) dst[m * (NITEM * AP) + a] = st.X[m][a];
  }

  // Y: item (cell, a): y-forward, diagonal scaling, y-backward, in place in T1, one m-column at a time.
  // The (cell, a) pair is decoded from the thread id independently of the (cell, j) decode of the other
  // phases so that the lanes of a warp read consecutive words.
  static PMG_HD void phase_y(const PmgApplyParams<P> &p, int tid, int cx0, int cy0, double *smem)
  {
    if (tid >= NITEM) return;
    int a, tcx, tcy;
    if (Y_A_FASTEST) { a = tid % N1; tcx = (tid / N1) % CXC; tcy = tid / (N1 * CXC); }
    else { tcx = tid % CXC; a = (tid / CXC) % N1; tcy = tid / (CXC * N1); }
    const int cx = cx0 - 1 + tcx, cy = cy0 - 1 + tcy;
    if (cx < 0 || cx >= p.nx || cy < 0 || cy >= p.ny) return;
    const double base = p.c[0] * p.lam[a];
    double *col = smem + t1_index(item_index(tcx, 0, tcy), 0, a);
    constexpr int JSTRIDE = CXC * AP;   // distance between consecutive j of one cell
    constexpr int MSTRIDE = NITEM * AP; // distance between consecutive m
#pragma unroll
    for (int m = 0; m < N1; ++m) {
      double T[N1], Y[N1];
#pragma unroll
      for (int j = 0; j < N1; ++j) T[j] = col[j * JSTRIDE + m * MSTRIDE];
#pragma unroll
      for (int bb = 0; bb < N1; ++bb) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < N1; ++j) s += p.S[bb * N1 + j] * T[j];
        Y[bb] = s * (base + p.c[1] * p.lam[bb] + p.c[2] * p.lam[m]);
      }
#pragma unroll
      for (int j = 0; j < N1; ++j) {
        double s = 0.0;
#pragma unroll
        for (int bb = 0; bb < N1; ++bb) s += p.S[bb * N1 + j] * Y[bb];
        col[j * JSTRIDE + m * MSTRIDE] = s;
      }
    }
  }

  static PMG_HD void phase_back_read(ThreadState &st, const double *smem)
  {
    if (!st.valid) return;
    const double *src = smem + t1_index(item_index(st.tcx, st.j, st.tcy), 0, 0);
#pragma unroll
    for (int m = 0; m < N1; ++m)
#pragma unroll
      for (int a = 0; a < N1; ++a) st.X[m][a] = src[m * (NITEM * AP) + a];
  }

  static PMG_HD void phase_back_write(const PmgApplyParams<P> &p, ThreadState &st, double *smem, bool first_layer, bool write_out)
  {
    if (!st.valid) return;
    if (!ALIAS) {
      const double *src = smem + t1_index(item_index(st.tcx, st.j, st.tcy), 0, 0);
#pragma unroll
      for (int m = 0; m < N1; ++m)
#pragma unroll
        for (int a = 0; a < N1; ++a) st.X[m][a] = src[m * (NITEM * AP) + a];
    }
#pragma unroll
    for (int k = 0; k < N1; ++k) {
      double xk[N1];
#pragma unroll
      for (int a = 0; a < N1; ++a) {
        double s = 0.0;
#pragma unroll
        for (int m = 0; m < N1; ++m) s += p.S[m * N1 + k] * st.X[m][a];
        xk[a] = s;
      }
      if (k == 0 && !first_layer) {
#pragma unroll
        for (int a = 0; a < N1; ++a) xk[a] += st.cout[a];
      }
      if (k == P) {
#pragma unroll
        for (int a = 0; a < N1; ++a) st.cout[a] = xk[a];
      } else if (write_out) {
        double *dst = smem + O_OFFSET + (k * OY + st.tcy * N1 + st.j) * OX + st.tcx * N1;
#pragma unroll
        for (int i = 0; i < N1; ++i) {
          double s = 0.0;
#pragma unroll
          for (int a = 0; a < N1; ++a) s += p.S[a * N1 + i] * xk[a];
          dst[i] = s;
        }
      }
    }
  }

  // flush of the top plane after the last layer: x-backward of the carried sums into O plane 0
  static PMG_HD void phase_flush(const PmgApplyParams<P> &p, ThreadState &st, double *smem)
  {
    if (!st.valid) return;
    double *dst = smem + O_OFFSET + (st.tcy * N1 + st.j) * OX + st.tcx * N1;
#pragma unroll
    for (int i = 0; i < N1; ++i) {
      double s = 0.0;
#pragma unroll
      for (int a = 0; a < N1; ++a) s += p.S[a * N1 + i] * st.cout[a];
      dst[i] = s;
    }
  }

  // E: owner epilogue for planes gz0 .. gz0+NPL-1 of this tile.  The iteration space is the
  // compile-time (BX*P+1) x (BY*P+1) grid of owned dof columns; everything that does not depend on the
  // plane (gather offsets, global offset, boundary flags, table index) is computed once per column, and
  // the global loads of all planes of a column are issued before any of them is used.
  template <int MODE, int NPL>
  static PMG_HD void epilogue_t(const PmgApplyParams<P> &p, int tid, const double *smem, int cx0, int cy0, int gz0)
  {
    const double *O = smem + O_OFFSET;
    const int gx_end = (cx0 + BX >= p.nx) ? p.Nx : (cx0 + BX) * P;
    const int gy_end = (cy0 + BY >= p.ny) ? p.Ny : (cy0 + BY) * P;
    constexpr int T = P + 2;
    const int64_t plane = (int64_t)p.Nx * p.Ny;
    const bool have_xold = (MODE == PMG_MODE_CHEB_STEP) && (p.xold != nullptr);
#pragma unroll
    for (int it = 0; it < EITER; ++it) {
      const int col = tid + it * NT;
      if (col >= ECOLS) break;
      const int ix = col % EW, iy = col / EW;
      const int gx = cx0 * P + ix, gy = cy0 * P + iy;
      if (gx >= gx_end || gy >= gy_end) continue;
      const int64_t g0 = (int64_t)(gz0 - p.z0) * plane + (int64_t)gy * p.Nx + gx;
      // issue every global load of this column first
      double uc[NPL], bb[NPL], xo[NPL];
      bool act[NPL];
#pragma unroll
      for (int k = 0; k < NPL; ++k) {
        act[k] = (gz0 + k >= p.z_own_lo) && (gz0 + k < p.z_own_hi);
        const int64_t g = act[k] ? g0 + k * plane : g0;
        uc[k] = p.u[g];
        bb[k] = (MODE != PMG_MODE_APPLY) ? p.b[g] : 0.0;
        xo[k] = have_xold ? p.xold[g] : 0.0;
      }
      // cell-local contributions: tile-local dof coordinates are (ix + P, iy + P)
      const int tx1 = ix / P + 1, il = ix % P, ty1 = iy / P + 1, jl = iy % P;
      const bool vx0 = (tx1 < CXC) && (cx0 - 1 + tx1 < p.nx);
      const bool vx1 = (il == 0) && (cx0 - 1 + tx1 - 1 >= 0);
      const bool vy0 = (ty1 < CYC) && (cy0 - 1 + ty1 < p.ny);
      const bool vy1 = (jl == 0) && (cy0 - 1 + ty1 - 1 >= 0);
      const int ox0 = tx1 * N1 + il, ox1 = (tx1 - 1) * N1 + P;
      const int oy0 = (ty1 * N1 + jl) * OX, oy1 = ((ty1 - 1) * N1 + P) * OX;
      const bool dirxy = (gx == 0 && (p.faces & 1u)) || (gx == p.Nx - 1 && (p.faces >> 1 & 1u)) ||
                         (gy == 0 && (p.faces >> 2 & 1u)) || (gy == p.Ny - 1 && (p.faces >> 3 & 1u));
      const int tbase = pmg_pos_type<P>(gx, p.Nx) + T * pmg_pos_type<P>(gy, p.Ny);
#pragma unroll
      for (int k = 0; k < NPL; ++k) {
        const int gz = gz0 + k;
        const double *Ok = O + k * (OX * OY);
        double y = 0.0;
        if (vx0 && vy0) y += Ok[oy0 + ox0];
        if (vx1 && vy0) y += Ok[oy0 + ox1];
        if (vx0 && vy1) y += Ok[oy1 + ox0];
        if (vx1 && vy1) y += Ok[oy1 + ox1];
        const bool dir = dirxy || (gz == 0 && (p.faces >> 4 & 1u)) || (gz == p.Nz - 1 && (p.faces >> 5 & 1u));
        const double Au = dir ? uc[k] : y;
        double r;
        if (MODE == PMG_MODE_APPLY) {
          r = Au;
        } else if (MODE == PMG_MODE_RESIDUAL) {
          r = bb[k] - Au;
        } else {
          double dinv;
          if (dir) dinv = 1.0;
          else if (p.dinv_vec) dinv = p.dinv_vec[g0 + k * plane];
          else dinv = p.dinv_tab[tbase + T * T * pmg_pos_type<P>(gz, p.Nz)];
          const double corr = p.f2 * dinv * (bb[k] - Au);
          if (MODE == PMG_MODE_CHEB_FIRST) r = uc[k] + corr;
          else r = uc[k] + p.f1 * (uc[k] - xo[k]) + corr;
        }
        if (act[k]) p.out[g0 + k * plane] = r;
      }
    }
  }

  template <int NPL>
  static PMG_HD void phase_epilogue(const PmgApplyParams<P> &p, int tid, const double *smem, int cx0, int cy0, int gz0)
  {
    switch (p.mode) {
      case PMG_MODE_APPLY: epilogue_t<PMG_MODE_APPLY, NPL>(p, tid, smem, cx0, cy0, gz0); break;
      case PMG_MODE_RESIDUAL: epilogue_t<PMG_MODE_RESIDUAL, NPL>(p, tid, smem, cx0, cy0, gz0); break;
      case PMG_MODE_CHEB_FIRST: epilogue_t<PMG_MODE_CHEB_FIRST, NPL>(p, tid, smem, cx0, cy0, gz0); break;
      default: epilogue_t<PMG_MODE_CHEB_STEP, NPL>(p, tid, smem, cx0, cy0, gz0); break;
    }
  }

  // ---- the tile program -----------------------------------------------------
  // Exec provides: template<F> void for_each_thread(F f)  with f(int tid, ThreadState&)
  //                void sync()
  template <class Exec>
  static PMG_HD void run(const PmgApplyParams<P> &p, Exec &ex, double *smem, int tile_x, int tile_y, int chunk)
  {
    const int cx0 = tile_x * BX, cy0 = tile_y * BY;
    const int cz_begin = p.cz_lo + chunk * p.layers_per_chunk;
    int cz_end = cz_begin + p.layers_per_chunk;
    if (cz_end > p.cz_hi) cz_end = p.cz_hi;
    if (cz_begin >= cz_end) return;
    // halo layer below the chunk (its data exist when the plane below is stored locally)
    const bool halo = (cz_begin > 0) && ((cz_begin - 1) * P >= p.z0);
    const int cz_first = halo ? cz_begin - 1 : cz_begin;

    ex.for_each_thread([&](int tid, ThreadState &st) { decode(tid, cx0, cy0, p, st); });

    for (int cz = cz_first; cz < cz_end; ++cz) {
      const bool first = (cz == cz_first);
      const bool write_out = (cz >= cz_begin);
      ex.for_each_thread([&](int, ThreadState &st) { phase_forward(p, st, smem, cz, first); });
      ex.sync();
      ex.for_each_thread([&](int tid, ThreadState &) { phase_y(p, tid, cx0, cy0, smem); });
      ex.sync();
      if (ALIAS) {
        ex.for_each_thread([&](int, ThreadState &st) { phase_back_read(st, smem); });
        ex.sync();
      }
      ex.for_each_thread([&](int, ThreadState &st) { phase_back_write(p, st, smem, first, write_out); });
      ex.sync();
      // !ALIAS: no barrier after the epilogue: the next layer's forward and y phases only touch T1, and the
      // two barriers they end with order this read of O before the next write to it
      if (write_out)
        ex.for_each_thread([&](int tid, ThreadState &) { phase_epilogue<P>(p, tid, smem, cx0, cy0, cz * P); });
      if (ALIAS) ex.sync();
    }
    // top plane of the slab (owned only by the chunk that ends at the top of the mesh)
    if (cz_end == p.cz_hi && cz_end * P < p.z_own_hi) {
      ex.sync();
      ex.for_each_thread([&](int, ThreadState &st) { phase_flush(p, st, smem); });
      ex.sync();
      ex.for_each_thread([&](int tid, ThreadState &) { phase_epilogue<1>(p, tid, smem, cx0, cy0, cz_end * P); });
    }
  }
};
