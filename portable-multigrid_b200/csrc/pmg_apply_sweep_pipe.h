// pmg_apply_sweep_pipe.h -- the line-marching apply kernel with its phases OVERLAPPED inside the CTA (experimental, not launched
// by libpmg.so yet: tools/exp/exp_pipe.cu times it, tests/emu/emu_sweep.cpp checks it against the oracle).
//
// Why (kernel K1 "v6"): on their own the kernel's three parts take 0.32 ms (staging), 0.45 ms (y and x sweeps) and
// 0.36 ms (z sweep + stores) for 100 M DoFs at Q4, together 0.88 ms -- they hardly overlap, because inside a CTA they are phases
// separated by block barriers and two of the three phases have work for only half of the CTA's warps.  Here the CTA has two
// groups of NG threads that work on DIFFERENT steps at the same time:
//     group YX (threads 0 .. NG-1)      step t:     stage u of step t + 1, y sweep, x sweep  -> C/D buffer t & 1
//     group Z  (threads NG .. 2 NG - 1) step t - 1: z sweep + epilogue + store               <- C/D buffer (t - 1) & 1
// with ONE block barrier per step (plus the YX group's own barrier between its two sweeps) instead of three, so that every warp
// has work all the time.  The phases themselves are the functions of PmgSweepTile (csrc/pmg_apply_sweep.h), instantiated for NG
// threads and called with group-local thread indices; only the buffers and the schedule differ:
//     u staging    3 buffers: step s in buffer (s + 1) % 3 (the z sweep of step s - 1 still reads u while step s + 1 is staged)
//     C/D          2 pairs
//     b / x_old    2 boxes (fused modes)
// Shared memory at Q4, 4 x 4 cell columns: 90 KB (APPLY), 127 KB (fused modes; 90 KB with EG = 1: b / x_old read from global
// memory by the Z group, whose load latency the YX group's arithmetic covers).
#pragma once
#include "pmg_apply_sweep.h"

template <int P, int BX, int BY, int LZ, int NG, int US = 0, int FM = -1, int RL = 0, int EG = 0>
struct PmgSweepPipe {
  using Base = PmgSweepTile<P, BX, BY, LZ, NG, US, FM, 1, RL, 0, EG>;
  using ThreadState = typename Base::ThreadState;
  using TileGeom = typename Base::TileGeom;
  static constexpr int NT = 2 * NG;
  static constexpr int CW = Base::CW, RW = Base::RW;
  static constexpr int NA = 3;
  static constexpr int ABUF = Base::ABUF, CBUF = Base::CBUF, EBUF = Base::EBUF;
  static constexpr int CD_OFFSET = NA * ABUF;           // pair i: C at CD_OFFSET + 2 i CBUF, D right after it
  static constexpr int E_OFFSET = CD_OFFSET + 4 * CBUF; // box i: b at E_OFFSET + 2 i EBUF, x_old right after it
  static constexpr int BAR_OFFSET_APPLY = E_OFFSET, BAR_OFFSET_EPI = E_OFFSET + 4 * EBUF;
  static constexpr int NBAR = NA;                       // one mbarrier per u staging buffer
  static PMG_HD constexpr int smem_doubles(bool epilogue_inputs) { return ((epilogue_inputs && !EG) ? BAR_OFFSET_EPI : BAR_OFFSET_APPLY) + NBAR; }
  static constexpr int SMEM_DOUBLES = (EG ? BAR_OFFSET_APPLY : BAR_OFFSET_EPI) + NBAR;
  static_assert(ABUF % 2 == 0 && CBUF % 2 == 0, "16-byte aligned staging buffers");
  static_assert(SMEM_DOUBLES * 8 <= 227 * 1024, "tile does not fit the 227 KB of shared memory of a CTA");

  // Exec provides for_each_thread(f) with f(tid in [0, 2 NG), ThreadState&), sync() and sync_some(n) = barrier of threads 0..n-1
  template <class Exec>
  static PMG_HD void run(const PmgSweepParams<P> &p, Exec &ex, double *smem, int tile_x, int tile_y, int chunk)
  {
    const TileGeom t = Base::geom(p, tile_x, tile_y);
    const int cz_begin = p.cz_lo + chunk * p.layers_per_chunk;
    int cz_end = cz_begin + p.layers_per_chunk;
    if (cz_end > p.cz_hi) cz_end = p.cz_hi;
    if (cz_begin >= cz_end) return;
    const bool halo = (cz_begin > 0) && ((cz_begin - 1) * P >= p.z0);
    const int cz_first = halo ? cz_begin - 1 : cz_begin;
    const int n_steps = (cz_end - cz_first + LZ - 1) / LZ;
    uint64_t *bars = (uint64_t *)(smem + (t.has_e ? BAR_OFFSET_EPI : BAR_OFFSET_APPLY));
    auto Abuf = [&](int i) { return smem + i * ABUF; };
    auto Cbuf = [&](int i) { return smem + CD_OFFSET + 2 * i * CBUF; };
    auto Dbuf = [&](int i) { return smem + CD_OFFSET + 2 * i * CBUF + CBUF; };
    auto Ebuf = [&](int i) { return smem + E_OFFSET + 2 * i * EBUF; };

    // prologue: plane 0 of the first layer -> u buffer 0 and C/D pair 1; the planes of step 0 -> u buffer 1
    int n0 = cz_end - cz_first; if (n0 > LZ) n0 = LZ;
    ex.for_each_thread([&](int tid, ThreadState &) {
      if (tid == 0) {
        for (int i = 0; i < NBAR; ++i) pmg_mbar_init(bars + i, Base::stage_arrivals(NG));
        pmg_mbar_init_fence();
      }
    });
    ex.sync();
    ex.for_each_thread([&](int tid, ThreadState &) {
      if (tid < NG) {
        Base::stage_u(p, t, tid, Abuf(0), bars + 0, cz_first * P, 1);
        Base::stage_u(p, t, tid, Abuf(1), bars + 1, cz_first * P + 1, n0 * P);
        if (tid >= NG - 32) {
          Base::load_u_async(p, t, tid - (NG - 32), Abuf(0), cz_first * P, 1);
          Base::load_u_async(p, t, tid - (NG - 32), Abuf(1), cz_first * P + 1, n0 * P);
          pmg_sweep_cp_async_wait_all();
        }
      }
    });
    ex.sync();
    ex.for_each_thread([&](int tid, ThreadState &st) {
      Base::decode(p, t, tid < NG ? tid : tid - NG, st);
      if (tid < NG) {
        pmg_mbar_wait(bars + 0, 0);
        Base::phase1(p, t, tid, Abuf(0), Cbuf(1), Dbuf(1), cz_first * P, 1);
      }
    });
    ex.sync();
    ex.for_each_thread([&](int tid, ThreadState &st) { if (tid < NG) Base::phase2(p, t, st, Cbuf(1), Dbuf(1), 1); });
    ex.sync();
    ex.for_each_thread([&](int tid, ThreadState &st) {
      if (tid >= NG) { pmg_mbar_wait(bars + 0, 0); Base::phase3_init(p, t, st, Abuf(0), Cbuf(1), Dbuf(1), cz_first * P); }
    });
    ex.sync();

    // steady state: in tick s the YX group works on step s and the Z group on step s - 1.  Step s has its u planes in buffer
    // (s + 1) % 3, as that buffer's use number (s + 1) / 3 (buffer 0 was used once by the prologue): the parity to wait for
    for (int s = 0; s <= n_steps; ++s) {
      const int cz = cz_first + s * LZ;                 // first layer of step s (YX group)
      int nlay = cz_end - cz; if (nlay > LZ) nlay = LZ;
      int nnext = cz_end - (cz + LZ); if (nnext > LZ) nnext = LZ;
      const int czp = cz - LZ;                          // first layer of step s - 1 (Z group)
      int nlayp = cz_end - czp; if (nlayp > LZ) nlayp = LZ;
      ex.for_each_thread([&](int tid, ThreadState &) {
        if (tid < NG && s < n_steps) {
          // u buffer (s + 2) % 3 was last read by the z sweep of step s - 2, in the previous tick
          if (nnext > 0) Base::stage_u(p, t, tid, Abuf((s + 2) % NA), bars + (s + 2) % NA, (cz + LZ) * P + 1, nnext * P);
          if (tid >= NG - 32) {
            if (t.has_e) Base::load_e_async(p, t, tid - (NG - 32), Ebuf(s & 1), cz * P, nlay * P);
            if (nnext > 0) Base::load_u_async(p, t, tid - (NG - 32), Abuf((s + 2) % NA), (cz + LZ) * P + 1, nnext * P);
          }
          pmg_mbar_wait(bars + (s + 1) % NA, ((s + 1) / NA) & 1);
          Base::phase1(p, t, tid, Abuf((s + 1) % NA), Cbuf(s & 1), Dbuf(s & 1), cz * P + 1, nlay * P);
        }
      });
      ex.sync_some(NG);
      ex.for_each_thread([&](int tid, ThreadState &st) {
        if (tid < NG) {
          if (s < n_steps) {
            Base::phase2(p, t, st, Cbuf(s & 1), Dbuf(s & 1), nlay * P);
            if (tid >= NG - 32) pmg_sweep_cp_async_wait_all();
          }
        } else if (s >= 1) {
          pmg_mbar_wait(bars + s % NA, (s / NA) & 1);   // step s - 1: buffer s % 3, use s / 3 (complete since the previous tick)
          Base::phase3(p, t, st, Abuf(s % NA), Cbuf((s - 1) & 1), Dbuf((s - 1) & 1), Ebuf((s - 1) & 1), czp, nlayp, cz_begin);
        }
      });
      ex.sync();
    }
    if (cz_end == p.cz_hi && cz_end * P < p.z_own_hi)
      ex.for_each_thread([&](int tid, ThreadState &st) { if (tid >= NG) Base::flush(p, t, st, cz_end * P); });
  }
};
