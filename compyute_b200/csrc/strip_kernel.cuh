// "Strip" (shared-halo) implicit-GEMM convolution for stride-1 / dilation-1 / same-padded KxK layers with few channels
// (bf16, tcgen05).  Replaces the tap-per-k-iteration im2col loads of tc_kernel<OP_CONV> where those are L2->SM bound.
//
// The activation tensor is staged ZERO-PADDED and channels-last, act_pad[B][Hp][Wp][Cp] (Hp = H + 2P, Wp = W + 2P), and
// read as one long matrix of pixel rows [R = B*Hp*Wp][Cp].  Output lane m is the padded-linear index of the window's
// top-left input pixel, m = b*Hp*Wp + h*Wp + w, so filter tap (j, k) of lane m reads row m + j*Wp + k: for a tile of 128
// consecutive lanes all K*K taps read from ONE strip of 128 + (K-1)*(Wp+1) consecutive rows.  The producer TMA-loads that
// strip once per (tile, 64-channel chunk); the MMA warp issues the K*K taps as UMMA instructions whose A descriptors start
// tap_off rows into the strip (the 128B swizzle is a function of the shared-memory address, so a descriptor may start at
// any 128-byte row of a tile that TMA wrote at a 1024-byte aligned base — tools/halo_probe.cu).  Lanes with h >= H or
// w >= W (the 2P pad columns / rows: 1 - H*W/(Hp*Wp) of the MMAs) are computed and dropped by the epilogue.
// Filter tiles [BN][64] per (tap, chunk) either stay RESIDENT in shared memory for the life of the persistent CTA
// (one N tile and T*chunks*BN*128 bytes fit) or stream through a ring like tc_kernel's B operand.
//
//   warp 0  TMA producer (strips + filter tiles)      warp 1  MMA issuer      warp 2  TMEM allocator      warps 4-7  epilogue
#pragma once
#include <cuda_bf16.h>

#include "strip_params.cuh"
#include "tc_kernel.cuh"  // col_sums_32x32, PTX wrappers

namespace cpt {
namespace tc {

template <int BN>
__global__ void __launch_bounds__((EpiCfg<BN>::THREADS), 1) strip_conv_kernel(const __grid_constant__ StripParams p) {
  constexpr uint32_t IDESC = make_idesc(true, false, false, 128, BN);
  constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  constexpr uint32_t B_BYTES = BN * 128;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // [0, 1024): barriers | strip units | filter tiles
  auto ufull = [&](int u) { return smem_base + 8u * u; };
  auto uempty = [&](int u) { return smem_base + 8u * (STRIP_MAX_UNITS + u); };
  auto bfull = [&](int s) { return smem_base + 8u * (2 * STRIP_MAX_UNITS + s); };
  auto bempty = [&](int s) { return smem_base + 8u * (2 * STRIP_MAX_UNITS + STRIP_MAX_BSTAGES + s); };
  auto tfull = [&](int a) { return smem_base + 8u * (2 * STRIP_MAX_UNITS + 2 * STRIP_MAX_BSTAGES + a); };
  auto tempty = [&](int a) { return smem_base + 8u * (2 * STRIP_MAX_UNITS + 2 * STRIP_MAX_BSTAGES + 2 + a); };
  const uint32_t tmem_slot = smem_base + 8u * (2 * STRIP_MAX_UNITS + 2 * STRIP_MAX_BSTAGES + 4);
  const uint32_t unit_base = smem_base + 1024u;
  const uint32_t b_base = unit_base + (uint32_t)(p.n_units * p.unit_bytes);
  auto unit_smem = [&](int u) { return unit_base + (uint32_t)(u * p.unit_bytes); };
  auto b_smem = [&](int s) { return b_base + (uint32_t)s * B_BYTES; };

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  if (warp == 0 && elect_one_sync()) {
    prefetch_tmap(&p.tmA);
    prefetch_tmap(&p.tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int u = 0; u < p.n_units; ++u) { mbar_init(ufull(u), 1); mbar_init(uempty(u), 1); }
    for (int s = 0; s < p.b_stages; ++s) { mbar_init(bfull(s), 1); mbar_init(bempty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), EpiCfg<BN>::WARPS); }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc<false>(tmem_slot, TMEM_COLS);
    tmem_relinquish<false>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int total_tiles = p.m_tiles * p.n_tiles;
  const int n_ctas = gridDim.x, cta = blockIdx.x;
  const int my_tiles = cta < total_tiles ? (total_tiles - cta + n_ctas - 1) / n_ctas : 0;
  const int T = p.T, cchunks = p.cchunks;

  // Slot / phase counters of the two rings are advanced incrementally and every quantity of the role loops is derived from
  // kernel parameters and loop counters only: no runtime division, so the compiler keeps addresses and descriptors on the
  // uniform datapath (a `k % n_units` here put them in vector registers: R2UR per MMA operand, ~420 cycles per tap).
  if (warp == 0) {
    // =========================== TMA producer ===========================
    const int total_units = my_tiles * cchunks;
    int iu = 0, iu_slot = 0, iu_ti = 0, iu_cc = 0;      // next strip to issue: sequence index, ring slot, (tile, chunk)
    uint32_t iu_phase = 0;
    int bslot = 0;
    uint32_t bphase = 0;
    bool ok = true;
    auto issue_strip = [&]() -> bool {
      const int t = cta + iu_ti * n_ctas;
      const int m0 = (t % p.m_tiles) * 128;
      if (!__all_sync(0xffffffffu, mbar_wait(uempty(iu_slot), iu_phase ^ 1))) { atomicExch(p.status, 11); return false; }
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(ufull(iu_slot), (uint32_t)p.unit_bytes);
        for (int l = 0; l < p.n_loads; ++l)
          tma_load_2d<false>(&p.tmA, ufull(iu_slot), unit_smem(iu_slot) + (uint32_t)(l * p.box_rows * 128), iu_cc * 64, m0 + l * p.box_rows);
      }
      __syncwarp();
      ++iu;
      if (++iu_slot == p.n_units) { iu_slot = 0; iu_phase ^= 1; }
      if (++iu_cc == cchunks) { iu_cc = 0; ++iu_ti; }
      return true;
    };
    const int la = T > 1 ? 1 : 0;
    if (total_units > 0) ok = issue_strip();
    for (int ti = 0; ti < my_tiles && ok; ++ti) {
      const int t = cta + ti * n_ctas;
      const int n0 = (t / p.m_tiles) * BN;
      const bool load_b = !p.resident || ti == 0;
      for (int cc = 0; cc < cchunks && ok; ++cc) {
        for (int tap = 0; tap < T && ok; ++tap) {
          if (tap == la && iu < total_units) { ok = issue_strip(); if (!ok) break; }
          if (load_b) {
            const int slot = p.resident ? cc * T + tap : bslot;
            if (!p.resident && !__all_sync(0xffffffffu, mbar_wait(bempty(slot), bphase ^ 1))) { atomicExch(p.status, 12); ok = false; break; }
            if (elect_one_sync()) {
              mbar_arrive_expect_tx(bfull(slot), B_BYTES);
              tma_load_2d<false>(&p.tmB, bfull(slot), b_smem(slot), tap * p.wk_cols + cc * 64, n0);
            }
            __syncwarp();
            if (!p.resident && ++bslot == p.b_stages) { bslot = 0; bphase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    int acc = 0, uslot = 0, bslot = 0;
    uint32_t acc_phase = 0, uphase = 0, bphase = 0;
    bool ok = true;
    const uint32_t a_lo_unit0 = smem_desc_lo(unit_smem(0), 16), b_lo_slot0 = smem_desc_lo(b_smem(0), 16);
    for (int ti = 0; ti < my_tiles && ok; ++ti) {
      if (!__all_sync(0xffffffffu, mbar_wait(tempty(acc), acc_phase ^ 1))) { atomicExch(p.status, 13); break; }
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
      for (int cc = 0; cc < cchunks && ok; ++cc) {
        if (!__all_sync(0xffffffffu, mbar_wait(ufull(uslot), uphase))) { atomicExch(p.status, 14); ok = false; break; }
        tc_fence_after();
        const uint32_t a_lo0 = a_lo_unit0 + (uint32_t)uslot * ((uint32_t)p.unit_bytes >> 4);
        for (int tap = 0; tap < T; ++tap) {
          const int bs = p.resident ? cc * T + tap : bslot;
          if (!p.resident || ti == 0) {
            if (!__all_sync(0xffffffffu, mbar_wait(bfull(bs), p.resident ? 0u : bphase))) { atomicExch(p.status, 15); ok = false; break; }
            tc_fence_after();
          }
          if (elect_one_sync()) {
            constexpr uint32_t HI = smem_desc_hi(1024);
            const uint32_t a_lo = a_lo0 + (uint32_t)p.tap_off[tap] * 8u;   // tap's row offset: rows x 128 B >> 4
            const uint32_t b_lo = b_lo_slot0 + (uint32_t)bs * (B_BYTES >> 4);
#pragma unroll
            for (int s = 0; s < 4; ++s)
              umma<true, false>(d_tmem, smem_desc_pack(a_lo + 2 * s, HI), smem_desc_pack(b_lo + 2 * s, HI), IDESC, (uint32_t)((cc | tap | s) != 0));
            if (!p.resident) umma_commit<false>(bempty(bs));
            if (tap == T - 1) umma_commit<false>(uempty(uslot));   // strip free once these MMAs have read it
          }
          __syncwarp();
          if (!p.resident && ++bslot == p.b_stages) { bslot = 0; bphase ^= 1; }
        }
        if (++uslot == p.n_units) { uslot = 0; uphase ^= 1; }
      }
      if (!ok) break;
      if (elect_one_sync()) umma_commit<false>(tfull(acc));
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // =========================== epilogue ===========================
    const int ew = warp & 3;
    constexpr int EPI_SPLIT = EpiCfg<BN>::WARPS / 4;  // warps sharing a TMEM lane quarter: each takes 1 / EPI_SPLIT of the columns
    const int half = (warp - 4) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    __shared__ float stat_acc[4][BN][2];
    const bool do_stats = p.stats != nullptr;
    int stat_n_tile = -1;
    auto stats_flush = [&]() {
      if (stat_n_tile >= 0) {
        float* dstp = p.stats + ((long long)(blockIdx.x * 4 + ew) * p.N) * 2;
        for (int c = half * (BN / EPI_SPLIT) + lane; c < (half + 1) * (BN / EPI_SPLIT); c += 32) {
          const int col = stat_n_tile * BN + c;
          if (col < p.N) { dstp[2 * col] = stat_acc[ew][c][0]; dstp[2 * col + 1] = stat_acc[ew][c][1]; }
        }
      }
    };
    for (int ti = 0; ti < my_tiles; ++ti) {
      const int t = cta + ti * n_ctas;
      const int m_tile = t % p.m_tiles, n_tile = t / p.m_tiles;
      const int m = m_tile * 128 + ew * 32 + lane, n0 = n_tile * BN;
      if (do_stats && n_tile != stat_n_tile) {
        stats_flush();
        stat_n_tile = n_tile;
        for (int c = half * (BN / EPI_SPLIT) + lane; c < (half + 1) * (BN / EPI_SPLIT); c += 32) stat_acc[ew][c][0] = stat_acc[ew][c][1] = 0.f;
        __syncwarp();
      }
      // lane -> (image, row, col) of the padded grid; pad rows / columns are dropped
      const int b = m / p.HpWp, rem = m - b * p.HpWp, h = rem / p.Wp, w = rem - h * p.Wp;
      const bool m_ok = m < p.M_lanes && h < p.H && w < p.W;
      float* dst = p.out + (long long)b * p.img_stride + (long long)h * p.W + w + (long long)n0 * p.col_stride;
      const long long cs = p.col_stride;
      if (!mbar_wait(tfull(acc), acc_phase)) { atomicExch(p.status, 16); break; }
      tc_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BN);
      constexpr int NCH = BN / 32;
      constexpr int CPW = NCH / EPI_SPLIT;
      const int c_lo = half * CPW;
      uint32_t va[32], vb[32];
      tmem_ld_32x32(tbase + (uint32_t)(c_lo * 32), va);
      auto release_acc = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty(acc));
      };
      auto store_chunk = [&](const uint32_t (&v)[32], int c) {
        const int cbase = n0 + c * 32;
        if (do_stats) {
          float a[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) a[j] = m_ok ? __uint_as_float(v[j]) : 0.f;
          const float s1 = col_sums_32x32(a, lane);
#pragma unroll
          for (int j = 0; j < 32; ++j) { const float x = m_ok ? __uint_as_float(v[j]) : 0.f; a[j] = x * x; }
          const float s2 = col_sums_32x32(a, lane);
          stat_acc[ew][c * 32 + lane][0] += s1;
          stat_acc[ew][c * 32 + lane][1] += s2;
        }
        float* q = dst + (long long)(c * 32) * cs;
        float bl = 0.f;
        if (p.bias && cbase + lane < p.N) bl = __ldg(p.bias + cbase + lane);
        if (cbase + 32 <= p.N) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float val = __uint_as_float(v[j]) + __shfl_sync(0xffffffffu, bl, j);
            if (m_ok) *q = val;
            q += cs;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float val = __uint_as_float(v[j]) + __shfl_sync(0xffffffffu, bl, j);
            if (m_ok && cbase + j < p.N) *q = val;
            q += cs;
          }
        }
      };
#pragma unroll 1
      for (int k = 0; k < CPW; k += 2) {
        tmem_ld_wait();
        if (k + 1 < CPW) tmem_ld_32x32(tbase + (uint32_t)((c_lo + k + 1) * 32), vb);
        if (k + 1 >= CPW) release_acc();
        store_chunk(va, c_lo + k);
        if (k + 1 < CPW) {
          tmem_ld_wait();
          if (k + 2 < CPW) tmem_ld_32x32(tbase + (uint32_t)((c_lo + k + 2) * 32), va);
          if (k + 2 >= CPW) release_acc();
          store_chunk(vb, c_lo + k + 1);
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (do_stats) { __syncwarp(); stats_flush(); }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<false>(tmem_base, TMEM_COLS);
  }
}

}  // namespace tc
}  // namespace cpt
