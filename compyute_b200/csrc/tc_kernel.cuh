// Persistent, warp-specialised tcgen05 GEMM / implicit-GEMM convolution kernel for sm_100a.
//
//   warp 0 (1 lane)  TMA producer: tiled 2-D or im2col 4-D loads into a ring of 128B-swizzled smem stages
//   warp 1 (1 lane)  MMA issuer:   tcgen05.mma (fp32 accumulate in TMEM), tcgen05.commit
//   warp 2           TMEM allocator (2 x BN columns: double-buffered accumulator)
//   warps 4-7        epilogue: tcgen05.ld -> registers -> (bias) -> global, overlapped with the next tile
//
// CTA2 = false: one CTA per SM computes a 128 x BN tile (cta_group::1, UMMA M=128).
// CTA2 = true : a cluster of two CTAs (an SM pair) computes a 256 x BN tile with cta_group::2 (UMMA M=256): each
//   CTA stages its own 128 rows of A and HALF of B (BN/2 columns); the leader CTA issues the MMAs, which read both
//   CTAs' shared memory and write both CTAs' TMEM.  Per SM this halves the B bytes written by TMA and read by the
//   tensor core, which is what bounds the 1-CTA kernel.
//
// The GEMM "M" (TMEM lane) dimension is always mapped onto the output's CONTIGUOUS dimension (pixels for
// NCHW conv outputs, the feature dimension for row-major Linear outputs, Ci for wgrad partials) so that the
// epilogue's 32-lane stores are 128-byte coalesced without a shared-memory transpose.
#pragma once
#include <cuda_bf16.h>

#include "tc_params.cuh"
#include "tc_ptx.cuh"

namespace cpt {
namespace tc {

// Column sums of a 32 x 32 block held one row per lane (a[j] = column j of this lane's row): five exchange rounds, each
// halving the columns a lane still owns; lane L ends up with the sum of column L.  31 shuffles instead of 32 x 5.
__device__ __forceinline__ float col_sums_32x32(float (&a)[32], int lane) {
#pragma unroll
  for (int r = 0; r < 5; ++r) {
    const int o = 16 >> r;  // exchange distance == number of columns kept
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (j < o) {
        const float send = up ? a[j] : a[j + o];
        const float keep = up ? a[j + o] : a[j];
        a[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
      }
    }
  }
  return a[0];
}

// X3 (fp32-exact mode, BF16 = false only): every fp32 operand element a is staged as two tf32 planes, hi = rna_tf32(a) and
// lo = rna_tf32(a - hi); a*b is evaluated as lo_a*hi_b + hi_a*lo_b + hi_a*hi_b (the dropped lo*lo term and the rounding of lo are
// ~2^-22 relative), fp32 accumulation in TMEM: three MMAs per k-step over a stage that holds both planes of both operands.
template <bool BF16, bool X3, bool A_MN, bool B_MN, int BN, int OP, bool CTA2>
__global__ void __launch_bounds__((EpiCfg<BN, X3>::THREADS), 1) tc_kernel(const __grid_constant__ TcParams p) {
  static_assert(!(BF16 && X3), "the hi/lo split applies to fp32 operands");
  using E = Elem<BF16>;
  using S = StageCfg<BN, CTA2, X3>;
  constexpr int STAGES = S::STAGES;
  constexpr int NCTA = CTA2 ? 2 : 1;
  constexpr uint32_t IDESC = make_idesc(BF16, A_MN, B_MN, 128 * NCTA, BN);
  constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  constexpr uint32_t CHUNK_BYTES = E::BK * 128;  // one MN-major chunk: BK k-rows x 128 B
  constexpr uint32_t MN_SBO = BF16 ? 1024u : 512u;
  constexpr uint32_t MN_LAYOUT = BF16 ? 2u : 1u;
  // X3: the TMEM accumulator rounds toward zero on every MMA (measured, tools/acc_rounding_probe.py: the relative error of
  // an all-positive dot product grows linearly, -3e-5 at K = 4096), which would break the 1e-5 contract of the exact mode
  // for long reductions.  The reduction is therefore cut into chunks of ACC_CHUNK k-iterations (256 elements): the MMA warp
  // hands each chunk's accumulator to the epilogue warps, which add it to a register accumulator with round-to-nearest
  // while the next chunk runs in the other TMEM stage.
  constexpr int ACC_CHUNK = X3 ? 8 : 0x3fffffff;
  static_assert(!X3 || BN <= 128, "X3 keeps BN fp32 accumulators per epilogue thread in registers");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * S::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  auto a_smem = [&](int s) { return smem_base + s * S::STAGE_BYTES; };
  auto b_smem = [&](int s) { return smem_base + s * S::STAGE_BYTES + S::A_BYTES; };

  // warp index / CTA rank through a shuffle: the compiler then knows they are warp-uniform and keeps the role loops on
  // the uniform datapath (UTMALDG / UTCHMMA operands live in uniform registers; a per-lane branch costs an
  // ELECT + R2UR waterfall per instruction, which made the single-thread MMA loop the bottleneck: ~750 cycles / k-iteration)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int cta_rank = CTA2 ? __shfl_sync(0xffffffffu, (int)cluster_ctarank(), 0) : 0;  // 0 = leader (issues the MMAs)
  const int group = blockIdx.x / NCTA, n_groups = gridDim.x / NCTA;

  if (warp == 0 && elect_one_sync()) {
    prefetch_tmap(&p.tmA);
    prefetch_tmap(&p.tmB);
    if (X3) { prefetch_tmap(&p.tmA2); prefetch_tmap(&p.tmB2); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), NCTA);   // one arrive(+expect_tx) per CTA's producer (on the leader's barrier)
      mbar_init(empty_bar(s), 1);     // tcgen05.commit (multicast to both CTAs in 2-CTA mode)
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), EpiCfg<BN, X3>::WARPS * NCTA);  // one arrive per epilogue warp of every CTA of the group
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc<CTA2>(tmem_slot, TMEM_COLS);
    tmem_relinquish<CTA2>();
  }
  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync();  // peer's barriers initialised and TMEM allocated before any cross-CTA traffic
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int total_tiles = p.m_tiles * p.n_tiles * p.z_tiles;

  auto tile_k_iters = [&](int z) -> int {
    if (OP == OP_CONV) return p.k_iters_total;
    const int split = (OP == OP_WGRAD) ? z / p.taps : z;
    const int rem = p.k_iters_total - split * p.k_iters_per_split;
    return rem < p.k_iters_per_split ? rem : p.k_iters_per_split;
  };

  if (warp == 0) {
    // =========================== TMA producer (every CTA; whole warp loops, one elected lane issues) ===========
    int stage = 0;
    uint32_t phase = 0;
    bool ok = true;
    for (int t = group; t < total_tiles && ok; t += n_groups) {
      const int m_tile = t % p.m_tiles, r = t / p.m_tiles, n_tile = r % p.n_tiles, z = r / p.n_tiles;
      const int m0 = m_tile * (128 * NCTA) + cta_rank * 128;   // this CTA's 128 rows of A
      const int n0 = n_tile * BN + cta_rank * S::BN_CTA;       // this CTA's columns of B
      const int iters = tile_k_iters(z);
      int cw = 0, ch = 0, cn = 0, k_begin = 0;
      // per-k-iteration indices are advanced incrementally (no integer division inside the loop: the producer is one warp
      // and its instruction count per k-iteration bounds the pipeline for small tiles)
      int tp = 0, cc = 0;                          // CONV: filter tap / channel chunk of iteration i
      int wb = 0, wpy = 0, wqx = 0, wj = 0, wkk = 0;  // WGRAD: (image, row, col) of the chunk's first pixel; tap (j, kk)
      if (OP == OP_CONV) {
        const int b = m0 / p.px_per_img, rem = m0 - b * p.px_per_img, py = rem / p.Wo, qx = rem - py * p.Wo;
        cw = qx * p.trav + p.lower_w;
        ch = py * p.trav + p.lower_h;
        cn = b;
      } else if (OP == OP_WGRAD) {
        const int tap = z % p.taps;
        k_begin = (z / p.taps) * p.k_iters_per_split;
        wj = tap / p.Kw; wkk = tap - wj * p.Kw;
        const int k0 = k_begin * E::BK;
        wb = k0 / p.px_per_img;
        const int rem = k0 - wb * p.px_per_img;
        wpy = rem / p.Wo; wqx = rem - wpy * p.Wo;
      } else {
        k_begin = z * p.k_iters_per_split;
      }
      for (int i = 0; i < iters; ++i) {
        if (!__all_sync(0xffffffffu, mbar_wait(empty_bar(stage), phase ^ 1))) { atomicExch(p.status, 1); ok = false; break; }
        if (elect_one_sync()) {
        // the full barrier lives in the leader CTA; TMA of both CTAs completes on it
        const uint32_t fb = CTA2 ? mapa_cluster(full_bar(stage), 0) : full_bar(stage);
        if (CTA2 && cta_rank != 0) mbar_arrive_expect_tx_cluster(fb, S::STAGE_BYTES);
        else mbar_arrive_expect_tx(full_bar(stage), S::STAGE_BYTES);
#pragma unroll
        for (int pl = 0; pl < (X3 ? 2 : 1); ++pl) {   // plane 0: the operands (X3: hi planes), plane 1 (X3 only): lo planes
        const CUtensorMap* tmA = pl ? &p.tmA2 : &p.tmA;
        const CUtensorMap* tmB = pl ? &p.tmB2 : &p.tmB;
        const uint32_t sa = a_smem(stage) + pl * S::PLANE_BYTES, sb = b_smem(stage) + pl * S::PLANE_BYTES;
        if (OP == OP_CONV) {
          // A: 128 output pixels x KC channels of filter tap (j, kk); B: weights [Co][tap][C] rows n0.., K-major
          tma_load_im2col_4d<CTA2>(tmA, fb, sa, cc * E::KC, cw, ch, cn, p.tap_w[tp], p.tap_h[tp]);
          tma_load_2d<CTA2>(tmB, fb, sb, tp * p.wk_cols + cc * E::KC, n0);
        } else if (OP == OP_WGRAD) {
          // reduction over output pixels: chunk of BK pixels starting at flattened pixel k0
          const int k0 = (k_begin + i) * E::BK;
#pragma unroll
          for (int c = 0; c < 128 / E::KC; ++c)
            tma_load_im2col_4d<CTA2>(tmA, fb, sa + c * CHUNK_BYTES, m0 + c * E::KC, wqx * p.conv_stride - p.pad,
                                     wpy * p.conv_stride - p.pad, wb, (uint16_t)(wkk * p.dil), (uint16_t)(wj * p.dil));
          if (p.b_im2col) {
#pragma unroll
            for (int c = 0; c < S::BN_CTA / E::KC; ++c)
              tma_load_im2col_4d<CTA2>(tmB, fb, sb + c * CHUNK_BYTES, n0 + c * E::KC, wqx + p.b_pad, wpy + p.b_pad, wb, (uint16_t)0, (uint16_t)0);
          } else {
#pragma unroll
            for (int c = 0; c < S::BN_CTA / E::KC; ++c)
              tma_load_2d<CTA2>(tmB, fb, sb + c * CHUNK_BYTES, n0 + c * E::KC, k0);
          }
        } else {
          const int kidx = k_begin + i;
          if (A_MN) {
#pragma unroll
            for (int c = 0; c < 128 / E::KC; ++c)
              tma_load_2d<CTA2>(tmA, fb, sa + c * CHUNK_BYTES, m0 + c * E::KC, kidx * E::BK);
          } else {
            tma_load_2d<CTA2>(tmA, fb, sa, kidx * E::KC, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int c = 0; c < S::BN_CTA / E::KC; ++c)
              tma_load_2d<CTA2>(tmB, fb, sb + c * CHUNK_BYTES, n0 + c * E::KC, kidx * E::BK);
          } else {
            tma_load_2d<CTA2>(tmB, fb, sb, kidx * E::KC, n0);
          }
        }
        }  // plane
        }  // elected lane
        __syncwarp();
        if (OP == OP_CONV) {
          if (++cc == p.cchunks) { cc = 0; ++tp; }
        } else if (OP == OP_WGRAD) {  // advance the chunk's first pixel by BK
          wqx += E::BK;
          while (wqx >= p.Wo) { wqx -= p.Wo; ++wpy; }
          while (wpy >= p.Ho) { wpy -= p.Ho; ++wb; }
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && cta_rank == 0) {
    // =========================== MMA issuer (leader CTA only; whole warp loops, one elected lane issues) =======
    // tcgen05.commit tracks the MMAs of the issuing THREAD: elect.sync picks the same lane for the same member mask
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    bool ok = true;
    const uint32_t a_lo_stage0 = smem_desc_lo(a_smem(0), A_MN ? CHUNK_BYTES : 16), b_lo_stage0 = smem_desc_lo(b_smem(0), B_MN ? CHUNK_BYTES : 16);
    for (int t = group; t < total_tiles && ok; t += n_groups) {
      const int z = (t / p.m_tiles) / p.n_tiles;
      const int iters = tile_k_iters(z);
      int i = 0;
      do {  // one pass per accumulator hand-off: the whole reduction, or ACC_CHUNK k-iterations of it in X3 mode
      if (!__all_sync(0xffffffffu, mbar_wait(tempty_bar(acc), acc_phase ^ 1))) { atomicExch(p.status, 2); ok = false; break; }
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
      const int i0 = i, i_end = (iters - i0 > ACC_CHUNK) ? i0 + ACC_CHUNK : iters;
      for (; i < i_end; ++i) {
        if (!__all_sync(0xffffffffu, mbar_wait(full_bar(stage), phase))) { atomicExch(p.status, 3); ok = false; break; }
        tc_fence_after();
        if (elect_one_sync()) {
        // descriptor low words of this stage (start address + LBO); per k-step only a constant is added (see tc_ptx.cuh)
        //   K-major : 32 B further inside the 128-B swizzle row;  MN-major: UMMA_K k-rows (x 128 B) further
        //   MN-major tf32 uses the 32-byte-atom swizzle: k-groups of 4 rows (512 B) instead of 8 rows (1024 B)
        constexpr uint32_t A_HI = A_MN ? smem_desc_hi(MN_SBO, MN_LAYOUT) : smem_desc_hi(1024), A_STEP = (A_MN ? E::UMMA_K * 128 : 32) >> 4;
        constexpr uint32_t B_HI = B_MN ? smem_desc_hi(MN_SBO, MN_LAYOUT) : smem_desc_hi(1024), B_STEP = (B_MN ? E::UMMA_K * 128 : 32) >> 4;
        const uint32_t a_lo = a_lo_stage0 + (uint32_t)stage * (S::STAGE_BYTES >> 4), b_lo = b_lo_stage0 + (uint32_t)stage * (S::STAGE_BYTES >> 4);
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const uint64_t da = smem_desc_pack(a_lo + s * A_STEP, A_HI), db = smem_desc_pack(b_lo + s * B_STEP, B_HI);
          const uint32_t accumulate = (uint32_t)(((i - i0) | s) != 0);
          if (X3) {  // small terms first: lo*hi, hi*lo, then hi*hi
            const uint64_t da2 = smem_desc_pack(a_lo + (S::PLANE_BYTES >> 4) + s * A_STEP, A_HI);
            const uint64_t db2 = smem_desc_pack(b_lo + (S::PLANE_BYTES >> 4) + s * B_STEP, B_HI);
            umma<BF16, CTA2>(d_tmem, da2, db, IDESC, accumulate);
            umma<BF16, CTA2>(d_tmem, da, db2, IDESC, 1u);
            umma<BF16, CTA2>(d_tmem, da, db, IDESC, 1u);
          } else {
            umma<BF16, CTA2>(d_tmem, da, db, IDESC, accumulate);
          }
        }
        umma_commit<CTA2>(empty_bar(stage));  // frees the smem stage (in both CTAs) once these MMAs have read it
        }  // elected lane
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (!ok) break;
      if (elect_one_sync()) umma_commit<CTA2>(tfull_bar(acc));  // accumulator complete -> epilogue warps of both CTAs
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      } while (i < iters);
    }
  } else if (warp >= 4) {
    // =========================== epilogue (every CTA: its own 128 TMEM lanes) ===========================
    const int ew = warp & 3;  // TMEM lane quarter this warp may access
    constexpr int EPI_SPLIT = EpiCfg<BN, X3>::WARPS / 4;  // warps sharing a lane quarter: each takes 1 / EPI_SPLIT of the columns
    const int half = (warp - 4) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    // batch statistics for a consuming BatchNorm (only the K-major x K-major instantiations that produce activations)
    constexpr bool STATS_OK = (OP == OP_CONV) || (OP == OP_GEMM && !A_MN && !B_MN);
    __shared__ float stat_acc[STATS_OK ? 4 : 1][STATS_OK ? BN : 1][2];
    const bool do_stats = STATS_OK && p.stats != nullptr;
    int stat_n_tile = -1;
    auto stats_flush = [&]() {  // this warp's column sums of the finished N-tile -> its slot of the partial buffer
      if (STATS_OK && stat_n_tile >= 0) {
        float* dstp = p.stats + ((long long)(blockIdx.x * 4 + ew) * p.N) * 2;
        for (int c = half * (BN / EPI_SPLIT) + lane; c < (half + 1) * (BN / EPI_SPLIT); c += 32) {
          const int col = stat_n_tile * BN + c;
          if (col < p.N) {
            dstp[2 * col] = stat_acc[STATS_OK ? ew : 0][STATS_OK ? c : 0][0];
            dstp[2 * col + 1] = stat_acc[STATS_OK ? ew : 0][STATS_OK ? c : 0][1];
          }
        }
      }
    };
    for (int t = group; t < total_tiles; t += n_groups) {
      const int m_tile = t % p.m_tiles, r = t / p.m_tiles, n_tile = r % p.n_tiles, z = r / p.n_tiles;
      const int m = m_tile * (128 * NCTA) + cta_rank * 128 + ew * 32 + lane, n0 = n_tile * BN;
      if (do_stats && n_tile != stat_n_tile) {
        stats_flush();
        stat_n_tile = n_tile;
        for (int c = half * (BN / EPI_SPLIT) + lane; c < (half + 1) * (BN / EPI_SPLIT); c += 32)
          stat_acc[STATS_OK ? ew : 0][STATS_OK ? c : 0][0] = stat_acc[STATS_OK ? ew : 0][STATS_OK ? c : 0][1] = 0.f;
        __syncwarp();
      }
      const int iters = tile_k_iters(z);
      long long z_off = 0;
      if (OP == OP_WGRAD) z_off = (long long)(z / p.taps) * p.split_stride + (long long)(z % p.taps) * p.tap_stride;
      else if (OP == OP_GEMM) z_off = (long long)z * p.split_stride;
      long long lane_off = m;
      if (p.lane_is_pixel) {
        const int b = m / p.px_per_img, rem = m - b * p.px_per_img, r = rem / p.Wo, c = rem - r * p.Wo;
        lane_off = (long long)b * p.img_stride + (long long)(p.out_r0 + p.out_s * r) * p.out_W + p.out_c0 + p.out_s * c;
      }
      const bool m_ok = m < p.M;
      const float lane_bias = (p.bias_mode == BIAS_LANE && m_ok) ? __ldg(p.bias + m) : 0.f;
      bool ok = X3 ? true : mbar_wait(tfull_bar(acc), acc_phase);
      if (!ok) { atomicExch(p.status, 4); break; }
      tc_fence_after();
      // Column n of this lane lives col_stride floats after column n-1: a running pointer, one 128-byte store per warp
      // and column.  Chunks of 32 columns are double-buffered in registers: tcgen05.ld of chunk c+1 is in flight while
      // chunk c is stored, and the accumulator stage is handed back to the MMA warp as soon as its last chunk is in
      // registers (before that chunk's stores).
      float* dst = p.out + z_off + lane_off + (long long)n0 * p.col_stride;
      const long long cs = p.col_stride;
      const uint32_t tbase = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BN);
      constexpr int NCH = BN / 32;
      constexpr int CPW = NCH / EPI_SPLIT;  // 32-column chunks per warp
      const int c_lo = half * CPW;
      const bool col_bias = p.bias_mode == BIAS_COL;
      uint32_t va[32], vb[32];
      if (!X3 && iters > 0) tmem_ld_32x32(tbase + (uint32_t)(c_lo * 32), va);
      auto release_acc = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {  // the MMA issuer (leader CTA) waits for all epilogue warps of the group
          if (CTA2 && cta_rank != 0) mbar_arrive_cluster(mapa_cluster(tempty_bar(acc), 0));
          else mbar_arrive(tempty_bar(acc));
        }
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
          stat_acc[STATS_OK ? ew : 0][STATS_OK ? c * 32 + lane : 0][0] += s1;   // lane L owns column c*32 + L of this warp's slot
          stat_acc[STATS_OK ? ew : 0][STATS_OK ? c * 32 + lane : 0][1] += s2;
        }
        if (p.out_bf16) {  // same addressing in 2-byte elements: 64-byte stores per warp and column
          __nv_bfloat16* qb = reinterpret_cast<__nv_bfloat16*>(p.out) + z_off + lane_off + (long long)cbase * cs;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (m_ok && cbase + j < p.N) *qb = __float2bfloat16_rn(__uint_as_float(v[j]));
            qb += cs;
          }
          return;
        }
        float* q = dst + (long long)(c * 32) * cs;
        if (p.relu == 1 || p.relu == 2) {  // Linear forward + ReLU: one ballot per column gives the mask word of this warp's 32 features
          __nv_bfloat16* lq = p.relu_lp ? reinterpret_cast<__nv_bfloat16*>(p.relu_lp) + z_off + lane_off + (long long)cbase * cs : nullptr;
          const int mw = m >> 5;  // m - lane is a multiple of 32: word index of this warp's features within a row
          if (p.relu == 2) {  // dgrad of the layer behind a ReLU: dx = acc * mask (activation_funcs.py:32-34)
            // lane L fetches the mask word of column cbase + L once per chunk (one load in flight per lane, issued before any
            // store); column j's word then comes by shuffle — a load per column would serialise behind the stores
            unsigned int myw = 0u;
            if (cbase + lane < p.N && m_ok) myw = __ldg(p.relu_mask + ((long long)(cbase + lane) * p.M >> 5) + mw);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const bool col_ok = cbase + j < p.N;
              const unsigned int bits = __shfl_sync(0xffffffffu, myw, j);
              const float val = __uint_as_float(v[j]) * (float)((bits >> lane) & 1u);
              if (col_ok && m_ok) {
                *q = val;
                if (lq) *lq = __float2bfloat16_rn(val);
              }
              q += cs;
              if (lq) lq += cs;
            }
            return;
          }
          unsigned int myw = 0u;  // lane L collects the mask word of column cbase + L: one store instruction per chunk
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float a = __uint_as_float(v[j]) + lane_bias;
            const float val = (a != a) ? a : fmaxf(a, 0.f);  // numpy.maximum propagates NaN (activation_funcs.py:27)
            const bool col_ok = cbase + j < p.N;
            const unsigned int bits = __ballot_sync(0xffffffffu, m_ok && val > 0.f);
            if (lane == j) myw = bits;
            if (col_ok && m_ok) {
              *q = val;
              if (lq) *lq = __float2bfloat16_rn(val);
            }
            q += cs;
            if (lq) lq += cs;
          }
          if (p.relu_mask && m_ok && cbase + lane < p.N) p.relu_mask[((long long)(cbase + lane) * p.M >> 5) + mw] = myw;
          return;
        }
        float bl = 0.f;  // this lane's column bias; column j's value is fetched with a shuffle (one LDG per chunk)
        if (col_bias && cbase + lane < p.N) bl = __ldg(p.bias + cbase + lane);
        // ReLU behind a convolution (relu == 3): one NaN-propagating max per value against 0, against -inf (identity) otherwise
        const float floor_v = p.relu == 3 ? 0.f : -INFINITY;
        if (cbase + 32 <= p.N) {
          if (col_bias) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float val = max_nan(__uint_as_float(v[j]) + __shfl_sync(0xffffffffu, bl, j), floor_v);
              if (m_ok) *q = val;
              q += cs;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (m_ok) *q = max_nan(__uint_as_float(v[j]) + lane_bias, floor_v);
              q += cs;
            }
          }
        } else {  // ragged last chunk of the N dimension
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float val = max_nan(__uint_as_float(v[j]) + lane_bias + __shfl_sync(0xffffffffu, bl, j), floor_v);
            if (m_ok && cbase + j < p.N) *q = val;
            q += cs;
          }
        }
      };
      if constexpr (X3) {
        // chunked reduction: add each hand-off's accumulator (TMEM, truncating) into registers (round to nearest)
        constexpr int XN = X3 ? NCH : 1;
        float r[XN][32];
#pragma unroll
        for (int c = 0; c < XN; ++c)
#pragma unroll
          for (int j = 0; j < 32; ++j) r[c][j] = 0.f;
        const int handoffs = iters > 0 ? (iters + ACC_CHUNK - 1) / ACC_CHUNK : 1;
        bool alive = true;
        for (int hnd = 0; hnd < handoffs; ++hnd) {
          if (!mbar_wait(tfull_bar(acc), acc_phase)) { atomicExch(p.status, 4); alive = false; break; }
          tc_fence_after();
          const uint32_t tb = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BN);
          if (iters > 0) {
            tmem_ld_32x32(tb, va);
#pragma unroll
            for (int c = 0; c < XN; c += 2) {
              tmem_ld_wait();
              tmem_ld_32x32(tb + (uint32_t)((c + 1) * 32), vb);
#pragma unroll
              for (int j = 0; j < 32; ++j) r[c][j] += __uint_as_float(va[j]);
              tmem_ld_wait();
              if (c + 2 < XN) tmem_ld_32x32(tb + (uint32_t)((c + 2) * 32), va);
#pragma unroll
              for (int j = 0; j < 32; ++j) r[c + 1][j] += __uint_as_float(vb[j]);
            }
          }
          release_acc();
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (!alive) break;
#pragma unroll
        for (int c = 0; c < XN; ++c) {
#pragma unroll
          for (int j = 0; j < 32; ++j) va[j] = __float_as_uint(r[c][j]);
          store_chunk(va, c);
        }
        continue;
      }
      if (iters == 0) {
#pragma unroll
        for (int j = 0; j < 32; ++j) va[j] = vb[j] = 0u;
      }
      // this warp's chunks c_lo .. c_lo + CPW - 1, the next one in flight (tcgen05.ld) while the current one is stored; the
      // accumulator stage goes back to the MMA warp as soon as the last chunk is in registers
#pragma unroll 1
      for (int k = 0; k < CPW; k += 2) {
        if (iters > 0) {
          tmem_ld_wait();                                   // chunk k in va
          if (k + 1 < CPW) tmem_ld_32x32(tbase + (uint32_t)((c_lo + k + 1) * 32), vb);
        }
        if (k + 1 >= CPW) release_acc();
        store_chunk(va, c_lo + k);
        if (k + 1 < CPW) {
          if (iters > 0) {
            tmem_ld_wait();                                 // chunk k+1 in vb
            if (k + 2 < CPW) tmem_ld_32x32(tbase + (uint32_t)((c_lo + k + 2) * 32), va);
          }
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
  if (CTA2) cluster_sync();  // the peer may still be reading this CTA's smem / signalling its barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<CTA2>(tmem_base, TMEM_COLS);
  }
}

}  // namespace tc
}  // namespace cpt
