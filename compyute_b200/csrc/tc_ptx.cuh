// Inline-PTX wrappers for the Blackwell (sm_100a) tensor-core path: mbarrier, TMA (tiled + im2col),
// tcgen05 (alloc / mma / commit / ld) and UMMA descriptor construction.  Hand-written; no CUTLASS.
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace cpt {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one lane of the (fully converged) warp; deterministic for a given member mask
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// ---- cluster (CTA pair) helpers
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same smem offset in CTA `cta` of the cluster
__device__ __forceinline__ uint32_t mapa_cluster(uint32_t local_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void mbar_arrive_expect_tx_cluster(uint32_t cluster_bar, uint32_t bytes) {
  // default .release.cta semantics on purpose: a cluster-scope release makes ptxas emit a MEMBAR + CCTL.IVALL per
  // stage on the producer thread, which halves the pipeline's throughput (measured); TMA data visibility is carried by
  // complete_tx, not by this arrive.
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: a wrong descriptor / byte count must surface as an error code, never as a hung GPU.
constexpr uint64_t kWaitTimeoutNs = 4000000000ull;  // 4 s
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return true;
  const uint64_t t0 = globaltimer_ns();
  while (true) {
#pragma unroll 1
    for (int i = 0; i < 64; ++i)
      if (mbar_try_wait(bar, parity)) return true;
    if (globaltimer_ns() - t0 > kWaitTimeoutNs) return false;
  }
}

// ------------------------------------------------------------------ TMA
__device__ __forceinline__ void prefetch_tmap(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
// CTA2: the load lands in this CTA's smem but completes on the LEADER CTA's mbarrier (`bar` is then a shared::cluster
// address obtained with mapa), which requires the .cta_group::2 form.
template <bool CTA2>
__device__ __forceinline__ void tma_load_2d(const void* tmap, uint32_t bar, uint32_t dst, int32_t c0, int32_t c1) {
  if (CTA2) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
  } else {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
  }
}
// im2col mode, 4-D (C, W, H, N): coordinates of the first pixel + filter-tap offsets (w, h)
template <bool CTA2>
__device__ __forceinline__ void tma_load_im2col_4d(const void* tmap, uint32_t bar, uint32_t dst, int32_t c, int32_t w,
                                                   int32_t h, int32_t n, uint16_t off_w, uint16_t off_h) {
  if (CTA2) {
    asm volatile(
        "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
        : "memory");
  } else {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
        : "memory");
  }
}

// ------------------------------------------------------------------ tcgen05
// CTA2: executed by the same warp of BOTH CTAs of the pair with the same smem offset (allocates in both TMEMs)
template <bool CTA2>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  if (CTA2) asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  else asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
template <bool CTA2>
__device__ __forceinline__ void tmem_relinquish() {
  if (CTA2) asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  else asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <bool CTA2>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if (CTA2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem];  kind::f16 covers bf16/fp16 inputs, kind::tf32 reads fp32 bits as tf32
// cta_group::2: issued by the leader CTA only; A rows 0-127 / B columns 0-BN/2 come from the leader's smem, rows 128-255 /
// the other B half from the peer's smem at the same offsets; D goes to the same TMEM address in both CTAs.
#define CPT_UMMA_ASM(GROUP, KIND)                                                                        \
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"                                       \
               "tcgen05.mma.cta_group::" GROUP ".kind::" KIND " [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), \
               "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)                                    \
               : "memory")
template <bool BF16, bool CTA2>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  if (BF16) {
    if (CTA2) CPT_UMMA_ASM("2", "f16");
    else CPT_UMMA_ASM("1", "f16");
  } else {
    if (CTA2) CPT_UMMA_ASM("2", "tf32");
    else CPT_UMMA_ASM("1", "tf32");
  }
}
// mbarrier arrives once every previously issued MMA of this thread has completed (implies fence::before_thread_sync)
// CTA2: the arrive is multicast to the barrier at the same offset in both CTAs of the pair
template <bool CTA2>
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  if (CTA2) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
  } else {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  }
}
// 32 lanes x 32 consecutive fp32 columns: thread i <- lane (base+i), v[j] <- column (base+j)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// max that propagates NaN like numpy.maximum (fmaxf returns the non-NaN operand); max_nan(x, -inf) == x for every x
__device__ __forceinline__ float max_nan(float a, float b) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------ descriptors
// Shared-memory matrix descriptor, SWIZZLE_128B (PTX ISA "tcgen05 matrix descriptor"):
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout type
//   layout 2 = SWIZZLE_128B (16-byte atoms, 8-row period); layout 1 = SWIZZLE_128B_BASE32B (32-byte atoms, 4-row
//   period) — the only swizzled layout the hardware accepts for MN-major 32-bit (tf32) operands.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type = 2) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | ((uint64_t)layout_type << 61);
}
// The same descriptor split into its two 32-bit halves.  Inside an MMA issue loop only the START ADDRESS field changes
// (bits [0,14) of the low word: stage base, + 32 B per k-step, + a tap's row offset ...), so a loop builds `lo` once per
// stage and adds small constants to it; `hi` (SBO, version, layout) is a compile-time constant.  One ADD per descriptor
// instead of the shift / mask / or chain of make_smem_desc — the single issuing warp's instruction count per k-iteration is
// what bounds small-N tiles.  The start field cannot carry: shared-memory addresses are < 2^18.
__device__ __forceinline__ constexpr uint32_t smem_desc_hi(uint32_t sbo_bytes, uint32_t layout_type = 2) {
  return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (layout_type << 29);
}
__device__ __forceinline__ uint32_t smem_desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__device__ __forceinline__ uint64_t smem_desc_pack(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}

// Instruction descriptor (kind::f16 / kind::tf32, fp32 accumulate):
//   [4,6) D fmt (1 = f32) | [7,10) A fmt | [10,13) B fmt (1 = bf16, 2 = tf32) | 15 A MN-major | 16 B MN-major |
//   [17,23) N>>3 | [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc(bool bf16, bool a_mn, bool b_mn, int M, int N) {
  return (1u << 4) | ((bf16 ? 1u : 2u) << 7) | ((bf16 ? 1u : 2u) << 10) | ((a_mn ? 1u : 0u) << 15) |
         ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

}  // namespace tc
}  // namespace cpt
