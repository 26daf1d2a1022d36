// Hardware probe for the strip ("shared halo") convolution kernels: can a tcgen05 shared-memory descriptor start at an
// arbitrary 128-byte ROW of a 128B-swizzled tile that TMA wrote at a 1024-byte aligned base?  (The K^2 filter taps of a
// 3x3 convolution are then K^2 descriptors into ONE strip of input pixels instead of K^2 im2col loads.)
//   test 1: K-major A (rows = pixels, 128 B = 64 bf16 channels), start = base + r*128, for two settings of the
//           descriptor's base_offset field (0, and (start >> 7) & 7 as the PTX ISA describes for unaligned patterns)
//   test 2: MN-major A (rows = k = pixels, 128 B = 64 channels = M), M = 128 made of TWO 64-channel chunks whose
//           "leading byte offset" is the distance between two filter taps in the same strip (wgrad: two taps per MMA)
// Exact integer data; prints one line per configuration.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o build/halo_probe tools/halo_probe.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../compyute_b200/csrc/tc_ptx.cuh"

using namespace cpt::tc;

constexpr int RS = 256;   // strip rows
constexpr int NB = 64;    // N
constexpr int NCFG_MAX = 64;

struct Cfg { int test, r, r2, variant; };
struct Params {
  CUtensorMap tmS, tmB;
  float* out;  // [ncfg][128][64]
  long long* cycles;  // [ncfg]: test 3 = cycles of TIMING_REPS x 4 MMAs (same descriptors as test 1 / 2, variant 0)
  int ncfg;
  Cfg cfg[NCFG_MAX];
};
constexpr int TIMING_REPS = 512;

__device__ __forceinline__ uint64_t desc_with_base(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t base_off) {
  return make_smem_desc(saddr, lbo, sbo, 2) | ((uint64_t)(base_off & 7u) << 49);
}

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sS = base, sB = base + RS * 128, bar = sB + NB * 128, bar2 = bar + 8, slot = bar + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc<false>(slot, 64); tmem_relinquish<false>(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(slot));
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, RS * 128 + NB * 128);
    tma_load_2d<false>(&p.tmS, bar, sS, 0, 0);
    tma_load_2d<false>(&p.tmB, bar, sB, 0, 0);
  }
  mbar_wait(bar, 0);
  uint32_t phase = 0;
  for (int c = 0; c < p.ncfg; ++c) {
    const Cfg cf = p.cfg[c];
    if (threadIdx.x == 0) {
      tc_fence_after();
      for (int s = 0; s < 4; ++s) {
        uint64_t da, db;
        uint32_t idesc;
        if (cf.test == 1) {
          const uint32_t sa = sS + cf.r * 128 + s * 32;
          da = desc_with_base(sa, 16, 1024, cf.variant ? ((sa >> 7) & 7u) : 0u);
          idesc = make_idesc(true, false, false, 128, NB);
        } else {
          const uint32_t sa = sS + cf.r * 128 + s * (16 * 128);
          da = desc_with_base(sa, (uint32_t)(cf.r2 - cf.r) * 128, 1024, cf.variant ? ((sa >> 7) & 7u) : 0u);
          idesc = make_idesc(true, true, false, 128, NB);
        }
        db = make_smem_desc(sB + s * 32, 16, 1024);
        umma<true, false>(tmem_base, da, db, idesc, (uint32_t)(s != 0));
      }
      umma_commit<false>(bar2);
    }
    mbar_wait(bar2, phase);
    phase ^= 1;
    tc_fence_after();
    uint32_t v[32];
    for (int ch = 0; ch < 2; ++ch) {
      tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + ch * 32, v);
      tmem_ld_wait();
      float* o = p.out + ((size_t)c * 128 + warp * 32 + lane) * NB + ch * 32;
      for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
  }
  // ---- test 3: MMA issue rate as a function of the descriptor start row (is an unaligned start a slow path?)
  for (int c = 0; c < p.ncfg; ++c) {
    const Cfg cf = p.cfg[c];
    if (cf.variant != 0) continue;
    if (threadIdx.x == 0) {
      tc_fence_after();
      const long long t0 = clock64();
      for (int rep = 0; rep < TIMING_REPS; ++rep) {
        for (int s = 0; s < 4; ++s) {
          uint64_t da;
          uint32_t idesc;
          if (cf.test == 1) {
            da = make_smem_desc(sS + cf.r * 128 + s * 32, 16, 1024);
            idesc = make_idesc(true, false, false, 128, NB);
          } else {
            da = make_smem_desc(sS + cf.r * 128 + s * (16 * 128), (uint32_t)(cf.r2 - cf.r) * 128, 1024, 2);
            idesc = make_idesc(true, true, false, 128, NB);
          }
          const uint64_t db = make_smem_desc(sB + s * 32, 16, 1024);
          umma<true, false>(tmem_base, da, db, idesc, 1u);
        }
      }
      umma_commit<false>(bar2);
      mbar_wait(bar2, phase);
      p.cycles[c] = clock64() - t0;
    }
    phase ^= 1;
    __syncthreads();
  }
  if (warp == 0) { tc_fence_after(); tmem_dealloc<false>(tmem_base, 64); }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { printf("no driver entry\n"); return 1; }
  auto encode = reinterpret_cast<EncodeTiledFn>(fn);
  std::vector<__nv_bfloat16> hS(RS * 64), hB(NB * 64);
  std::vector<float> fS(RS * 64), fB(NB * 64);
  srand(1);
  for (int i = 0; i < RS * 64; ++i) { fS[i] = (float)(rand() % 5 - 2); hS[i] = __float2bfloat16(fS[i]); }
  for (int i = 0; i < NB * 64; ++i) { fB[i] = (float)(rand() % 3 - 1); hB[i] = __float2bfloat16(fB[i]); }
  __nv_bfloat16 *dS, *dB;
  float* dOut;
  cudaMalloc(&dS, hS.size() * 2); cudaMalloc(&dB, hB.size() * 2);
  cudaMemcpy(dS, hS.data(), hS.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  Params p{};
  auto mk = [&](CUtensorMap* m, void* ptr, int rows) {
    cuuint64_t dims[2] = {64, (cuuint64_t)rows};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, (cuuint32_t)rows};
    cuuint32_t es[2] = {1, 1};
    return encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  if (mk(&p.tmS, dS, RS) != CUDA_SUCCESS || mk(&p.tmB, dB, NB) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
  const int rows1[] = {0, 8, 1, 2, 3, 4, 7, 9, 58, 59, 117, 118};
  int n = 0;
  for (int v = 0; v < 2; ++v)
    for (int r : rows1) p.cfg[n++] = Cfg{1, r, 0, v};
  const int pairs2[][2] = {{0, 8}, {0, 1}, {1, 2}, {3, 59}, {58, 117}, {5, 64}, {0, 64}};
  for (int v = 0; v < 2; ++v)
    for (auto& pr : pairs2) p.cfg[n++] = Cfg{2, pr[0], pr[1], v};
  p.ncfg = n;
  cudaMalloc(&dOut, (size_t)n * 128 * NB * 4);
  cudaMemset(dOut, 0xff, (size_t)n * 128 * NB * 4);
  p.out = dOut;
  long long* dCyc;
  cudaMalloc(&dCyc, NCFG_MAX * sizeof(long long));
  cudaMemset(dCyc, 0, NCFG_MAX * sizeof(long long));
  p.cycles = dCyc;
  const int smem = RS * 128 + NB * 128 + 1024 + 64;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_kernel<<<1, 128, smem>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> out((size_t)n * 128 * NB);
  cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost);
  for (int c = 0; c < n; ++c) {
    const Cfg cf = p.cfg[c];
    int bad = 0, first = -1;
    for (int m = 0; m < 128; ++m)
      for (int nn = 0; nn < NB; ++nn) {
        float ref = 0.f;
        for (int k = 0; k < 64; ++k) {
          float a;
          if (cf.test == 1) a = fS[(cf.r + m) * 64 + k];                       // A[m][k] = strip row r+m, channel k
          else a = m < 64 ? fS[(cf.r + k) * 64 + m] : fS[(cf.r2 + k) * 64 + (m - 64)];  // A[m][k] = strip row (tap)+k, channel m
          ref += a * fB[nn * 64 + k];
        }
        if (out[((size_t)c * 128 + m) * NB + nn] != ref) { if (first < 0) first = m * NB + nn; ++bad; }
      }
    printf("test %d r=%3d r2=%3d base_offset=%s : %s (%d bad, first m=%d n=%d)\n", cf.test, cf.r, cf.r2, cf.variant ? "(addr>>7)&7" : "0",
           bad ? "MISMATCH" : "exact", bad, first < 0 ? -1 : first / NB, first < 0 ? -1 : first % NB);
  }
  std::vector<long long> cyc(NCFG_MAX);
  cudaMemcpy(cyc.data(), dCyc, NCFG_MAX * sizeof(long long), cudaMemcpyDeviceToHost);
  for (int c = 0; c < n; ++c)
    if (p.cfg[c].variant == 0)
      printf("timing test %d r=%3d r2=%3d : %.1f cycles per MMA (M=128 N=64 K=16 bf16; floor 32)\n", p.cfg[c].test, p.cfg[c].r, p.cfg[c].r2,
             (double)cyc[c] / (TIMING_REPS * 4));
  return 0;
}
