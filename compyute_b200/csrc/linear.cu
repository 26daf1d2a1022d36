// Linear forward / dgrad / wgrad(+db): C-ABI entry points and the exact-fp32 (FFMA) GEMM problem.
// Reference: compyute/nn/functional/linear_funcs.py:11-35.  Tensor-core modes go to tc_gemm.cu.
#include "simt_gemm.cuh"
#include "tc.cuh"

namespace cpt {

// C[m][n] = Σ_k A(m,k) B(k,n) (+ bias[n]);  A(m,k) = a[m*sam + k*sak], B(k,n) = b[k*sbk + n*sbn]
template <bool A_MN, bool B_MN>
struct GemmP {
  const float* a; const float* b; const float* bias; float* c;
  int M, N, K; int64_t sam, sak, sbk, sbn; int vec;
  static constexpr bool A_MN_CONTIG = A_MN, B_MN_CONTIG = B_MN, OUT_M_CONTIG = false;
  struct RowA { const float* base; };
  struct ColB { const float* base; };
  __device__ RowA rowA(int m) const { return RowA{m < M ? a + (int64_t)m * sam : nullptr}; }
  __device__ float loadA(const RowA& r, int k) const { return r.base ? __ldg(r.base + (int64_t)k * sak) : 0.f; }
  __device__ ColB colB(int n) const { return ColB{n < N ? b + (int64_t)n * sbn : nullptr}; }
  __device__ float loadB(const ColB& cb, int k) const { return cb.base ? __ldg(cb.base + (int64_t)k * sbk) : 0.f; }
  __device__ void store4(int split, int m, int n, const float v[4]) const {
    if (m >= M || n >= N) return;
    float* dst = c + ((int64_t)split * M + m) * N + n;
    if (vec && n + 3 < N) {
      float4 o = make_float4(v[0], v[1], v[2], v[3]);
      if (bias) { o.x += __ldg(bias + n); o.y += __ldg(bias + n + 1); o.z += __ldg(bias + n + 2); o.w += __ldg(bias + n + 3); }
      *reinterpret_cast<float4*>(dst) = o;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (n + i < N) dst[i] = v[i] + (bias ? __ldg(bias + n + i) : 0.f);
    }
  }
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int check_linear(const char* who, int64_t N, int In, int Out) {
  CPT_REQUIRE(N > 0 && In > 0 && Out > 0, CPT_ERR_INVALID, "%s: non-positive dimension", who);
  CPT_REQUIRE(N < (1LL << 31), CPT_ERR_UNSUPPORTED, "%s: N exceeds int32", who);
  return CPT_OK;
}

}  // namespace cpt

using namespace cpt;

extern "C" {

size_t cpt_linear_workspace_size(int op, int64_t N, int In, int Out, int mode) {
  if (N <= 0 || In <= 0 || Out <= 0) return 0;
  if (mode != CPT_MODE_FP32) return tc::linear_workspace_size(op, N, In, Out, mode);
  if (op == CPT_OP_WGRAD) {
    const int splits = sg_pick_splits(Out, In, (int)N);
    return align_up((size_t)(splits > 1 ? splits : 0) * Out * In * sizeof(float), 256) +
           align_up((size_t)Out * 64 * sizeof(float), 256) + 256;
  }
  return 256;
}

int cpt_linear_fwd(const float* x, const float* w, const float* bias, float* y, int64_t N, int In, int Out, int mode,
                   void* ws, size_t ws_bytes, void* stream) {
  if (int e = check_linear("linear_fwd", N, In, Out)) return e;
  CPT_REQUIRE(x && w && y, CPT_ERR_INVALID, "linear_fwd: null tensor");
  if (mode != CPT_MODE_FP32) return tc::linear_fwd(x, w, bias, y, N, In, Out, mode, ws, ws_bytes, as_stream(stream));
  GemmP<false, false> p{x, w, bias, y, (int)N, Out, In, In, 1, 1, In, (Out % 4 == 0) && aligned16(y)};
  CPT_CUDA(sg_launch(p, 1, as_stream(stream)));
  return CPT_OK;
}

int cpt_linear_dgrad(const float* dy, const float* w, float* dx, int64_t N, int In, int Out, int mode, void* ws,
                     size_t ws_bytes, void* stream) {
  if (int e = check_linear("linear_dgrad", N, In, Out)) return e;
  CPT_REQUIRE(dy && w && dx, CPT_ERR_INVALID, "linear_dgrad: null tensor");
  if (mode != CPT_MODE_FP32) return tc::linear_dgrad(dy, w, dx, N, In, Out, mode, ws, ws_bytes, as_stream(stream));
  // dx[m][n=in] = Σ_k dy[m][k] w[k][n]
  GemmP<false, true> p{dy, w, nullptr, dx, (int)N, In, Out, Out, 1, In, 1, (In % 4 == 0) && aligned16(dx)};
  CPT_CUDA(sg_launch(p, 1, as_stream(stream)));
  return CPT_OK;
}

int cpt_linear_wgrad(const float* x, const float* dy, float* dw, float* db, int64_t N, int In, int Out, int mode,
                     void* ws, size_t ws_bytes, void* stream) {
  if (int e = check_linear("linear_wgrad", N, In, Out)) return e;
  CPT_REQUIRE(x && dy && dw, CPT_ERR_INVALID, "linear_wgrad: null tensor");
  if (mode != CPT_MODE_FP32) return tc::linear_wgrad(x, dy, dw, db, N, In, Out, mode, ws, ws_bytes, as_stream(stream));
  CPT_REQUIRE(ws && ws_bytes >= cpt_linear_workspace_size(CPT_OP_WGRAD, N, In, Out, mode), CPT_ERR_WORKSPACE,
              "linear_wgrad: workspace too small");
  cudaStream_t st = as_stream(stream);
  // dw[m=out][n=in] = Σ_k dy[k][m] x[k][n]
  const int splits = sg_pick_splits(Out, In, (int)N);
  const size_t part_bytes = align_up((size_t)(splits > 1 ? splits : 0) * Out * In * sizeof(float), 256);
  float* out = splits > 1 ? reinterpret_cast<float*>(ws) : dw;
  GemmP<true, true> p{dy, x, nullptr, out, Out, In, (int)N, 1, Out, In, 1, (In % 4 == 0) && aligned16(out)};
  CPT_CUDA(sg_launch(p, splits, st));
  if (splits > 1) {
    launch_reduce_splits(reinterpret_cast<float*>(ws), dw, (int64_t)Out * In, splits, st);
    CPT_LAUNCH_CHECK("linear_wgrad reduce");
  }
  if (db) return channel_sum(dy, db, (int)N, Out, 1, reinterpret_cast<char*>(ws) + part_bytes, st);
  return CPT_OK;
}

}  // extern "C"
