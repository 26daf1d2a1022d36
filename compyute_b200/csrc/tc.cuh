// Tensor-core (tcgen05 / TMEM / TMA) back end: declarations shared with the C-ABI dispatchers.
#pragma once
#include "common.cuh"

namespace cpt {
namespace tc {

size_t conv_workspace_size(int op, const cpt_conv2d_desc* d, int mode);
int conv_fprop(const cpt_conv2d_desc* d, const float* x, const float* w, const float* bias, float* y, int mode, void* ws,
               size_t ws_bytes, cudaStream_t st);
int conv_dgrad(const cpt_conv2d_desc* d, const float* dy, const float* w, float* dx, int mode, void* ws, size_t ws_bytes,
               cudaStream_t st);
int conv_wgrad(const cpt_conv2d_desc* d, const float* x, const float* dy, float* dw, float* db, int mode, void* ws,
               size_t ws_bytes, cudaStream_t st);

size_t linear_workspace_size(int op, int64_t N, int In, int Out, int mode);
int linear_fwd(const float* x, const float* w, const float* bias, float* y, int64_t N, int In, int Out, int mode, void* ws,
               size_t ws_bytes, cudaStream_t st);
int linear_dgrad(const float* dy, const float* w, float* dx, int64_t N, int In, int Out, int mode, void* ws,
                 size_t ws_bytes, cudaStream_t st);
int linear_wgrad(const float* x, const float* dy, float* dw, float* db, int64_t N, int In, int Out, int mode, void* ws,
                 size_t ws_bytes, cudaStream_t st);

}  // namespace tc
}  // namespace cpt
