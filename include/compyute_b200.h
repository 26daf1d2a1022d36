/*
 * compyute_b200 — C ABI of the B200-native CNN-training hot path (libcompyute_b200.so).
 *
 * This is the drop-in boundary.  The reference (dakofler/Compyute v0.1.8) has no
 * FFI layer: its operator interface is the Python `Function` protocol
 * (compyute/nn/functional/functions.py:37-54) whose bodies call NumPy/CuPy through
 * `Device.module` (compyute/backend.py:31,60).  Each entry point below replaces the array-library
 * calls made by one reference `XxxFn.forward/backward`; the citation on each declaration is the
 * reference code it stands in for.  `compyute_b200/_lib.py` binds them with ctypes; a maintainer
 * of the reference would add the same stub.
 *
 * Conventions
 *  - All tensor pointers are DEVICE pointers (e.g. CuPy `arr.data.ptr`, torch `t.data_ptr()`,
 *    `__cuda_array_interface__['data'][0]`), fp32, C-contiguous, NCHW / OIHW, unless stated.
 *  - `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy default stream).  Every call
 *    is asynchronous on `stream`; nothing allocates, synchronises or takes ownership.
 *  - `ws` / `ws_bytes`: caller-owned scratch, size from the matching `*_workspace_size` call.
 *  - `mode`: arithmetic of the contractions (CPT_MODE_*).  FP32 = exact fp32 FFMA (matches the
 *    reference at 1e-5); TF32 / BF16 = tcgen05 tensor cores, fp32 accumulate (looser, stated
 *    tolerances in tests/).
 *  - Return value: 0 on success, a negative CPT_ERR_* otherwise; `cpt_last_error()` returns a
 *    thread-local message.  Error mapping to the reference's exceptions is in INTEGRATION.md.
 */
#ifndef COMPYUTE_B200_H
#define COMPYUTE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CPT_OK 0
#define CPT_ERR_INVALID (-1)     /* bad shape / argument          -> ShapeError / ValueError  */
#define CPT_ERR_CUDA (-2)        /* CUDA runtime / driver failure -> CUDARuntimeError         */
#define CPT_ERR_UNSUPPORTED (-3) /* configuration not implemented -> NotImplementedError      */
#define CPT_ERR_WORKSPACE (-4)   /* ws_bytes too small            -> ValueError               */

#define CPT_MODE_FP32 0
#define CPT_MODE_TF32 1
#define CPT_MODE_BF16 2
/* fp32-exact contraction on the tensor cores ("3xTF32"): every fp32 operand a is staged as two tf32 planes, hi = rna(a) and
 * lo = rna(a - hi); a*b = lo*hi + hi*lo + hi*hi with fp32 accumulation (the dropped terms are ~2^-22 relative).  Meets the
 * reference's allclose(1e-5) like CPT_MODE_FP32 (the FFMA kernels), which stays the fallback for geometries outside the TMA
 * limits.  Channels-last / workspace sizes of this mode hold both planes. */
#define CPT_MODE_FP32X3 3

#define CPT_OP_FPROP 0
#define CPT_OP_DGRAD 1
#define CPT_OP_WGRAD 2

/* ---- library ------------------------------------------------------------------------------ */
const char* cpt_last_error(void);
int cpt_version(void);
/* compyute/backend.py:69-86 (CUDA.properties / mem_info) */
int cpt_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, size_t* smem_optin,
                    size_t* total_mem);

/* ---- Conv2D: compyute/nn/functional/convolution_funcs.py:218-410 --------------------------- */
typedef struct cpt_conv2d_desc {
  int32_t B, Ci, H, W; /* input  (B, Ci, H, W)                                        */
  int32_t Co, K;       /* filter (Co, Ci, K, K); square kernel like the reference      */
  int32_t pad, stride, dil;
} cpt_conv2d_desc;

/* Output extent: Ho = (H + 2 pad - dil (K-1) - 1) / stride + 1 (shape_ops.py:297). */
int cpt_conv2d_out_shape(const cpt_conv2d_desc* d, int* Ho, int* Wo);
size_t cpt_conv2d_workspace_size(int op, const cpt_conv2d_desc* d, int mode);

/* Conv2DFn.forward :222-241  (dilate f, pad x, window view, einsum 'biyxjk,oijk->boyx', + b).
 * bias may be NULL.  y: (B, Co, Ho, Wo). */
int cpt_conv2d_fprop(const cpt_conv2d_desc* d, const float* x, const float* w, const float* bias,
                     float* y, int mode, void* ws, size_t ws_bytes, void* stream);
/* RawConv2DFn.backward dx branch :390-403 + Pad2DFn.backward :351-355.  dx: (B, Ci, H, W). */
int cpt_conv2d_dgrad(const cpt_conv2d_desc* d, const float* dy, const float* w, float* dx,
                     int mode, void* ws, size_t ws_bytes, void* stream);
/* RawConv2DFn.backward dW branch :405-408 + Dilation2DFn.backward :312-316 + db = dy.sum((0,2,3))
 * :252.  db may be NULL.  dw: (Co, Ci, K, K), db: (Co,). */
int cpt_conv2d_wgrad(const cpt_conv2d_desc* d, const float* x, const float* dy, float* dw,
                     float* db, int mode, void* ws, size_t ws_bytes, void* stream);

/* Tensor-core modes stage activations as channels-last (NHWC, C padded to a multiple of 8) in
 * bf16 (BF16) or tf32-rounded fp32 (TF32) so that TMA im2col can feed tcgen05.  These expose the
 * staging so a caller can keep x_cl from forward for wgrad (what Conv2DFn's cache does). */
size_t cpt_channels_last_bytes(int B, int C, int H, int W, int mode);
/* chan_sum (C floats) may be NULL; if given it receives the per-channel sum of src (fuses db = dy.sum((0,2,3)) into
 * the staging pass).  With ws >= cpt_to_channels_last_workspace_size the sums are reduced deterministically (per-block
 * partials + fixed-order pass) and chan_sum is overwritten; with ws = NULL they are atomically ADDED to chan_sum, which
 * the caller must have zeroed. */
size_t cpt_to_channels_last_workspace_size(int B, int C, int H, int W);
int cpt_to_channels_last(const float* src, void* dst, int B, int C, int H, int W, int mode,
                         float* chan_sum, void* ws, size_t ws_bytes, void* stream);
int cpt_conv2d_fprop_cl(const cpt_conv2d_desc* d, const void* x_cl, const float* w,
                        const float* bias, float* y, int mode, void* ws, size_t ws_bytes,
                        void* stream);
int cpt_conv2d_dgrad_cl(const cpt_conv2d_desc* d, const void* dy_cl, const float* w, float* dx,
                        int mode, void* ws, size_t ws_bytes, void* stream);
/* 1 if cpt_conv2d_dgrad_cl can run this geometry on the tensor-core path (stride classes within the TMA im2col
 * limits), else 0: the caller then uses cpt_conv2d_dgrad with CPT_MODE_FP32 on the fp32 dy. */
int cpt_conv2d_dgrad_cl_supported(const cpt_conv2d_desc* d, int mode);
int cpt_conv2d_wgrad_cl(const cpt_conv2d_desc* d, const void* x_cl, const void* dy_cl, float* dw,
                        int mode, void* ws, size_t ws_bytes, void* stream);

/* Strip ("shared halo") path for stride-1, dilation-1, same-padded odd-K layers with Ci, Co multiples of 64 and <= 128
 * (larger layers are tensor-pipe bound on the im2col path), bf16 mode.  OPT-IN (cpt_conv2d_set_strip_enabled): parity-green and
 * 6x lighter on L2->SM traffic, but not faster than the im2col kernels on B200 (DESIGN.md §4 has the measurements).  Replaces the same reference computation as the
 * *_cl entry points (convolution_funcs.py:222-241 forward, :390-403 dX, :405-408 dW); what changes is the operand layout:
 * activations are staged ZERO-PADDED channels-last, act_pad[B][H+2P][W+2P][C] bf16, so that the K*K filter taps of 128
 * consecutive output positions read ONE strip of input rows from shared memory (each activation crosses L2->SM once
 * instead of once per tap) and the filter tiles stay resident in shared memory.
 *   cpt_conv2d_strip_supported       1 when the geometry / mode is covered
 *   cpt_channels_last_padded_bytes   bytes of act_pad
 *   cpt_to_channels_last_padded      x (NCHW fp32) -> act_pad (+ optional per-channel sums = db when staging dy)
 *   cpt_conv2d_fprop_strip           y = conv(x_pad, w) + bias; stats != NULL: also the batch-statistics partials of
 *                                    cpt_conv2d_fprop_cl_stats
 *   cpt_conv2d_dgrad_strip           dx from dy_pad (the forward strip with reversed taps)
 *   cpt_conv2d_wgrad_padded          dw from x_pad and dy_pad (im2col maps over the padded tensors; split-K, fixed order) */
int cpt_conv2d_set_strip_enabled(int enabled);   /* opt-in switch (default: env CPT_STRIP, off); returns the previous setting */
int cpt_conv2d_strip_supported(const cpt_conv2d_desc* d, int mode);
size_t cpt_channels_last_padded_bytes(int B, int C, int H, int W, int pad);
int cpt_to_channels_last_padded(const float* src, void* dst, int B, int C, int H, int W, int pad, float* chan_sum, void* ws,
                                size_t ws_bytes, void* stream);
size_t cpt_conv2d_strip_workspace_size(int op, const cpt_conv2d_desc* d);
int cpt_conv2d_fprop_strip(const cpt_conv2d_desc* d, const void* x_pad, const float* w, const float* bias, float* y, float* stats,
                           void* ws, size_t ws_bytes, void* stream);
int cpt_conv2d_dgrad_strip(const cpt_conv2d_desc* d, const void* dy_pad, const float* w, float* dx, void* ws, size_t ws_bytes,
                           void* stream);
int cpt_conv2d_wgrad_padded(const cpt_conv2d_desc* d, const void* x_pad, const void* dy_pad, float* dw, void* ws, size_t ws_bytes,
                            void* stream);

/* Packed-K path for first layers (Ci <= 16: ResNet stem 3->64 7x7/s2, VGG / MNIST conv1), bf16 mode.  The reference
 * evaluates these layers like any other (window view + einsum, convolution_funcs.py:357-408); the im2col-TMA path would
 * spend one 64-channel k-iteration per tap on 1-3 real channels.  Here the patches are written once as an explicit bf16
 * matrix col[(b, ho, wo)][c*K*K + j*K + kk] (row pitch round_up(Ci*K*K, 8)) and the three passes are dense GEMMs over it:
 *   cpt_conv2d_packed_bytes          bytes of `col`, or 0 when the geometry / mode is not covered (use the *_cl path)
 *   cpt_conv2d_im2col_pack           x (NCHW fp32) -> col                                   (kept for wgrad)
 *   cpt_conv2d_fprop_packed          y = col . w^T + bias                                    :222-241
 *   cpt_conv2d_dgrad_packed          dcol = dy_cl . w, dx = col2im(dcol) (fixed-order gather)  :390-403
 *   cpt_conv2d_wgrad_packed          dw = dy_cl^T . col (split-K, fixed-order reduce)         :405-408
 * dy_cl is the channels-last bf16 staging of dy made by cpt_to_channels_last (which also yields db). */
size_t cpt_conv2d_packed_bytes(const cpt_conv2d_desc* d, int mode);
size_t cpt_conv2d_packed_workspace_size(int op, const cpt_conv2d_desc* d);
int cpt_conv2d_im2col_pack(const cpt_conv2d_desc* d, const float* x, void* col, void* stream);
int cpt_conv2d_fprop_packed(const cpt_conv2d_desc* d, const void* col, const float* w,
                            const float* bias, float* y, void* ws, size_t ws_bytes, void* stream);
int cpt_conv2d_dgrad_packed(const cpt_conv2d_desc* d, const void* dy_cl, const float* w, float* dx,
                            void* ws, size_t ws_bytes, void* stream);
int cpt_conv2d_wgrad_packed(const cpt_conv2d_desc* d, const void* col, const void* dy_cl, float* dw,
                            void* ws, size_t ws_bytes, void* stream);

/* ---- Linear: compyute/nn/functional/linear_funcs.py:11-35 ---------------------------------- */
/* x: (N, In) (leading dims flattened), w: (Out, In), bias: (Out,) or NULL, y: (N, Out). */
size_t cpt_linear_workspace_size(int op, int64_t N, int In, int Out, int mode);
/* LinearFn.forward :15-23   y = x @ w.T (+ b) */
int cpt_linear_fwd(const float* x, const float* w, const float* bias, float* y, int64_t N, int In,
                   int Out, int mode, void* ws, size_t ws_bytes, void* stream);
/* LinearFn.backward :31     dx = dy @ w */
int cpt_linear_dgrad(const float* dy, const float* w, float* dx, int64_t N, int In, int Out,
                     int mode, void* ws, size_t ws_bytes, void* stream);
/* LinearFn.backward :32-33  dw = (dy.T @ x).sum(leading), db = dy.sum(leading); db may be NULL */
int cpt_linear_wgrad(const float* x, const float* dy, float* dw, float* db, int64_t N, int In,
                     int Out, int mode, void* ws, size_t ws_bytes, void* stream);

/* bf16 tensor-core mode with operands staged once by the caller (what LinearFn's cache does: x_bf16 made in forward is
 * reused by wgrad, dy_bf16 is shared by dgrad and wgrad, w_bf16 by forward and dgrad).  Staged layout: row-major bf16,
 * row pitch = cols rounded up to 8 elements, zero padded (cpt_cast_bf16). */
size_t cpt_cast_bf16_bytes(int64_t rows, int cols);
int cpt_cast_bf16(const float* src, void* dst, int64_t rows, int cols, void* stream);
int cpt_linear_fwd_bf16(const void* x_bf, const void* w_bf, const float* bias, float* y, int64_t N, int In,
                        int Out, void* stream);
int cpt_linear_dgrad_bf16(const void* dy_bf, const void* w_bf, float* dx, int64_t N, int In, int Out,
                          void* stream);
/* ws: room for up to 16 split-K partials (16 * Out * In floats), see cpt_linear_workspace_size(CPT_OP_WGRAD, ...) */
int cpt_linear_wgrad_bf16(const void* x_bf, const void* dy_bf, float* dw, int64_t N, int In, int Out, void* ws,
                          size_t ws_bytes, void* stream);
/* out[c] = sum over n, hw of x[n][c][hw] (db of Linear: HW = 1; of Conv2D: HW = Ho*Wo).  ws: C * 64 floats. */
int cpt_channel_sum(const float* x, float* out, int N, int C, int HW, void* ws, size_t ws_bytes, void* stream);

/* ---- Pooling: compyute/nn/functional/pooling_funcs.py -------------------------------------- */
/* MaxPooling2DFn.forward :71-76 — stride = k, no padding, floor.  y: (B, C, H/k, W/k). */
int cpt_maxpool2d_fwd(const float* x, float* y, int B, int C, int H, int W, int k, void* stream);
/* MaxPooling2DFn.backward :79-82 — equality mask: every tied maximum receives dy; uncovered tail
 * rows/cols get 0; non-selected positions are dy*0 (so -0.0 where dy < 0), bit-exact. */
int cpt_maxpool2d_bwd(const float* x, const float* y, const float* dy, float* dx, int B, int C,
                      int H, int W, int k, void* stream);
/* AvgPooling2DFn :111-121 */
int cpt_avgpool2d_fwd(const float* x, float* y, int B, int C, int H, int W, int k, void* stream);
int cpt_avgpool2d_bwd(const float* dy, float* dx, int B, int C, int H, int W, int k, void* stream);

/* ---- BatchNorm 1-D/2-D: compyute/nn/functional/normalization_funcs.py:10-177 ---------------- */
/* x viewed as (N, C, HW): BatchNorm2D → (B, C, H*W); BatchNorm1D 2-D input → (B, C, 1), 3-D → (B, C, S).
 * Training forward :139-147: batch mean / biased var for normalisation, running stats updated with
 * the UNBIASED variance.  save_mean/save_rstd (C floats each) replace the reference's cached
 * (std, x_norm): backward recomputes x_norm from x. */
size_t cpt_bn_workspace_size(int N, int C, int HW);
int cpt_bn_fwd_train(const float* x, const float* w, const float* b, const float* rmean,
                     const float* rvar, float* y, float* rmean_out, float* rvar_out,
                     float* save_mean, float* save_rstd, int N, int C, int HW, float m, float eps,
                     void* ws, size_t ws_bytes, void* stream);
/* Inference forward :148-152 (running stats). */
int cpt_bn_fwd_eval(const float* x, const float* w, const float* b, const float* rmean,
                    const float* rvar, float* y, float* save_mean, float* save_rstd, int N, int C,
                    int HW, float eps, void* stream);
/* backward :161-177: dx = w/(std n) (n dy − Σdy − x̂ Σ(dy x̂)), dw = Σ(dy x̂), db = Σdy. */
int cpt_bn_bwd(const float* x, const float* dy, const float* w, const float* save_mean,
               const float* save_rstd, float* dx, float* dw, float* db, int N, int C, int HW,
               void* ws, size_t ws_bytes, void* stream);
/* Fused variants for the Sequential peephole BatchNorm -> ReLU (normalizations.py:150-171 followed by activations.py:114-120):
 * act = CPT_ACT_RELU applies y = max(bn(x), 0) in the same pass; backward recomputes the ReLU mask (y > 0) from
 * (x, mean, rstd, w, b) with the forward's own expression, so no mask is stored and dy*mask is folded into both backward passes.
 * act = CPT_ACT_NONE is identical to the plain entry points above. */
#define CPT_ACT_NONE 0
#define CPT_ACT_RELU 1
int cpt_bn_act_fwd_train(const float* x, const float* w, const float* b, const float* rmean,
                         const float* rvar, float* y, float* rmean_out, float* rvar_out,
                         float* save_mean, float* save_rstd, int N, int C, int HW, float m, float eps,
                         int act, void* ws, size_t ws_bytes, void* stream);
int cpt_bn_act_fwd_eval(const float* x, const float* w, const float* b, const float* rmean,
                        const float* rvar, float* y, float* save_mean, float* save_rstd, int N, int C,
                        int HW, float eps, int act, void* stream);
int cpt_bn_act_bwd(const float* x, const float* dy, const float* w, const float* b,
                   const float* save_mean, const float* save_rstd, float* dx, float* dw, float* db,
                   int N, int C, int HW, int act, void* ws, size_t ws_bytes, void* stream);
/* Producer-side staging for the tensor-core convolutions: same computation as cpt_bn_act_*, but the apply pass also writes
 * the channels-last bf16 copy (layout of cpt_to_channels_last, CPT_MODE_BF16, with H*W = HW) of its output — y_cl for the
 * convolution that consumes y in forward, dx_cl for the backward of the convolution that produced x.  dx_chan_sum (C floats,
 * may be NULL) receives the per-channel sums of dx, i.e. that convolution's bias gradient (convolution_funcs.py:252).
 * fp32 results are bit-identical to cpt_bn_act_*. */
size_t cpt_bn_cl_workspace_size(int N, int C, int HW);
int cpt_bn_act_fwd_train_cl(const float* x, const float* w, const float* b, const float* rmean,
                            const float* rvar, float* y, void* y_cl, float* rmean_out, float* rvar_out,
                            float* save_mean, float* save_rstd, int N, int C, int HW, float m,
                            float eps, int act, void* ws, size_t ws_bytes, void* stream);
int cpt_bn_act_fwd_eval_cl(const float* x, const float* w, const float* b, const float* rmean,
                           const float* rvar, float* y, void* y_cl, float* save_mean, float* save_rstd,
                           int N, int C, int HW, float eps, int act, void* stream);
int cpt_bn_act_bwd_cl(const float* x, const float* dy, const float* w, const float* b,
                      const float* save_mean, const float* save_rstd, float* dx, void* dx_cl,
                      float* dx_chan_sum, float* dw, float* db, int N, int C, int HW, int act, void* ws,
                      size_t ws_bytes, void* stream);
/* Batch statistics from the producing convolution's epilogue (tensor-core modes).  cpt_conv2d_fprop_*_stats are the forward
 * passes above that additionally leave, in `stats` (cpt_conv2d_stats_bytes bytes = cpt_conv2d_stats_slots() x Co x 2 floats),
 * per-epilogue-warp column sums (Σ a, Σ a²) of the bias-free accumulators; cpt_bn_act_fwd_train_presum is the training forward
 * of normalization_funcs.py:138-147 that adds the slots in fixed order (double) instead of reading x a first time:
 * mean = Σa/n + conv_bias, var = Σa²/n - (Σa/n)².  y_cl may be NULL.  Not bit-identical to the two-pass statistics (different
 * summation order / variance formula), so only the tf32 / bf16 modes use it. */
size_t cpt_conv2d_stats_bytes(const cpt_conv2d_desc* d);
int cpt_conv2d_stats_slots(void);
int cpt_conv2d_fprop_cl_stats(const cpt_conv2d_desc* d, const void* x_cl, const float* w,
                              const float* bias, float* y, float* stats, int mode, void* ws,
                              size_t ws_bytes, void* stream);
int cpt_conv2d_fprop_packed_stats(const cpt_conv2d_desc* d, const void* col, const float* w,
                                  const float* bias, float* y, float* stats, void* ws, size_t ws_bytes,
                                  void* stream);
/* ---- Conv2D -> ReLU pairs ----
 * Replaces ReLUFn.forward behind Conv2DFn.forward (compyute/nn/functional/activation_funcs.py:26-29 after
 * convolution_funcs.py:237-238) and ReLUFn.backward in front of Conv2DFn.backward (activation_funcs.py:32-34 before
 * convolution_funcs.py:244-252):
 *   cpt_conv2d_fprop_cl_relu / _packed_relu   y = max(conv(x, w) + bias, 0) from the GEMM epilogue (NaN propagates like
 *                                             numpy.maximum); nothing else is written — the mask is y > 0
 *   cpt_to_channels_last_gated                staging of dy for dgrad / wgrad with dy * (gate > 0) applied in the same pass
 *                                             (gate = the y above); chan_sum = db of the gated values
 * Bit-identical to the two separate layers. */
int cpt_conv2d_fprop_cl_relu(const cpt_conv2d_desc* d, const void* x_cl, const float* w, const float* bias, float* y, int mode,
                             void* ws, size_t ws_bytes, void* stream);
int cpt_conv2d_fprop_packed_relu(const cpt_conv2d_desc* d, const void* col, const float* w, const float* bias, float* y, void* ws,
                                 size_t ws_bytes, void* stream);
int cpt_to_channels_last_gated(const float* src, const float* gate, void* dst, int B, int C, int H, int W, int mode,
                               float* chan_sum, void* ws, size_t ws_bytes, void* stream);
int cpt_bn_act_fwd_train_presum(const float* x, const float* w, const float* b, const float* rmean,
                                const float* rvar, float* y, void* y_cl, float* rmean_out,
                                float* rvar_out, float* save_mean, float* save_rstd, int N, int C, int HW,
                                float m, float eps, int act, const float* stats, int stat_slots,
                                const float* conv_bias, void* stream);

/* Tail of a residual block in one pass: y = relu(bn(x) + skip), i.e. BatchNorm2D.forward (normalizations.py:150-165), the
 * `y += skip` of ResidualConnection.forward (containers.py:153-157) and ReLUFn.forward (activation_funcs.py:26-29).  The batch
 * statistics come from a statistics-only call of the forward entry points above (cpt_bn_act_fwd_train / _presum / _merged /
 * _fwd_eval with y == NULL), this call applies them.  mask: the bit-packed (y > 0) of cpt_relu_fwd (same layout, consumed by
 * cpt_relu_bwd), may be NULL.  Bit-identical to the three separate passes; needs H*W % 4 == 0 and < 2^31 elements
 * (CPT_ERR_UNSUPPORTED otherwise: the caller then runs the separate passes). */
int cpt_bn_add_relu_apply(const float* x, const float* skip, const float* w, const float* b,
                          const float* save_mean, const float* save_rstd, float* y, uint8_t* mask, int N,
                          int C, int HW, void* stream);
/* BatchNorm2D -> ReLU -> MaxPooling2D(kernel 2) as one forward pass and one backward pass pair (normalizations.py:150-171,
 * activations.py:114-120, pooling_funcs.py:71-82): the full-resolution activation is neither written nor read.  Forward: the
 * statistics come from a statistics-only call of the forward entry points above (y == NULL); y is the POOLED output
 * (N, C, H/2, W/2).  Backward: dy_pool is the gradient of the pooled output; the pooling tie mask and the ReLU mask are
 * recomputed from x.  Bit-identical to the three separate layers (same expressions, same summation order).  Needs H even and
 * W % 4 == 0 (CPT_ERR_UNSUPPORTED otherwise).  ws: cpt_bn_workspace_size(N, C, H*W). */
int cpt_bn_relu_pool2_fwd(const float* x, const float* w, const float* b, const float* save_mean,
                          const float* save_rstd, float* y, int N, int C, int H, int W, void* stream);
int cpt_bn_relu_pool2_bwd(const float* x, const float* dy_pool, const float* w, const float* b,
                          const float* save_mean, const float* save_rstd, float* dx, float* dw, float* db,
                          int N, int C, int H, int W, void* ws, size_t ws_bytes, void* stream);
/* Synchronised BatchNorm for the batch-sharded data-parallel mode: the statistics of
 * normalization_funcs.py:139-147 / the sums of :169-175 taken over the GLOBAL batch.  Each pass is split around the
 * collective the caller runs (torch.distributed / NCCL):
 *   forward : cpt_bn_local_stats -> all-gather of stats[3][C] = (mean, M2 = Σ(x-mean)², count) -> cpt_bn_act_fwd_train_merged,
 *             which merges the `world` triples in rank order (Chan's pairwise update, double) — the same arithmetic on the
 *             same values on every rank, so replicas stay bit-identical — writes the global element count to the device
 *             float *global_count and applies (y_cl may be NULL);
 *   backward: cpt_bn_act_bwd_local_sums -> SUM all-reduce of sums[2][C] = (Σdy, Σdy·x̂) -> cpt_bn_act_bwd_apply_global with the
 *             forward's device-resident global count (no host synchronisation anywhere).  dw / db are the LOCAL sums (the gradient exchange averages them like every other
 *             parameter gradient).  dx_cl / dx_chan_sum as in cpt_bn_act_bwd_cl, may be NULL.
 * ws: cpt_bn_workspace_size (cpt_bn_cl_workspace_size when dx_cl is given). */
int cpt_bn_local_stats(const float* x, float* stats, int N, int C, int HW, void* ws, size_t ws_bytes, void* stream);
int cpt_bn_act_fwd_train_merged(const float* x, const float* w, const float* b, const float* rmean,
                                const float* rvar, float* y, void* y_cl, float* rmean_out,
                                float* rvar_out, float* save_mean, float* save_rstd, int N, int C, int HW,
                                float m, float eps, int act, const float* gathered, int world,
                                float* global_count, void* stream);
int cpt_bn_act_bwd_local_sums(const float* x, const float* dy, const float* w, const float* b,
                              const float* save_mean, const float* save_rstd, float* sums, float* dw,
                              float* db, int N, int C, int HW, int act, void* ws, size_t ws_bytes,
                              void* stream);
int cpt_bn_act_bwd_apply_global(const float* x, const float* dy, const float* w, const float* b,
                                const float* save_mean, const float* save_rstd, const float* sums,
                                const float* global_count, float* dx, void* dx_cl, float* dx_chan_sum, int N,
                                int C, int HW, int act, void* ws, size_t ws_bytes, void* stream);

/* ---- activations / elementwise ------------------------------------------------------------- */
/* ReLUFn.forward activation_funcs.py:26-29.  mask = bit-packed (y > 0), 4 * ((n + 31) / 32) bytes, 4-byte aligned, in a
 * layout private to cpt_relu_fwd / cpt_relu_bwd; may be NULL. */
int cpt_relu_fwd(const float* x, float* y, uint8_t* mask, int64_t n, void* stream);
/* Linear + ReLU in one pass (linear_funcs.py:15-23 followed by activation_funcs.py:26-29; bf16 mode, Out % 32 == 0): the GEMM
 * epilogue writes y = max(x @ w.T + b, 0), optionally the same values as bf16 rows (y_bf16: the x operand of a following
 * Linear layer, layout of cpt_cast_bf16) and the mask (y > 0) in PLAIN bit order — element e of the (N, Out) output is bit
 * e % 32 of 32-bit word e / 32; 4 * ceil(N*Out / 32) bytes.  cpt_relu_bwd_plain is ReLUFn.backward for that mask layout
 * (dx_bf16 optional, as cpt_relu_bwd_lp).  Bit-identical to cpt_linear_fwd_bf16 followed by cpt_relu_fwd_lp. */
int cpt_linear_relu_fwd_bf16(const void* x_bf, const void* w_bf, const float* bias, float* y, void* y_bf16,
                             uint8_t* mask, int64_t N, int In, int Out, void* stream);
int cpt_relu_bwd_plain(const float* dy, const uint8_t* mask, float* dx, void* dx_bf16, int64_t n, void* stream);
/* LinearFn.backward's dx = dy @ w (linear_funcs.py:31) followed by the backward of the ReLU that produced this layer's input
 * (activation_funcs.py:32-34) in the dgrad epilogue: dx = (dy @ w) * mask, mask in plain bit order over the (N, In) input (as
 * written by cpt_linear_relu_fwd_bf16 of the previous layer), In % 32 == 0; dx_bf16 (optional) = the same values as bf16 rows,
 * the dy operand of the previous Linear layer's backward.  Bit-identical to cpt_linear_dgrad_bf16 + cpt_relu_bwd_plain. */
int cpt_linear_dgrad_relu_bf16(const void* dy_bf, const void* w_bf, const uint8_t* mask, float* dx, void* dx_bf16,
                               int64_t N, int In, int Out, void* stream);
/* ReLUFn.backward :32-34   dx = dy * mask */
int cpt_relu_bwd(const float* dy, const uint8_t* mask, float* dx, int64_t n, void* stream);
/* Same passes, additionally writing the result as bf16 in the same linear order (y_bf16 / dx_bf16, 8-byte aligned, may be NULL):
 * the pre-cast operand of a neighbouring Linear layer in bf16 mode (cpt_linear_*_bf16), when the feature count is a multiple
 * of 8 so that the bf16 row pitch equals the row length. */
int cpt_relu_fwd_lp(const float* x, float* y, uint8_t* mask, void* y_bf16, int64_t n, void* stream);
int cpt_relu_bwd_lp(const float* dy, const uint8_t* mask, float* dx, void* dx_bf16, int64_t n, void* stream);
/* a += b  (ResidualConnection containers.py:153-162; grad accumulation module.py:399-400) */
int cpt_add_inplace(float* a, const float* b, int64_t n, void* stream);
/* y = alpha * x (+ y if accumulate) */
int cpt_axpby(float* y, const float* x, float alpha, int accumulate, int64_t n, void* stream);
int cpt_fill(float* a, float value, int64_t n, void* stream);
/* is_nan(x).any() without a host sync (module.py:332,366): ORs 1 into *flag when any NaN. */
int cpt_isnan_flag(const float* x, int64_t n, int* flag, void* stream);
/* out[0] = Σ x   (fp32 tree; used for loss/metric scalars) */
int cpt_sum(const float* x, int64_t n, float* out, void* stream);

/* ---- loss / regularisation (closing the train step on the device) ---------- */
/* CrossEntropyLossFn.forward loss_funcs.py:57-64: probs = softmax(logits), loss = -mean log(p_t + eta).
 * targets int32.  loss: 1 float; row_loss: B floats of scratch (per-sample losses, reduced in a fixed order). */
int cpt_softmax_ce_fwd(const float* logits, const int32_t* targets, float* probs, float* loss,
                       float* row_loss, int B, int NC, float eta, void* stream);
/* backward :67-69: dlogits = (probs - onehot) / B */
int cpt_softmax_ce_bwd(const float* probs, const int32_t* targets, float* dlogits, int B, int NC,
                       void* stream);
/* count of argmax(logits) == target → *correct (int, pre-zeroed by the call)  metric_funcs.py:10-25 */
int cpt_accuracy_count(const float* logits, const int32_t* targets, int* correct, int B, int NC,
                       void* stream);
/* DropoutFn regularization_funcs.py:15-32; mask int8 Bernoulli(1-p) from a counter-based RNG.  live_seed (device, may be
 * NULL) is mixed into the seed at run time so that a CUDA-graph replay draws a fresh mask. */
int cpt_dropout_fwd(const float* x, float* y, int8_t* mask, int64_t n, float p, uint64_t seed,
                    const uint64_t* live_seed, void* stream);
int cpt_dropout_bwd(const float* dy, const int8_t* mask, float* dx, int64_t n, float p,
                    void* stream);

/* ---- optimizers: compyute/nn/optimizers.py --------------------------------------------------- */
typedef struct cpt_param_entry {
  float* p;       /* parameter, updated in place (optimizers.py:174,269) */
  const float* g; /* gradient (already all-reduced SUM over ranks in DP mode) */
  float* m;       /* Adam first moment / SGD velocity (may be NULL for plain SGD) */
  float* v;       /* Adam second moment (NULL for SGD) */
  int64_t n;      /* elements; 0 = skip (p.grad is None, optimizers.py:155,247) */
} cpt_param_entry;

/* Multi-tensor fused steps.  `table` is a DEVICE array of n_entries entries.  grad_scale
 * multiplies every gradient first (1/world_size in data-parallel mode).
 * Adam.step :241-271 (decoupled = 0), AdamW.step :335-362 (decoupled = 1).  m_div = 1 - beta1^t,
 * v_div = 1 - beta2^t are computed by the caller in double like the reference (:243-244).
 * live_scalars (device, may be NULL): when given, the per-step scalars are read from device memory instead of the
 * by-value arguments — {lr, m_div, v_div} for Adam/AdamW, {lr, m_div, v_div, mu, mu_next, g_div} for NAdam, {lr} for SGD —
 * so a CUDA graph that captured the step keeps following LR schedulers and bias correction. */
int cpt_adam_step(const cpt_param_entry* table, int n_entries, int64_t max_n, float lr, float beta1,
                  float beta2, float eps, float weight_decay, float m_div, float v_div,
                  float grad_scale, int decoupled, const float* live_scalars, void* stream);
/* NAdam.step :437-475.  mu, mu_next, m_div = 1 - mu_prod*mu_next, g_div = 1 - mu_prod, v_div = 1 - beta2^t are computed
 * by the caller in double like the reference (:438-447). */
int cpt_nadam_step(const cpt_param_entry* table, int n_entries, int64_t max_n, float lr, float beta1,
                   float beta2, float eps, float weight_decay, float mu, float mu_next, float m_div,
                   float g_div, float v_div, float grad_scale, const float* live_scalars, void* stream);
/* Writes n <= 8 host floats / one uint64 into device memory with a 1-block kernel (values travel in the launch
 * parameters): how the live scalars / Dropout replay counter are refreshed before a CUDA-graph replay. */
int cpt_set_live_scalars(float* dst, const float* host_values, int n, void* stream);
int cpt_set_u64(uint64_t* dst, uint64_t value, void* stream);
/* SGD.step :152-176 (momentum / nesterov / L2 weight decay). */
int cpt_sgd_step(const cpt_param_entry* table, int n_entries, int64_t max_n, float lr,
                 float momentum, int nesterov, float weight_decay, float grad_scale,
                 const float* live_scalars, void* stream);

/* ---- generic device-tensor operators -------------------------------------------
 * What the reference reaches as `Tensor.data <op> other` / `Tensor.data.<reduction>(dim)` (compyute/tensors.py:196-292,
 * 552-682) and `device.module.<fn>` (compyute/tensor_ops/ modules) on CuPy arrays.  Values are fp32 unless a dtype code says
 * otherwise; booleans are uint8 0/1 (NumPy bool layout).  Shapes of up to CPT_MAX_DIMS dims, strides in ELEMENTS,
 * innermost dim last; `dims`/strides/`reduced` are HOST arrays read during the call. */
#define CPT_MAX_DIMS 6
#define CPT_DT_F32 0
#define CPT_DT_I32 1
#define CPT_DT_I64 2
#define CPT_DT_U8 3 /* bool */
#define CPT_DT_F64 4
/* binary ops: arithmetic → fp32 out, comparisons → uint8 out  (tensors.py:196-292, selection_ops.py:89-144) */
#define CPT_EW_ADD 0
#define CPT_EW_SUB 1
#define CPT_EW_MUL 2
#define CPT_EW_DIV 3
#define CPT_EW_POW 4
#define CPT_EW_MAX 5
#define CPT_EW_MIN 6
#define CPT_EW_FLOORDIV 7
#define CPT_EW_MOD 8
#define CPT_EW_LT 9
#define CPT_EW_GT 10
#define CPT_EW_LE 11
#define CPT_EW_GE 12
#define CPT_EW_EQ 13
#define CPT_EW_NE 14
/* out[i] = a[ia] op b[ib] with NumPy broadcasting: `dims` is the OUTPUT shape, sa / sb the operands' element strides over
 * it (0 on broadcast dims); out is C-contiguous and may alias a (in-place operators).  b == NULL: the second operand is
 * `scalar` (scalar_mode 1: a op scalar, 2: scalar op a; a must then be contiguous). */
int cpt_ew_binary(int op, void* out, const float* a, const float* b, float scalar, int scalar_mode, int ndim,
                  const int64_t* dims, const int64_t* sa, const int64_t* sb, void* stream);
/* unary ops (tensors.py:267,543; unary_ops.py:33-438); ISNAN writes uint8; CLIP uses p0 = min, p1 = max; ROUND p0 = 10^decimals */
#define CPT_UN_NEG 0
#define CPT_UN_ABS 1
#define CPT_UN_EXP 2
#define CPT_UN_LOG 3
#define CPT_UN_LOG2 4
#define CPT_UN_LOG10 5
#define CPT_UN_SQRT 6
#define CPT_UN_TANH 7
#define CPT_UN_SIN 8
#define CPT_UN_COS 9
#define CPT_UN_TAN 10
#define CPT_UN_SINH 11
#define CPT_UN_COSH 12
#define CPT_UN_CLIP 13
#define CPT_UN_ISNAN 14
#define CPT_UN_ROUND 15
#define CPT_UN_SQUARE 16
#define CPT_UN_RECIP 17
int cpt_ew_unary(int op, void* out, const float* a, float p0, float p1, int64_t n, void* stream);
/* boolean arrays: op 0 and, 1 or, 2 xor, 3 not (b ignored)   tensors.py:270 (__invert__) */
int cpt_logic(int op, uint8_t* out, const uint8_t* a, const uint8_t* b, int64_t n, void* stream);
/* reductions over the axes flagged in `reduced` (tensors.py:552-682, reduction_ops.py, selection_ops.py:22-127).
 * x: C-contiguous fp32 (CPT_DT_F32) or bool (CPT_DT_U8).  out: C-contiguous over the kept axes — fp32 for SUM (× scale, so
 * mean = SUM with scale 1/n) / SUMSQ / PROD / MAX / MIN (NaN propagates like NumPy), uint8 for ANY / ALL, int64 for COUNT
 * (non-zeros) and ARGMAX (first maximum, NaN counts as the maximum; one axis or the whole tensor).  Deterministic: the
 * reduced range is split over the grid and the partials are combined in fixed order (ws from cpt_reduce_workspace_size). */
#define CPT_RED_SUM 0
#define CPT_RED_SUMSQ 1
#define CPT_RED_PROD 2
#define CPT_RED_MAX 3
#define CPT_RED_MIN 4
#define CPT_RED_ANY 5
#define CPT_RED_ALL 6
#define CPT_RED_COUNT 7
#define CPT_RED_ARGMAX 8
size_t cpt_reduce_workspace_size(int64_t n_out);
int cpt_reduce(int op, void* out, const void* x, int x_dtype, int ndim, const int64_t* dims, const int32_t* reduced,
               float scale, void* ws, size_t ws_bytes, void* stream);
/* dst[Σ i_k·dst_strides[k]] = src[Σ i_k·src_strides[k]] over `dims` — slicing / __setitem__ / permute / flip / concat / pad /
 * broadcast_to (tensors.py:176-183, 615-667; shape_ops.py:34-458).  elem_size 1, 4 or 8 bytes; strides may be negative
 * or 0 (source only); base offsets are folded into the pointers by the caller. */
int cpt_strided_copy(void* dst, const void* src, int elem_size, int ndim, const int64_t* dims,
                     const int64_t* dst_strides, const int64_t* src_strides, void* stream);
/* dst[i, :] = src[idx[i], :] — integer-array indexing of the leading axis (Dataloader batch gather
 * nn/utils/dataloaders.py:65-66, one-hot identity(n)[t] preprocessing/basic.py:125).  Negative indices wrap; an index
 * out of range sets *err_flag (device int, may be NULL) and is clamped. */
int cpt_gather_rows(void* dst, const void* src, const void* idx, int idx_dtype, int64_t n_idx, int64_t row_bytes,
                    int64_t n_src_rows, int* err_flag, void* stream);
/* astype (tensors.py:372-455) between CPT_DT_* types; → bool is `!= 0` */
int cpt_cast(void* dst, int dst_dtype, const void* src, int src_dtype, int64_t n, void* stream);
/* arange (creation_ops.py:24-57): dst[i] = start + i·step */
int cpt_arange(void* dst, int dtype, double start, double step, int64_t n, void* stream);
/* random/random.py:54-184 on a counter-based device RNG (statistical contract, not NumPy's stream): kind 0 uniform
 * [p0, p1), 1 normal(mean p0, std p1), 2 integers in [p0, p1) stored as fp32 */
int cpt_random_fill(float* dst, int64_t n, int kind, float p0, float p1, uint64_t seed, void* stream);

/* ---- diagnostics ----------------------------------------------------------------------------- */
/* Synchronises the device and returns (then clears) the tensor-core pipeline watchdog flag: 0 = healthy,
 * non-zero = an mbarrier wait timed out inside a tcgen05 kernel (results invalid).  Test/debug helper. */
int cpt_tc_check_status(void);
/* Number of kernels this library has launched since it was loaded (bench.py's gpu_launches). */
uint64_t cpt_launch_count(void);
/* The tensor-core kernels are persistent (one CTA per SM).  n > 0 makes them leave n SMs (rounded down to whole pairs)
 * free, so that a concurrent kernel of another stream — the NCCL all-reduce of the overlapped data-parallel exchange —
 * finds an SM to run on; 0 restores the full grid.  Takes effect for subsequent launches. */
int cpt_tc_reserve_sms(int n);

/* ---------------------------------------------------------------------------------------------------------------------
 * Data parallelism (the reference is single-device: compyute/backend.py:88-89).  One process per GPU.
 *
 * (1) Gradient exchange over NCCL: SUM all-reduce of the flat gradient arena before the update (hook points
 *     compyute/nn/modules/module.py:392-400, nn/optimizers.py:152,241).  libnccl.so.2 is bound at run time (dlopen).
 *     Rank 0 calls cpt_nccl_unique_id and ships the 128 bytes to the other ranks (any side channel), every rank calls
 *     cpt_nccl_init; cpt_nccl_allreduce_sum_f32 is in place and asynchronous on `stream`. */
int cpt_nccl_unique_id(void* out128);
int cpt_nccl_init(int rank, int world, const void* unique_id128);
int cpt_nccl_world_size(void);
int cpt_nccl_allreduce_sum_f32(float* ptr, int64_t count, void* stream);
int cpt_nccl_destroy(void);

/* (2) Optimizer step fused with the exchange over NVLink / NVSwitch peer memory (csrc/dp_step.cu): the gradient arena G and
 *     the parameter arena P are symmetric allocations (same layout on every rank, mapped into every peer and, on NVSwitch,
 *     into a multicast object).  Rank r owns the shard [shard_off, shard_off + shard_elems) of both arenas; one kernel reads the
 *     sum of its gradient shard over all ranks (multimem.ld_reduce through the multicast mapping, or `world` peer loads in rank
 *     order when g_mc == NULL), applies SGD (optimizers.py:152-176) / Adam / AdamW (:241-271, :335-362) to its shard with moment
 *     buffers that exist for the shard only, and writes the new parameters to every replica (multimem.st / peer stores).
 *     Ordering across ranks is the caller's: all gradients written before the call, all replicas complete before P is read
 *     again (a symmetric-memory barrier on `stream` before and after).  pre_reduced != 0: g_local already holds the averaged
 *     global gradient (an earlier all-reduce, e.g. for gradient clipping); grad_scale is then applied to it as is. */
typedef struct cpt_dp_view {
  float* p_local;            /* this rank's parameter arena                                   */
  float* p_mc;               /* multicast address of the parameter arenas, or NULL            */
  const uint64_t* p_peers;   /* device array of `world` arena base addresses (used if !p_mc)  */
  const float* g_local;      /* this rank's gradient arena                                    */
  const float* g_mc;         /* multicast address of the gradient arenas, or NULL             */
  const uint64_t* g_peers;   /* device array of `world` arena base addresses (used if !g_mc)  */
  int64_t shard_off;         /* first element of this rank's shard (multiple of 4)            */
  int64_t shard_elems;       /* elements in the shard (multiple of 4)                         */
  int32_t world;
  int32_t pre_reduced;
  int32_t max_ctas_per_sm;   /* 0: default grid; n > 0: at most n CTAs per SM (a step overlapped with other kernels) */
  int32_t reserved;
} cpt_dp_view;
int cpt_dp_adam_step(const cpt_dp_view* view, float* m, float* v, float lr, float beta1, float beta2, float eps, float weight_decay,
                     float m_div, float v_div, float grad_scale, int decoupled, const float* live_scalars, void* stream);
int cpt_dp_sgd_step(const cpt_dp_view* view, float* velocity, float lr, float momentum, int nesterov, float weight_decay,
                    float grad_scale, const float* live_scalars, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COMPYUTE_B200_H */
