"""CPU oracle: a NumPy restatement of Compyute's CNN-training hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``compyute_b200/`` imports this module; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may.  The product path is the CUDA library and fails loudly without it.

Every function restates the algorithm of one reference function (cited as file:line relative to
``/root/reference``) with the *same evaluation strategy* -- ``as_strided`` window views and
un-optimised ``numpy.einsum`` for the convolutions, ``@`` for Linear, ``ndarray.mean/var`` for
BatchNorm -- so that (a) results agree with the reference to the last bit on the same NumPy and
(b) timing it is a fair stand-in ("port") for the reference's CPU path on a box where
``/root/reference`` does not exist.

Parity pinning: ``oracle/gen_golden.py`` imports the real reference (through an import shim for the
absent ``cupy``/``tensorboardX``) and stores its outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this restatement against those vectors bit-for-bit
(``numpy.array_equal``) on every run, so the oracle is pinned to reference outputs, not to itself.

Conventions: plain ``numpy.ndarray`` in and out, fp32 unless stated; ``cache`` is a python list used
as a LIFO stack exactly like ``FunctionCache`` (nn/functional/functions.py:12-28).
"""

from __future__ import annotations

import math

import numpy as np
from numpy.lib.stride_tricks import as_strided

# --------------------------------------------------------------------------------------
# helpers  (compyute/tensor_ops/shape_ops.py)
# --------------------------------------------------------------------------------------


def window_view2d(x: np.ndarray, window: int, stride: int = 1) -> np.ndarray:
    """Zero-copy (…, Y, X, window, window) view.  shape_ops.py:278-305 (``pooling2d``).

    Like the reference, *both* output extents are derived from ``x.shape[-1]`` (square inputs).
    """
    out = (x.shape[-1] - window) // stride + 1
    s = x.strides
    shape = (*x.shape[:-2], out, out, window, window)
    strides = (*s[:-2], s[-2] * stride, s[-1] * stride, s[-2], s[-1])
    return as_strided(x, shape, strides)


def repeat2d(x: np.ndarray, n: int) -> np.ndarray:
    """Nearest-neighbour upsampling by ``n`` in the last two dims.  shape_ops.py:335-359."""
    s = x.strides
    v = as_strided(x, (*x.shape[:-1], n, x.shape[-1], n), (*s[:-1], 0, s[-1], 0))
    return v.reshape((*v.shape[:-4], v.shape[-4] * n, v.shape[-2] * n))


def pad_to_shape(x: np.ndarray, shape: tuple[int, ...]) -> np.ndarray:
    """Zero-pad at the *end* of every dim up to ``shape``.  shape_ops.py:209-227."""
    if x.shape == tuple(shape):
        return x
    return np.pad(x, tuple((0, shape[i] - x.shape[i]) for i in range(x.ndim)))


def dilate2d(x: np.ndarray, dilation: int) -> np.ndarray:
    """Insert ``dilation-1`` zeros between elements.  convolution_funcs.py:298-309."""
    if dilation == 1:
        return x
    h = dilation * (x.shape[-2] - 1) + 1
    w = dilation * (x.shape[-1] - 1) + 1
    y = np.zeros((*x.shape[:-2], h, w), dtype=x.dtype)
    y[..., ::dilation, ::dilation] = x
    return y


def pad2d(x: np.ndarray, padding: int) -> np.ndarray:
    """Symmetric zero padding of the last two dims.  convolution_funcs.py:341-348."""
    if padding == 0:
        return x
    widths = tuple([(0, 0)] * (x.ndim - 2) + [(padding, padding)] * 2)
    return np.pad(x, widths)


def same_padding(kernel_size: int, dilation: int) -> int:
    """``padding="same"`` → symmetric int.  nn/modules/convolutions.py:23-28."""
    return (kernel_size * dilation - 1) // 2


# --------------------------------------------------------------------------------------
# Conv2D  (compyute/nn/functional/convolution_funcs.py:218-410)
# --------------------------------------------------------------------------------------


def conv2d_forward(cache, x, f, b, padding: int, stride: int, dilation: int):
    """``Conv2DFn.forward`` (:222-241): dilate f → pad x → window view → einsum → + bias."""
    if x.ndim != 4:
        raise ValueError(f"Expected input to be 4D, got {x.ndim}D.")
    f_d = dilate2d(f, dilation)  # :298-309
    x_p = pad2d(x, padding)  # :341-348
    win = window_view2d(x_p, f_d.shape[-1], stride)  # (B, Ci, Y, X, Fy, Fx) :380
    y = np.ascontiguousarray(np.einsum("biyxjk,oijk->boyx", win, f_d))  # :381-382
    if b is not None:
        y += b.reshape(-1, 1, 1)  # :237-238
    cache.append((x_p, f_d, stride, padding, dilation, b is not None))
    return y


def conv2d_backward(cache, dy):
    """``Conv2DFn.backward`` (:244-254) incl. ``RawConv2DFn.backward`` (:387-410)."""
    x_p, f_d, stride, padding, dilation, has_bias = cache.pop()
    kd = f_d.shape[-1]

    g = dilate2d(dy, stride)  # :391  zeros where the stride skipped
    t = x_p.shape[-1] - kd + 1
    g = pad_to_shape(g, (*g.shape[:-2], t, t))  # :394-395
    g = pad2d(g, kd - 1)  # :398  full padding

    # input grads :401-403
    win = window_view2d(g, kd)  # (B, Co, Y, X, Fy, Fx)
    dx = np.ascontiguousarray(np.einsum("boyxjk,oijk->biyx", win, np.flip(f_d, (-2, -1))))

    # filter grads :406-408
    win = window_view2d(g, x_p.shape[-1])  # (B, Co, Fy, Fx, Y, X)
    df = np.einsum("bojkyx,biyx->oijk", win, x_p)
    df = np.ascontiguousarray(np.flip(df, (-2, -1)))

    if padding != 0:  # Pad2DFn.backward :351-355
        dx = np.ascontiguousarray(dx[..., padding:-padding, padding:-padding])
    if dilation != 1:  # Dilation2DFn.backward :312-316
        df = df[..., ::dilation, ::dilation]
    db = dy.sum((0, 2, 3)) if has_bias else None  # :252
    return dx, df, db


# --------------------------------------------------------------------------------------
# Linear  (compyute/nn/functional/linear_funcs.py:11-35)
# --------------------------------------------------------------------------------------


def linear_forward(cache, x, w, b):
    """``LinearFn.forward`` (:15-23): ``x @ w.T (+ b)``."""
    y = x @ np.swapaxes(w, -1, -2)
    if b is not None:
        y += b
    cache.append((x, w, b is not None))
    return y


def linear_backward(cache, dy):
    """``LinearFn.backward`` (:26-35); ``.T`` = last-two-dims transpose (tensors.py:131-135)."""
    x, w, has_bias = cache.pop()
    dx = dy @ w
    dw = (np.swapaxes(dy, -1, -2) @ x).sum(tuple(range(dy.ndim - 2)))
    db = dy.sum(tuple(range(dy.ndim - 1))) if has_bias else None
    return dx, dw, db


# --------------------------------------------------------------------------------------
# Pooling  (compyute/nn/functional/pooling_funcs.py)
# --------------------------------------------------------------------------------------


def _upsample2d(x, scaling, target_shape):
    """``Upsample2DFn.forward`` (:20-34)."""
    y = repeat2d(x, scaling)
    if y.shape != tuple(target_shape):
        y = pad_to_shape(y, tuple(target_shape))
    return y


def maxpool2d_forward(cache, x, kernel_size: int):
    """``MaxPooling2DFn.forward`` (:71-76): stride = kernel, no padding, floor."""
    if x.ndim != 4:
        raise ValueError(f"Expected input to be 4D, got {x.ndim}D.")
    y = window_view2d(x, kernel_size, kernel_size).max((-2, -1))
    cache.append((x, kernel_size, y))
    return y


def maxpool2d_backward(cache, dy):
    """``MaxPooling2DFn.backward`` (:79-82): equality mask -- every tied maximum gets dy."""
    x, k, y = cache.pop()
    mask = _upsample2d(y, k, x.shape) == x
    return _upsample2d(dy, k, x.shape) * mask


def avgpool2d_forward(cache, x, kernel_size: int):
    """``AvgPooling2DFn.forward`` (:111-116)."""
    if x.ndim != 4:
        raise ValueError(f"Expected input to be 4D, got {x.ndim}D.")
    y = window_view2d(x, kernel_size, kernel_size).mean((-2, -1))
    cache.append((x.shape, kernel_size))
    return y


def avgpool2d_backward(cache, dy):
    """``AvgPooling2DFn.backward`` (:119-121)."""
    x_shape, k = cache.pop()
    return _upsample2d(dy / (k * k), k, x_shape)


# --------------------------------------------------------------------------------------
# BatchNorm  (compyute/nn/functional/normalization_funcs.py:10-177)
# --------------------------------------------------------------------------------------


def batchnorm_forward(cache, x, rmean, rvar, w, b, m: float, eps: float, training: bool):
    """``BatchNorm1DFn.forward`` (:14-52) for 2-D/3-D x, ``BatchNorm2DFn.forward`` (:124-159) for 4-D."""
    if x.ndim not in (2, 3, 4):
        raise ValueError(f"Expected 2D/3D/4D input, got {x.ndim}D.")
    dims = {2: (0,), 3: (0, 2), 4: (0, 2, 3)}[x.ndim]
    tail = (1,) * (x.ndim - 2)
    if training:
        mean = x.mean(dims, keepdims=True)
        std = np.sqrt(x.var(dims, keepdims=True) + eps)
        x_norm = (x - mean) / std
        rmean = rmean * (1 - m) + mean.squeeze() * m
        rvar = rvar * (1 - m) + x.var(dims, ddof=1) * m  # unbiased for the running stat
    else:
        mean = rmean.reshape(*rmean.shape, *tail)
        std = np.sqrt(rvar.reshape(*rvar.shape, *tail) + eps)
        x_norm = (x - mean) / std
    w_ = w.reshape(*w.shape, *tail)
    y = w_ * x_norm + b.reshape(*b.shape, *tail)
    cache.append((w_, dims, std, x_norm))
    return y, rmean, rvar


def batchnorm_backward(cache, dy):
    """``BatchNorm{1,2}DFn.backward`` (:54-70, :161-177)."""
    w_, dims, std, x_norm = cache.pop()
    n = float(dy.size / dy.shape[1])
    dy_sum = dy.sum(dims, keepdims=True)
    dy_xn_sum = (dy * x_norm).sum(dims, keepdims=True)
    dx = w_ / (std * n) * (n * dy - dy_sum - x_norm * dy_xn_sum)
    return dx, dy_xn_sum.squeeze(), dy_sum.squeeze()


# --------------------------------------------------------------------------------------
# ReLU / Dropout / Flatten / CE  (activation_funcs.py:22-34, regularization_funcs.py:11-32,
#                                 shape_funcs.py:8-19, loss_funcs.py:53-69)
# --------------------------------------------------------------------------------------


def relu_forward(cache, x):
    y = np.maximum(x, 0.0)
    cache.append((y > 0.0,))  # mask from the *output* → grad 0 at x == 0
    return y


def relu_backward(cache, dy):
    (mask,) = cache.pop()
    return dy * mask


def dropout_forward(cache, x, p: float, training: bool, mask=None):
    """``DropoutFn.forward``; ``mask`` (int8 Bernoulli(1-p)) may be injected for determinism."""
    if not training or p == 0.0:
        cache.append((False, p, None))
        return x
    keep = 1.0 - p
    if mask is None:
        mask = (np.random.random(x.shape) < keep).astype(np.int8)  # random.py bernoulli
    y = x * mask / keep
    cache.append((True, keep, mask))
    return y


def dropout_backward(cache, dy):
    training, keep, mask = cache.pop()
    if not training:
        return dy
    return dy * mask / keep


def softmax(x):
    """activation_funcs.py ``SoftmaxFn.forward``: exp(x - max) / sum."""
    e = np.exp(x - x.max(-1, keepdims=True))
    return e / e.sum(-1, keepdims=True)


def cross_entropy_forward(cache, logits, targets, eta: float = 1e-8):
    """``CrossEntropyLossFn.forward`` (loss_funcs.py:57-64); one-hot = identity(n)[t]."""
    probs = softmax(logits)
    onehot = np.identity(logits.shape[-1], dtype=probs.dtype)[targets]
    loss = -(np.log(probs + eta) * onehot).sum(-1).mean()
    cache.append((onehot, probs))
    return loss


def cross_entropy_backward(cache):
    """``CrossEntropyLossFn.backward`` (:67-69): (p - t) / B_local."""
    onehot, probs = cache.pop()
    return (probs - onehot) / float(math.prod(onehot.shape[:-1]))


# --------------------------------------------------------------------------------------
# Optimizers  (compyute/nn/optimizers.py)
# --------------------------------------------------------------------------------------


class SGD:
    """``SGD.step`` (:152-176).  ``params``/``grads`` are lists of ndarrays updated in place."""

    def __init__(self, lr=1e-3, momentum=0.0, nesterov=False, weight_decay=0.0):
        self.lr, self.momentum, self.nesterov, self.weight_decay = lr, momentum, nesterov, weight_decay
        self.t = 1
        self.state: dict[int, dict[str, np.ndarray]] = {}

    def step(self, params, grads):
        for i, (p, g0) in enumerate(zip(params, grads)):
            if g0 is None:
                continue
            st = self.state.setdefault(i, {})
            g = g0.copy()
            if self.weight_decay > 0.0:
                g += self.weight_decay * p
            if self.momentum > 0.0:
                v = self.momentum * st.get("v", 0.0) + g
                st["v"] = v
                g = g + self.momentum * v if self.nesterov else v
            p -= self.lr * g
        self.t += 1


class Adam:
    """``Adam.step`` (:241-271); ``decoupled=True`` gives ``AdamW.step`` (:335-362)."""

    def __init__(self, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, decoupled=False):
        self.lr, self.beta1, self.beta2, self.eps = lr, beta1, beta2, eps
        self.weight_decay, self.decoupled = weight_decay, decoupled
        self.t = 1
        self.state: dict[int, dict[str, np.ndarray]] = {}

    def step(self, params, grads):
        m_div = 1.0 - self.beta1**self.t
        v_div = 1.0 - self.beta2**self.t
        for i, (p, g) in enumerate(zip(params, grads)):
            if g is None:
                continue
            st = self.state.setdefault(i, {})
            if self.decoupled:
                p *= 1.0 - self.lr * self.weight_decay  # optimizers.py:344
            elif self.weight_decay != 0.0:
                g = g + self.weight_decay * p
            m = self.beta1 * st.get("m", 0.0) + (1.0 - self.beta1) * g
            st["m"] = m.copy()
            v = self.beta2 * st.get("v", 0.0) + (1.0 - self.beta2) * g**2
            st["v"] = v.copy()
            m = m / m_div
            v = v / v_div
            p -= (self.lr * m / (np.sqrt(v) + self.eps)).astype(p.dtype)
        self.t += 1


class NAdam:
    """``NAdam.step`` (optimizers.py:437-475)."""

    def __init__(self, lr=2e-3, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, momentum_decay=4e-3):
        self.lr, self.beta1, self.beta2, self.eps = lr, beta1, beta2, eps
        self.weight_decay, self.momentum_decay = weight_decay, momentum_decay
        self.t, self.mu_prod = 1, 1.0
        self.state: dict[int, dict[str, np.ndarray]] = {}

    def step(self, params, grads):
        mu = self.beta1 * (1.0 - 0.5 * 0.96 ** (self.t * self.momentum_decay))
        mu_next = self.beta1 * (1.0 - 0.5 * 0.96 ** ((self.t + 1) * self.momentum_decay))
        self.mu_prod *= mu
        m_div = 1.0 - self.mu_prod * mu_next
        g_div = 1.0 - self.mu_prod
        v_div = 1.0 - self.beta2**self.t
        for i, (p, g) in enumerate(zip(params, grads)):
            if g is None:
                continue
            st = self.state.setdefault(i, {})
            if self.weight_decay != 0.0:
                g = g + self.weight_decay * p
            m = self.beta1 * st.get("m", 0.0) + (1.0 - self.beta1) * g
            st["m"] = m.copy()
            v = self.beta2 * st.get("v", 0.0) + (1.0 - self.beta2) * g**2
            st["v"] = v.copy()
            m = mu_next * m / m_div + (1.0 - mu) * g / g_div
            v = v / v_div
            p -= (self.lr * m / (np.sqrt(v) + self.eps)).astype(p.dtype)
        self.t += 1


# --------------------------------------------------------------------------------------
# Direct-form equations -- an independent second statement, small cases only
# --------------------------------------------------------------------------------------


def conv2d_direct(x, f, b, padding, stride, dilation):
    """y[b,o,p,q] = bias[o] + Σ xp[b,i,p·s+j·d,q·s+k·d]·f[o,i,j,k]; fp64 loops, tiny inputs only."""
    B, Ci, H, W = x.shape
    Co, _, K, _ = f.shape
    Ho = (H + 2 * padding - dilation * (K - 1) - 1) // stride + 1
    Wo = (W + 2 * padding - dilation * (K - 1) - 1) // stride + 1
    xp = np.pad(x.astype(np.float64), ((0, 0), (0, 0), (padding, padding), (padding, padding)))
    y = np.zeros((B, Co, Ho, Wo))
    for j in range(K):
        for k in range(K):
            patch = xp[:, :, j * dilation : j * dilation + (Ho - 1) * stride + 1 : stride,
                       k * dilation : k * dilation + (Wo - 1) * stride + 1 : stride]
            y += np.einsum("bipq,oi->bopq", patch, f[:, :, j, k].astype(np.float64))
    if b is not None:
        y += b.reshape(1, -1, 1, 1)
    return y


def conv2d_direct_backward(x, f, dy, padding, stride, dilation):
    """dgrad / wgrad / bgrad direct forms, fp64, tiny inputs only."""
    B, Ci, H, W = x.shape
    Co, _, K, _ = f.shape
    _, _, Ho, Wo = dy.shape
    xp = np.pad(x.astype(np.float64), ((0, 0), (0, 0), (padding, padding), (padding, padding)))
    dxp = np.zeros_like(xp)
    df = np.zeros(f.shape)
    g = dy.astype(np.float64)
    for j in range(K):
        for k in range(K):
            sl = (slice(None), slice(None),
                  slice(j * dilation, j * dilation + (Ho - 1) * stride + 1, stride),
                  slice(k * dilation, k * dilation + (Wo - 1) * stride + 1, stride))
            dxp[sl] += np.einsum("bopq,oi->bipq", g, f[:, :, j, k].astype(np.float64))
            df[:, :, j, k] = np.einsum("bopq,bipq->oi", g, xp[sl])
    dx = dxp[:, :, padding : padding + H, padding : padding + W]
    return dx, df, g.sum((0, 2, 3))
