"""CPU oracle, model level: interprets a bench_workloads spec with the function-level oracle (oracle/compyute_ref.py),
reproducing what the reference's Sequential / ResidualConnection / BatchNorm modules + CrossEntropyLoss + Adam do
(compyute/nn/modules/containers.py:35-45, 153-162; normalizations.py:150-171; module.py:392-400; trainer.py:118-135).

TEST INFRASTRUCTURE ONLY (see compyute_ref.py).  Parameters are kept in the flat order of ``Module.get_parameters()``
(own parameters, then children depth-first: residual block before projection), so they can be exchanged with a
compyute_b200 model one-to-one."""

from __future__ import annotations

import numpy as np

from . import compyute_ref as R


def round_operand(a: np.ndarray, how: str) -> np.ndarray:
    """Operand rounding of the tensor-core compute modes, emulated on fp32 bits: "bf16" = round to nearest even to 8
    significand bits (cvt.rn.bf16), "tf32_rna" = round to nearest, ties away, to 11 bits (cvt.rna.tf32, the staging kernels of
    the convolutions), "tf32_trunc" = the low 13 bits ignored (what tcgen05 kind::tf32 does with raw fp32 bits: Linear layers)."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    if how == "bf16":
        u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    elif how == "tf32_rna":
        u = (u + 0x1000) & 0xFFFFE000
    elif how == "tf32_trunc":
        u = u & 0xFFFFE000
    else:
        raise ValueError(how)
    return u.astype(np.uint32).view(np.float32).reshape(a.shape)


_ROUNDING = {None: (None, None), "bf16": ("bf16", "bf16"), "tf32": ("tf32_rna", "tf32_trunc")}  # mode -> (conv, linear)


class _Layer:
    rounding = None  # compute mode whose operand rounding the contractions emulate (None = the reference's fp32)

    def _rnd(self, a):
        how = _ROUNDING[self.rounding][0 if self.kind == "conv" else 1]
        return a if how is None else round_operand(a, how)

    def __init__(self, spec, rng_params):
        self.spec = spec
        self.kind = spec[0]
        self.params: list[np.ndarray] = []   # trainable, in the module's registration order (w, b)
        self.buffers: list[np.ndarray] = []  # rmean, rvar
        self.grads: list = []
        self.cache: list = []
        self.block = self.proj = None
        if self.kind == "residual":
            self.block = [_Layer(s, rng_params) for s in spec[1]]
            self.proj = [_Layer(s, rng_params) for s in spec[2]] if spec[2] else None

    def leaves(self):
        if self.kind == "residual":
            for l in self.block:
                yield from l.leaves()
            if self.proj:
                for l in self.proj:
                    yield from l.leaves()
        else:
            yield self

    def forward(self, x, training):
        k, s, c = self.kind, self.spec, self.cache
        if k == "conv":  # the cache keeps the (rounded) operands: backward contracts those, like the device path
            return R.conv2d_forward(c, self._rnd(x), self._rnd(self.params[0]), self.params[1] if s[6] else None, s[4], s[5], 1)
        if k == "linear":
            return R.linear_forward(c, self._rnd(x), self._rnd(self.params[0]), self.params[1] if s[3] else None)
        if k in ("bn2d", "bn1d"):
            y, rm, rv = R.batchnorm_forward(c, x, self.buffers[0], self.buffers[1], self.params[0], self.params[1], 0.1, 1e-5, training)
            self.buffers = [rm, rv]
            return y
        if k == "relu":
            return R.relu_forward(c, x)
        if k == "maxpool":
            return R.maxpool2d_forward(c, x, s[1])
        if k == "avgpool":
            return R.avgpool2d_forward(c, x, s[1])
        if k == "dropout":
            return R.dropout_forward(c, x, s[1], training)
        if k == "flatten":
            c.append(x.shape)
            return x.reshape(x.shape[0], -1)
        if k == "residual":
            y = x
            for l in self.block:
                y = l.forward(y, training)
            r = x
            if self.proj:
                for l in self.proj:
                    r = l.forward(r, training)
            return y + r
        raise ValueError(k)

    def backward(self, dy):
        k, s, c = self.kind, self.spec, self.cache
        if k == "conv":
            dx, dw, db = R.conv2d_backward(c, self._rnd(dy))
            if self.rounding and s[6]:
                db = dy.sum((0, 2, 3))  # the bias gradient is an fp32 reduction of the unrounded dy in every mode
            self.grads = [dw] + ([db] if s[6] else [])
            return dx
        if k == "linear":
            dx, dw, db = R.linear_backward(c, self._rnd(dy))
            if self.rounding and s[3]:
                db = dy.reshape(-1, dy.shape[-1]).sum(0)
            self.grads = [dw] + ([db] if s[3] else [])
            return dx
        if k in ("bn2d", "bn1d"):
            dx, dw, db = R.batchnorm_backward(c, dy)
            self.grads = [dw, db]
            return dx
        if k == "relu":
            return R.relu_backward(c, dy)
        if k == "maxpool":
            return R.maxpool2d_backward(c, dy)
        if k == "avgpool":
            return R.avgpool2d_backward(c, dy)
        if k == "dropout":
            return R.dropout_backward(c, dy)
        if k == "flatten":
            return dy.reshape(c.pop())
        if k == "residual":
            dx = dy
            for l in reversed(self.block):
                dx = l.backward(dx)
            r = dy
            if self.proj:
                for l in reversed(self.proj):
                    r = l.backward(r)
            return dx + r
        raise ValueError(k)


class RefModel:
    """Sequential model over a spec.  ``params`` / ``buffers``: flat lists in get_parameters() / get_buffers() order."""

    def __init__(self, spec, params, buffers, operand_rounding=None):
        """``operand_rounding`` = "bf16" / "tf32": Conv2D / Linear round their operands (x, w, dy) like the device's tensor-core
        modes before contracting in fp32 — the oracle for model-level parity in those modes (decisions such as ReLU masks and
        max-pool winners then agree, so the comparison is tight); None = the reference's fp32 arithmetic."""
        self.layers = [_Layer(s, None) for s in spec]
        for l in self.leaves():
            l.rounding = operand_rounding
        pi, bi = iter(params), iter(buffers)
        for l in self.leaves():
            s = l.spec
            if l.kind == "conv":
                l.params = [next(pi)] + ([next(pi)] if s[6] else [])
            elif l.kind == "linear":
                l.params = [next(pi)] + ([next(pi)] if s[3] else [])
            elif l.kind in ("bn2d", "bn1d"):
                l.params = [next(pi), next(pi)]
                l.buffers = [next(bi), next(bi)]

    def leaves(self):
        for l in self.layers:
            yield from l.leaves()

    def parameters(self):
        return [p for l in self.leaves() for p in l.params]

    def buffers(self):
        return [b for l in self.leaves() for b in l.buffers]

    def gradients(self):
        return [g for l in self.leaves() if l.params for g in l.grads]

    def forward(self, x, training=True):
        for l in self.layers:
            x = l.forward(x, training)
        return x

    def backward(self, dy):
        for l in reversed(self.layers):
            dy = l.backward(dy)
        return dy

    def train_steps(self, x, t, steps, lr=1e-3, optimizer="adam", **kw):
        """README.md:170-183 loop with CrossEntropyLoss + Adam / SGD; returns the loss trace."""
        opt = R.Adam(lr=lr, **kw) if optimizer == "adam" else R.SGD(lr=lr, **kw)
        losses = []
        for _ in range(steps):
            lc = []
            logits = self.forward(x, True)
            losses.append(float(R.cross_entropy_forward(lc, logits, t)))
            self.backward(R.cross_entropy_backward(lc))
            opt.step(self.parameters(), self.gradients())
        return losses
