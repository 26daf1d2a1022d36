"""Containers.  Sequential (compyute/nn/modules/containers.py:14-45) and ResidualConnection (:120-162)."""

from __future__ import annotations

from typing import Optional

from ...tensors import DeviceArray, Tensor
from ..functional.activation_funcs import FUSED_INTO_PRODUCER
from ..functional.normalization_funcs import pool2_fusion_supported, residual_tail_supported
from .module import Module, ModuleList, get_debug_mode

_fusion = True
_epilogue_stats = True


def set_epilogue_stats_enabled(enabled: bool) -> None:
    """BatchNorm batch statistics from the producing convolution's epilogue (tensor-core modes).  Unlike the other
    peepholes this changes the summation order / variance formula of the statistics (not bit-identical to the two-pass
    computation, well inside the tf32 / bf16 tolerances); on by default."""
    global _epilogue_stats
    _epilogue_stats = bool(enabled)


def get_epilogue_stats_enabled() -> bool:
    return _epilogue_stats


_conv_relu = True


def set_conv_relu_fusion_enabled(enabled: bool) -> None:
    """Conv2D -> ReLU pairs from the convolution's epilogue (``Sequential._conv_relu_fusable``).  On by default; a separate
    switch so that its effect can be measured on its own (``bench.py --no-conv-relu-fusion``)."""
    global _conv_relu
    _conv_relu = bool(enabled)


def get_conv_relu_fusion_enabled() -> bool:
    return _conv_relu


def set_fusion_enabled(enabled: bool) -> None:
    """Peephole fusion inside ``Sequential`` (BatchNorm -> ReLU in one pass).  On by default; results are identical to the
    unfused layers (the mask is recomputed from the forward's own expression), it only removes memory passes."""
    global _fusion
    _fusion = bool(enabled)


def get_fusion_enabled() -> bool:
    return _fusion

__all__ = ["Sequential", "ResidualConnection", "EmptyContainerError", "set_fusion_enabled", "get_fusion_enabled",
           "set_epilogue_stats_enabled", "get_epilogue_stats_enabled", "set_conv_relu_fusion_enabled", "get_conv_relu_fusion_enabled"]


class EmptyContainerError(Exception):
    """Container built without modules (containers.py:165)."""


class Sequential(Module):
    """y = f_n(... f_1(x)); backward walks the layers in reverse."""

    def __init__(self, *modules: Module, label: Optional[str] = None) -> None:
        super().__init__(label)
        if not modules:
            raise EmptyContainerError()
        self.layers = ModuleList(modules)

    def _plan_staging_hints(self) -> None:
        """Tells every BatchNorm2D of this container whether its neighbours are convolutions, so that it writes their
        channels-last bf16 operand in its own apply pass (forward: the consumer of its — possibly ReLU-fused — output;
        backward: the convolution that produced its input).  Recomputed when the layer list or the fusion switch changes."""
        from .layers import BatchNorm2D, Conv2D, Linear, ReLU
        key = (tuple(id(m) for m in self.layers), _fusion, _epilogue_stats)
        if getattr(self, "_hint_key", None) == key:
            return
        object.__setattr__(self, "_hint_key", key)

        def wants_cl(m) -> bool:  # does m feed its input straight into a Conv2D (and into nothing that mutates it)?
            if isinstance(m, Conv2D):
                return True
            if isinstance(m, Sequential):
                return wants_cl(m.layers[0])
            if isinstance(m, ResidualConnection):
                return wants_cl(m.residual_block) and (m.residual_proj is None or wants_cl(m.residual_proj))
            return False

        n = len(self.layers)
        for i, m in enumerate(self.layers):
            if isinstance(m, Conv2D):  # its epilogue sums the batch statistics of a BatchNorm2D that follows
                m._emit_stats = _fusion and _epilogue_stats and i + 1 < n and isinstance(self.layers[i + 1], BatchNorm2D)
            if type(m) is ReLU:  # ReLU between Linear layers writes their bf16 operands (x of the next, dy of the previous)
                m._emit_lp_fwd = _fusion and i + 1 < n and isinstance(self.layers[i + 1], Linear)
                m._emit_lp_bwd = _fusion and i > 0 and isinstance(self.layers[i - 1], Linear)
            if not isinstance(m, BatchNorm2D):
                continue
            j = i + 1
            if j < n and type(self.layers[j]) is ReLU:
                j = j + 1 if _fusion else n  # unfused: the BatchNorm output feeds the ReLU, not the convolution
            m._emit_cl_fwd = _fusion and j < n and wants_cl(self.layers[j])
            prev = self.layers[i - 1] if i > 0 else None
            m._emit_cl_bwd = _fusion and isinstance(prev, Conv2D)
            m._emit_cl_bwd_sum = bool(m._emit_cl_bwd and prev.b is not None)

    def _fusable(self, i: int, x: Tensor) -> bool:
        """layers[i] is a BatchNorm directly followed by a ReLU and nothing observes the tensor between them."""
        from .layers import ReLU, _BatchNorm
        if not _fusion or i + 1 >= len(self.layers) or get_debug_mode():
            return False
        a, b = self.layers[i], self.layers[i + 1]
        return (isinstance(a, _BatchNorm) and type(b) is ReLU and not a.retain_values and not b.retain_values
                and a.is_training == b.is_training and isinstance(x.data, DeviceArray))

    def _residual_fusable(self, i: int, x: Tensor) -> bool:
        """layers[i] is a ResidualConnection whose block ends in a BatchNorm2D, directly followed by a ReLU: the BatchNorm
        apply, the ``y += skip`` and the ReLU run as one pass (cpt_bn_add_relu_apply) — bit-identical, 16 B/element less traffic."""
        from .layers import BatchNorm2D, ReLU
        if not _fusion or i + 1 >= len(self.layers) or get_debug_mode() or not isinstance(x.data, DeviceArray):
            return False
        rc, relu = self.layers[i], self.layers[i + 1]
        if not isinstance(rc, ResidualConnection) or type(relu) is not ReLU or not isinstance(rc.residual_block, Sequential):
            return False
        bn = rc.residual_block.layers[-1]
        mods = (rc, rc.residual_block, bn, relu)
        return (type(bn) is BatchNorm2D and not any(m.retain_values for m in mods)
                and len({m.is_training for m in mods}) == 1)

    def _pool_fusable(self, i: int, x: Tensor) -> bool:
        """layers[i : i + 3] is BatchNorm2D -> ReLU -> MaxPooling2D(2): one forward pass that writes only the pooled tensor, one
        backward pass pair that recomputes both masks from x (cpt_bn_relu_pool2_*) — bit-identical to the three layers."""
        from .layers import BatchNorm2D, MaxPooling2D, ReLU
        if not _fusion or i + 2 >= len(self.layers) or get_debug_mode() or not isinstance(x.data, DeviceArray):
            return False
        bn, relu, pool = self.layers[i], self.layers[i + 1], self.layers[i + 2]
        if type(bn) is not BatchNorm2D or type(relu) is not ReLU or type(pool) is not MaxPooling2D or pool.kernel_size != 2:
            return False
        mods = (bn, relu, pool)
        return (not any(m.retain_values for m in mods) and len({m.is_training for m in mods}) == 1
                and pool2_fusion_supported(x))

    def _linear_relu_fusable(self, i: int, x: Tensor) -> bool:
        """layers[i : i + 2] is Linear -> ReLU in bf16 mode with Out % 32 == 0: the ReLU (+ its mask, + the bf16 rows of a following
        Linear) comes out of the GEMM epilogue (cpt_linear_relu_fwd_bf16) — bit-identical, one pass over y less each way."""
        from ..functional.linear_funcs import LinearFn
        from .layers import Linear, ReLU
        if not _fusion or i + 1 >= len(self.layers) or get_debug_mode() or not isinstance(x.data, DeviceArray):
            return False
        lin, relu = self.layers[i], self.layers[i + 1]
        if type(lin) is not Linear or type(relu) is not ReLU:
            return False
        return (not lin.retain_values and not relu.retain_values and lin.is_training == relu.is_training
                and LinearFn.relu_fusable(x, lin.w))

    def _conv_relu_fusable(self, i: int, x: Tensor) -> bool:
        """layers[i : i + 2] is Conv2D -> ReLU on the tensor-core path, followed by another layer of this container: the ReLU
        comes out of the convolution's epilogue and its backward is folded into the staging of dy (cpt_conv2d_fprop_cl_relu,
        cpt_to_channels_last_gated) — bit-identical, 12 B/element less traffic.  The backward mask is the fused output itself
        (y > 0), so the pair must not end the container: an enclosing ResidualConnection adds the skip branch in place."""
        from ..functional.convolution_funcs import Conv2DFn
        from .layers import Conv2D, ReLU
        if not _fusion or not _conv_relu or i + 2 >= len(self.layers) or get_debug_mode() or not isinstance(x.data, DeviceArray):
            return False
        conv, relu = self.layers[i], self.layers[i + 1]
        if type(conv) is not Conv2D or type(relu) is not ReLU:
            return False
        return (not conv.retain_values and not relu.retain_values and conv.is_training == relu.is_training
                and Conv2DFn.relu_fusable(x, conv.w, conv.padding, conv.stride, conv.dilation))

    def _run(self, x: Tensor, tail=None) -> Tensor:
        """The layer walk.  ``tail = (skip_fn, relu)``: this container is the block of a fused residual connection — its last
        BatchNorm2D evaluates ``relu(bn(x) + skip_fn())``."""
        self._plan_staging_hints()
        i, n = 0, len(self.layers)
        while i < n:
            layer = self.layers[i]
            if tail is not None and i == n - 1:
                skip = tail[0]()
                if residual_tail_supported(x) and skip.shape == x.shape:
                    return layer.forward_add_relu(x, skip, tail[1])
                y = layer(x)  # shapes the fused kernel does not cover: the three separate passes
                y += skip
                return tail[1](y)
            if self._pool_fusable(i, x):
                x = layer.forward_relu_pool2(x)
                self.layers[i + 1].fcache.push(FUSED_INTO_PRODUCER)                            # ReLU.backward passes dy through
                self.layers[i + 2].fcache.push(FUSED_INTO_PRODUCER, self.layers[i + 2].kernel_size, None)  # so does MaxPooling2D.backward
                i += 3
            elif self._fusable(i, x):
                x = layer.forward_relu(x)
                self.layers[i + 1].fcache.push(FUSED_INTO_PRODUCER)  # its backward is folded into the BatchNorm's
                i += 2
            elif self._linear_relu_fusable(i, x):
                x = layer.forward_relu(x, self.layers[i + 1])
                i += 2
            elif self._conv_relu_fusable(i, x):
                x = layer.forward_relu(x)
                self.layers[i + 1].fcache.push(FUSED_INTO_PRODUCER)  # its backward is folded into the convolution's dy staging
                i += 2
            elif self._residual_fusable(i, x):
                x = layer.forward_relu(x, self.layers[i + 1])
                i += 2
            else:
                x = layer(x)
                i += 1
        return x

    @Module.register_forward
    def forward(self, x: Tensor) -> Tensor:
        return self._run(x)

    @Module.register_backward
    def backward(self, dy: Tensor) -> Tensor:
        from ..functional.activation_funcs import PlainMask
        from .layers import Linear, ReLU
        i = len(self.layers) - 1
        while i >= 0:
            layer = self.layers[i]
            prev = self.layers[i - 1] if i > 0 else None
            if (_fusion and type(layer) is Linear and type(prev) is ReLU and not get_debug_mode() and prev.is_training and layer.is_training
                    and not layer.retain_values and not prev.retain_values
                    and prev.fcache.cache and isinstance(prev.fcache.cache[-1][0], PlainMask)):
                dy = layer.backward_relu(dy, prev)  # the ReLU's dx * mask comes out of this layer's dgrad epilogue
                i -= 2
            else:
                dy = layer.backward(dy)  # a ReLU whose cache entry is FUSED_INTO_PRODUCER passes dy through
                i -= 1
        return dy


class ResidualConnection(Module):
    """y = f(x) + proj(x) (or + x); the adds are ``cpt_add_inplace`` launches (containers.py:153-162)."""

    def __init__(self, *modules: Module, residual_proj: Optional[Module] = None, label: Optional[str] = None) -> None:
        if not modules:
            raise EmptyContainerError()
        super().__init__(label)
        self.residual_block = modules[0] if len(modules) == 1 else Sequential(*modules)
        self.residual_proj = residual_proj

    @Module.register_forward
    def forward(self, x: Tensor) -> Tensor:
        y = self.residual_block(x)
        y += self.residual_proj(x) if self.residual_proj else x
        return y

    def forward_relu(self, x: Tensor, relu: Module) -> Tensor:
        """``relu(self(x))`` with the block's last BatchNorm2D, the residual add and the ReLU as one pass (called by the
        enclosing Sequential, see ``Sequential._residual_fusable``).  The skip branch is evaluated when the block reaches its
        last layer, i.e. in the reference's order (block first, then projection)."""
        return self.residual_block._run(x, tail=((lambda: self.residual_proj(x) if self.residual_proj else x), relu))

    @Module.register_backward
    def backward(self, dy: Tensor) -> Tensor:
        dx = self.residual_block.backward(dy)
        dx += self.residual_proj.backward(dy) if self.residual_proj else dy
        return dx
