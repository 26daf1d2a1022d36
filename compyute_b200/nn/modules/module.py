"""Module base class.  Public behaviour of compyute/nn/modules/module.py:24-400 for the in-scope surface:
``fcache``, ``training()/inference()``, the ``__setattr__`` registry, parameters / buffers / state dict,
``register_forward/backward`` wrappers and ``update_parameter_grad``.

Difference from the reference, on purpose: the wrappers' NaN assert (module.py:332, 366) costs a full reduction and
a host sync per layer on a GPU; it runs only in debug mode here (``COMPYUTE_DEBUG=1`` / ``set_debug_mode``), where
it uses the ``cpt_isnan_flag`` kernel.
"""

from __future__ import annotations

import os
import time
from collections import OrderedDict
from functools import wraps
from typing import Any, Callable, Iterable, Iterator, Optional

from ...backend import Device, DeviceError
from ...tensors import DeviceArray, Tensor
from ..functional.functions import FunctionCache, PseudoCache
from ..parameter import Buffer, Parameter

__all__ = ["Module", "ModuleList", "Identity", "set_debug_mode", "get_debug_mode"]

_debug = bool(os.environ.get("COMPYUTE_DEBUG", False))


def set_debug_mode(active: bool) -> None:
    """compyute/utils.py:19-27"""
    global _debug
    _debug = bool(active)


def get_debug_mode() -> bool:
    return _debug


def _has_nan(t: Optional[Tensor]) -> bool:
    if t is None:
        return False
    if isinstance(t.data, DeviceArray):
        return t.data.has_nan()
    import numpy as np
    return bool(np.isnan(t.data).any())


class Module:
    """Neural network base module."""

    def __init__(self, label: Optional[str] = None) -> None:
        self.label = label or self.__class__.__name__
        self.fcache = FunctionCache()
        self.x: Optional[Tensor] = None
        self.y: Optional[Tensor] = None
        self._is_training = True
        self._retain_values = False
        self._trainable = True
        self._parameters: OrderedDict[str, Parameter] = OrderedDict()
        self._buffers: OrderedDict[str, Buffer] = OrderedDict()
        self._modules: OrderedDict[str, Module] = OrderedDict()

    # ---- device / mode -----------------------------------------------------------------------
    @property
    def device(self) -> Device:
        try:
            return next(self.get_parameters()).device
        except StopIteration:
            raise ValueError("Module has no parameters.")

    def to_device(self, device: Device) -> None:
        """Moves every Tensor attribute, then recurses (module.py:57-70)."""
        for t in vars(self).values():
            if isinstance(t, Tensor):
                t.ito_device(device)
        for m in self.get_modules(recursive=False):
            m.to_device(device)

    @property
    def retain_values(self) -> bool:
        return self._retain_values

    @retain_values.setter
    def retain_values(self, value: bool) -> None:
        self._retain_values = value
        for m in self.get_modules(recursive=False):
            m.retain_values = value

    @property
    def trainable(self) -> bool:
        return self._trainable

    @trainable.setter
    def trainable(self, value: bool) -> None:
        self._trainable = value
        for m in self.get_modules(recursive=False):
            m.trainable = value

    @property
    def is_training(self) -> bool:
        return self._is_training

    def training(self) -> None:
        """Training mode: a real FunctionCache (module.py:122-128)."""
        self._is_training = True
        self.fcache = FunctionCache()
        for m in self.get_modules(recursive=False):
            m.training()

    def inference(self) -> None:
        """Inference mode: a PseudoCache that drops pushes (module.py:130-136)."""
        self._is_training = False
        self.fcache = PseudoCache()
        for m in self.get_modules(recursive=False):
            m.inference()

    # ---- registry ----------------------------------------------------------------------------
    def __setattr__(self, name: str, value: Any) -> None:
        if isinstance(value, Parameter):
            self._parameters[name] = value
        elif isinstance(value, Buffer):
            self._buffers[name] = value
        elif isinstance(value, Module):
            self._modules[name] = value
        elif isinstance(value, ModuleList):
            for i, m in enumerate(value):
                self._modules[f"{name}.{i}"] = m
        super().__setattr__(name, value)

    def __bool__(self) -> bool:
        return True

    def __repr__(self) -> str:
        skip = {"label", "fcache", "x", "y"}
        attrs = [f"{k}={v}" for k, v in vars(self).items()
                 if not k.startswith("_") and k not in skip and not isinstance(v, (Tensor, Module, ModuleList)) and v is not None]
        s = f"{self.label}(" + ", ".join(attrs) + ")"
        for m in self.get_modules(recursive=False):
            s += "\n" + repr(m)
        return s

    @property
    def n_modules(self) -> int:
        return len(self._modules)

    def get_modules(self, recursive: bool = True) -> Iterator["Module"]:
        for m in self._modules.values():
            yield m
            if recursive:
                yield from m.get_modules()

    def get_parameters(self, recursive: bool = True) -> Iterator[Parameter]:
        yield from self._parameters.values()
        if recursive:
            for m in self.get_modules():
                yield from m.get_parameters(recursive=False)

    def get_buffers(self, recursive: bool = True) -> Iterator[Buffer]:
        yield from self._buffers.values()
        if recursive:
            for m in self.get_modules():
                yield from m.get_buffers(recursive=False)

    def get_state_dict(self) -> "OrderedDict[str, Tensor]":
        """Own parameters, own buffers, then children prefixed ``<attr>.`` (module.py:224-248)."""
        sd: OrderedDict[str, Tensor] = OrderedDict()
        sd.update(self._parameters)
        sd.update(self._buffers)
        for name, m in self._modules.items():
            for k, v in m.get_state_dict().items():
                sd[f"{name}.{k}"] = v
        return sd

    def load_state_dict(self, state_dict: "OrderedDict[str, Tensor]") -> None:
        """Ordered, key-checked, device-checked; rebinds ``.data`` (module.py:250-273)."""
        for (k1, v1), (k2, v2) in zip(self.get_state_dict().items(), state_dict.items()):
            if k1 != k2:
                raise ValueError(f"State dict key mismatch: {k1} != {k2}")
            if v1.device != v2.device:
                raise DeviceError(f"Device mismatch. Module device: {v1.device}, state dict device: {v2.device}")
            v1.data = v2.data
            v1.grad = v2.grad

    # ---- call protocol -----------------------------------------------------------------------
    def __call__(self, x: Tensor) -> Tensor:
        return self.forward(x)

    def forward(self, x: Tensor) -> Tensor:
        raise NotImplementedError

    def backward(self, dy: Tensor) -> Tensor:
        raise NotImplementedError

    @staticmethod
    def register_forward(fwd_fn: Callable) -> Callable:
        """Forward wrapper: optional debug timing, retain_values, NaN check in debug mode (module.py:309-335)."""

        @wraps(fwd_fn)
        def wrapper(m: "Module", x: Tensor) -> Tensor:
            if _debug:
                t0 = time.perf_counter()
                y = fwd_fn(m, x)
                from ...backend import synchronize
                synchronize()
                print(f"{m.label:20s} | fwd | {str(x.dtype):15s} | {str(y.dtype):15s} | dt={(time.perf_counter() - t0) * 1e3:>10.4f} ms")
                assert not _has_nan(y), repr(m)
            else:
                y = fwd_fn(m, x)
            if m.retain_values:
                m.x, m.y = x, y
            return y

        return wrapper

    @staticmethod
    def register_backward(bwd_fn: Callable) -> Callable:
        """Backward wrapper: training-mode check, debug timing, retain grads (module.py:337-369)."""

        @wraps(bwd_fn)
        def wrapper(m: "Module", dy: Tensor) -> Tensor:
            if not m.is_training:
                raise AttributeError(f"{m.label} is not in training mode.")
            if _debug:
                t0 = time.perf_counter()
                dx = bwd_fn(m, dy)
                from ...backend import synchronize
                synchronize()
                print(f"{m.label:20s} | bwd | {str(dy.dtype):15s} | dt={(time.perf_counter() - t0) * 1e3:>10.4f} ms")
                assert not _has_nan(dx)
            else:
                dx = bwd_fn(m, dy)
            if m.retain_values and m.x is not None and m.y is not None:
                m.x.grad, m.y.grad = dx, dy
            return dx

        return wrapper

    def clean(self, force: bool = False) -> None:
        """Drops cached intermediates and gradients (module.py:371-390)."""
        self.fcache.cache.clear()
        if not self._retain_values or force:
            self.x = self.y = None
            for p in self.get_parameters(recursive=False):
                p.grad = None
        for m in self.get_modules(recursive=False):
            m.clean(force)

    def update_parameter_grad(self, parameter: Optional[Parameter], grad: Optional[Tensor]) -> None:
        """First gradient is stored by reference, later ones accumulate with ``+=`` (module.py:392-400)."""
        if self.trainable and parameter is not None and grad is not None:
            first = parameter.grad is None
            if first:
                parameter.grad = grad
            else:
                parameter.grad += grad
            if getattr(parameter, "grad_slot", None) is not None:
                from ..optimizers import notify_grad_written
                notify_grad_written(parameter, first)  # overlapped data-parallel exchange (Optimizer.overlap_grad_sync)

    @staticmethod
    def grad_slot(parameter: Optional[Parameter]):
        """Arena slot to write this parameter's gradient into directly, or None.  Only used for the FIRST gradient of
        a step (``p.grad is None``): a shared parameter's later contributions go through ``+=`` as in the reference."""
        if parameter is None or parameter.grad is not None:
            return None
        return getattr(parameter, "grad_slot", None)


class Identity(Module):
    def forward(self, x: Tensor) -> Tensor:
        return x

    def backward(self, dy: Tensor) -> Tensor:
        return dy


class ModuleList(list):
    """List of modules registered as ``<attr>.<i>`` (module.py ModuleList)."""

    def __init__(self, modules: Iterable[Module]) -> None:
        super().__init__(modules)
