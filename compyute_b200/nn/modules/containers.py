"""Containers.  Sequential (compyute/nn/modules/containers.py:14-45) and ResidualConnection (:120-162)."""

from __future__ import annotations

from typing import Optional

from ...tensors import Tensor
from .module import Module, ModuleList

__all__ = ["Sequential", "ResidualConnection", "EmptyContainerError"]


class EmptyContainerError(Exception):
    """Container built without modules (containers.py:165)."""


class Sequential(Module):
    """y = f_n(... f_1(x)); backward walks the layers in reverse."""

    def __init__(self, *modules: Module, label: Optional[str] = None) -> None:
        super().__init__(label)
        if not modules:
            raise EmptyContainerError()
        self.layers = ModuleList(modules)

    @Module.register_forward
    def forward(self, x: Tensor) -> Tensor:
        for layer in self.layers:
            x = layer(x)
        return x

    @Module.register_backward
    def backward(self, dy: Tensor) -> Tensor:
        for layer in reversed(self.layers):
            dy = layer.backward(dy)
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

    @Module.register_backward
    def backward(self, dy: Tensor) -> Tensor:
        dx = self.residual_block.backward(dy)
        dx += self.residual_proj.backward(dy) if self.residual_proj else dy
        return dx
