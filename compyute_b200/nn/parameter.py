"""Parameter / Buffer.  Mirrors compyute/nn/parameter.py:9-41."""

from __future__ import annotations

import numpy as np

from ..tensors import Tensor

__all__ = ["Parameter", "Buffer"]


class Parameter(Tensor):
    """Trainable tensor; must be floating point (parameter.py:25-28)."""

    def __init__(self, data: Tensor) -> None:
        if not np.issubdtype(np.dtype(data.dtype), np.floating):
            raise TypeError("Invalid data type for parameter. Must be float.")
        super().__init__(data.data)
        self.grad_slot = None  # DeviceArray view into a flat gradient arena (data-parallel mode), else None

    def __getstate__(self):
        """Checkpoints (``cp.save(model.get_state_dict())``, utils.py:44-73) carry the value only: the gradient and its arena
        slot belong to the running optimizer (a pickled slot would be a detached host copy of the whole gradient)."""
        state = dict(self.__dict__)
        state["grad_slot"] = None
        state["grad"] = None
        return state


class Buffer(Tensor):
    """Non-trainable state (running statistics)."""

    def __init__(self, data: Tensor) -> None:
        super().__init__(data.data)
