"""compyute/tensors.py of the reference."""

from typing import Union

from compyute_b200.tensors import DeviceArray, ShapeError, Tensor, tensor  # noqa: F401

ShapeLike = tuple[int, ...]
AxisLike = Union[int, tuple[int, ...]]
