"""compyute/typing.py of the reference: dtype names.  Device tensors are NumPy-typed here, so the names are NumPy dtypes."""

import numpy as _np

bool_ = _np.bool_
int8, int16, int32, int64 = _np.int8, _np.int16, _np.int32, _np.int64
float16, float32, float64 = _np.float16, _np.float32, _np.float64
uint8 = _np.uint8
DType = type(_np.float32)
ScalarLike = (int, float)

__all__ = ["bool_", "int8", "int16", "int32", "int64", "float16", "float32", "float64", "uint8"]
