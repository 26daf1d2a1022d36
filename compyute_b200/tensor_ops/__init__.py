"""``compyute.tensor_ops`` for tensors on ``cuda``: the reference's function names and signatures
(compyute/tensor_ops/{creation,unary,reduction,selection,shape,multiary}_ops.py) over the kernels of csrc/tensor_ops.cu.

On ``cpu`` tensors (host staging) the functions evaluate with NumPy exactly like the reference does; on ``cuda`` tensors
every call is one or a few launches of our library — no CuPy.  Functions of the reference that are not on or near the
CNN training path (FFT, einsum, topk, unique, tril/triu, histogram, complex parts) raise ``NotImplementedError`` on device
tensors."""

from .ops import *  # noqa: F401,F403
from .ops import __all__  # noqa: F401
