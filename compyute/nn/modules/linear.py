"""compyute/nn/modules/linear.py of the reference: in-scope layers from compyute_b200, the rest as raising stubs."""

from compyute_b200.nn import Linear  # noqa: F401

from .._stubs import *  # noqa: F401,F403
