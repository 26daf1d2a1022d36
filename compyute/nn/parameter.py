"""compyute/nn/parameter.py of the reference."""

from compyute_b200.nn.parameter import Buffer, Parameter  # noqa: F401
