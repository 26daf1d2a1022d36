"""compyute/nn/functional of the reference."""

from compyute_b200.nn.functional import *  # noqa: F401,F403
