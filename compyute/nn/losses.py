"""compyute/nn/losses.py of the reference."""

from compyute_b200.nn.losses import *  # noqa: F401,F403
