"""compyute/nn/optimizers.py of the reference."""

from compyute_b200.nn.optimizers import *  # noqa: F401,F403
from compyute_b200.nn.optimizers import SGD, Adam, AdamW, NAdam, Optimizer  # noqa: F401
