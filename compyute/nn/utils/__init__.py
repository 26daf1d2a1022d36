"""compyute/nn/utils of the reference (dataloaders, lr schedulers, training utilities)."""

from compyute_b200.nn.utils import *  # noqa: F401,F403
