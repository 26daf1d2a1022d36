"""compyute/nn/modules of the reference."""

from compyute_b200.nn.modules import *  # noqa: F401,F403

from . import activations, containers, convolutions, linear, module, normalizations, poolings, regularizations, reshapes  # noqa: F401,E402
