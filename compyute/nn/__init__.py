"""compyute/nn of the reference: modules, functional, losses, optimizers, parameters, utils."""

from compyute_b200.nn import *  # noqa: F401,F403
from compyute_b200.nn import functional, optimizers  # noqa: F401
from compyute_b200.nn.parameter import Buffer, Parameter  # noqa: F401

from . import modules, parameter, utils  # noqa: F401,E402
from ._stubs import *  # noqa: F401,F403,E402
