"""compyute_b200.nn — modules, functional, optimizers, losses of the CNN hot path (same names as compyute.nn)."""

from . import functional, optimizers, utils
from .losses import *
from .modules import *
from .parameter import *
