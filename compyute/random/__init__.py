"""compyute/random of the reference."""

from .random import *  # noqa: F401,F403
from . import random  # noqa: F401
