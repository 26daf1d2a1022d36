"""compyute/random/random.py of the reference: ``seed`` / ``set_seed`` and the tensor constructors."""

from compyute_b200.random import *  # noqa: F401,F403
from compyute_b200.random import bernoulli, normal, permutation, random, seed, set_seed, shuffle, uniform, uniform_int  # noqa: F401
