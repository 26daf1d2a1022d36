"""Import stub: the reference's normalizations_test imports torchtune for its (out-of-scope) RMSNorm case."""
from . import modules  # noqa: F401
