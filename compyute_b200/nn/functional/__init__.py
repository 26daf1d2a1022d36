"""compyute_b200.nn.functional — same names as compyute.nn.functional for the CNN hot path."""

from .activation_funcs import *
from .convolution_funcs import *
from .functions import *
from .linear_funcs import *
from .loss_funcs import *
from .normalization_funcs import *
from .pooling_funcs import *
from .regularization_funcs import *
from .shape_funcs import *
