from .containers import *
from .layers import *
from .module import *
