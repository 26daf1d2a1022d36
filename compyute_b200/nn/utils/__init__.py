"""Host-side utilities around the path (compyute/nn/utils): input pipeline, gradient clipping, LR schedulers."""

from . import lr_schedulers
from .dataloaders import *
from .lr_schedulers import *
from .training import *
