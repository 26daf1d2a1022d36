"""Reference names outside the hot path (out of scope).  They exist so that reference code which imports them
BY NAME next to in-scope layers keeps importing; using one raises instead of silently computing something else."""

_OUT_OF_SCOPE = ["Conv1D", "ConvTranspose1D", "ConvTranspose2D", "Upsample2D", "LayerNorm", "RMSNorm", "GELU", "FastGELU", "LeakyReLU",
                 "Sigmoid", "SiLU", "Softmax", "Tanh", "Embedding", "GRU", "LSTM", "Recurrent", "ParallelConcat", "ParallelAdd",
                 "BCELoss", "DiceLoss", "MSELoss", "Reshape", "Slice", "DenseBlock", "Convolution1DBlock", "Convolution2DBlock"]


def _stub(name):
    def __init__(self, *a, **k):
        raise NotImplementedError(f"compyute.nn.{name} is outside the B200 hot path (SURVEY §8); use the reference for it")
    return type(name, (), {"__init__": __init__, "__doc__": f"Out-of-scope reference layer {name} (raises when instantiated)."})


import compyute_b200.nn as _nn

__all__ = []
for _n in _OUT_OF_SCOPE:
    if not hasattr(_nn, _n):
        globals()[_n] = _stub(_n)
        __all__.append(_n)
