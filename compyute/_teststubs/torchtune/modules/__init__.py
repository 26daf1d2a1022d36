class RMSNorm:  # pragma: no cover
    def __init__(self, *a, **k):
        raise NotImplementedError("torchtune is not installed; RMSNorm is outside the hot path")
