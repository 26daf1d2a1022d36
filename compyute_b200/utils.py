"""``save`` / ``load`` — pickle, like compyute/utils.py:44-73.  Device arrays are written as host copies and come back
as device arrays (see ``DeviceArray.__reduce__``)."""

from __future__ import annotations

import pickle
from typing import Any

__all__ = ["save", "load"]


def save(obj: object, filepath: str) -> None:
    with open(filepath, "wb") as f:
        pickle.dump(obj, f)


def load(filepath: str) -> Any:
    with open(filepath, "rb") as f:
        return pickle.load(f)
