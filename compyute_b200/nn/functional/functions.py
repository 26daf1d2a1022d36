"""Function protocol and the per-module LIFO cache — the drop-in boundary of the hot path.

Behavioural mirror of compyute/nn/functional/functions.py:12-78: ``push`` stores one tuple per call (dropped
while caching is disabled or on a ``PseudoCache``), ``pop`` returns the most recent tuple.  A fused
``XxxFn`` here pushes exactly one tuple and pops exactly that tuple, so a module's cache ends balanced
just like with the reference's nested sub-Functions.
"""

from __future__ import annotations

from contextlib import contextmanager
from typing import Any

__all__ = ["Function", "FunctionCache", "PseudoCache", "no_caching", "get_caching_enabled", "set_caching_enabled"]

_caching = True


def get_caching_enabled() -> bool:
    return _caching


def set_caching_enabled(enabled: bool) -> None:
    global _caching
    _caching = bool(enabled)


@contextmanager
def no_caching():
    """Disables caching for gradient computation inside the block (functions.py:71-78)."""
    set_caching_enabled(False)
    try:
        yield
    finally:
        set_caching_enabled(True)


class FunctionCache:
    """Stack of tuples saved by ``forward`` for ``backward``."""

    def __init__(self) -> None:
        self.cache: list[tuple[Any, ...]] = []

    def push(self, *items: Any) -> None:
        if _caching:
            self.cache.append(items)

    def pop(self) -> tuple[Any, ...]:
        return self.cache.pop()


class PseudoCache(FunctionCache):
    """Placeholder used in inference mode / by the user-level wrappers: ``push`` is a no-op."""

    def push(self, *items: Any) -> None:
        return None


class Function:
    """Stateless op with static ``forward(cache, ...)`` / ``backward(cache, dy)`` (functions.py:37-54)."""

    def __init__(self) -> None:
        raise NotImplementedError("Function cannot be instantiated.")

    @staticmethod
    def forward(*args: Any, **kwargs: Any) -> Any:
        raise NotImplementedError

    @staticmethod
    def backward(*args: Any, **kwargs: Any) -> Any:
        raise NotImplementedError
