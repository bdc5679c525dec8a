"""Kernel compute engines (reference: computations/__init__.py)."""

from .lazy import LazyKernelComputation

__all__ = ["LazyKernelComputation"]
