"""Experimental routines (reference: pulsarbat/contrib/)."""

from .misc import istft, stft  # noqa: F401

__all__ = ["stft", "istft"]
