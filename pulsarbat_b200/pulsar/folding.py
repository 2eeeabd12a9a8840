"""Phase-binned folding on the GPU (builder-defined: the reference only lists an absent
``pulsar/folding.py`` in setup.cfg:60; semantics in SURVEY.md 8a row F and oracle.fold)."""

import numpy as np

from .. import kernels
from .. import units as u
from ..core import Signal

__all__ = ["fold"]


def fold(z, predictor, nbin, *, profile=None, counts=None, want_bins=False):
    """Fold a real-valued signal into ``nbin`` pulse-phase bins.

    ``predictor`` is a :class:`PhasePredictor` (its ``phasepol(z.start_time)`` supplies the phase
    polynomial, reference predictor.py:149-160) or a plain coefficient sequence in ascending powers
    of seconds since the first sample.  Returns ``(profile, counts)`` -- profile has shape
    (nbin,) + z.sample_shape, float32; counts is (nbin,) int64 and exact -- plus the per-sample bin
    index when ``want_bins``.  Pass ``profile``/``counts`` from an earlier call to accumulate.
    """
    if not isinstance(z, Signal):
        raise TypeError("z must be a Signal.")
    if hasattr(predictor, "phasepol"):
        if z.start_time is None:
            raise ValueError("folding with a predictor needs a signal with a start_time")
        coeffs, _ = predictor.phasepol(z.start_time)
    else:
        coeffs = np.asarray(predictor, dtype=np.float64)
    if np.iscomplexobj(np.empty(0, dtype=z.dtype)):
        raise TypeError("fold expects real-valued (intensity) data")
    return kernels.fold(z.data, coeffs, float(u.to_value(z.sample_rate, u.Hz)), int(nbin),
                        profile=profile, counts=counts, want_bins=want_bins)
