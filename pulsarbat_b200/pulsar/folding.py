"""Phase-binned folding on the GPU (builder-defined: the reference only lists an absent
``pulsar/folding.py`` in setup.cfg:60; semantics in SURVEY.md 8a row F and oracle.fold)."""

import math

import numpy as np

from .. import kernels
from .. import units as u
from ..core import Signal

__all__ = ["fold", "fold_segments"]


def fold_segments(z, predictor):
    """Split ``z`` where the predictor switches polyco entry (reference predictor.py:108-119 picks
    the entry with ``searchsorted(span_ends, t)``): a list of ``(first_sample, nsamples, coeffs)``
    with ``coeffs`` = ``predictor.phasepol(time of first_sample)[0]`` (predictor.py:149-160), the
    phase polynomial in seconds since that sample."""
    if z.start_time is None:
        raise ValueError("folding with a predictor needs a signal with a start_time")
    n = len(z)
    sr = float(u.to_value(z.sample_rate, u.Hz))
    t0 = z.start_time
    ends = [e.tmid + e.span / 2 for e in predictor.entries]
    segs, first = [], 0
    while first < n:
        t_first = t0 + (first / sr) * u.s
        idx, _ = predictor._index_and_dt(t_first)
        # samples up to and including the span end belong to entry idx (searchsorted side='left')
        left = float((ends[idx] - t_first).to_value(u.s)) * sr
        count = n - first if idx == len(predictor.entries) - 1 else \
            max(1, min(n - first, int(math.floor(left + 1e-9)) + 1))
        coeffs, _ = predictor.phasepol(t_first)
        segs.append((first, count, np.asarray(coeffs, dtype=np.float64)))
        first += count
    return segs


def fold(z, predictor, nbin, *, profile=None, counts=None, want_bins=False):
    """Fold a real-valued signal into ``nbin`` pulse-phase bins.

    ``predictor`` is a :class:`PhasePredictor` or a plain coefficient sequence in ascending powers
    of seconds since the first sample.  With a predictor the signal is folded entry by entry
    (:func:`fold_segments`): every stretch uses the phase polynomial of the polyco entry the
    reference would pick for its samples, re-centred on the stretch's first sample, so signals
    longer than one polyco span fold correctly.  Returns ``(profile, counts)`` -- profile has shape
    (nbin,) + z.sample_shape, float32; counts is (nbin,) int64 and exact -- plus the per-sample bin
    index when ``want_bins``.  Pass ``profile``/``counts`` from an earlier call to accumulate.
    """
    if not isinstance(z, Signal):
        raise TypeError("z must be a Signal.")
    if np.iscomplexobj(np.empty(0, dtype=z.dtype)):
        raise TypeError("fold expects real-valued (intensity) data")
    sr = float(u.to_value(z.sample_rate, u.Hz))
    if not hasattr(predictor, "phasepol"):
        coeffs = np.asarray(predictor, dtype=np.float64)
        return kernels.fold(z.data, coeffs, sr, int(nbin), profile=profile, counts=counts,
                            want_bins=want_bins)
    bins = []
    for first, count, coeffs in fold_segments(z, predictor):
        res = kernels.fold(z.data[first:first + count], coeffs, sr, int(nbin), profile=profile,
                           counts=counts, want_bins=want_bins)
        profile, counts = res[0], res[1]
        if want_bins:
            bins.append(np.asarray(res[2]))
    if want_bins:
        return profile, counts, np.concatenate(bins) if bins else np.empty(0, np.int32)
    return profile, counts
