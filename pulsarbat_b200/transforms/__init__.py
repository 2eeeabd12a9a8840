"""Transforms on the hot path (reference: pulsarbat/transforms/)."""

from .dedispersion import (DM, DispersionMeasure, coherent_dedispersion,  # noqa: F401
                           dedisperse_detect, overlap_save_dedispersion)

__all__ = ["DM", "DispersionMeasure", "coherent_dedispersion", "dedisperse_detect",
           "overlap_save_dedispersion"]
