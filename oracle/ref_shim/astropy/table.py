"""TEST INFRASTRUCTURE ONLY -- stand-in for ``astropy.table.QTable`` as far as the reference's
``PhasePredictor`` (pulsar/predictor.py:53-160) uses it: built from a list of row dicts (or from
another table), columns by name as arrays -- ``Time`` and ``Quantity`` columns keep their type,
anything else becomes a numpy array (object dtype for the polynomials) -- ``colnames``, ``len``."""

import numpy as np

from . import units as u
from .time import Time

__all__ = ["QTable"]


class QTable:
    def __init__(self, data=None, *args, descriptions=None, **kwargs):
        self._cols = {}
        self.descriptions = descriptions
        if isinstance(data, QTable):
            self._cols = dict(data._cols)
        elif data:
            rows = list(data)
            for name in rows[0]:
                vals = [r[name] for r in rows]
                if isinstance(vals[0], Time):
                    col = Time(vals)
                elif isinstance(vals[0], u.Quantity):
                    unit = vals[0].unit
                    col = u.Quantity(np.array([v.to_value(unit) for v in vals]), unit)
                elif isinstance(vals[0], (str, int, float, np.integer, np.floating)):
                    col = np.array(vals)
                else:
                    col = np.empty(len(vals), dtype=object)
                    col[:] = vals
                self._cols[name] = col

    @property
    def colnames(self):
        return list(self._cols)

    def __getitem__(self, name):
        return self._cols[name]

    def __len__(self):
        return len(next(iter(self._cols.values()))) if self._cols else 0
