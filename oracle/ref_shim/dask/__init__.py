"""TEST INFRASTRUCTURE ONLY -- import-time stand-in for dask (absent from the build image).  The
reference only needs the NAMES at import time (``isinstance(x, dask.array.Array)`` checks, which
are False for numpy input); every lazy code path raises here."""

from . import array  # noqa: F401


def _absent(*a, **k):
    raise RuntimeError("dask stub: the lazy (dask) code path is not available in this image")


delayed = compute = persist = _absent
