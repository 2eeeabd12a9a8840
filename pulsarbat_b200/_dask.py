"""Lazy (dask) inputs: the per-chunk route of the reference, with GPU chunk functions.

The reference applies array functions to dask-backed signals chunk by chunk
(``transforms/transforms.py:49-50``: ``da.map_blocks(func, x.data, ...)``; ``fft.py:40-43``:
``da.fft.fft_wrap``), each chunk function receiving a numpy block and being called concurrently
from dask's thread pool.  This module does the same with the GPU kernels as chunk functions:
nothing is computed until the caller computes the returned dask array, every chunk goes host ->
GPU -> host through ``kernels.*`` (thread-safe: cached plans are pinned and serialised per plan),
and what must be GLOBAL across chunks -- the reference frequency, the crop and every chunk's own
channel frequencies -- is fixed when the graph is built (``core.py:479-484`` re-centres a
frequency slice, so a chunk must never derive them from itself).

dask is an optional dependency: when it cannot be imported no input is ever a dask array and none
of this runs.  Only the stable core of its API is used (``Array.chunks / rechunk / __getitem__``,
``map_blocks(func, arr, dtype=, chunks=)``, ``concatenate``).
"""

import functools

import numpy as np

__all__ = ["dask_array", "is_dask", "map_channel_chunks", "map_time_chunks", "map_whole_axis"]


def dask_array():
    """The ``dask.array`` module, or None when dask is not installed."""
    try:
        import dask.array as da
    except Exception:
        return None
    return da


def is_dask(x):
    da = dask_array()
    return da is not None and isinstance(x, da.Array)


def _bounds(sizes):
    out, lo = [], 0
    for n in sizes:
        out.append((lo, lo + int(n)))
        lo += int(n)
    return out


def map_channel_chunks(x, fn, *, out_rows, out_dtype, drop_trailing=False):
    """Apply ``fn(block, lo, hi)`` to every channel chunk of ``x`` (nsamp, nchan, ...).

    The time axis and the trailing axes are brought into one chunk each (every column needs all
    of its samples; a lane pair is the two pols of a channel); the channel axis keeps the caller's
    chunking, which is what shards the work.  ``fn`` gets the numpy block and its global channel
    range and returns (out_rows, hi - lo[, trailing...]).
    """
    da = dask_array()
    one = {0: -1, **{ax: -1 for ax in range(2, x.ndim)}}
    x = x.rechunk(one)
    trailing = () if drop_trailing else tuple((int(s),) for s in x.shape[2:])
    pieces = []
    for lo, hi in _bounds(x.chunks[1]):
        sub = x[:, lo:hi]
        pieces.append(da.map_blocks(functools.partial(fn, lo=lo, hi=hi), sub, dtype=out_dtype,
                                    chunks=((int(out_rows),), (hi - lo,)) + trailing,
                                    **({"drop_axis": list(range(2, x.ndim))}
                                       if drop_trailing and x.ndim > 2 else {})))
    return pieces[0] if len(pieces) == 1 else da.concatenate(pieces, axis=1)


def map_time_chunks(x, fn, *, multiple, rows_out, cols_out, out_dtype):
    """Apply ``fn(block)`` to chunks along TIME of ``x`` (nsamp, nchan, ...), each a multiple of
    ``multiple`` samples long (STFT segments never straddle a chunk); the other axes are one chunk.
    ``rows_out(n)`` / ``cols_out`` give the output extent of a chunk of n samples."""
    da = dask_array()
    sizes = [int(n) for n in x.chunks[0]]
    if any(n % multiple for n in sizes):
        # re-cut the time axis on segment boundaries, keeping roughly the caller's chunk length
        per = max(multiple, (max(sizes) // multiple) * multiple)
        x = x.rechunk({0: per})
        sizes = [int(n) for n in x.chunks[0]]
    x = x.rechunk({ax: -1 for ax in range(1, x.ndim)})
    trailing = tuple((int(s),) for s in x.shape[2:])
    return da.map_blocks(fn, x, dtype=out_dtype,
                         chunks=(tuple(rows_out(n) for n in sizes), (int(cols_out),)) + trailing)


def map_whole_axis(x, fn, axis, out_dtype):
    """``fn(block)`` on chunks of ``x`` that hold the whole of ``axis`` (what dask's ``fft_wrap``
    requires of an FFT axis); shape-preserving."""
    da = dask_array()
    axis = axis % x.ndim
    if len(x.chunks[axis]) != 1:
        raise ValueError(
            f"Dask array only supports taking an FFT along an axis that has a single chunk. An "
            f"FFT operation was tried on axis {axis}, which has chunks {x.chunks[axis]}. To change "
            "the array's chunks use dask.Array.rechunk.")
    return da.map_blocks(fn, x, dtype=out_dtype, chunks=x.chunks)


def map_elementwise(x, fn, *, out_dtype, chunks=None, **kw):
    da = dask_array()
    return da.map_blocks(fn, x, dtype=out_dtype, **({} if chunks is None else {"chunks": chunks}),
                         **kw)


def real_dtype_of(cdtype):
    return np.float64 if np.dtype(cdtype) == np.complex128 else np.float32
