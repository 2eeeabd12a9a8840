"""TEST INFRASTRUCTURE ONLY: a tiny stand-in for ``dask.array`` (dask is not installable in this
image).  It implements the handful of calls ``pulsarbat_b200/_dask.py`` relies on -- ``Array``
with ``chunks / rechunk / __getitem__ / reshape / compute``, ``map_blocks(func, arr, dtype=,
chunks=, drop_axis=)``, ``concatenate`` and ``from_array`` -- with dask's semantics: lazily (the
chunk functions run at ``compute()``, from a thread pool like dask's threaded scheduler) and chunk
by chunk (``func`` only ever sees one numpy block)."""

import itertools
import sys
import types
from concurrent.futures import ThreadPoolExecutor

import numpy as np


def _norm_chunks(shape, chunks):
    out = []
    for n, c in zip(shape, chunks):
        if isinstance(c, (tuple, list)):
            assert sum(c) == n, (c, n)
            out.append(tuple(int(v) for v in c))
        else:
            c = n if c in (-1, None) else int(c)
            out.append(tuple([c] * (n // c) + ([n % c] if n % c else [])) or (0,))
    return tuple(out)


class Array:
    def __init__(self, fn, shape, dtype, chunks):
        self._fn = fn                      # () -> numpy array (the whole thing)
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.chunks = _norm_chunks(self.shape, chunks)
        self.ndim = len(self.shape)

    def compute(self):
        a = np.asarray(self._fn())
        assert a.shape == self.shape, (a.shape, self.shape)
        return a

    def __array__(self, dtype=None, copy=None):
        a = self.compute()
        return a if dtype is None else a.astype(dtype)

    def __len__(self):
        return self.shape[0]

    def astype(self, dtype, casting="unsafe"):
        if not np.can_cast(self.dtype, dtype, casting=casting):
            raise TypeError("cannot cast")
        return Array(lambda: self.compute().astype(dtype), self.shape, dtype, self.chunks)

    def rechunk(self, spec):
        ch = list(self.chunks)
        if isinstance(spec, dict):
            for ax, c in spec.items():
                ch[ax] = c
        else:
            ch = list(spec)
        return Array(self._fn, self.shape, self.dtype, ch)

    def reshape(self, shape):
        shape = tuple(shape)
        return Array(lambda: self.compute().reshape(shape), shape, self.dtype, shape)

    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        idx = idx + (slice(None),) * (self.ndim - len(idx))
        shape, chunks = [], []
        for n, sl, ch in zip(self.shape, idx, self.chunks):
            lo, hi, step = sl.indices(n)
            assert step == 1
            shape.append(max(0, hi - lo))
            # chunk boundaries clipped to [lo, hi)
            edges, pos, out = np.cumsum((0,) + ch), 0, []
            for a, b in zip(edges[:-1], edges[1:]):
                a, b = max(a, lo), min(b, hi)
                if b > a:
                    out.append(int(b - a))
            chunks.append(tuple(out) or (0,))
        return Array(lambda: self.compute()[idx], shape, self.dtype, chunks)


def from_array(a, chunks):
    a = np.asarray(a)
    return Array(lambda: a, a.shape, a.dtype, chunks)


def map_blocks(func, arr, dtype=None, chunks=None, drop_axis=None, **kwargs):
    drop = [] if drop_axis is None else ([drop_axis] if isinstance(drop_axis, int) else
                                         list(drop_axis))
    out_chunks = chunks
    if out_chunks is None:
        out_chunks = tuple(c for ax, c in enumerate(arr.chunks) if ax not in drop)
    out_shape = tuple(sum(c) for c in out_chunks)
    in_edges = [np.cumsum((0,) + c) for c in arr.chunks]
    out_edges = [np.cumsum((0,) + tuple(c)) for c in out_chunks]
    kept = [ax for ax in range(arr.ndim) if ax not in drop]
    assert len(kept) == len(out_chunks)
    for ax_out, ax_in in enumerate(kept):
        assert len(out_chunks[ax_out]) == len(arr.chunks[ax_in]), "block counts must agree"
    for ax in drop:
        assert len(arr.chunks[ax]) == 1, "dropped axes must be a single chunk"

    def run():
        src = arr.compute()
        out = np.empty(out_shape, dtype if dtype is not None else arr.dtype)
        blocks = list(itertools.product(*[range(len(c)) for c in arr.chunks]))

        def one(bidx):
            isl = tuple(slice(in_edges[ax][b], in_edges[ax][b + 1]) for ax, b in enumerate(bidx))
            res = np.asarray(func(src[isl], **kwargs))
            osl = tuple(slice(out_edges[i][bidx[ax]], out_edges[i][bidx[ax] + 1])
                        for i, ax in enumerate(kept))
            assert res.shape == out[osl].shape, (res.shape, out[osl].shape)
            out[osl] = res
        with ThreadPoolExecutor(4) as ex:      # chunk functions run concurrently, like dask threads
            list(ex.map(one, blocks))
        return out
    return Array(run, out_shape, dtype if dtype is not None else arr.dtype, out_chunks)


def concatenate(arrs, axis=0):
    shape = list(arrs[0].shape)
    shape[axis] = sum(a.shape[axis] for a in arrs)
    chunks = list(arrs[0].chunks)
    chunks[axis] = tuple(itertools.chain.from_iterable(a.chunks[axis] for a in arrs))
    return Array(lambda: np.concatenate([a.compute() for a in arrs], axis=axis), shape,
                 arrs[0].dtype, chunks)


class _FFT:
    @staticmethod
    def fft_wrap(fn):
        def wrapped(a, *args, **kw):
            return Array(lambda: fn(a.compute(), *args, **kw), a.shape,
                         np.result_type(a.dtype, np.complex64), a.chunks)
        return wrapped


def install(monkeypatch):
    """Make ``import dask.array`` resolve to this module for the duration of a test."""
    dask = types.ModuleType("dask")
    da = types.ModuleType("dask.array")
    for k, v in dict(Array=Array, from_array=from_array, map_blocks=map_blocks,
                     concatenate=concatenate, fft=_FFT()).items():
        setattr(da, k, v)
    dask.array = da
    monkeypatch.setitem(sys.modules, "dask", dask)
    monkeypatch.setitem(sys.modules, "dask.array", da)
    return da
