"""GPU-resident arrays that can live inside a Signal.

``Signal.__init__`` only touches ``ndim``, ``shape``, ``dtype`` and ``astype`` of its data
(reference core.py:59-97), so a thin wrapper exposing that surface is a legal ``Signal.data``.
The storage is a torch CUDA tensor (PyTorch is the device-memory and stream plumbing here, not
the compute); complex data is kept as torch.complex64.
"""

import os
import warnings

import numpy as np

__all__ = ["DeviceArray"]

#: implicit device -> host copies (np.asarray(device_array)) above this size warn once per call site
_D2H_WARN_BYTES = int(float(os.environ.get("PBK_D2H_WARN_BYTES", 256 << 20)))

_NP2T = None


def _maps():
    global _NP2T
    import torch
    if _NP2T is None:
        _NP2T = {np.dtype(np.complex64): torch.complex64, np.dtype(np.complex128): torch.complex128,
                 np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
                 np.dtype(np.int8): torch.int8, np.dtype(np.uint8): torch.uint8,
                 np.dtype(np.int32): torch.int32,
                 np.dtype(np.int64): torch.int64}
    return _NP2T


class DeviceArray:
    """A C-contiguous array in HBM with a numpy-like face."""

    def __init__(self, tensor):
        import torch
        if not isinstance(tensor, torch.Tensor) or not tensor.is_cuda:
            raise TypeError("DeviceArray wraps a CUDA torch.Tensor")
        self.tensor = tensor

    # -- construction -----------------------------------------------------------------------
    @classmethod
    def from_numpy(cls, a, device=0):
        import torch
        t = torch.from_numpy(np.ascontiguousarray(a)).to(f"cuda:{device}", non_blocking=False)
        return cls(t)

    @classmethod
    def from_dlpack(cls, obj):
        import torch
        return cls(torch.from_dlpack(obj))

    @classmethod
    def empty(cls, shape, dtype, device=0):
        import torch
        return cls(torch.empty(tuple(shape), dtype=_maps()[np.dtype(dtype)],
                               device=f"cuda:{device}"))

    # -- numpy-like face --------------------------------------------------------------------
    @property
    def shape(self):
        return tuple(self.tensor.shape)

    @property
    def ndim(self):
        return self.tensor.ndim

    @property
    def dtype(self):
        for k, v in _maps().items():
            if v == self.tensor.dtype:
                return k
        raise TypeError(f"unsupported tensor dtype {self.tensor.dtype}")

    @property
    def device(self):
        return self.tensor.device.index or 0

    @property
    def ptr(self):
        return self.tensor.data_ptr()

    def contiguous(self):
        return self if self.tensor.is_contiguous() else DeviceArray(self.tensor.contiguous())

    def astype(self, dtype, casting="unsafe"):
        dtype = np.dtype(dtype)
        if not np.can_cast(self.dtype, dtype, casting=casting):
            raise TypeError(f"Cannot cast from {self.dtype} to {dtype} with casting '{casting}'")
        return DeviceArray(self.tensor.to(_maps()[dtype]))

    def __getitem__(self, idx):
        return DeviceArray(self.tensor[idx])

    def __len__(self):
        return self.tensor.shape[0]

    def numpy(self):
        """Explicit device -> host copy (never warns)."""
        return self.tensor.detach().cpu().numpy()

    def __array__(self, dtype=None, copy=None):
        # np.asarray(signal) on a resident array is a full device -> host copy: harmless for a
        # profile, a hidden multi-gigabyte PCIe transfer for a voltage block -- say so
        nbytes = self.tensor.numel() * self.tensor.element_size()
        if nbytes > _D2H_WARN_BYTES:
            warnings.warn(f"implicit device->host copy of {nbytes / 2**20:.0f} MiB from a "
                          "DeviceArray (np.asarray / __array__); call .numpy() to make it explicit "
                          "or raise $PBK_D2H_WARN_BYTES", ResourceWarning, stacklevel=2)
        a = self.numpy()
        return a if dtype is None else a.astype(dtype)

    def __dlpack__(self, stream=None):
        return self.tensor.__dlpack__(stream=stream)

    def __dlpack_device__(self):
        return self.tensor.__dlpack_device__()

    def __repr__(self):
        return f"DeviceArray(shape={self.shape}, dtype={self.dtype}, device=cuda:{self.device})"
