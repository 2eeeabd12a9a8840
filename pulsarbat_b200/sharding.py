"""One-process-per-GPU sharding of the baseband hot path.

Every (channel, pol) column and every overlap-save time block is independent for dedispersion,
channelize and detection, so ranks never exchange voltages: each rank takes a contiguous channel
range (or a set of time blocks) and runs the single-GPU kernels on it.  The one exchange step is
folding with time-sharded input: every rank folds its own samples into a full-size profile and
the profiles are summed with ONE all-reduce (NCCL over NVLink on the GPU box, gloo in the CPU
tests).

The reference has no multi-process code; what must stay identical across shards is spelled out
in its single-process semantics:
  * ``ref_freq`` defaults to the signal's centre frequency (dedispersion.py:118-119) and slicing
    channels re-centres it (core.py:479-484), so shards always get the GLOBAL ``ref_freq``;
  * the crop (dedispersion.py:127-131) is computed from the whole band's edges, so shards use
    the GLOBAL (start, stop) and concatenate along frequency afterwards (transforms.py:126-141).
"""

import numpy as np

from . import units as u

__all__ = ["channel_range", "shard_channels", "block_ranges", "time_block_shards",
           "dedispersion_crop", "allreduce_profiles", "fold_sharded", "bind_host_to_device",
           "numa_node_cpus"]


def _parse_cpulist(text):
    """'0-3,8,10-11' -> [0, 1, 2, 3, 8, 10, 11] (the kernel's cpulist format)."""
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def numa_node_cpus(pci_bus_id, sysfs="/sys"):
    """(node, cpus) of the NUMA node a PCI device hangs off, read from sysfs; (None, []) when the
    platform does not say (single-node hosts report -1)."""
    try:
        with open(f"{sysfs}/bus/pci/devices/{pci_bus_id.lower()}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None, []
        with open(f"{sysfs}/devices/system/node/node{node}/cpulist") as f:
            return node, _parse_cpulist(f.read())
    except (OSError, ValueError):
        return None, []


def bind_host_to_device(device):
    """Pin the calling process to the CPUs of ``device``'s NUMA node.

    With one process per GPU every rank stages its blocks through page-locked host memory; the
    kernel places those pages on the node of the allocating thread, so binding first keeps each
    rank's PCIe traffic off the inter-socket link.  Returns the node, or None when nothing was
    changed (unknown topology, or the node has no CPU this process may run on)."""
    import ctypes
    import os
    from . import _lib as L
    buf = ctypes.create_string_buffer(32)
    try:
        L.check(L.lib().pbk_device_pci_bus_id(int(device), buf, 32))
    except L.PbkError:
        return None
    node, cpus = numa_node_cpus(buf.value.decode())
    if node is None or not hasattr(os, "sched_setaffinity"):
        return None
    allowed = set(os.sched_getaffinity(0)) & set(cpus)
    if not allowed:
        return None
    os.sched_setaffinity(0, allowed)
    return node


def channel_range(nchan, world_size, rank):
    """Contiguous [lo, hi) channel range of ``rank``; the first ``nchan % world_size`` ranks get
    one extra channel, empty ranges are allowed when world_size > nchan."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(nchan, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def dedispersion_crop(z, dm, ref_freq=None):
    """Global (start, stop, ref_freq) of ``coherent_dedispersion(z, dm)`` computed on the FULL
    band exactly as dedispersion.py:118-131; pass these to every shard."""
    from .transforms.dedispersion import _as_dm, crop_range
    if ref_freq is None:
        ref_freq = z.center_freq
    start, stop = crop_range(z, _as_dm(dm), ref_freq)
    return start, stop, ref_freq


def shard_channels(z, world_size, rank):
    """This rank's channel slice of a RadioSignal with its own (re-centred) metadata
    (core.py:479-498) -- remember to keep using the global ``ref_freq``."""
    lo, hi = channel_range(z.nchan, world_size, rank)
    return z[:, lo:hi], (lo, hi)


def block_ranges(nsamp, block_len, valid):
    """Start offsets of the overlap-save blocks of a stream of ``nsamp`` samples: blocks of
    ``block_len`` advancing by ``valid`` = block_len - sweep."""
    if valid <= 0:
        raise ValueError("block length does not exceed the dispersion sweep")
    out, b = [], 0
    while b + block_len <= nsamp:
        out.append(b)
        b += valid
    return out


def time_block_shards(nblocks, world_size, rank):
    """Block indices of ``rank``: contiguous ranges, so each rank reads one contiguous piece of
    the stream (plus the overlap)."""
    lo, hi = channel_range(nblocks, world_size, rank)
    return list(range(lo, hi))


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def allreduce_profiles(profile, counts):
    """Sum folded profiles (float32) and counts (int64) over all ranks, in place.

    Accepts numpy arrays, torch tensors (CPU for gloo, CUDA for NCCL) or DeviceArrays and returns
    objects of the same kind.  Without an initialised process group this is the identity, so
    single-GPU code can call it unconditionally.  The counts are integers: their sum is exact and
    identical on every rank whatever the reduction order.
    """
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return profile, counts
    import torch

    def as_tensor(a):
        if hasattr(a, "tensor"):
            return a.tensor, None
        if isinstance(a, torch.Tensor):
            return a, None
        arr = np.ascontiguousarray(a)
        t = torch.from_numpy(arr)
        if dist.get_backend() == "nccl":
            t = t.cuda()
        return t, arr

    tp, hp = as_tensor(profile)
    tc, hc = as_tensor(counts)
    dist.all_reduce(tp, op=dist.ReduceOp.SUM)
    dist.all_reduce(tc, op=dist.ReduceOp.SUM)
    if hp is not None:
        hp[...] = tp.cpu().numpy()
        profile = hp
    if hc is not None:
        hc[...] = tc.cpu().numpy()
        counts = hc
    return profile, counts


def fold_sharded(z, coeffs, nbin, *, first_sample, fold_fn=None):
    """Fold this rank's time slice ``z`` (whose first sample is sample ``first_sample`` of the
    whole stream) and all-reduce the result.  ``coeffs`` is the phase polynomial of the WHOLE
    stream (ascending powers of seconds since ITS first sample), so every rank evaluates the same
    phase for the same absolute sample and the reduced counts equal the single-rank counts bit
    for bit.  ``fold_fn`` defaults to the GPU kernel wrapper."""
    if fold_fn is None:
        from . import kernels
        fold_fn = kernels.fold
    sr = float(u.to_value(z.sample_rate, u.Hz))
    profile, counts = fold_fn(z.data, np.asarray(coeffs, dtype=np.float64), sr, int(nbin),
                              n0=int(first_sample))
    return allreduce_profiles(profile, counts)
