"""Host-side multi-GPU logic on CPU: channel/time-block sharding and the fold all-reduce with a
world_size-2 gloo group (the GPU box runs the same code over NCCL)."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import pulsarbat_b200 as pb
from oracle import pbk_oracle as orc
from pulsarbat_b200 import sharding as sh


@pytest.mark.parametrize("nchan, world", [(64, 8), (1024, 8), (10, 4), (3, 8), (128, 1)])
def test_channel_ranges_partition(nchan, world):
    got = []
    for r in range(world):
        lo, hi = sh.channel_range(nchan, world, r)
        assert 0 <= lo <= hi <= nchan
        got += list(range(lo, hi))
    assert got == list(range(nchan))
    sizes = [np.diff(sh.channel_range(nchan, world, r))[0] for r in range(world)]
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sh.channel_range(nchan, world, world)


def test_channel_shards_keep_global_frequencies_and_crop():
    x = np.zeros((4096, 16, 2), np.complex64)
    z = pb.DualPolarizationSignal(x, sample_rate=1e6 * pb.units.Hz,
                                  center_freq=600e6 * pb.units.Hz, pol_type="linear")
    start, stop, ref = sh.dedispersion_crop(z, pb.DM(0.5))
    o_start, o_stop = orc.crop_range(0.5, 4096, 600e6, 1e6, 16, 600e6)
    assert (start, stop) == (o_start, o_stop)
    freqs = []
    for r in range(4):
        zs, (lo, hi) = sh.shard_channels(z, 4, r)
        assert zs.shape == (4096, hi - lo, 2)
        freqs += list(zs.channel_freqs_hz)
        # slicing re-centres the shard (core.py:479-484): the default ref_freq would differ
        assert float(pb.units.to_value(zs.center_freq, pb.units.Hz)) != 600e6
    assert np.allclose(freqs, z.channel_freqs_hz, rtol=0, atol=1e-6)
    assert float(pb.units.to_value(ref, pb.units.Hz)) == 600e6


def test_block_ranges_and_time_shards():
    starts = sh.block_ranges(10_000, 4096, 3000)
    assert starts == [0, 3000]
    assert sh.block_ranges(4096, 4096, 1) == [0]
    with pytest.raises(ValueError):
        sh.block_ranges(4096, 1024, 0)
    blocks = [b for r in range(3) for b in sh.time_block_shards(8, 3, r)]
    assert blocks == list(range(8))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nsamp, nelem, nbin, coeffs, sr, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(16)
        x = rng.random((nsamp, nelem), dtype=np.float32)      # same stream on every rank
        lo, hi = sh.channel_range(nsamp, world, rank)          # contiguous time slice
        z = pb.Signal(x[lo:hi], sample_rate=sr * pb.units.Hz)

        def fold_cpu(data, c, rate, nb, n0=0):
            prof, cnt = orc.fold(np.asarray(data), c, rate, nb, n0=n0)
            return prof.astype(np.float32), cnt.astype(np.int64)

        prof, cnt = sh.fold_sharded(z, coeffs, nbin, first_sample=lo, fold_fn=fold_cpu)
        if rank == 0:
            np.save(out + ".prof.npy", prof)
            np.save(out + ".cnt.npy", cnt)
        # torch tensors take the same path
        t = torch.ones(4)
        c = torch.ones(4, dtype=torch.int64)
        sh.allreduce_profiles(t, c)
        assert float(t[0]) == world and int(c[0]) == world
    finally:
        dist.destroy_process_group()


def test_fold_allreduce_world2_gloo(tmp_path):
    nsamp, nelem, nbin, sr = 20_000, 6, 64, 1e4
    coeffs = [0.123, 29.7, 1e-6]
    out = str(tmp_path / "fold")
    mp.spawn(_worker, args=(2, _free_port(), nsamp, nelem, nbin, coeffs, sr, out), nprocs=2,
             join=True)
    rng = np.random.default_rng(16)
    x = rng.random((nsamp, nelem), dtype=np.float32)
    want_p, want_c = orc.fold(x, coeffs, sr, nbin)
    got_p, got_c = np.load(out + ".prof.npy"), np.load(out + ".cnt.npy")
    assert np.array_equal(got_c, want_c)                       # counts: bit-exact
    assert np.allclose(got_p, want_p, rtol=1e-5, atol=1e-5)    # float sums: reduction order


def test_allreduce_is_identity_without_process_group():
    p, c = np.ones((4, 2), np.float32), np.ones(4, np.int64)
    p2, c2 = sh.allreduce_profiles(p, c)
    assert p2 is p and c2 is c


def test_numa_node_lookup_from_sysfs(tmp_path):
    """bind_host_to_device reads the GPU's NUMA node and its cpulist from sysfs."""
    from pulsarbat_b200 import sharding
    dev = tmp_path / "bus" / "pci" / "devices" / "0000:1b:00.0"
    dev.mkdir(parents=True)
    (dev / "numa_node").write_text("1\n")
    node = tmp_path / "devices" / "system" / "node" / "node1"
    node.mkdir(parents=True)
    (node / "cpulist").write_text("16-19,48,50-51\n")
    assert sharding.numa_node_cpus("0000:1B:00.0", sysfs=str(tmp_path)) == (
        1, [16, 17, 18, 19, 48, 50, 51])
    (dev / "numa_node").write_text("-1\n")
    assert sharding.numa_node_cpus("0000:1b:00.0", sysfs=str(tmp_path)) == (None, [])
    assert sharding.numa_node_cpus("0000:ff:00.0", sysfs=str(tmp_path)) == (None, [])
