"""TEST INFRASTRUCTURE: one rank of the multi-GPU parity run (tests/test_gpu_multirank.py launches
it with torch.distributed.run, one process per GPU, NCCL).  Every rank runs the CUDA kernels on
ITS shard through the public API, the one exchange step (the all-reduce of folded profiles and
counts) goes over NCCL, and the result is compared with the single-process oracle on the whole
stream: counts and bins bit-exact, profiles and voltages <= 1e-5 relative RMS."""

import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import pulsarbat_b200 as pb                      # noqa: E402
from pulsarbat_b200 import sharding as sh        # noqa: E402
from oracle import pbk_oracle as orc             # noqa: E402

TOL = 1e-5


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64 if not np.iscomplexobj(b) else np.complex128), np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(np.ravel(b)), 1e-300))


def main():
    out_path = sys.argv[1]
    world, rank = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ["PBK_DEVICE"] = str(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    report = {"world": world, "backend": dist.get_backend(), "checks": {}}
    try:
        coeffs = [0.123, 29.7, 1e-6]
        # ---- 1. CUDA fold of time slices + NCCL all-reduce vs the oracle on the whole stream ----
        for tag, (nsamp, sshape, nbin, sr) in {"wide": (40_000, (96,), 128, 1e4),
                                               "narrow": (60_001, (2,), 64, 1e4),
                                               "dualpol": (30_000, (40, 2), 1024, 3e3)}.items():
            x = np.random.default_rng(16).random((nsamp,) + sshape, dtype=np.float32)
            lo, hi = sh.channel_range(nsamp, world, rank)            # contiguous time slice
            z = pb.Signal(pb.DeviceArray.from_numpy(x[lo:hi], local), sample_rate=sr * pb.units.Hz)
            prof, cnt = sh.fold_sharded(z, coeffs, nbin, first_sample=lo)     # CUDA fold + NCCL
            want_p, want_c = orc.fold(x, coeffs, sr, nbin)
            _, _, bins = pb.kernels.fold(z.data, coeffs, sr, nbin, n0=lo, want_bins=True)
            ok_bins = np.array_equal(np.asarray(bins), orc.fold_bins(nsamp, coeffs, sr, nbin)[lo:hi])
            e = relerr(np.asarray(prof), want_p)
            report["checks"][f"fold_{tag}"] = dict(
                counts_equal=bool(np.array_equal(np.asarray(cnt), want_c)), bins_equal=bool(ok_bins),
                profile_relerr=e)
            assert np.array_equal(np.asarray(cnt), want_c), tag
            assert ok_bins and e <= TOL, (tag, e)

        # ---- 2. cfg4 in small: channelize -> detect -> x16 channel sum -> fold -> all-reduce ------
        nper, fsum, nbin, seg_per_rank, npol = 1024, 16, 64, 96, 2
        n_all = world * seg_per_rank * nper
        rng = np.random.default_rng(17)
        xs = (rng.standard_normal((n_all, 1, npol)) + 1j * rng.standard_normal((n_all, 1, npol))
              ).astype(np.complex64)
        sr = 4e6
        mine = xs[rank * seg_per_rank * nper:(rank + 1) * seg_per_rank * nper]
        zc = pb.kernels.stft(pb.DeviceArray.from_numpy(mine, local), nper)
        inten = pb.kernels.detect(zc, freq_sum=fsum)
        prof, cnt = pb.kernels.fold(inten, coeffs, sr / nper, nbin, n0=rank * seg_per_rank)
        prof, cnt = sh.allreduce_profiles(prof, cnt)
        ch = orc.stft(xs.astype(np.complex128), nper)                 # (segments, 1024, 2)
        iw = orc.to_intensity(ch).reshape(ch.shape[0], nper // fsum, fsum, npol).sum(2)
        want_p, want_c = orc.fold(iw, coeffs, sr / nper, nbin)
        e = relerr(np.asarray(prof), want_p)
        report["checks"]["cfg4_small"] = dict(
            counts_equal=bool(np.array_equal(np.asarray(cnt), want_c)), profile_relerr=e,
            counts_sum=int(np.asarray(cnt).sum()))
        assert np.array_equal(np.asarray(cnt), want_c) and e <= TOL, e

        # ---- 3. channel-sharded dedispersion with the GLOBAL ref_freq and crop (cfg3 in small) ---
        N, C = 8192, 16
        xr = np.random.default_rng(18).integers(-127, 128, (N, C, 2, 2), dtype=np.int8)
        xc = (xr[..., 0].astype(np.float64) + 1j * xr[..., 1]).astype(np.complex128)
        z = pb.DualPolarizationSignal(xc.astype(np.complex64), sample_rate=1e6 * pb.units.Hz,
                                      center_freq=600e6 * pb.units.Hz, pol_type="linear")
        start, stop, ref = sh.dedispersion_crop(z, pb.DM(2.0))
        lo, hi = sh.channel_range(C, world, rank)
        y = pb.kernels.dedisperse(pb.DeviceArray.from_numpy(xr[:, lo:hi], local), dm=2.0,
                                  sample_rate_hz=1e6, chan_freq_hz=z.channel_freqs_hz[lo:hi],
                                  ref_freq_hz=600e6, crop=(start, stop), raw="int8")
        want, s0, s1 = orc.coherent_dedispersion(xc, 2.0, sample_rate=1e6, center_freq=600e6)
        e = relerr(np.asarray(y), want[:, lo:hi])
        report["checks"]["shard_dedisp"] = dict(relerr=e, crop=[start, stop],
                                                crop_equal=bool((start, stop) == (s0, s1)))
        assert (start, stop) == (s0, s1) and e <= TOL, e
        report["ok"] = True
    except BaseException as exc:                                 # noqa: BLE001
        report["ok"] = False
        report["error"] = f"{type(exc).__name__}: {exc}"
        raise
    finally:
        with open(f"{out_path}.rank{rank}.json", "w") as f:
            json.dump(report, f)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
