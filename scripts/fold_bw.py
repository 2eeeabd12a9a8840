"""Fold kernel throughput probe: (nsamp, row_elems) float32 -> (nbin, row_elems) profile."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pulsarbat_b200 as pb  # noqa: E402

dev = torch.device("cuda:0")
for (n, e, nbin, f0, sr) in [(2 ** 22, 128, 1024, 29.7, 6.25e6), (2 ** 22, 128, 1024, 641.9, 6.25e6),
                             (2 ** 16, 1024, 1024, 29.7, 97656.25), (2 ** 20, 2048, 256, 641.9, 1e5),
                             (2 ** 24, 2, 1024, 29.7, 6.25e6), (2 ** 26, 1, 1024, 29.7, 16e6),
                             (2 ** 26, 2, 1024, 29.7, 16e6), (2 ** 25, 4, 1024, 29.7, 16e6),
                             (2 ** 25, 2, 1024, 641.9, 1e5), (2 ** 24, 3, 1024, 29.7, 6.25e6),
                             (2 ** 24, 8, 1024, 29.7, 6.25e6), (2 ** 24, 16, 1024, 29.7, 6.25e6),
                             (2 ** 23, 32, 1024, 29.7, 6.25e6), (2 ** 23, 12, 1024, 29.7, 6.25e6)]:
    x = pb.DeviceArray(torch.rand((n, e), device=dev, dtype=torch.float32))
    coeffs = [0.123, f0, 1e-6]
    for _ in range(2):
        prof, cnt = pb.kernels.fold(x, coeffs, sr, nbin)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prof = pb.DeviceArray(torch.zeros((nbin, e), device=dev, dtype=torch.float32))
    cnt = pb.DeviceArray(torch.zeros((nbin,), device=dev, dtype=torch.int64))
    e0.record()
    for _ in range(5):
        pb.kernels.fold(x, coeffs, sr, nbin, profile=prof, counts=cnt)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    tot = float(prof.tensor.sum()) / 5
    ref = float(x.tensor.sum())
    print(f"fold n=2^{int(np.log2(n))} elems={e} nbin={nbin} f0={f0}: {ms:.3f} ms -> "
          f"{n * e * 4 / ms / 1e6:.0f} GB/s; sum check {tot / ref:.6f}; counts {int(cnt.tensor.sum()) // 5 == n}",
          flush=True)
