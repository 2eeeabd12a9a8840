"""Channelizer throughput probe: stft of (N, C, P) complex64 with nperseg n (device resident)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pulsarbat_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda:0")
for (N, C, P, n) in [(2 ** 22, 64, 2, 2 ** 10), (2 ** 22, 64, 2, 2 ** 6), (2 ** 22, 16, 2, 2 ** 16),
                     (2 ** 26, 1, 2, 2 ** 16), (2 ** 26, 1, 1, 2 ** 16), (2 ** 24, 8, 2, 2 ** 12)]:
    nseg = N // n
    x = torch.randn((N, C, P, 2), device=dev, dtype=torch.float32)
    y = torch.empty_like(x)
    st = torch.cuda.current_stream().cuda_stream
    for inverse in (False, True):
        plan = L.STFTPlan(nseg, n, C, P, inverse=inverse)
        for _ in range(2):
            plan.exec_device(x.data_ptr(), y.data_ptr(), st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            plan.exec_device(x.data_ptr(), y.data_ptr(), st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"{'istft' if inverse else 'stft '} N=2^{int(np.log2(N))} C={C} P={P} n=2^{int(np.log2(n))}: "
              f"{ms:.3f} ms -> {N * C * P / ms / 1e6:.1f} Gsamples/s, {2 * x.numel() * 4 / ms / 1e6:.0f} GB/s "
              f"algorithmic | {plan.describe()}", flush=True)
        plan.destroy()
    del x, y
