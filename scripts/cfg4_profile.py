"""Developer probe: per-launch device time of the fused cfg4 block pipeline
(kernels.stft_fold: bins + counts | FWD pass | FWD-last pass with detect + channel sum + fold).
usage: cfg4_profile.py [npol]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pulsarbat_b200 import _lib as L  # noqa: E402

npol = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n_per, nper, fsum, nbin, sr = 2 ** 26, 2 ** 16, 64, 1024, 400e6
nseg = n_per // nper
x = torch.randn((n_per, 1, npol, 2), device="cuda", dtype=torch.float32)
plan = L.STFTDetectPlan(nseg, nper, 1, npol, L.OUT_INTENSITY, fsum)
prof = torch.zeros((nbin, nper // fsum, npol), device="cuda", dtype=torch.float32)
cnt = torch.zeros((nbin,), device="cuda", dtype=torch.int64)
st = torch.cuda.current_stream().cuda_stream
coeffs = [0.123, 29.7, 1e-6]
for _ in range(3):
    plan.fold_device(x.data_ptr(), prof.data_ptr(), cnt.data_ptr(), coeffs, sr / nper, 0, nbin, st)
torch.cuda.synchronize()
K = 10
plan.profile(K)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(K):
    plan.fold_device(x.data_ptr(), prof.data_ptr(), cnt.data_ptr(), coeffs, sr / nper, 0, nbin, st)
e1.record()
torch.cuda.synchronize()
seg = np.array([plan.profile_read(i) for i in range(K)]).mean(axis=0)
gb = x.numel() * 4 / 1e9
print(f"cfg4 block 2^26 x {npol} pol: {e0.elapsed_time(e1) / K:.3f} ms per block; passes "
      f"{np.round(seg, 4).tolist()} ms [{plan.describe()}]; pass 1 moves {2 * gb:.2f} GB "
      f"({2 * gb / seg[0]:.0f} GB/s), pass 2 reads {gb:.2f} GB ({gb / seg[1]:.0f} GB/s)")
