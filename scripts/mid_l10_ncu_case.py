"""One plan with a 2^10-point MID pass at cfg5's row shape (32 chan x 2 pol), for ncu:
   ncu --set full -k regex:pass_kernel -c 10 python scripts/mid_l10_ncu_case.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pulsarbat_b200 import _lib as L  # noqa: E402

os.environ.setdefault("PBK_LEVELS", "6,8,10")
N, C, P, sr, fcen = 2 ** 24, 32, 2, 1.5625e6, 600e6
freqs = fcen + sr * (np.arange(C) + 0.5 - C / 2)
plan = L.DedispPlan(nsamp=N, nchan=C, npol=P, dm=300.0, sample_rate_hz=sr, ref_freq_hz=fcen,
                    chan_freq_hz=freqs, crop=(0, N), out_kind=1)
x = torch.randn((N, C, P, 2), device="cuda")
out = torch.empty(plan.out_rows * plan.row_elems * plan.elem_bytes, device="cuda", dtype=torch.uint8)
for _ in range(2):
    plan.exec_device(x.data_ptr(), out.data_ptr(), None, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print(plan.describe())
