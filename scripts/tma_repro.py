"""Developer repro: one plan execution with the TMA kernels (for compute-sanitizer).
usage: tma_repro.py log2N C P out_kind ds [levels] [tma_kinds]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pulsarbat_b200 import _lib as L  # noqa: E402

n, C, P, ok, ds = (int(v) for v in sys.argv[1:6])
if len(sys.argv) > 6 and sys.argv[6] != "auto":
    os.environ["PBK_LEVELS"] = sys.argv[6]
os.environ["PBK_TMA"] = sys.argv[7] if len(sys.argv) > 7 else "1"
N = 2 ** n
sr, fcen = 6.25e6, 600e6
freqs = fcen + sr * (np.arange(C) + 0.5 - C / 2)
plan = L.DedispPlan(nsamp=N, nchan=C, npol=P, dm=3.0, sample_rate_hz=sr, ref_freq_hz=fcen,
                    chan_freq_hz=freqs, crop=(0, N), out_kind=ok, downsample=ds)
print(plan.describe(), flush=True)
x = torch.randn((N, C, P, 2), device="cuda", dtype=torch.float32)
nout = plan.out_rows * plan.row_elems * plan.elem_bytes
out = torch.zeros(max(nout, 16), device="cuda", dtype=torch.uint8)
plan.exec_device(x.data_ptr(), out.data_ptr(), None, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("ok", float(out[:nout].view(torch.float32).abs().sum()))
