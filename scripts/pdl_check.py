"""Latency of one plan execution with a cold L2, with and without programmatic dependent launch
between the passes, and replayed from a CUDA graph.  Run twice: PBK_PDL=0 and PBK_PDL=1 (the
switch is read once per process).  No per-launch events here: an event between two kernels
serialises them, which is exactly what is being measured.

  PBK_PDL=0 python scripts/pdl_check.py ; PBK_PDL=1 python scripts/pdl_check.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pulsarbat_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda:0")
stream = torch.cuda.Stream()
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
CASES = [("cfg1 2^20 x 1", 2 ** 20, 1, 1, 71.0, 16e6, 400e6),
         ("2^18 x 8 x 2", 2 ** 18, 8, 2, 20.0, 6.25e6, 600e6),
         ("cfg1x256 2^20 x 256", 2 ** 20, 256, 1, 71.0, 16e6, 6e9)]
for name, N, C, P, dm, sr, fcen in CASES:
    freqs = fcen + sr * (np.arange(C) + 0.5 - C / 2)
    plan = L.DedispPlan(nsamp=N, nchan=C, npol=P, dm=dm, sample_rate_hz=sr, ref_freq_hz=fcen,
                        chan_freq_hz=freqs, crop=(0, N))
    x = torch.randn((N, C, P, 2), device=dev)
    y = torch.empty_like(x)
    with torch.cuda.stream(stream):
        st = stream.cuda_stream
        for _ in range(5):
            plan.exec_device(x.data_ptr(), y.data_ptr(), None, st)
        stream.synchronize()
        want = y.clone()

        def timed(run, reps=40):
            ts = []
            for _ in range(reps):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                run()
                b.record(stream)
                stream.synchronize()
                ts.append(a.elapsed_time(b) * 1e3)
            return float(np.median(ts)), float(np.min(ts))

        direct = timed(lambda: plan.exec_device(x.data_ptr(), y.data_ptr(), None, st))
        same = bool(torch.equal(y, want))
        graph = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(graph, stream=stream):
                plan.exec_device(x.data_ptr(), y.data_ptr(), None,
                                 torch.cuda.current_stream().cuda_stream)
            y.zero_()
            replay = timed(graph.replay)
            gsame = bool(torch.equal(y, want))
        except Exception as e:                                  # noqa: BLE001
            replay, gsame = (float("nan"), float("nan")), f"capture failed: {e}"
    print(f"PBK_PDL={os.environ.get('PBK_PDL', '1')} {name}: [{plan.describe()}] "
          f"launches median {direct[0]:.1f} us (min {direct[1]:.1f}), output equal {same}; "
          f"graph replay median {replay[0]:.1f} us (min {replay[1]:.1f}), output equal {gsame}",
          flush=True)
    plan.destroy()
