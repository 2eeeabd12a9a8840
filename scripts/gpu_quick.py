"""Developer timing probe (not the benchmark contract -- see bench.py).

Times device-resident dedispersion plans with CUDA events on torch's current stream and prints
one line per configuration.  Usage: python scripts/gpu_quick.py [cfg ...]
"""

import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pulsarbat_b200 import _lib as L  # noqa: E402


def chan_freqs(fcen, bw, nchan):
    return fcen + bw * (np.arange(nchan) + 0.5 - nchan / 2)


def time_plan(name, N, C, P, dm, sr, fcen, out_kind=0, downsample=1, in_dtype=0,
              iters=int(os.environ.get("PBK_QUICK_ITERS", "5"))):
    dev = torch.device("cuda:0")
    freqs = chan_freqs(fcen, sr, C)
    plan = L.DedispPlan(nsamp=N, nchan=C, npol=P, dm=dm, sample_rate_hz=sr, ref_freq_hz=fcen,
                        chan_freq_hz=freqs, crop=(0, N), in_dtype=in_dtype, out_kind=out_kind,
                        downsample=downsample)
    if in_dtype == 0:
        x = torch.randn((N, C, P, 2), device=dev, dtype=torch.float32)
    else:
        x = torch.randint(-127, 128, (N, C, P, 2), device=dev, dtype=torch.int8)
    nout = plan.out_rows * plan.row_elems * plan.elem_bytes
    out = torch.empty(max(nout, 16), device=dev, dtype=torch.uint8)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        plan.exec_device(x.data_ptr(), out.data_ptr(), None, st)
    torch.cuda.synchronize()
    plan.profile(iters)
    ts = []
    flush = (torch.empty(256 << 20, device=dev, dtype=torch.uint8)
             if os.environ.get("PBK_QUICK_FLUSH") else None)     # cold-L2 timing of small plans
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if flush is not None:
            flush.zero_()
        e0.record()
        plan.exec_device(x.data_ptr(), out.data_ptr(), None, st)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    nsmp = N * C * P
    in_b = x.numel() * x.element_size()
    print(f"{name}: N=2^{int(np.log2(N))} C={C} P={P} levels={plan.info()['levels']} "
          f"median {ms:.3f} ms best {min(ts):.3f} ms -> {nsmp / ms / 1e6:.1f} Gsamples/s, "
          f"alg {(in_b + nout) / ms / 1e6:.0f} GB/s", flush=True)
    seg = np.array([plan.profile_read(i) for i in range(iters)]).mean(axis=0)
    names = plan.describe().split(";") + (
        ["downsample"] if downsample > 1 and ":timesum" not in plan.describe() else [])
    for nm, t in zip(names, seg):
        print(f"      {t:7.3f} ms  {nm}", flush=True)
    plan.destroy()
    del x, out
    torch.cuda.empty_cache()
    return ms


CFGS = {
    "cfg1": dict(N=2 ** 20, C=1, P=1, dm=71.0, sr=16e6, fcen=400e6),
    "cfg1x256": dict(N=2 ** 20, C=256, P=1, dm=71.0, sr=16e6, fcen=400e6),
    "mid": dict(N=2 ** 20, C=64, P=2, dm=100.0, sr=6.25e6, fcen=600e6),
    "cfg2_c64": dict(N=2 ** 22, C=64, P=2, dm=100.0, sr=6.25e6, fcen=600e6),
    "cfg2": dict(N=2 ** 22, C=64, P=2, dm=100.0, sr=6.25e6, fcen=600e6, out_kind=2,
                 downsample=64),
    "cfg3_1gpu": dict(N=2 ** 22, C=128, P=2, dm=100.0, sr=390625.0, fcen=600e6, in_dtype=1),
    "cfg5_shard": dict(N=2 ** 26, C=32, P=2, dm=1000.0, sr=400e6 / 256, fcen=410e6, out_kind=1),
    "n12": dict(N=2 ** 12, C=4096, P=2, dm=1.0, sr=1e6, fcen=1e9),
    "n16": dict(N=2 ** 16, C=1024, P=2, dm=1.0, sr=1e6, fcen=1e9),
}

if __name__ == "__main__":
    names = sys.argv[1:] or ["cfg1", "mid", "cfg2_c64", "cfg2"]
    print(torch.cuda.get_device_name(0), flush=True)
    for nme in names:
        t0 = time.time()
        try:
            time_plan(nme, **CFGS[nme])
        except Exception as e:  # keep going: this is a probe
            print(f"{nme}: FAILED {type(e).__name__}: {e}", flush=True)
        print(f"   (wall {time.time() - t0:.1f}s)", flush=True)
