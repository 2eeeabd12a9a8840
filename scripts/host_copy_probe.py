"""Host-array calls with PAGEABLE numpy blocks: plain cudaMemcpy staging (PBK_BOUNCE=0) against
the multi-threaded bounce pipeline of csrc/pbk_hostcopy.h, for several thread counts."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pulsarbat_b200 as pb  # noqa: E402

N, C, P = 2 ** 22, 16, 2
sr, fcen, dm = 6.25e6, 600e6, 10.0
rng = np.random.default_rng(1)
x = rng.standard_normal((N, C, P, 2), dtype=np.float32).view(np.complex64).reshape(N, C, P)
freqs = fcen + sr * (np.arange(C) + 0.5 - C / 2)
kw = dict(dm=dm, sample_rate_hz=sr, chan_freq_hz=freqs, ref_freq_hz=fcen, crop=None)
gb = x.nbytes / 1e9
print(f"block {gb:.2f} GB complex64 in, {gb:.2f} GB out, pageable numpy both ways", flush=True)
ref = None
for mode, threads in [("0", None), ("1", 1), ("1", 2), ("1", 4), ("1", 8), ("1", 12), ("1", 16)]:
    os.environ["PBK_BOUNCE"] = mode
    if threads:
        os.environ["PBK_BOUNCE_THREADS"] = str(threads)
    y = pb.kernels.dedisperse(x, **kw)         # warm-up (plan, buffers, lanes)
    ts = []
    for _ in range(3):
        del y
        t0 = time.perf_counter()
        y = pb.kernels.dedisperse(x, **kw)     # fresh pageable result every call
        ts.append(time.perf_counter() - t0)
    if ref is None:
        ref = y.copy()
    assert np.array_equal(y, ref)
    t = min(ts)
    print(f"PBK_BOUNCE={mode} threads={threads}: {t * 1e3:8.1f} ms per call, "
          f"{2 * gb / t:6.1f} GB/s host<->device (in+out), {N * C * P / t / 1e9:.2f} Gsamples/s",
          flush=True)
