"""Small end-to-end run for compute-sanitizer: one fast-kernel plan (3 levels) and one generic plan."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pulsarbat_b200 import _lib as L  # noqa: E402

rng = np.random.default_rng(0)
for (N, C, P, lv) in [(2 ** 18, 32, 2, "6,6,6"), (2 ** 14, 64, 2, None), (2 ** 12, 3, 1, None)]:
    if lv:
        os.environ["PBK_LEVELS"] = lv
    else:
        os.environ.pop("PBK_LEVELS", None)
    x = (rng.standard_normal((N, C, P)) + 1j * rng.standard_normal((N, C, P))).astype(np.complex64)
    freqs = 600e6 + 1e6 * (np.arange(C) + 0.5 - C / 2)
    for kind, ds in [(0, 1), (2 if P == 2 else 1, 4)]:
        plan = L.DedispPlan(nsamp=N, nchan=C, npol=P, dm=1.0, sample_rate_hz=1e6,
                            ref_freq_hz=600e6, chan_freq_hz=freqs, crop=(5, N - 7), out_kind=kind,
                            downsample=ds)
        out = plan.exec_host(x, plan.out_array())
        print(N, C, P, plan.describe()[:80], float(np.abs(out).sum()), flush=True)
        plan.destroy()
