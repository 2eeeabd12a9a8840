"""The warp-private time-summing last pass (pbk_tsumw.cuh) against the thread-group TMA kernel and
the LDG kernel on the same plans: outputs must be EQUAL bit for bit; per-pass times of all three."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from tma_check import run  # noqa: E402

shapes = [
    # N, C, P, out_kind, ds, crop, levels
    (2 ** 18, 16, 2, 2, 16, None, None),
    (2 ** 20, 32, 2, 1, 8, (1000, 2 ** 20 - 3000), "8,6,6"),
    (2 ** 20, 32, 2, 2, 64, (12345, 2 ** 20 - 777), "8,6,6"),
    (2 ** 21, 64, 1, 1, 4, None, "8,7,6"),
    (2 ** 22, 64, 2, 2, 64, None, None),
    (2 ** 22, 64, 2, 1, 64, None, None),
]
bad = 0
for (N, C, P, ok, ds, crop, levels) in shapes:
    if levels:
        os.environ["PBK_LEVELS"] = levels
    os.environ.pop("PBK_TSUMW", None)
    a, ta, da = run(N, C, P, 3.0, 6.25e6, 600e6, ok, ds, "0", crop, iters=5)
    b, tb, db = run(N, C, P, 3.0, 6.25e6, 600e6, ok, ds, "1", crop, iters=5)
    os.environ["PBK_TSUMW"] = "1"
    c, tc, dc = run(N, C, P, 3.0, 6.25e6, 600e6, ok, ds, "1", crop, iters=5)
    os.environ.pop("PBK_LEVELS", None)
    same = bool(torch.equal(a, c)) and bool(torch.equal(b, c))
    bad += not same
    err = float((a - c).abs().max() / a.abs().max())
    print(f"N=2^{int(np.log2(N))} C={C} P={P} out={ok} ds={ds} crop={crop} levels={levels or 'auto'}: "
          f"{'EQUAL' if same else 'DIFFERENT (max rel %.1e)' % err}; last pass ldg {ta[-1]:.4f} tma "
          f"{tb[-1]:.4f} warp {tc[-1]:.4f} ms [{dc.split(';')[-1]}]", flush=True)
print("FAILED" if bad else "all equal")
sys.exit(1 if bad else 0)
