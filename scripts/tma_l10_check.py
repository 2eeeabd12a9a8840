"""The TMA-pipelined 2^10-point MID pass (pbk_tma_l10.cu) against the LDG kernel on the same plans:
outputs must be EQUAL bit for bit; prints the per-pass times of both.

    python scripts/tma_l10_check.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from tma_check import run  # noqa: E402

shapes = [
    # N, C, P, dm, sr, fcen, out_kind, ds, crop, levels
    (2 ** 18, 32, 2, 3.0, 6.25e6, 600e6, 0, 1, None, None),
    (2 ** 18, 64, 1, 3.0, 6.25e6, 600e6, 0, 1, (77, 2 ** 18 - 101), "8,10"),     # two chirps per pair
    (2 ** 20, 4, 2, 3.0, 6.25e6, 600e6, 1, 1, None, "10,10"),                    # one tile per row
    (2 ** 22, 32, 2, 30.0, 1.5625e6, 600e6, 1, 1, (100000, 2 ** 22 - 300000), "6,6,10"),
    (2 ** 24, 32, 2, 300.0, 1.5625e6, 600e6, 1, 1, None, "6,8,10"),              # cfg5-like rows
]
bad = 0
for (N, C, P, dm, sr, fcen, ok, ds, crop, levels) in shapes:
    if levels:
        os.environ["PBK_LEVELS"] = levels
    a, ta, da = run(N, C, P, dm, sr, fcen, ok, ds, "0", crop)
    b, tb, db = run(N, C, P, dm, sr, fcen, ok, ds, "mid", crop)
    os.environ.pop("PBK_LEVELS", None)
    same = bool(torch.equal(a, b))
    bad += not same
    print(f"N=2^{int(np.log2(N))} C={C} P={P} out={ok} crop={crop} levels={levels or 'auto'}: "
          f"{'EQUAL' if same else 'DIFFERENT'}\n   ldg {np.round(ta, 4).tolist()} [{da}]\n"
          f"   tma {np.round(tb, 4).tolist()} [{db}]", flush=True)
print("FAILED" if bad else "all equal")
sys.exit(1 if bad else 0)
