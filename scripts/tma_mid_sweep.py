"""MID pass on the TMA ring vs the LDG kernel for every tile length (level split forced so that
the last level is 2^l): per-pass times, outputs compared bit for bit."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from tma_check import run  # noqa: E402

for l3, n in ((6, 22), (7, 23), (8, 22), (9, 23)):
    N = 2 ** n
    os.environ["PBK_LEVELS"] = {6: "8,8,6", 7: "8,8,7", 8: "7,7,8", 9: "7,7,9"}[l3]
    for rep in range(2):
        a, ta, da = run(N, 64, 2, 50.0, 6.25e6, 600e6, 2, 64, "fwd,inv,tsum,final", None, iters=5)
        b, tb, db = run(N, 64, 2, 50.0, 6.25e6, 600e6, 2, 64, "fwd,inv,mid,tsum,final", None, iters=5)
        err = float((a - b).abs().max() / a.abs().max())
        print(f"MID 2^{l3} (N=2^{n} x 64 x 2): ldg {ta[2]:.3f} ms, tma {tb[2]:.3f} ms; step {ta.sum():.3f} / "
              f"{tb.sum():.3f}; max rel diff {err:.1e} [{db.split(';')[2]}]", flush=True)
