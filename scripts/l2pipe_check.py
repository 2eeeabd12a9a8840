"""Developer check: the L2-resident pipeline of the three middle passes (PBK_L2PIPE=1,
csrc/pbk_l2pipe.cuh) against three separate launches: outputs must be EQUAL bit for bit (same
per-tile arithmetic); prints per-segment timings of both schedules.

    python scripts/l2pipe_check.py [quick]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pulsarbat_b200 import _lib as L  # noqa: E402


def run(N, C, P, dm, sr, fcen, out_kind, ds, env, crop=None, iters=5):
    for k in ("PBK_L2PIPE", "PBK_TMA", "PBK_LEVELS", "PBK_NO_PINGPONG"):
        os.environ.pop(k, None)
    os.environ.update(env)
    dev = torch.device("cuda:0")
    freqs = fcen + sr * (np.arange(C) + 0.5 - C / 2)
    plan = L.DedispPlan(nsamp=N, nchan=C, npol=P, dm=dm, sample_rate_hz=sr, ref_freq_hz=fcen,
                        chan_freq_hz=freqs, crop=crop or (0, N), out_kind=out_kind, downsample=ds)
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    x = torch.randn((N, C, P, 2), device=dev, dtype=torch.float32, generator=g)
    nout = plan.out_rows * plan.row_elems * plan.elem_bytes
    out = torch.zeros(max(nout, 16), device=dev, dtype=torch.uint8)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        plan.exec_device(x.data_ptr(), out.data_ptr(), None, st)
    torch.cuda.synchronize()
    plan.profile(iters)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        plan.exec_device(x.data_ptr(), out.data_ptr(), None, st)
    e1.record()
    torch.cuda.synchronize()
    seg = np.array([plan.profile_read(i) for i in range(iters)]).mean(axis=0)
    desc = plan.describe()
    plan.destroy()
    dt = torch.float32 if out_kind else torch.complex64
    return out[:nout].view(dt).clone(), seg, e0.elapsed_time(e1) / iters, desc


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    shapes = [(2 ** 20, 64, 2, 3.0, 6.25e6, 600e6, 0, 1, (77, 2 ** 20 - 101)),
              (2 ** 20, 64, 1, 3.0, 6.25e6, 600e6, 1, 1, None)]
    if not quick:
        shapes += [(2 ** 22, 64, 2, 100.0, 6.25e6, 600e6, 2, 64, None),          # cfg2
                   (2 ** 22, 32, 2, 100.0, 6.25e6, 600e6, 0, 1, None),           # 8 MB blocks
                   (2 ** 22, 128, 2, 100.0, 390625.0, 600e6 - 175e6, 0, 1, (196979, 3631508)),
                   (2 ** 20, 256, 1, 71.0, 16e6, 6e9, 0, 1, None)]               # cfg1 x 256
    bad = 0
    for (N, C, P, dm, sr, fcen, ok, ds, crop) in shapes:
        a, ta, ma, da = run(N, C, P, dm, sr, fcen, ok, ds, {}, crop)
        for extra in ({}, {"PBK_NO_PINGPONG": "1"}):
            b, tb, mb, db = run(N, C, P, dm, sr, fcen, ok, ds, {"PBK_L2PIPE": "1", **extra}, crop)
            same = bool(torch.equal(a, b))
            bad += not same
            print(f"N=2^{int(np.log2(N))} C={C} P={P} out={ok} ds={ds} {extra}: "
                  f"{'EQUAL' if same else 'MISMATCH maxdiff %.3e' % float((a - b).abs().max())}\n"
                  f"    plain  {np.round(ta, 3).tolist()} total {ma:.3f} ms\n"
                  f"    l2pipe {np.round(tb, 3).tolist()} total {mb:.3f} ms   [{db}]", flush=True)
    print("l2pipe_check:", "FAILED" if bad else "all ok")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
