"""Developer check: the TMA-pipelined pass kernels (pbk_tma.cuh) against the LDG kernels
(pbk_fast.cuh) on the same plans.  The arithmetic per element is identical, so complex64 and
intensity outputs must be EQUAL bit for bit; the fused time sum adds its (at most two) group
contributions with float atomics, whose order is free, so it is compared to 1e-6.
Prints one line per shape with the per-pass timings of both variants.

    python scripts/tma_check.py [quick]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pulsarbat_b200 import _lib as L  # noqa: E402


def run(N, C, P, dm, sr, fcen, out_kind, ds, tma, crop=None, iters=3):
    os.environ["PBK_TMA"] = tma
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
    plan.exec_device(x.data_ptr(), out.data_ptr(), None, st)
    torch.cuda.synchronize()
    plan.profile(iters)
    for _ in range(iters):
        plan.exec_device(x.data_ptr(), out.data_ptr(), None, st)
    torch.cuda.synchronize()
    seg = np.array([plan.profile_read(i) for i in range(iters)]).min(axis=0)
    desc = plan.describe()
    plan.destroy()
    dt = torch.float32 if out_kind else torch.complex64
    return out[:nout].view(dt).clone(), seg, desc


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    shapes = [
        # N, C, P, dm, sr, fcen, out_kind, ds, crop
        (2 ** 16, 16, 2, 3.0, 6.25e6, 600e6, 0, 1, None),             # l8,l8 / small
        (2 ** 18, 16, 2, 3.0, 6.25e6, 600e6, 2, 16, None),            # tsum stokes
        (2 ** 18, 32, 2, 3.0, 6.25e6, 600e6, 1, 8, (1000, 2 ** 18 - 3000)),   # tsum intensity, crop
        (2 ** 20, 64, 1, 3.0, 6.25e6, 600e6, 0, 1, (77, 2 ** 20 - 101)),      # single pol (TWOCH mid)
        (2 ** 19, 8, 2, 3.0, 6.25e6, 600e6, 1, 1, None),              # W = 16 levels
        (2 ** 21, 32, 2, 10.0, 6.25e6, 600e6, 2, 64, None),
    ]
    if not quick:
        shapes += [(2 ** 22, 64, 2, 100.0, 6.25e6, 600e6, 2, 64, None),      # cfg2
                   (2 ** 22, 64, 2, 100.0, 6.25e6, 600e6, 0, 1, None),       # cfg2 voltages
                   (2 ** 22, 128, 2, 100.0, 390625.0, 600e6 - 175e6, 0, 1, (196979, 3631508))]
    bad = 0
    for (N, C, P, dm, sr, fcen, ok, ds, crop) in shapes:
        a, ta, da = run(N, C, P, dm, sr, fcen, ok, ds, "0", crop)
        for levels in ([None] if quick else [None, "7,7,%d" % (int(np.log2(N)) - 14),
                                             "9,%d,6" % (int(np.log2(N)) - 15)]):
            if levels:
                if int(levels.split(",")[-1]) < 4 and levels.startswith("7"):
                    continue
                os.environ["PBK_LEVELS"] = levels
                a, ta, da = run(N, C, P, dm, sr, fcen, ok, ds, "0", crop)
            b, tb, db = run(N, C, P, dm, sr, fcen, ok, ds, "1", crop)
            os.environ.pop("PBK_LEVELS", None)
            same = bool(torch.equal(a, b))
            if ds > 1:
                err = float((a - b).abs().max() / a.abs().max())
                good = err < 1e-6
            else:
                err, good = (0.0 if same else float((a - b).abs().max())), same
            bad += not good
            ntma = db.count("tma-r16")
            print(f"N=2^{int(np.log2(N))} C={C} P={P} out={ok} ds={ds} crop={crop} "
                  f"levels={levels or 'auto'}: tma passes {ntma}/{len(db.split(';'))} "
                  f"{'EQUAL' if same else f'maxdiff {err:.2e}'} {'ok' if good else 'MISMATCH'}\n"
                  f"    ldg {np.round(ta, 3).tolist()} sum {ta.sum():.3f} ms\n"
                  f"    tma {np.round(tb, 3).tolist()} sum {tb.sum():.3f} ms   [{db}]", flush=True)
    print("tma_check:", "FAILED" if bad else "all ok")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
