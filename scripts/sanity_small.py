"""Small end-to-end run of every kernel family (a tool target: ncu, asserts; compute-sanitizer is
closed on the GPU pool, so memory safety is checked by the guard-band tests in
tests/test_gpu_kernels.py instead)."""
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

# ---- the other kernel families, small, through the Python layer over the C ABI ---------------
os.environ.pop("PBK_LEVELS", None)
import pulsarbat_b200 as pb  # noqa: E402
from pulsarbat_b200 import kernels as K  # noqa: E402

u = pb.units
kw = dict(dm=1.0, sample_rate_hz=1e6, ref_freq_hz=600e6)
N, C = 2 ** 15, 32
freqs = 600e6 + 1e6 * (np.arange(C) + 0.5 - C / 2)
x = (rng.standard_normal((N, C, 2)) + 1j * rng.standard_normal((N, C, 2))).astype(np.complex64)
# fused time sum in the last pass, aligned and unaligned crop starts
for crop in [(0, N), (37, N - 11)]:
    y = K.dedisperse(x, chan_freq_hz=freqs, crop=crop, out_kind=L.OUT_STOKES_I, downsample=8, **kw)
    print("tsum", crop, y.shape, float(y.sum()), flush=True)
# raw baseband decoded in the first pass
r8 = rng.integers(-127, 128, (N, C, 2, 2), dtype=np.int8)
print("int8", float(np.abs(K.dedisperse(r8, chan_freq_hz=freqs, raw="int8", **kw)).sum()), flush=True)
r4 = rng.integers(0, 256, (N, C, 2), dtype=np.uint8)
print("u4", float(np.abs(K.dedisperse(r4, chan_freq_hz=freqs, raw="u4", **kw)).sum()), flush=True)
r2 = rng.integers(0, 256, (N, C), dtype=np.uint8)
print("u2", float(np.abs(K.dedisperse(r2, chan_freq_hz=freqs, raw="u2", raw_shape=(C, 2), **kw)).sum()),
      flush=True)
# few channels (narrow tiles), single polarisation (two channels per lane pair), odd lane count
for shp in [(2 ** 14, 4, 2), (2 ** 14, 16), (2 ** 12, 5, 2), (4233, 3, 2), (1023,)]:
    xs = (rng.standard_normal(shp) + 1j * rng.standard_normal(shp)).astype(np.complex64)
    if len(shp) > 1:
        f = 600e6 + 1e6 * (np.arange(shp[1]) + 0.5 - shp[1] / 2)
        print("dd", shp, float(np.abs(K.dedisperse(xs, chan_freq_hz=f, **kw)).sum()), flush=True)
    print("fft", shp, float(np.abs(K.fft(xs)).sum()), float(np.abs(K.fft(xs, inverse=True)).sum()),
          flush=True)
# channelizer / unchannelizer, power-of-two and Bluestein segment lengths
xc = (rng.standard_normal((4224, 4, 2)) + 1j * rng.standard_normal((4224, 4, 2))).astype(np.complex64)
for n in (32, 33, 128):
    yc = K.stft(xc, n)
    print("stft", n, yc.shape, float(np.abs(K.istft(yc, n) - xc).max()), flush=True)
# shifts, mixing, analytic signal, gather, detection, Stokes, pol basis, time sum
print("ramp", float(np.abs(K.phase_ramp(xc.reshape(4224, 8), shift_samples=np.linspace(-3, 3, 8))).sum()))
print("mix", float(np.abs(K.mix(xc.reshape(4224, 8), np.linspace(-0.1, 0.1, 8))).sum()))
print("r2c", float(np.abs(K.analytic_decimate(rng.standard_normal((4096, 6)).astype(np.float32))).sum()))
print("roll", float(K.shift_channels(rng.standard_normal((4096, 8)).astype(np.float32),
                                     np.arange(8) * 3, 4000).sum()))
print("det", float(K.detect(x, stokes=True, downsample=4).sum()), float(K.detect(x).sum()))
print("stokes", float(K.stokes(x, "linear").sum()), float(np.abs(K.pol_basis(x, True)).sum()))
print("sum", float(K.downsample(np.abs(x) ** 2, 16).sum()), flush=True)
# fold: wide rows, narrow rows; device phase predictor
inten = rng.standard_normal((2 ** 15, 256)).astype(np.float32)
for d in (inten, inten[:, :2].copy()):
    prof, cnt, bins = K.fold(d, [0.123, 29.7, 1e-6], 1e4, 128, want_bins=True)
    print("fold", d.shape, float(prof.sum()), int(cnt.sum()), int(bins.max()), flush=True)
pi, pf = K.predict_phase([0.3, 641.9, 1e-9, 1e-13], 146774936445, nsamp=4096, sample_rate_hz=1e3)
print("phase", int(pi[-1]), float(pf[-1]), flush=True)
print("sanity ok")
