"""GPU tests of the public API (the mirror of the reference's interface): numpy blocks, device
arrays (DLPack / torch tensors), concurrent calls from a thread pool (what dask's threaded
scheduler does, reference transforms/transforms.py:49-50), overlap-save, polarisation, folding."""

import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest
import scipy.fft

from oracle import pbk_oracle as orc

pytestmark = pytest.mark.gpu


def relerr(a, b):
    a = np.asarray(a).astype(np.complex128 if np.iscomplexobj(a) else np.float64)
    b = np.asarray(b).astype(a.dtype)
    assert a.size == b.size, (a.shape, b.shape)
    b = b.reshape(a.shape)
    nb = np.linalg.norm(b.ravel())
    return np.linalg.norm((a - b).ravel()) / (nb if nb > 0 else 1.0)


def crandn(rng, shape):
    x = np.empty(shape, np.complex64)
    x.real = rng.standard_normal(shape, dtype=np.float32)
    x.imag = rng.standard_normal(shape, dtype=np.float32)
    return x


def _dualpol(pb, x, sr, fcen, **kw):
    u = pb.units
    return pb.DualPolarizationSignal(x, sample_rate=sr * u.Hz, center_freq=fcen * u.Hz,
                                     pol_type="linear", **kw)


def test_device_array_round_trip_matches_numpy_path():
    import torch

    import pulsarbat_b200 as pb
    rng = np.random.default_rng(2)
    N, C = 2 ** 15, 8
    sr, fcen, dm = 6.25e6, 625e6, 1.5
    x = crandn(rng, (N, C, 2))
    t0 = pb.Time(58245.0)
    z = _dualpol(pb, x, sr, fcen, start_time=t0)
    zd = z.to_device(0)
    assert isinstance(zd.data, pb.DeviceArray) and zd.shape == z.shape and zd.dtype == z.dtype
    y_host = pb.coherent_dedispersion(z, pb.DM(dm))
    y_dev = pb.coherent_dedispersion(zd, pb.DM(dm))
    assert isinstance(y_dev.data, pb.DeviceArray)
    assert type(y_dev) is type(z) and y_dev.shape == y_host.shape
    assert y_dev.start_time.isclose(y_host.start_time)
    assert np.array_equal(np.asarray(y_dev.data), np.asarray(y_host.data))   # same kernels
    want, s0, s1 = orc.coherent_dedispersion(x.astype(np.complex128), dm, sample_rate=sr,
                                             center_freq=fcen)
    assert y_host.shape[0] == s1 - s0
    assert relerr(np.asarray(y_dev.data), want) < 1e-5
    # detection and Stokes stay on the device
    i_dev = y_dev.to_intensity()
    assert isinstance(i_dev.data, pb.DeviceArray)
    assert relerr(np.asarray(i_dev.data), orc.to_intensity(want)) < 1e-5
    s_dev = y_dev.to_stokes()
    assert relerr(np.asarray(s_dev.data), orc.to_stokes(want, "linear")) < 1e-5
    # a torch tensor enters through DLPack without a copy
    t = torch.view_as_complex(torch.randn((4096, 4, 2, 2), device="cuda:0"))
    da = pb.DeviceArray.from_dlpack(t)
    assert da.ptr == t.data_ptr() and da.shape == (4096, 4, 2)
    zt = _dualpol(pb, da, 1e6, 600e6)
    yt = pb.coherent_dedispersion(zt, pb.DM(0.2))
    wt, _, _ = orc.coherent_dedispersion(t.cpu().numpy().astype(np.complex128), 0.2,
                                         sample_rate=1e6, center_freq=600e6)
    assert relerr(np.asarray(yt.data), wt) < 1e-5
    assert np.asarray(yt.compute().data).shape == wt.shape       # Signal.compute() -> numpy


def test_concurrent_block_calls_like_dask_threads():
    """Channel chunks of one signal dedispersed from 8 threads at once (plan cache + per-plan
    locks), each with the GLOBAL ref_freq and crop: concatenating the blocks along frequency
    reproduces the whole-band result."""
    import pulsarbat_b200 as pb
    from pulsarbat_b200 import sharding
    rng = np.random.default_rng(3)
    N, C = 2 ** 14, 32
    sr, fcen, dm = 1e6, 600e6, 0.4
    x = crandn(rng, (N, C, 2))
    z = _dualpol(pb, x, sr, fcen)
    whole = pb.coherent_dedispersion(z, pb.DM(dm))
    start, stop, ref = sharding.dedispersion_crop(z, pb.DM(dm))

    def block(r):
        zs, (lo, hi) = sharding.shard_channels(z, 8, r)
        y = pb.kernels.dedisperse(zs.data, dm=dm, sample_rate_hz=sr,
                                  chan_freq_hz=zs.channel_freqs_hz, ref_freq_hz=fcen,
                                  crop=(start, stop))
        return lo, np.asarray(y)

    for _ in range(3):
        with ThreadPoolExecutor(max_workers=8) as ex:
            parts = sorted(ex.map(block, range(8)), key=lambda t: t[0])
        got = np.concatenate([p for _, p in parts], axis=1)
        assert got.shape == whole.shape
        assert relerr(got, np.asarray(whole.data)) < 1e-6
    want, _, _ = orc.coherent_dedispersion(x.astype(np.complex128), dm, sample_rate=sr,
                                           center_freq=fcen)
    assert relerr(got, want) < 1e-5


def test_overlap_save_matches_blockwise_oracle():
    import pulsarbat_b200 as pb
    rng = np.random.default_rng(15)
    N, C, L = 2 ** 15, 4, 2 ** 13
    sr, fcen, dm = 1e6, 600e6, 0.3
    x = crandn(rng, (N, C, 2))
    z = _dualpol(pb, x, sr, fcen, start_time=pb.Time(58000.0))
    y = pb.overlap_save_dedispersion(z, pb.DM(dm), L)
    want = orc.overlap_save_dedispersion(x.astype(np.complex128), dm, L, sample_rate=sr,
                                         center_freq=fcen)
    want = want[0] if isinstance(want, tuple) else want
    assert y.shape == want.shape
    assert relerr(np.asarray(y.data), want) < 1e-5


def test_polarisation_known_vectors_and_round_trip():
    """reference tests/test_polarization.py:34-68 (hand-computed Stokes vectors) on the GPU."""
    import pulsarbat_b200 as pb
    u = pb.units
    x = np.array([[[1 + 0j, 0j]], [[1 + 0j, 1 + 0j]], [[1 + 0j, 1j]]], np.complex64)
    z = pb.DualPolarizationSignal(x, sample_rate=1 * u.Hz, center_freq=1e9 * u.Hz,
                                  pol_type="linear")
    s = np.asarray(z.to_stokes().data)[:, 0]
    assert np.allclose(s, [[1, 1, 0, 0], [2, 0, 2, 0], [2, 0, 0, 2]])
    zc = z.to_circular()
    assert zc.pol_type == "circular"
    assert np.allclose(np.asarray(zc.to_stokes().data)[:, 0], s, atol=1e-6)
    assert np.allclose(np.asarray(zc.to_linear().data), x, atol=1e-6)
    rng = np.random.default_rng(9)
    r = crandn(rng, (1000, 6, 2))
    zr = pb.DualPolarizationSignal(r, sample_rate=1 * u.Hz, center_freq=1e9 * u.Hz,
                                   pol_type="circular")
    assert relerr(np.asarray(zr.to_stokes().data), orc.to_stokes(r, "circular")) < 1e-6
    assert relerr(np.asarray(zr.to_linear().data), orc.to_linear(r, "circular")) < 1e-6


def test_fold_with_polyco_predictor():
    """Fold through PhasePredictor.phasepol (reference predictor.py:149-160) with the reference's
    own polyco fixture; bins and counts bit-exact against the oracle."""
    import pulsarbat_b200 as pb
    u = pb.units
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "golden", "timing.dat")) as f:
        text = f.read()
    pred = pb.PhasePredictor.from_polyco(os.path.join(here, "golden", "timing.dat"))
    entries = orc.parse_polyco(text)
    t0 = pb.Time(58245.375)
    sr, nsamp, nbin = 5e4, 200_000, 128
    rng = np.random.default_rng(4)
    x = rng.random((nsamp, 3), dtype=np.float32)
    z = pb.Signal(x, sample_rate=sr * u.Hz, start_time=t0)
    prof, counts, bins = pb.fold(z, pred, nbin, want_bins=True)
    coeffs, _ = orc.phasepol(entries, (58245, 0.375))
    ref_bins = orc.fold_bins(nsamp, coeffs, sr, nbin)
    assert np.array_equal(bins, ref_bins)
    want_p, want_c = orc.fold(x, coeffs, sr, nbin)
    assert np.array_equal(counts, want_c)
    assert relerr(prof, want_p) < 1e-5


def test_real_to_complex_matches_reference_kats():
    """reference utils.py:15-65 / tests/test_utils.py:20-68 on the GPU."""
    import pulsarbat_b200 as pb
    N = 512
    t = np.linspace(0, 2 * np.pi, N, endpoint=False)
    for w in [1, 2, 127, 128, 129, 254, 255]:
        for p in [-np.pi, -np.pi / 2, 0, np.pi / 2]:
            x = np.cos(w * t + p)
            y = np.exp(1j * ((w - N / 4) * t[::2] + p))
            z = pb.utils.real_to_complex(x)
            assert z.dtype == np.complex128 and np.allclose(z, y, atol=2e-5)
    ws = [1, 2, 3, 4]
    t = np.linspace(0, 2 * np.pi, 128, endpoint=False)
    x = np.stack([np.cos(w * t) for w in ws], axis=0)
    y = np.stack([np.exp(1j * (w - 32) * t[::2]) for w in ws], axis=0)
    assert np.allclose(pb.utils.real_to_complex(x, axis=1), y, atol=2e-5)
    assert np.allclose(pb.utils.real_to_complex(x.T.copy(), axis=0), y.T, atol=2e-5)
    rng = np.random.default_rng(12)
    r = rng.standard_normal((2 ** 14, 3, 2)).astype(np.float32)
    got = pb.utils.real_to_complex(r)
    assert got.dtype == np.complex64 and got.shape == (2 ** 13, 3, 2)
    assert relerr(got, orc.real_to_complex(r.astype(np.float64))) < 1e-5
    gd = pb.utils.real_to_complex(pb.DeviceArray.from_numpy(r))
    assert isinstance(gd, pb.DeviceArray) and np.array_equal(np.asarray(gd), got)
    with pytest.raises(ValueError):
        pb.utils.real_to_complex(np.ones((128, 4), dtype=complex))
    e = np.zeros((0, 2))
    assert np.array_equal(pb.utils.real_to_complex(e), e)
    assert pb.utils.real_to_complex(np.ones(32, np.float32)).dtype == np.complex64


def test_shifts_and_real_to_complex_any_length():
    """The reference's own odd-length cases: freq_shift at N = 1023 (tests/test_transforms.py:398),
    real_to_complex at N = 511 (tests/test_utils.py:29), time_shift at a non power of two."""
    import pulsarbat_b200 as pb
    u = pb.units
    N = 1023
    n = np.arange(N) / N
    fs = np.array([[-52, -45.4], [-25.5, 34], [14, -36.9], [45.1, 27]])
    tones = np.exp(2j * np.pi * fs[None] * n[:, None, None]).astype(np.complex64)
    x = pb.BasebandSignal(tones, sample_rate=N * u.Hz, center_freq=1e6 * u.Hz)
    assert np.allclose(np.asarray(pb.freq_shift(x, -fs * u.Hz).data), 1, atol=3e-5)
    rng = np.random.default_rng(7)
    z = crandn(rng, (N, 4, 2))
    zs = pb.BasebandSignal(z, sample_rate=N * u.Hz, center_freq=1e6 * u.Hz)
    for shift in [49.0, np.array([5.0, -6.0, 300.25, -8.0])]:
        want = orc.freq_shift(z.astype(np.complex128), np.asarray(shift) / N)
        assert relerr(np.asarray(pb.freq_shift(zs, shift * u.Hz).data), want) < 1e-5
    zt = crandn(rng, (4095, 3))
    st = pb.Signal(zt, sample_rate=1e3 * u.Hz)
    for shift in [7, -3.5, np.array([1.25, -20.0, 11.0])]:
        want, a, b = orc.time_shift(zt.astype(np.complex128), shift)
        got = pb.time_shift(st, shift)
        assert relerr(np.asarray(got.data), want) < 1e-5
        assert np.array_equal(np.asarray(got.data) == 0, want == 0)
    t = np.linspace(0, 2 * np.pi, 511, endpoint=False)
    for w in [1, 2, 127, 128, 129, 254, 255]:
        for p in [-np.pi, 0, np.pi / 2]:
            y = np.exp(1j * ((w - 511 / 4) * t[::2] + p))
            assert np.allclose(pb.utils.real_to_complex(np.cos(w * t + p)), y, atol=3e-5)


def test_block_stream_matches_single_calls():
    """streaming.dedisperse_blocks (copy of block i+1 overlapped with the kernels of block i)
    returns exactly what kernels.dedisperse returns block by block."""
    import pulsarbat_b200 as pb
    from pulsarbat_b200 import _lib as L
    rng = np.random.default_rng(11)
    N, C = 2 ** 14, 16
    sr, fcen, dm = 1e6, 600e6, 0.3
    freqs = orc.channel_freqs(fcen, sr, C)
    blocks = [crandn(rng, (N, C, 2)) for _ in range(5)]
    for kind, ds in [(L.OUT_C64, 1), (L.OUT_STOKES_I, 4)]:
        kw = dict(dm=dm, sample_rate_hz=sr, chan_freq_hz=freqs, ref_freq_hz=fcen,
                  crop=(100, N - 300), out_kind=kind, downsample=ds)
        got = list(pb.streaming.dedisperse_blocks(iter(blocks), **kw))
        assert len(got) == 5
        for b, g in zip(blocks, got):
            assert np.array_equal(g, np.asarray(pb.kernels.dedisperse(b, **kw)))
    assert list(pb.streaming.dedisperse_blocks(iter([]), dm=dm, sample_rate_hz=sr,
                                               chan_freq_hz=freqs, ref_freq_hz=fcen)) == []
    # blocks large enough (64 MiB) for the pageable bounce pipeline of the upload
    big = [crandn(rng, (2 ** 18, C, 2)) for _ in range(3)]
    kwb = dict(dm=dm, sample_rate_hz=sr, chan_freq_hz=freqs, ref_freq_hz=fcen,
               crop=(1000, 2 ** 18 - 3000), out_kind=L.OUT_STOKES_I, downsample=8)
    gotb = list(pb.streaming.dedisperse_blocks(iter(big), **kwb))
    for g, b in zip(gotb, big):
        assert np.array_equal(g, pb.kernels.dedisperse(b, **kwb))
    one = list(pb.streaming.dedisperse_blocks([blocks[0]], dm=dm, sample_rate_hz=sr,
                                              chan_freq_hz=freqs, ref_freq_hz=fcen))
    assert len(one) == 1 and one[0].shape == (N, C, 2)


def test_fold_across_polyco_spans():
    """A signal longer than one polyco entry (90 min spans in the reference's fixture): each
    stretch is folded with the entry the reference would select (predictor.py:108-119).  Bins are
    bit-exact against the oracle stretch by stretch, and a pulse train generated from the
    per-sample phase of the ORIGINAL polynomials stays in its phase window across the boundaries."""
    import pulsarbat_b200 as pb
    from pulsarbat_b200.pulsar.folding import fold_segments
    u = pb.units
    here = os.path.dirname(os.path.abspath(__file__))
    path = os.path.join(here, "golden", "timing.dat")
    pred = pb.PhasePredictor.from_polyco(path)
    with open(path) as f:
        entries = orc.parse_polyco(f.read())
    t0 = pb.Time(pred.entries[0].tmid.mjd + 0.02)
    sr, nsamp, nbin = 500.0, 6_000_000, 64               # 200 minutes: four entries
    z0 = pb.Signal(np.zeros((nsamp, 1), np.float32), sample_rate=sr * u.Hz, start_time=t0)
    segs = fold_segments(z0, pred)
    assert len(segs) == 4 and sum(c for _, c, _ in segs) == nsamp
    # independent per-sample phase from the original polynomials, entry by entry
    frac = np.empty(nsamp)
    for first, count, _ in segs:
        tf = t0 + (first / sr) * u.s
        _, fr = orc.predict_phase(entries, (tf.jd1, tf.jd2), np.arange(count) / sr)
        frac[first:first + count] = fr - np.floor(fr)
    x = ((frac >= 0.2) & (frac < 0.25)).astype(np.float32)[:, None]
    z = pb.Signal(x, sample_rate=sr * u.Hz, start_time=t0)
    prof, counts, bins = pb.fold(z, pred, nbin, want_bins=True)
    ref_bins = np.concatenate([
        orc.fold_bins(c, orc.phasepol(entries, ((t0 + (f / sr) * u.s).jd1,
                                                (t0 + (f / sr) * u.s).jd2))[0], sr, nbin)
        for f, c, _ in segs])
    assert np.array_equal(bins, ref_bins)
    assert np.array_equal(counts, np.bincount(ref_bins, minlength=nbin))
    inside = prof[12:16, 0].sum()
    assert inside / prof.sum() > 0.999 and prof.sum() == pytest.approx(x.sum(), rel=1e-6)


def test_explicit_chirp_equals_generated_chirp():
    """reference tests/test_dedispersion.py:141-164 through the public API: the chirp returned by
    ``DM.chirp_from_signal`` passed back as ``chirp=`` gives the same result as the generated one,
    for every reference frequency, also when squeezed to 2-D; and it matches the oracle's
    complex64 chirp of dedispersion.py:19-23."""
    import pulsarbat_b200 as pb
    u = pb.units
    rng = np.random.default_rng(5)
    shape = (8192, 4, 2)
    x = crandn(rng, shape)
    z = pb.DualPolarizationSignal(x, sample_rate=1e6 * u.Hz, center_freq=1e9 * u.Hz,
                                  pol_type="linear")
    dm = pb.DM(10.0)
    for rf in [z.center_freq, z.min_freq, z.max_freq]:
        chirp = dm.chirp_from_signal(z, ref_freq=rf)
        assert chirp.shape == (8192, 4, 1) and chirp.dtype == np.complex64
        want_c = orc.chirp_from_signal(10.0, 8192, 1e6, z.channel_freqs_hz,
                                       float(u.to_value(rf, u.Hz)))
        assert np.max(np.abs(chirp[:, :, 0] - want_c)) < 2e-6
        y1 = pb.coherent_dedispersion(z, dm, ref_freq=rf)
        y2 = pb.coherent_dedispersion(z, dm, ref_freq=rf, chirp=chirp)
        y3 = pb.coherent_dedispersion(z, dm, ref_freq=rf, chirp=chirp[:, :, 0])
        assert y1.shape == y2.shape == y3.shape
        assert relerr(np.asarray(y2.data), np.asarray(y1.data)) < 2e-6
        assert np.array_equal(np.asarray(y3.data), np.asarray(y2.data))
    one = dm.chirp_function(8192, z.dt, z.channel_freqs[0], z.center_freq)
    assert one.shape == (8192,)
    assert np.max(np.abs(one - orc.transfer_function(10.0, 8192, 1e6, z.channel_freqs_hz[0],
                                                     1e9))) < 2e-6


def test_pinned_result_arrays():
    """kernels.pinned_results(True): results live in page-locked memory, values unchanged."""
    import pulsarbat_b200 as pb
    rng = np.random.default_rng(12)
    N, C = 2 ** 12, 4
    sr, fcen, dm = 1e6, 600e6, 0.2
    x = crandn(rng, (N, C, 2))
    kw = dict(dm=dm, sample_rate_hz=sr, chan_freq_hz=orc.channel_freqs(fcen, sr, C),
              ref_freq_hz=fcen, crop=(10, N - 20))
    plain = pb.kernels.dedisperse(x, **kw)
    old = pb.kernels.pinned_results(True)
    try:
        assert old is False
        pinned = pb.kernels.dedisperse(x, **kw)
        spec = pb.kernels.fft(x)
    finally:
        assert pb.kernels.pinned_results(old) is True
    assert pinned.flags.c_contiguous and pinned.flags.writeable and pinned.dtype == plain.dtype
    assert np.array_equal(pinned, plain)
    assert spec.shape == x.shape
    pinned += 1          # ordinary numpy array semantics
    keep = pinned[5:7].copy()
    del pinned
    assert np.isfinite(keep).all()


def test_device_phase_prediction_bit_exact():
    """pbk_phase_predict = PhasePredictor.__call__ (reference pulsar/predictor.py:121-147) for the
    samples of a block: bit-equal to numpy's polyval of the entry polynomial and to the int/frac
    split of pulsar/phase.py, for given offsets and for generated sample times, on host arrays and
    device arrays; across polyco entries it follows the entry the reference selects."""
    import pulsarbat_b200 as pb
    u = pb.units
    here = os.path.dirname(os.path.abspath(__file__))
    path = os.path.join(here, "golden", "timing.dat")
    pred = pb.PhasePredictor.from_polyco(path)
    with open(path) as f:
        entries = orc.parse_polyco(f.read())
    e = pred.entries[3]
    rng = np.random.default_rng(23)
    dt = rng.uniform(-2700.0, 2700.0, 100_001)
    want = orc.polyval_numpy(dt, e.poly.coef)
    pi, pf = pb.kernels.predict_phase(e.poly.coef, e.rphase, dt_s=dt)
    assert pi.dtype == np.int64 and pf.dtype == np.float64
    assert np.array_equal(pi, e.rphase + np.rint(want).astype(np.int64))
    assert np.array_equal(pf, want - np.rint(want))
    assert np.all(np.abs(pf) <= 0.5)
    di, df = pb.kernels.predict_phase(e.poly.coef, e.rphase, dt_s=pb.DeviceArray.from_numpy(dt))
    assert np.array_equal(np.asarray(di), pi) and np.array_equal(np.asarray(df), pf)
    # generated sample times: dt0 + (n0 + i) / sample_rate
    sr, n0, n = 1234.5, 77, 50_000
    gi, gf = pb.kernels.predict_phase(e.poly.coef, e.rphase, nsamp=n, dt0_s=-100.25,
                                      sample_rate_hz=sr, n0=n0)
    w2 = orc.polyval_numpy(-100.25 + (n0 + np.arange(n)) / sr, e.poly.coef)
    assert np.array_equal(gi, e.rphase + np.rint(w2).astype(np.int64))
    assert np.array_equal(gf, w2 - np.rint(w2))
    # a block spanning four entries, against the oracle's predictor stretch by stretch
    t0 = pb.Time(pred.entries[0].tmid.mjd + 0.02)
    sr, nsamp = 500.0, 6_000_000
    si, sf = pred.sample_phases(t0, nsamp, sr * u.Hz)
    from pulsarbat_b200.pulsar.folding import fold_segments
    z0 = pb.Signal(np.zeros((nsamp, 1), np.float32), sample_rate=sr * u.Hz, start_time=t0)
    segs = fold_segments(z0, pred)
    assert len(segs) == 4
    for first, count, _ in segs:
        tf = t0 + (first / sr) * u.s
        oi, of = orc.predict_phase(entries, (tf.jd1, tf.jd2), np.arange(count) / sr)
        assert np.array_equal(si[first:first + count], oi)
        # the kernel adds the sample offset to dt before the polynomial exactly as the oracle does
        assert np.array_equal(sf[first:first + count], of)
    with pytest.raises(ValueError):
        pred.sample_phases(pb.Time(pred.entries[-1].tmid.mjd + 1.0), 10, sr * u.Hz)
    ddi, ddf = pred.sample_phases(t0, 100_000, sr * u.Hz, on_device=True)
    assert np.array_equal(np.asarray(ddi), si[:100_000]) and np.array_equal(np.asarray(ddf), sf[:100_000])


def test_raw_packed_blocks_through_the_api():
    """kernels.dedisperse / streaming.dedisperse_blocks with raw="u4" / "u2" / "int8" blocks give
    what the complex64 path gives on the decoded samples (fused Stokes I + time sum included)."""
    import pulsarbat_b200 as pb
    from pulsarbat_b200 import _lib as L
    rng = np.random.default_rng(31)
    N, C = 2 ** 14, 16
    sr, fcen, dm = 1e6, 600e6, 0.3
    kw = dict(dm=dm, sample_rate_hz=sr, chan_freq_hz=orc.channel_freqs(fcen, sr, C),
              ref_freq_hz=fcen, crop=(128, N - 256), out_kind=L.OUT_STOKES_I, downsample=8)
    raw4 = [rng.integers(0, 256, size=(N, C, 2), dtype=np.uint8) for _ in range(3)]
    raw2 = [rng.integers(0, 256, size=(N, C), dtype=np.uint8) for _ in range(3)]
    for blocks, kind, dec, extra in [
            (raw4, "u4", orc.unpack_u4, {}),
            (raw2, "u2", lambda r: orc.unpack_u2(r).reshape(N, C, 2), {"raw_shape": (C, 2)})]:
        want = [pb.kernels.dedisperse(dec(b), **kw) for b in blocks]
        one = pb.kernels.dedisperse(blocks[0], raw=kind, **extra, **kw)
        assert one.shape == want[0].shape and np.array_equal(one, want[0])
        got = list(pb.streaming.dedisperse_blocks(iter(blocks), raw=kind, **extra, **kw))
        assert len(got) == 3
        for g, w in zip(got, want):
            assert np.array_equal(g, w)
        dev = pb.kernels.dedisperse(pb.DeviceArray.from_numpy(blocks[1]), raw=kind, **extra, **kw)
        assert np.array_equal(np.asarray(dev), want[1])
    with pytest.raises(ValueError):
        pb.kernels.dedisperse(raw2[0], raw="u2", **kw)               # raw_shape missing
    with pytest.raises(ValueError):
        pb.kernels.dedisperse(raw2[0], raw="u3", **kw)


# ------------------------------------------------------------------------------------------
# complex128: FP64 arithmetic end to end (csrc/pbk_f64.cuh), never narrowed to complex64
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", [4, 8, 15, 16, 23, 42])
def test_complex128_reversibility_kat_at_the_reference_tolerance(seed):
    """reference tests/test_dedispersion.py:73-98, verbatim tolerance: dedisperse by +DM then -DM
    and recover the input to atol 3e-8 -- only possible if complex128 is computed in FP64."""
    import scipy.signal
    import pulsarbat_b200 as pb
    u = pb.units
    N, M = 2 ** 18, 2 ** 12
    R = np.random.default_rng(seed=seed)
    x = R.standard_normal(N) + 1j * R.standard_normal(N)
    x *= np.exp(-(((np.arange(N) - N // 2) / M) ** 2))
    x = scipy.signal.sosfilt(scipy.signal.butter(10, 0.45, "lowpass", fs=1.0, output="sos"), x)
    sig = pb.BasebandSignal(x.reshape(-1, 1), sample_rate=400 * u.MHz, center_freq=600 * u.MHz,
                            start_time=pb.Time(56000.0))
    assert sig.dtype == np.complex128
    temp = pb.coherent_dedispersion(sig, pb.DM(0.01))
    sig2 = pb.coherent_dedispersion(temp, -pb.DM(0.01))
    assert temp.dtype == np.complex128 and sig2.dtype == np.complex128
    toffset = sig2.start_time - sig.start_time
    noffset = int(np.rint(float((toffset * sig.sample_rate).to_value(u.one))))
    res = np.asarray(sig[noffset:noffset + len(sig2)].data) - np.asarray(sig2.data)
    assert np.allclose(res, 0, atol=3e-8), float(np.abs(res).max())


def test_complex128_paths_match_the_oracle_to_double_precision():
    import pulsarbat_b200 as pb
    from pulsarbat_b200 import kernels
    rng = np.random.default_rng(128)
    N, C = 4096, 3
    x = rng.standard_normal((N, C, 2)) + 1j * rng.standard_normal((N, C, 2))
    freqs = orc.channel_freqs(600e6, 1e6, C)
    want, s0, s1 = orc.coherent_dedispersion(x, 2.0, sample_rate=1e6, center_freq=600e6)
    for src in (x, pb.DeviceArray.from_numpy(x)):
        got = kernels.dedisperse(src, dm=2.0, sample_rate_hz=1e6, chan_freq_hz=freqs,
                                 ref_freq_hz=600e6, crop=(s0, s1))
        # (the chirp is rounded to complex64 on both sides, dedispersion.py:23: where the FP64
        # phases differ in the last place a few of its float32 values differ by one ulp)
        assert got.dtype == np.complex128 and relerr(np.asarray(got), want) < 1e-9
    chirp = orc.chirp_from_signal(2.0, N, 1e6, freqs, 600e6)
    got = kernels.dedisperse(x, dm=2.0, sample_rate_hz=1e6, chan_freq_hz=freqs, ref_freq_hz=600e6,
                             crop=(s0, s1), chirp_array=chirp)
    assert relerr(got, want) < 1e-12
    gi = kernels.dedisperse(x, dm=2.0, sample_rate_hz=1e6, chan_freq_hz=freqs, ref_freq_hz=600e6,
                            crop=(s0, s1), out_kind=2, downsample=4)
    wi = orc.downsample(orc.stokes_I(want), 4)
    assert gi.dtype == np.float64 and relerr(gi, wi) < 1e-9
    # plain transforms, channelizer, detection
    for n in (2, 8, 64, 1024, 2 ** 15):
        y = rng.standard_normal((3, n, 5)) + 1j * rng.standard_normal((3, n, 5))
        assert relerr(kernels.fft(y, axis=1), scipy.fft.fft(y, axis=1)) < 1e-13
        assert relerr(kernels.fft(y, axis=1, inverse=True), scipy.fft.ifft(y, axis=1)) < 1e-13
    xs = rng.standard_normal((16 * 64, 3, 2)) + 1j * rng.standard_normal((16 * 64, 3, 2))
    ys = kernels.stft(xs, 64)
    assert ys.dtype == np.complex128 and relerr(ys, orc.stft(xs, 64)) < 1e-13
    assert relerr(kernels.istft(ys, 64), xs) < 1e-13
    assert relerr(kernels.detect(xs), orc.to_intensity(xs)) < 1e-15
    assert relerr(kernels.detect(xs, stokes=True, downsample=4, freq_sum=3),
                  orc.stokes_I(xs).reshape(256, 4, 1, 3).sum(axis=(1, 3))) < 1e-14
    # any length (Bluestein in FP64): the reference's own tests use 4224, 4233, nperseg = 33
    for n in (3, 33, 1000, 4233):
        y = rng.standard_normal((2, n, 3)) + 1j * rng.standard_normal((2, n, 3))
        assert relerr(kernels.fft(y, axis=1), scipy.fft.fft(y, axis=1)) < 1e-12
        assert relerr(kernels.fft(y, axis=1, inverse=True), scipy.fft.ifft(y, axis=1)) < 1e-12
    xo = rng.standard_normal((4224, 2)) + 1j * rng.standard_normal((4224, 2))
    fo = orc.channel_freqs(600e6, 1e6, 2)
    wo, a0, a1 = orc.coherent_dedispersion(xo, 2.0, sample_rate=1e6, center_freq=600e6)
    go = kernels.dedisperse(xo, dm=2.0, sample_rate_hz=1e6, chan_freq_hz=fo, ref_freq_hz=600e6,
                            crop=(a0, a1))
    assert go.dtype == np.complex128 and relerr(go, wo) < 1e-9
    x33 = rng.standard_normal((33 * 8, 2, 2)) + 1j * rng.standard_normal((33 * 8, 2, 2))
    y33 = kernels.stft(x33, 33)
    assert relerr(y33, orc.stft(x33, 33)) < 1e-12 and relerr(kernels.istft(y33, 33), x33) < 1e-12
