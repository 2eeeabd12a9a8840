"""Pin the CPU oracle against the reference's own known-answer tests (no GPU).

Each test cites the reference test it restates (paths under /root/reference).  The oracle is
only trusted as a checker for the CUDA path because these pass.
"""

import os

import numpy as np
import pytest
import scipy.signal

from oracle import pbk_oracle as orc

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


# ---- tests/test_dedispersion.py:12-32 ---------------------------------------------------
def test_delay_constants():
    dm = 2.41e-4
    for f in [0.1, 1.0, 10.0]:
        assert np.isclose(orc.time_delay(dm, f * 1e6, np.inf), 1 / f / f)
        assert np.isclose(orc.time_delay(dm, np.inf, f * 1e6), -(1 / f / f))
    assert np.isclose(orc.time_delay(dm, 2e6, 1e6), -0.75)
    for sr in [1e6, 10e6, 1e3]:
        assert np.isclose(orc.sample_delay(dm, 1e6, np.inf, sr), sr)
    for a in [10, 20, 100]:
        assert np.isclose(orc.time_delay(2.41e-4 * a, 1e6, np.inf), a)


# ---- tests/test_dedispersion.py:35-71 ---------------------------------------------------
@pytest.mark.parametrize("dm", [10.0, 50.0, 100.0])
def test_crop_inequalities(dm):
    rng = np.random.default_rng(1)
    shape = (8192, 4)
    fcen, sr = 1e9, 1e6
    x = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    fmin, fmax = orc.band_edges(fcen, sr, 4)
    for ref in [fmin, fcen, fmax]:
        y, start, stop = orc.coherent_dedispersion(x, dm, sample_rate=sr, center_freq=fcen,
                                                   ref_freq=ref)
        assert len(x) - len(y) >= orc.sample_delay(dm, fmin, fmax, sr)
        assert start >= orc.sample_delay(dm, ref, fmax, sr)
        assert len(y) == stop - start


# ---- tests/test_dedispersion.py:73-98 ---------------------------------------------------
@pytest.mark.parametrize("seed", [4, 8, 15, 16, 23, 42])
def test_reversibility(seed):
    ref, sr, dm = 600e6, 400e6, 0.01
    N, M = 2 ** 18, 2 ** 12
    R = np.random.default_rng(seed=seed)
    x = R.standard_normal(N) + 1j * R.standard_normal(N)
    x *= np.exp(-(((np.arange(N) - N // 2) / M) ** 2))
    sos = scipy.signal.butter(10, 0.45, "lowpass", fs=1.0, output="sos")
    x = scipy.signal.sosfilt(sos, x).reshape(-1, 1)

    t, s1, _ = orc.coherent_dedispersion(x, dm, sample_rate=sr, center_freq=ref)
    y, s2, _ = orc.coherent_dedispersion(t, -dm, sample_rate=sr, center_freq=ref)
    noffset = s1 + s2
    assert np.allclose(x[noffset:noffset + len(y)] - y, 0, atol=3e-8)


# ---- tests/test_dedispersion.py:100-139 (known answer) ----------------------------------
@pytest.mark.parametrize("dm", [0.01, 0.02])
def test_gabor_collapse(dm):
    ref, sr = 600e6, 400e6
    index, N, width = 100000, 2 ** 18, 256
    t = np.arange(N) / sr
    t0 = t[index]
    x = np.zeros(N, dtype=np.complex128)
    for df in np.linspace(-3 * sr / 8, 3 * sr / 8, 13):
        dt = orc.time_delay(dm, ref + df, ref)
        a = 2j * np.pi * (t - (t0 + dt)) * df - ((t - (t0 + dt)) / (width / sr)) ** 2
        x += np.exp(a)

    y, noffset, _ = orc.coherent_dedispersion(x.reshape(-1, 1), dm, sample_rate=sr,
                                              center_freq=ref)
    id1, id2 = index - noffset - 8 * width, index - noffset + 8 * width
    p1 = (np.abs(x) ** 2).sum()
    p2 = (np.abs(y[id1:id2]) ** 2).sum()
    assert np.allclose(p1, p2)
    assert np.allclose(y[id2:], 0)
    assert np.allclose(y[:id1], 0)


# ---- tests/test_dedispersion.py:141-164 -------------------------------------------------
@pytest.mark.parametrize("dm", [10, 20, 50])
def test_explicit_chirp(dm):
    rng = np.random.default_rng(3)
    shape = (8192, 4, 2)
    x = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    sr, fcen = 1e6, 1e9
    freqs = orc.channel_freqs(fcen, sr, 4)
    fmin, fmax = orc.band_edges(fcen, sr, 4)
    for rf in [fcen, fmin, fmax]:
        chirp = orc.chirp_from_signal(dm, 8192, sr, freqs, rf)
        assert chirp.shape == (8192, 4) and chirp.dtype == np.complex64
        y1, *_ = orc.coherent_dedispersion(x, dm, sample_rate=sr, center_freq=fcen, ref_freq=rf)
        y2, *_ = orc.coherent_dedispersion(x, dm, sample_rate=sr, center_freq=fcen, ref_freq=rf,
                                           chirp=chirp)
        assert np.allclose(y1, y2)


# ---- tests/test_contrib.py:22-51 --------------------------------------------------------
@pytest.mark.parametrize("shape", [(4224, 4, 2), (4233, 3, 2)])
def test_stft_reversibility(shape):
    rng = np.random.default_rng(5)
    x = np.exp(1j * rng.uniform(-np.pi, np.pi, shape))
    for n in [33, 32, shape[0]]:
        keep = x.copy()
        y = orc.istft(orc.stft(x, n), n)
        assert np.array_equal(x, keep)
        assert np.allclose(x[: len(y)], y)
        assert len(y) == shape[0] - shape[0] % n


def test_stft_single_tone():
    x = np.exp(2j * np.pi * np.arange(1024) * 0.25)[:, None]
    for n in [32, 64, 512, 1024]:
        y = orc.stft(x, n)
        a = np.zeros_like(y)
        a[:, 3 * n // 4] = 1.0
        assert np.allclose(a, y)


# ---- tests/test_polarization.py:34-68 (known answer) ------------------------------------
def test_stokes_known_vectors():
    x = np.array([[[1 + 1j, 2 + 1j]], [[3 + 0j, 0 + 4j]], [[0 + 2j, 3 + 1j]]],
                 dtype=np.complex128)
    lin = np.array([[[7, -3, 6, -2]], [[25, -7, 0, 24]], [[14, -6, 4, -12]]])
    cir = np.array([[[7, 6, -2, -3]], [[25, 0, 24, -7]], [[14, 4, -12, -6]]])
    for pol_type, stokes in zip(["linear", "circular"], [lin, cir]):
        y_lin = orc.to_stokes(orc.to_linear(x, pol_type), "linear")
        y_cir = orc.to_stokes(orc.to_circular(x, pol_type), "circular")
        assert np.allclose(y_lin, stokes)
        assert np.allclose(y_cir, stokes)
        assert np.allclose(orc.stokes_I(x), stokes[..., 0])


# ---- tests/test_radio_signal.py:142-172 -------------------------------------------------
@pytest.mark.parametrize("A", [1, 4, 10])
@pytest.mark.parametrize("in_dtype, out_dtype",
                         [(np.complex64, np.float32), (np.complex128, np.float64)])
def test_intensity(A, in_dtype, out_dtype):
    rng = np.random.default_rng(7)
    z = (A * np.exp(1j * rng.uniform(-np.pi, np.pi, (1024, 8, 2)))).astype(in_dtype)
    zi = orc.to_intensity(z)
    assert zi.dtype == out_dtype
    assert np.allclose(A ** 2, zi)


# ---- tests/test_phase_predictor.py:39-75 (known answer) ---------------------------------
def test_phase_predictor_constants():
    with open(os.path.join(GOLDEN, "timing.dat")) as f:
        entries = orc.parse_polyco(f.read())
    assert len(entries) == 16
    t = (58245, 0.375)
    pi, pf = orc.predict_phase(entries, t)
    assert int(pi) == 146774936445
    assert np.isclose(float(pf), 0.058161699852649296)
    assert np.isclose(orc.spin_freq(entries, t), 641.973647812571, rtol=1e-8)
    assert np.isclose(orc.spin_freq(entries, t, n=1), -6.635997412662843e-08, rtol=1e-8)

    pi, pf = orc.predict_phase(entries, t, np.arange(10000) * 1e-6)
    assert int(pi[-1]) == 146774936451
    assert np.isclose(pf[-1], 0.4772562027766636)

    with pytest.raises(ValueError):
        orc.predict_phase(entries, (60000, 0.0))


ENTRY_TEXT = """B1937+21    7-May-18   0.00   58245.00000000000   71.020168
 146754136477.666475  641.928232294317   ao   90   12   327.000
 -3.17034199847385061e-07  2.76360291261698521e+00  8.05424212611731503e-05
 -1.14853014406135967e-07 -1.39769248548540950e-10  6.39552923641417649e-13
  3.19619782475082226e-15 -5.35166928586675360e-16 -4.58065943719761444e-19
  2.26855569952374124e-19 -4.63024751309515689e-23 -3.42478559749583972e-23
"""


# ---- tests/test_phase_predictor.py:77-95 ------------------------------------------------
def test_phasepol():
    entries = orc.parse_polyco(ENTRY_TEXT)
    t = entries[0]["tmid"]
    pi, pf = orc.predict_phase(entries, t)
    # Phase(146754136477, 0.666475) + (-3.17e-07): int rounds to nearest, frac in [-.5, .5]
    assert int(pi) == 146754136478
    assert np.isclose(float(pf), 0.666475 - 3.17034199847385061e-07 - 1, atol=1e-8)

    coef, ref = orc.phasepol(entries, t)
    for off in [1.0, 8.0, 0.001]:
        pi, pf = orc.predict_phase(entries, t, off)
        want = float(pi - ref) + float(pf)
        got = float(orc.polyval_numpy(off, coef))
        assert abs(want - got) < 1e-8


# ---- builder-defined ops: internal consistency (no reference test exists) ----------------
def test_downsample_and_unpack():
    rng = np.random.default_rng(11)
    x = rng.standard_normal((1003, 3, 2)).astype(np.float32)
    d = orc.downsample(x, 64)
    assert d.shape == (15, 3, 2)
    assert np.allclose(d[2], x[128:192].astype(np.float64).sum(0))
    raw = rng.integers(-128, 128, (50, 2, 2, 2), dtype=np.int8)
    z = orc.unpack_int8(raw)
    assert z.dtype == np.complex64 and z.shape == (50, 2, 2)
    assert z[3, 1, 0] == complex(raw[3, 1, 0, 0], raw[3, 1, 0, 1])


def test_fold_counts_and_sum():
    rng = np.random.default_rng(13)
    x = rng.random((5000, 4)).astype(np.float32)
    coef = np.array([0.123, 29.7, 1e-6])
    prof, counts = orc.fold(x, coef, 1e4, 64)
    assert counts.sum() == 5000 and prof.shape == (64, 4)
    assert np.allclose(prof.sum(0), x.astype(np.float64).sum(0))
    bins = orc.fold_bins(5000, coef, 1e4, 64)
    assert bins.min() >= 0 and bins.max() < 64
    assert bins[0] == int(0.123 * 64)


def test_overlap_save_matches_blockwise_definition():
    rng = np.random.default_rng(17)
    N, L = 2 ** 14, 2 ** 12
    x = (rng.standard_normal((N, 2)) + 1j * rng.standard_normal((N, 2))).astype(np.complex64)
    kw = dict(sample_rate=1e6, center_freq=1e9)
    y, first = orc.overlap_save_dedispersion(x, 30.0, L, **kw)
    start, stop = orc.crop_range(30.0, L, 1e9, 1e6, 2, 1e9)
    valid = stop - start
    assert first == start and len(y) % valid == 0
    # interior samples agree with one long transform up to the chirp's out-of-sweep tails
    full, s_full, e_full = orc.coherent_dedispersion(x, 30.0, **kw)
    lo = max(first, s_full)
    hi = min(first + len(y), e_full)
    a = y[lo - first:hi - first]
    b = full[lo - s_full:hi - s_full]
    assert np.linalg.norm(a - b) / np.linalg.norm(b) < 0.05


@pytest.mark.parametrize("dm", [50.0, 100.0, 200.0])
def test_incoherent_dedispersion_crop(dm):
    """reference tests/test_dedispersion.py:167-189: the crop equals the rounded band-edge delays."""
    sr, ref, bw = 1e3, 1e9, 8e6
    x = np.random.default_rng(1).standard_normal((8192, 32, 4))
    y, crop_before, delays = orc.incoherent_dedispersion(x, dm, sample_rate=sr, center_freq=ref,
                                                         chan_bw=bw)
    freqs = orc.channel_freqs(ref, bw, 32)
    d_top = np.round(orc.sample_delay(dm, ref, freqs[-1], sr))
    d_bot = np.round(orc.sample_delay(dm, freqs[0], ref, sr))
    assert x.shape[0] - y.shape[0] == int(d_top + d_bot)
    # an impulse at the dispersed arrival time of every channel lines up after dedispersion
    imp = np.zeros((8192, 32))
    n0 = 4000
    raw = np.round(orc.sample_delay(dm, freqs, ref, sr)).astype(int)
    imp[n0 + raw, np.arange(32)] = 1.0
    yi, cb, _ = orc.incoherent_dedispersion(imp, dm, sample_rate=sr, center_freq=ref, chan_bw=bw)
    rows = np.argmax(yi, axis=0)
    assert np.all(rows == rows[0]) and rows[0] == n0 - cb


# ---------------------------------------------------------------- time_shift / freq_shift KATs
def _impulse(N, t0):
    """reference tests/test_transforms.py:27-31."""
    n = (np.arange(N) - N // 2) / N
    x = np.exp(-2j * np.pi * np.asarray(t0) * n)
    return np.fft.ifft(np.fft.ifftshift(x, axes=(-1,))).astype(np.complex128)


def _sinusoid(N, f0):
    """reference tests/test_transforms.py:34-37."""
    n = np.arange(N) / N
    return np.exp(2j * np.pi * np.asarray(f0) * n).astype(np.complex128)


@pytest.mark.parametrize("shape", [(4096, 4, 2), (4096, 4), (4096,)])
def test_time_shift_impulses(shape):
    """reference tests/test_transforms.py:346-378: fractionally delayed impulses line up."""
    rng = np.random.default_rng(5)
    N = shape[0]
    for lo, hi in [(-20, 20), (0, 20), (-20, 0)]:
        shift = rng.uniform(lo, hi, shape[1:])
        x = np.moveaxis(_impulse(N, 100 - shift[..., None]), -1, 0)
        y, a, b = orc.time_shift(x, shift)
        z = np.zeros_like(y)
        z[100] = 1.0
        assert np.allclose(y, z)
        assert a == max(0, int(np.ceil(shift.max())))
        assert N + b == N + min(0, int(np.floor(shift.min())))


def test_time_shift_integer_rolls():
    """reference tests/test_transforms.py:312-344."""
    rng = np.random.default_rng(6)
    z = rng.standard_normal((4096, 4, 2)) + 1j * rng.standard_normal((4096, 4, 2))
    for n in [-12, -3, 4, 13]:
        y, _, _ = orc.time_shift(z, n)
        if n < 0:
            assert np.allclose(y[:n], z[-n:], atol=1e-6) and np.allclose(y[n:], 0)
        else:
            assert np.allclose(y[n:], z[:-n], atol=1e-6) and np.allclose(y[:n], 0)


def test_freq_shift_tones_and_zeroing():
    """reference tests/test_transforms.py:420-475."""
    N = 1024
    fs = np.array([[-52, -45.4], [-25.5, 34], [14, -36.9], [45.1, 27]])
    x = _sinusoid(N, fs[None].T).T
    assert np.allclose(orc.freq_shift(x, -fs / N), 1)
    x = np.zeros((N, 4, 2), np.complex128)
    x[0] = 1
    shift = np.array([[10, -10], [20.5, -20.5], [-600, 600], [2000, -2000]])
    y = np.fft.fftshift(np.abs(scipy.fft.fft(orc.freq_shift(x, shift / N), axis=0)), axes=(0,))
    for i in range(4):
        for j in range(2):
            s = shift[i, j]
            if s < 0:
                s = int(np.floor(s))
                assert np.allclose(y[s:, i, j], 0) and np.allclose(y[:s, i, j], 1)
            else:
                s = int(np.ceil(s))
                assert np.allclose(y[:s, i, j], 0) and np.allclose(y[s:, i, j], 1)


def test_real_to_complex_theoretical():
    """reference tests/test_utils.py:20-68: a real tone at w becomes a complex tone at w - N/4."""
    for N in [511, 512]:
        t = np.linspace(0, 2 * np.pi, N, endpoint=False)
        for w in [1, 2, 127, 128, 129, 254, 255]:
            for p in [-np.pi, -np.pi / 2, 0, np.pi / 2]:
                x = np.cos(w * t + p)
                y = np.exp(1j * ((w - (len(t) / 4)) * t[::2] + p))
                assert np.allclose(orc.real_to_complex(x), y)
    x = np.zeros((0, 2))
    assert np.array_equal(orc.real_to_complex(x), x)
    with pytest.raises(ValueError):
        orc.real_to_complex(np.ones((8, 2), complex))
    assert orc.real_to_complex(np.ones(32, np.float32)).dtype == np.complex64


def test_packed_sample_decoders():
    """Builder-defined raw formats (include/pbk.h PBK_U4X2 / PBK_U2X2): offset-binary 4-bit and the
    four-level 2-bit code, first sample in the least significant bits."""
    b = np.array([[0x00, 0xF8, 0x7F, 0x19]], dtype=np.uint8)
    assert np.array_equal(orc.unpack_u4(b), np.array([[-8 - 8j, 0 + 7j, 7 - 1j, 1 - 7j]],
                                                     dtype=np.complex64))
    hi = np.float32(3.3359)
    # byte 0b11_10_01_00: sample 0 = (code 0, code 1), sample 1 = (code 2, code 3)
    got = orc.unpack_u2(np.array([[0b11100100]], dtype=np.uint8))
    assert got.shape == (1, 2)
    assert np.array_equal(got, np.array([[-hi - 1j, 1 + 1j * hi]], dtype=np.complex64))
