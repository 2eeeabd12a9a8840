"""CPU-only tests of the host-side drop-in logic added in round 2: the pinned plan cache, the
per-generator stream states, ``pb.fft`` pass-through, the lazy (dask) per-chunk route and the
argument checks in front of the raw-pointer calls.  Kernels are replaced by the oracle where a
test needs numbers (tests may use the oracle; the product never does)."""

import threading
import time

import numpy as np
import pytest

import pulsarbat_b200 as pb
from pulsarbat_b200 import kernels, streaming
from pulsarbat_b200 import units as u
from oracle import pbk_oracle as orc

import fake_dask


# ------------------------------------------------------------------------------------------
# plan cache (ADVICE: eviction race, creation under the global lock, cap by bytes)
# ------------------------------------------------------------------------------------------
class FakePlan:
    def __init__(self, nbytes=0, delay=0.0):
        time.sleep(delay)
        self.nbytes, self.destroyed = nbytes, False

    def info(self):
        return {"workspace_bytes": self.nbytes}

    def destroy(self):
        self.destroyed = True


@pytest.fixture
def clean_cache(monkeypatch):
    kernels.clear_plan_cache()
    monkeypatch.setattr(kernels, "_MAX_PLANS", 3)
    monkeypatch.setenv("PBK_PLAN_CACHE_BYTES", str(100))
    yield
    kernels.clear_plan_cache()


def test_plan_cache_never_evicts_an_entry_in_use(clean_cache):
    made = {}

    def factory(k, nbytes=10):
        def f():
            made[k] = FakePlan(nbytes)
            return made[k]
        return f
    with kernels._use_plan("a", factory("a")) as pa:
        for k in "bcdef":                       # five more plans while "a" is pinned
            with kernels._use_plan(k, factory(k)):
                pass
        assert not pa.destroyed                 # still usable by its holder
        assert "a" in kernels._cache
    assert len(kernels._cache) <= 3
    assert sum(p.destroyed for p in made.values()) >= 3
    with kernels._use_plan("a", factory("a2")) as again:   # and still cached afterwards
        assert again is pa


def test_plan_cache_is_capped_by_workspace_bytes(clean_cache):
    plans = []

    def factory(nbytes):
        def f():
            plans.append(FakePlan(nbytes))
            return plans[-1]
        return f
    with kernels._use_plan("big1", factory(60)):
        pass
    with kernels._use_plan("big2", factory(60)):     # 120 > cap 100: the idle one goes
        pass
    assert plans[0].destroyed and not plans[1].destroyed
    with kernels._use_plan("huge", factory(500)) as p:   # a plan above the cap is still admitted
        assert not p.destroyed
    assert list(kernels._cache) == ["huge"]


def test_plan_cache_builds_outside_the_lock_and_only_once(clean_cache):
    built, got = [], []

    def factory():
        built.append(1)
        return FakePlan(1, delay=0.2)

    def other():
        built.append(2)
        return FakePlan(1)

    def worker():
        with kernels._use_plan("slow", factory) as p:
            got.append(p)
    ts = [threading.Thread(target=worker) for _ in range(4)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    time.sleep(0.05)
    with kernels._use_plan("fast", other):        # not blocked by the slow build
        fast_dt = time.perf_counter() - t0
    for t in ts:
        t.join()
    assert built.count(1) == 1 and len(set(map(id, got))) == 1
    assert fast_dt < 0.18


def test_plan_cache_failed_build_is_not_cached(clean_cache):
    def bad():
        raise pb.PbkError(-3, "boom")
    with pytest.raises(pb.PbkError):
        with kernels._use_plan("x", bad):
            pass
    assert "x" not in kernels._cache
    with kernels._use_plan("x", lambda: FakePlan()) as p:
        assert isinstance(p, FakePlan)


def test_clear_plan_cache_defers_plans_in_use(clean_cache):
    with kernels._use_plan("a", lambda: FakePlan()) as p:
        kernels.clear_plan_cache()
        assert not p.destroyed
    assert p.destroyed


# ------------------------------------------------------------------------------------------
# stream states (ADVICE: one process-global slot shared by live generators)
# ------------------------------------------------------------------------------------------
class FakeState:
    def __init__(self):
        self.destroyed = False

    def quiesce(self):
        pass

    def destroy(self):
        self.destroyed = True


def test_stream_states_are_owned_and_pooled(monkeypatch):
    streaming.clear_stream_cache()
    monkeypatch.setattr(streaming, "_MAX_IDLE", 2)
    a = streaming._checkout("k", FakeState)
    b = streaming._checkout("k", FakeState)      # a second live generator gets its OWN state
    assert a is not b
    streaming._checkin("k", a)
    assert streaming._checkout("k", FakeState) is a       # reused once returned
    streaming._checkin("k", a)
    streaming._checkin("k", b)
    c = FakeState()
    streaming._checkin("other", c)               # pool limit: the oldest idle state is freed
    assert a.destroyed and not b.destroyed and not c.destroyed
    d = streaming._checkout("k", FakeState)
    assert d is b
    streaming.clear_stream_cache()               # only idle states are touched
    assert c.destroyed and not d.destroyed
    streaming._checkin("k", d)
    streaming.clear_stream_cache()
    assert d.destroyed


# ------------------------------------------------------------------------------------------
# pb.fft: the 12 off-path names pass through to scipy (reference fft.py:8-43)
# ------------------------------------------------------------------------------------------
def test_fft_module_lists_and_forwards_every_reference_name():
    import scipy.fft
    names = ["fft", "fft2", "fftn", "ifft", "ifft2", "ifftn", "rfft", "rfft2", "rfftn", "irfft",
             "irfft2", "irfftn", "hfft", "ihfft"]
    assert dir(pb.fft) == sorted(names)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((8, 6))
    z = x + 1j * rng.standard_normal((8, 6))
    assert np.array_equal(pb.fft.rfft(x, axis=0), scipy.fft.rfft(x, axis=0))
    assert np.array_equal(pb.fft.fft2(z), scipy.fft.fft2(z))
    assert np.array_equal(pb.fft.irfft(z, n=9, axis=1), scipy.fft.irfft(z, n=9, axis=1))
    assert np.array_equal(pb.fft.ihfft(x[0]), scipy.fft.ihfft(x[0]))
    assert pb.fft.rfft.__name__ == "rfft"
    with pytest.raises(AttributeError):
        pb.fft.fftshift                      # not in the reference's list either (fft.py:31-32)

    class Dev:                               # a DeviceArray is never copied to the host silently
        tensor = object()
    with pytest.raises(pb.PbkUnsupported):
        pb.fft.rfft(Dev())


# ------------------------------------------------------------------------------------------
# lazy (dask) inputs: per-chunk route with global ref_freq / crop
# ------------------------------------------------------------------------------------------
@pytest.fixture
def oracle_kernels(monkeypatch):
    """kernels.* replaced by oracle arithmetic, recording how each chunk was called."""
    calls = []

    def dedisperse(data, *, dm, sample_rate_hz, chan_freq_hz, ref_freq_hz, crop=None,
                   out_kind=0, downsample=1, chirp_array=None, int8=False, **kw):
        assert isinstance(data, np.ndarray), "chunk functions receive numpy blocks"
        calls.append(dict(shape=data.shape, freqs=np.array(chan_freq_hz), ref=ref_freq_hz,
                          crop=crop, thread=threading.get_ident()))
        x = data.astype(np.complex128)
        chirp = orc.chirp_from_signal(dm, x.shape[0], sample_rate_hz, np.asarray(chan_freq_hz),
                                      ref_freq_hz)
        if chirp_array is not None:
            chirp = np.asarray(chirp_array)
        chirp = chirp.reshape(chirp.shape[:2] + (1,) * (x.ndim - 2))
        import scipy.fft
        y = scipy.fft.ifft(scipy.fft.fft(x, axis=0) * chirp, axis=0)[crop[0]:crop[1]]
        if out_kind == 1:
            y = orc.to_intensity(y)
        elif out_kind == 2:
            y = orc.stokes_I(y)
        if downsample > 1:
            y = orc.downsample(y, downsample)
        return y.astype(np.complex64 if out_kind == 0 else np.float32)

    monkeypatch.setattr(kernels, "dedisperse", dedisperse)
    monkeypatch.setattr(kernels, "stft", lambda x, n, **kw: orc.stft(np.asarray(x), n).astype(np.complex64))
    monkeypatch.setattr(kernels, "istft", lambda x, n, **kw: orc.istft(np.asarray(x), n).astype(np.complex64))
    monkeypatch.setattr(kernels, "detect", lambda x, stokes=False, **kw: (
        orc.stokes_I(np.asarray(x)) if stokes else orc.to_intensity(np.asarray(x))).astype(np.float32))
    monkeypatch.setattr(kernels, "fft", lambda x, axis=0, inverse=False, **kw: (
        np.fft.ifft(x, axis=axis) if inverse else np.fft.fft(x, axis=axis)).astype(np.complex64))
    return calls


def _noise(shape, seed=3):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)


def test_dask_dedispersion_is_lazy_chunked_and_uses_global_ref_and_crop(monkeypatch, oracle_kernels):
    da = fake_dask.install(monkeypatch)
    x = _noise((2048, 8, 2))
    kw = dict(sample_rate=1 * u.MHz, center_freq=600 * u.MHz, pol_type="linear",
              freq_align="bottom", start_time=pb.Time(58245, 0.375))
    z_np = pb.DualPolarizationSignal(x, **kw)
    z_da = pb.DualPolarizationSignal(da.from_array(x, chunks=(512, 3, 1)), **kw)
    want = pb.coherent_dedispersion(z_np, pb.DM(5.0))
    n_eager = len(oracle_kernels)
    got = pb.coherent_dedispersion(z_da, pb.DM(5.0))
    assert isinstance(got.data, da.Array) and len(oracle_kernels) == n_eager     # nothing ran yet
    assert got.shape == want.shape and got.start_time == want.start_time
    assert got.center_freq == want.center_freq and got.freq_align == want.freq_align
    y = np.asarray(got.data)
    np.testing.assert_allclose(y, np.asarray(want.data), rtol=0, atol=1e-5)
    chunk_calls = oracle_kernels[n_eager:]
    assert [c["shape"] for c in chunk_calls] == [(2048, 3, 2), (2048, 3, 2), (2048, 2, 2)] or \
        sorted(c["shape"][1] for c in chunk_calls) == [2, 3, 3]
    full = z_np.channel_freqs_hz
    for c in chunk_calls:
        assert c["ref"] == 600e6 and c["crop"] == oracle_kernels[0]["crop"]      # global, not per chunk
        i = int(np.argmin(np.abs(full - c["freqs"][0])))
        assert np.array_equal(c["freqs"], full[i:i + len(c["freqs"])])           # ITS channels


def test_dask_detect_stft_and_fft_routes(monkeypatch, oracle_kernels):
    da = fake_dask.install(monkeypatch)
    x = _noise((1024, 4, 2), seed=5)
    kw = dict(sample_rate=1 * u.MHz, center_freq=600 * u.MHz, pol_type="linear")
    z_np = pb.DualPolarizationSignal(x, **kw)
    z_da = pb.DualPolarizationSignal(da.from_array(x, chunks=(300, 2, 1)), **kw)
    for fn in (lambda z: z.to_intensity(), lambda z: z.to_stokes_I(),
               lambda z: pb.dedisperse_detect(z, pb.DM(2.0), stokes_I=True, downsample=4),
               lambda z: pb.dedisperse_detect(z, pb.DM(2.0), downsample=2),
               lambda z: pb.contrib.stft(z, nperseg=32),
               lambda z: pb.contrib.istft(pb.contrib.stft(z, nperseg=32), nperseg=32)):
        a, b = fn(z_np), fn(z_da)
        assert isinstance(b.data, da.Array) and type(a) is type(b) and a.shape == b.shape
        np.testing.assert_allclose(np.asarray(b.data), np.asarray(a.data), rtol=1e-5, atol=1e-5)
        assert a.sample_rate == b.sample_rate
    xd = da.from_array(x, chunks=(1024, 2, 1))
    np.testing.assert_allclose(np.asarray(pb.fft.fft(xd, axis=0)), np.fft.fft(x, axis=0),
                               rtol=1e-4, atol=1e-3)
    with pytest.raises(ValueError, match="single chunk"):
        pb.fft.ifft(da.from_array(x, chunks=(512, 4, 2)), axis=0)
    assert isinstance(pb.fft.rfft(da.from_array(x.real, chunks=(1024, 2, 1)), axis=0), da.Array)


# ------------------------------------------------------------------------------------------
# argument checks in front of raw-pointer calls
# ------------------------------------------------------------------------------------------
def test_fold_rejects_bad_accumulators_and_coefficients():
    x = np.ones((64, 4), np.float32)
    with pytest.raises(ValueError, match="finite"):
        kernels.fold(x, [0.0, np.nan], 1e3, 16)
    with pytest.raises(ValueError, match="profile"):
        kernels.fold(x, [0.0, 1.0], 1e3, 16, profile=np.zeros((16, 4), np.float64))
    with pytest.raises(ValueError, match="profile"):
        kernels.fold(x, [0.0, 1.0], 1e3, 16, profile=np.zeros((8, 4), np.float32))
    with pytest.raises(ValueError, match="counts"):
        kernels.fold(x, [0.0, 1.0], 1e3, 16, counts=np.zeros((16,), np.int32))


def test_device_array_warns_on_large_implicit_copy(monkeypatch):
    from pulsarbat_b200 import device

    class T:
        def numel(self):
            return 1 << 20

        def element_size(self):
            return 8
    d = device.DeviceArray.__new__(device.DeviceArray)
    d.tensor = T()
    monkeypatch.setattr(device, "_D2H_WARN_BYTES", 1 << 20)
    monkeypatch.setattr(device.DeviceArray, "numpy", lambda self: np.zeros(4))
    with pytest.warns(ResourceWarning, match="implicit device->host copy"):
        np.asarray(d)
    monkeypatch.setattr(device, "_D2H_WARN_BYTES", 1 << 30)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        np.asarray(d)
