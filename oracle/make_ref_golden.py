"""TEST INFRASTRUCTURE ONLY -- freeze outputs of the REFERENCE ITSELF into tests/golden.

Runs the reference's own source files from /root/reference (through oracle/ref_run.py: the
unmodified hot-path modules, imported where they lie, against the unit/dask stand-ins of
oracle/ref_shim) on small seeded inputs and writes inputs, parameters and reference outputs to
tests/golden/ref_golden.npz.  /root/reference does not exist on the GPU box, so the GPU parity
tests (tests/test_ref_golden.py) compare the CUDA path with these frozen vectors; the CPU tests
compare the oracle restatement with them and, when /root/reference is present, with the live
reference as well.

    python -m oracle.make_ref_golden        # from the repo root, in the build container
"""

import json
import os

import numpy as np

from oracle import ref_run

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "tests", "golden", "ref_golden.npz")

T0 = (58245, 0.375)        # start time of every fixture signal: MJD day, day fraction


def cnoise(rng, shape, dtype=np.complex64):
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(dtype)


# name -> (signal class, shape, dtype, sample_rate MHz, center_freq MHz, freq_align, DM,
#          ref_freq: None | "top" | "bottom" | MHz number)
DEDISP_CASES = {
    "dd_dualpol":   ("DualPolarizationSignal", (2048, 4, 2), "complex64", 1.0, 600.0, "center", 5.0, None),
    "dd_ref_top":   ("BasebandSignal", (2048, 3), "complex64", 2.0, 800.0, "center", 3.0, "top"),
    "dd_ref_bot":   ("BasebandSignal", (2048, 3), "complex64", 2.0, 800.0, "center", 3.0, "bottom"),
    "dd_c128_bot":  ("BasebandSignal", (2048, 4), "complex128", 1.0, 400.0, "bottom", 1.0, 401.5),
    "dd_align_top": ("BasebandSignal", (2048, 2), "complex64", 1.0, 1400.0, "top", 40.0, None),
    "dd_cfg1_like": ("BasebandSignal", (8192, 1), "complex64", 16.0, 400.0, "center", 0.02, None),
    "dd_neg_dm":    ("BasebandSignal", (2048, 2), "complex64", 1.0, 600.0, "center", -4.0, None),
}


def make_signal(pb, u, Time, kind, x, sr, fc, align):
    kw = dict(sample_rate=sr * u.MHz, center_freq=fc * u.MHz, freq_align=align,
              start_time=Time(*T0))
    if kind == "DualPolarizationSignal":
        kw["pol_type"] = "linear"
    return getattr(pb, kind)(x, **kw)


def ref_freq_of(z, u, ref):
    if ref is None:
        return None
    if ref == "top":
        return z.max_freq
    if ref == "bottom":
        return z.min_freq
    return ref * u.MHz


def main():
    pb, u, Time = ref_run.load()
    out, meta = {}, {"T0": T0, "dedisp": {}}
    t0 = Time(*T0)

    # ---- coherent dedispersion (dedispersion.py:81-133) ------------------------------------
    for i, (name, (kind, shape, dt, sr, fc, align, dm, ref)) in enumerate(DEDISP_CASES.items()):
        x = cnoise(np.random.default_rng(100 + i), shape, np.dtype(dt))
        z = make_signal(pb, u, Time, kind, x, sr, fc, align)
        rf = ref_freq_of(z, u, ref)
        y = pb.coherent_dedispersion(z, pb.DM(dm), ref_freq=rf)
        assert type(y) is type(z) and y.dtype == z.dtype
        out[name + "_x"], out[name + "_y"] = x, np.asarray(y.data)
        meta["dedisp"][name] = dict(
            kind=kind, sample_rate_hz=sr * 1e6, center_freq_hz=fc * 1e6, freq_align=align, dm=dm,
            ref_freq_hz=None if rf is None else float(rf.to_value(u.Hz)),
            start_shift_s=float((y.start_time - t0).to_value(u.s)), nout=len(y),
            out_sample_rate_hz=float(y.sample_rate.to_value(u.Hz)),
            out_center_freq_hz=float(y.center_freq.to_value(u.Hz)))

    # explicit chirp == implicit (dedispersion.py:121-124); a squeezed (N, C) chirp broadcasts
    z = make_signal(pb, u, Time, "DualPolarizationSignal", out["dd_dualpol_x"], 1.0, 600.0,
                    "center")
    chirp = pb.DM(5.0).chirp_from_signal(z)
    out["dd_dualpol_chirp"] = np.asarray(chirp)
    y2 = pb.coherent_dedispersion(z, pb.DM(5.0), chirp=np.asarray(chirp)[:, :, 0])
    assert np.array_equal(np.asarray(y2.data), out["dd_dualpol_y"])

    # ---- the chirp at BASELINE sizes, sampled (dedispersion.py:19-23) -----------------------
    rng = np.random.default_rng(7)
    zc = pb.BasebandSignal(np.zeros((8, 64), np.complex64), sample_rate=6.25 * u.MHz,
                           center_freq=600 * u.MHz, freq_align="center")
    freqs = zc.channel_freqs
    out["chan_freqs_cfg2_hz"] = freqs.to_value(u.Hz)
    idx = np.sort(rng.choice(2 ** 22, 1024, replace=False))
    idx[:4], idx[-4:] = [0, 1, 2 ** 21 - 1, 2 ** 21], [2 ** 21 + 1, 2 ** 22 - 3, 2 ** 22 - 2, 2 ** 22 - 1]
    out["chirp_cfg2_idx"], chans = idx, [0, 31, 63]
    out["chirp_cfg2_chan"] = np.array(chans)
    dt2 = (1 / (6.25 * u.MHz)).to(u.s)
    out["chirp_cfg2_val"] = np.stack(
        [pb.DM(100.0).chirp_function(2 ** 22, dt2, freqs[c], 600 * u.MHz)[idx] for c in chans], 1)
    dt1 = (1 / (16 * u.MHz)).to(u.s)
    idx1 = np.sort(rng.choice(2 ** 20, 1024, replace=False))
    out["chirp_cfg1_idx"] = idx1
    out["chirp_cfg1_val"] = pb.DM(71.0).chirp_function(2 ** 20, dt1, 400 * u.MHz, 400 * u.MHz)[idx1]
    # delays (dedispersion.py:32-42) at cfg3 / cfg5 band edges -> the integer crops of SURVEY 8d
    crops = {}
    for tag, (n, c, sr, dm) in {"cfg3": (2 ** 22, 1024, 0.390625, 100.0),
                                "cfg5": (2 ** 26, 256, 1.5625, 1000.0),
                                "cfg2": (2 ** 22, 64, 6.25, 100.0),
                                "cfg1": (2 ** 20, 1, 16.0, 71.0)}.items():
        fc = 400.0 if tag == "cfg1" else 600.0
        ze = pb.BasebandSignal(np.zeros((2, c), np.complex64), sample_rate=sr * u.MHz,
                               center_freq=fc * u.MHz)
        d_top = pb.DM(dm).sample_delay(ze.max_freq, ze.center_freq, ze.sample_rate)
        d_bot = pb.DM(dm).sample_delay(ze.min_freq, ze.center_freq, ze.sample_rate)
        import math
        crops[tag] = dict(delay_top=float(d_top), delay_bot=float(d_bot),
                          start=math.ceil(-min(0, d_top, d_bot)),
                          stop=n - math.ceil(+max(0, d_top, d_bot)))
    meta["crops"] = crops

    # ---- channelize / unchannelize (contrib/misc.py:17-93) ----------------------------------
    xs = cnoise(np.random.default_rng(200), (1056, 3, 2))
    out["stft_x"] = xs
    meta["stft"] = {}
    for n in (32, 33, 1056):
        zs = make_signal(pb, u, Time, "DualPolarizationSignal", xs.copy(), 1.0, 600.0, "center")
        ys = pb.contrib.stft(zs, nperseg=n)
        out[f"stft_y{n}"] = np.asarray(ys.data).copy()
        meta["stft"][str(n)] = dict(sample_rate_hz=float(ys.sample_rate.to_value(u.Hz)),
                                    freq_align=ys.freq_align, nchan=ys.nchan,
                                    center_freq_hz=float(ys.center_freq.to_value(u.Hz)))
        zi = pb.contrib.istft(ys, nperseg=n)      # NB misc.py:82-83 scales ys.data in place
        out[f"istft_y{n}"] = np.asarray(zi.data)
        meta["stft"][str(n)].update(inv_sample_rate_hz=float(zi.sample_rate.to_value(u.Hz)),
                                    inv_freq_align=zi.freq_align, inv_nchan=zi.nchan)

    # freq_align 'bottom' / 'top' with an EVEN number of channels: misc.py:41 slices both axes, and
    # the frequency slice re-centres center_freq (core.py:479-484) before `like` sets freq_align,
    # so the output's center_freq moves by half a coarse channel -- frozen here as metadata
    xa = cnoise(np.random.default_rng(201), (256, 4))
    out["stft_align_x"] = xa
    meta["stft_align"] = {}
    for align in ("bottom", "top", "center"):
        for n in (32, 33):
            za = make_signal(pb, u, Time, "BasebandSignal", xa.copy(), 1.0, 400.0, align)
            ya = pb.contrib.stft(za, nperseg=n)
            tag = f"{align}_{n}"
            out[f"stft_align_{tag}_freqs_hz"] = ya.channel_freqs.to_value(u.Hz)
            if align == "bottom" and n == 32:
                out["stft_align_y"] = np.asarray(ya.data).copy()
            meta["stft_align"][tag] = dict(
                in_align=align, nperseg=n, freq_align=ya.freq_align, nchan=ya.nchan,
                center_freq_hz=float(ya.center_freq.to_value(u.Hz)),
                sample_rate_hz=float(ya.sample_rate.to_value(u.Hz)))

    # ---- detection (core.py:766-774, 882-966) ----------------------------------------------
    xp = cnoise(np.random.default_rng(300), (256, 3, 2))
    out["pol_x"] = xp
    for pt in ("linear", "circular"):
        zp = pb.DualPolarizationSignal(xp, sample_rate=1 * u.MHz, center_freq=600 * u.MHz,
                                       pol_type=pt)
        out[f"pol_{pt}_intensity"] = np.asarray(zp.to_intensity().data)
        out[f"pol_{pt}_stokes"] = np.asarray(zp.to_stokes().data)
        out[f"pol_{pt}_to_linear"] = np.asarray(zp.to_linear().data)
        out[f"pol_{pt}_to_circular"] = np.asarray(zp.to_circular().data)

    # ---- time_shift / freq_shift (transforms.py:211-361) ------------------------------------
    xt = cnoise(np.random.default_rng(400), (1024, 3, 2))
    out["shift_x"] = xt
    zt = make_signal(pb, u, Time, "DualPolarizationSignal", xt, 1.0, 600.0, "center")
    meta["time_shift"], meta["freq_shift"] = {}, {}
    for tag, sh in {"scalar": 2.5, "neg": -7.25, "perchan": [-3.25, 0.5, 7.0]}.items():
        yt = pb.time_shift(zt, sh)
        out[f"tshift_{tag}"] = np.asarray(yt.data)
        yc = pb.time_shift(zt, sh, crop=True)
        meta["time_shift"][tag] = dict(shift=sh, ncrop=len(yc),
                                       crop_shift_s=float((yc.start_time - t0).to_value(u.s)))
    yq = pb.time_shift(zt, 2.5 * u.us)
    assert np.array_equal(np.asarray(yq.data), out["tshift_scalar"])
    for tag, sh in {"scalar": 12.5, "neg": -100.0, "perchan": [30.0, -0.25, 250.0]}.items():
        yf = pb.freq_shift(zt, np.array(sh) * u.kHz)
        out[f"fshift_{tag}"] = np.asarray(yf.data)
        meta["freq_shift"][tag] = dict(shift_hz=(np.array(sh) * 1e3).tolist())

    # ---- incoherent dedispersion (dedispersion.py:136-177) ----------------------------------
    xi = np.random.default_rng(500).standard_normal((4096, 8)).astype(np.float32) ** 2
    out["incoh_x"] = xi
    meta["incoh"] = {}
    for tag, (dm, ref) in {"centre": (30.0, None), "top": (30.0, "top"), "bot": (12.0, "bottom")}.items():
        zi = pb.IntensitySignal(xi, sample_rate=1 * u.kHz, center_freq=600 * u.MHz,
                                chan_bw=10 * u.MHz, freq_align="center", start_time=Time(*T0))
        rf = ref_freq_of(zi, u, ref)
        yi = pb.incoherent_dedispersion(zi, pb.DM(dm), ref_freq=rf)
        out[f"incoh_{tag}"] = np.asarray(yi.data)
        meta["incoh"][tag] = dict(dm=dm, ref_freq_hz=None if rf is None else float(rf.to_value(u.Hz)),
                                  start_shift_s=float((yi.start_time - t0).to_value(u.s)))

    # ---- real_to_complex (utils.py:15-65) ---------------------------------------------------
    xr = np.random.default_rng(600).standard_normal((1024, 3)).astype(np.float32)
    out["r2c_x32"], out["r2c_y32"] = xr, pb.utils.real_to_complex(xr, axis=0)
    xr64 = np.random.default_rng(601).standard_normal(1023)
    out["r2c_x64"], out["r2c_y64"] = xr64, pb.utils.real_to_complex(xr64)

    # ---- phase predictor (pulsar/predictor.py:108-160, 276-306) on the reference's own polyco
    # fixture: raw (integer, fractional) parts as PhasePredictor.__call__ hands them to pb.Phase
    PhasePredictor, _, _, _ = ref_run.load_predictor()
    polyco = os.path.join(ref_run.REF_ROOT, "tests", "data", "timing.dat")
    if not os.path.isfile(polyco):      # a pip install of the reference carries no tests: use the
        polyco = os.path.join(os.path.dirname(OUT), "timing.dat")       # byte-identical copy
    pred = PhasePredictor.from_polyco(polyco)
    t1 = Time("58245.375", format="mjd", precision=9)
    ph = pred(t1)
    out["pred_scalar"] = np.array([float(ph.phase1), float(ph.phase2),
                                   pred.f0(t1).to_value(u.cycle / u.s),
                                   pred.f0(t1, n=1).to_value(u.cycle / u.s ** 2)])
    out["pred_rphase"] = np.asarray(pred["rphase"], dtype=np.int64)
    ph = pred(t1 + np.arange(10000) * u.us)
    out["pred_us_ph1"], out["pred_us_ph2"] = ph.phase1.astype(np.int64), ph.phase2
    (a, b), = pred.intervals
    span_s = float((b - a).to_value(u.s))
    offs = (np.arange(400) + 0.37) / 400 * span_s                  # across all 16 entries
    ts = a + offs * u.s
    ph = pred(ts)
    out["pred_span_mjd_int"], out["pred_span_mjd_frac"] = np.asarray(ts.jd1), np.asarray(ts.jd2)
    out["pred_span_ph1"], out["pred_span_ph2"] = ph.phase1.astype(np.int64), ph.phase2
    out["pred_span_f0"] = pred.f0(ts).to_value(u.cycle / u.s)
    pol_t = [Time("58245.375", format="mjd"), Time("58245.0", format="mjd"),
             Time(58245.0, 0.6180339887, format="mjd"), ts[123], ts[377]]
    out["pred_phasepol_mjd"] = np.array([[float(t.jd1), float(t.jd2)] for t in pol_t])
    pols = [pred.phasepol(t) for t in pol_t]
    out["pred_phasepol_coef"] = np.stack([q.coef for q, _ in pols])
    out["pred_phasepol_ref"] = np.array([int(r.phase1) for _, r in pols], dtype=np.int64)

    out["meta_json"] = np.array(json.dumps(meta))
    np.savez(OUT, **out)
    print(f"wrote {OUT}: {len(out)} arrays, {os.path.getsize(OUT) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
