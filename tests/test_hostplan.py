"""The native host planner (csrc/ms_hostplan.cpp) against its specification, the Python planner: identical job tables,
field for field, bit for bit, on the canonical configs and on randomised parameter sets of the family it covers; and its
restatement of numpy's random streams against numpy itself."""
import subprocess
import os

import numpy as np
import pytest

from audio_suite_b200 import configs, hostplan, plan as P, tables as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def built():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "audio_suite_b200", "csrc"), "libms_hostplan.so"])
    hostplan._lib = None
    assert hostplan.lib() is not None


def _same(a, b, path):
    if isinstance(a, np.ndarray):
        assert isinstance(b, np.ndarray) and a.dtype == b.dtype and a.shape == b.shape, (path, a.dtype, getattr(b, "dtype", None), a.shape, getattr(b, "shape", None))
        if a.dtype.names:
            for f in a.dtype.names:
                if f.startswith("_pad") or f == "h":
                    continue
                assert np.array_equal(a[f], b[f]), (path, f, a[f][:4], b[f][:4])
        else:
            assert np.array_equal(a, b, equal_nan=True), (path, a.ravel()[:6], b.ravel()[:6])
    elif isinstance(a, tuple):
        assert len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            _same(x, y, f"{path}[{i}]")
    else:
        assert a == b, (path, a, b)


def assert_same_tables(ps):
    want = T.pack_chunk([P.plan_render(p) for p in ps])
    got = hostplan.plan_chunk(ps)
    for name in want.__dataclass_fields__:
        _same(getattr(want, name), getattr(got, name), name)


def test_numpy_streams_restated():
    l = hostplan.lib()
    rng = np.random.default_rng(7)
    for seed in [0, 1, 12345, 2 ** 31 + 5, 2 ** 32 + 17, 2 ** 40 + 3] + [int(s) for s in rng.integers(0, 2 ** 62, 20)]:
        st = np.random.PCG64(seed).state["state"]
        got = hostplan.pcg64_states([seed])[0]
        assert [int(v) for v in got] == [st["state"] >> 64, st["state"] & (2 ** 64 - 1), st["inc"] >> 64, st["inc"] & (2 ** 64 - 1)], seed
        out = np.zeros(4000)
        l.ms_hp_draws(seed, 0, 0, 4000, out.ctypes.data)
        assert np.array_equal(out, np.random.default_rng(seed).random(4000))
        l.ms_hp_draws(seed, 2, 0, 4000, out.ctypes.data)           # exponential ziggurat incl. its wedge / tail branches
        assert np.array_equal(out, np.random.default_rng(seed).exponential(1.0, 4000))
        for high in (1, 2, 3, 1000, 96000, 2 ** 31 + 11, 2 ** 32):
            l.ms_hp_draws(seed, 1, high, 1001, out.ctypes.data)     # odd count: a 32-bit half stays buffered
            assert np.array_equal(out[:1001], np.random.default_rng(seed).integers(0, high, size=1001).astype(np.float64)), (seed, high)


def test_canonical_configs_and_sweep_members():
    ir = configs.synth_ir(5.0, 48000, 303)
    assert_same_tables([configs.canonical(n) for n in ("C1", "C1b", "C2", "C3")])
    assert_same_tables([configs.c5_params(i, shared_ir=ir) for i in range(96)])
    c4 = configs.canonical("C4")
    c4["out_dur_s"] = 30.0
    assert_same_tables([c4])


def test_randomised_family():
    rng = np.random.default_rng(2026)
    irs = [configs.synth_ir(0.2, 48000, 3), configs.synth_ir(0.05, 48000, 4, channels=1), np.ones(7), None]
    ps = []
    for i in range(160):
        ir = irs[int(rng.integers(0, len(irs)))]
        ps.append(configs.with_defaults(
            base_sr=int(rng.choice([44100, 48000, 96000])), out_dur_s=float(rng.choice([0.001, 0.05003, 0.3, 1.0, 2.5])),
            time_unfold=float(rng.choice([1.0, 16.0, 25.5, 100.0, 700.0])), gen_mode=str(rng.choice(configs.BASIC_MODES)),
            micro_ms=float(rng.choice([0.05, 0.3333, 1.25, 10.0])), seed=int(rng.integers(0, 2_000_000_000)),
            dust_density=float(rng.uniform(0, 0.2)), noise_tilt=float(rng.uniform(-12, 12)), ring_hz=float(10 ** rng.uniform(1, 5)),
            partial_stretch=float(rng.choice([1.0, 1.0 + 1e-10, 0.25, 2.5, 4.0])), nl_warp_on=bool(rng.integers(0, 2)),
            unfold_mode=str(rng.choice(["Classic reinterpret", "Multi-band unfold"])), mb_roll=float(rng.choice([0.0, 2000.0])),
            bandlimit_on=bool(rng.integers(0, 2)), bandlimit_roll_hz=float(rng.choice([0.0, 2500.0])),
            event_process=str(rng.choice(["Single", "Poisson"])), grains_per_sec=float(rng.choice([0.0, 3.0, 40.0])),
            max_grains=int(rng.choice([1, 5, 4000])), grain_amp_rand=float(rng.uniform(0, 1)), grain_offset_on=bool(rng.integers(0, 2)),
            grain_offset_max_ms=float(rng.choice([0.0, 0.4, 60.0])), bp_density=str(rng.choice(["", "0:18, 4:40, 8:14"])),
            bp_unfold=str(rng.choice(["", "0:20, 1:33.3, 2:25"])), bp_cutoff=str(rng.choice(["", "0:16000, 2:6000"])),
            bp_stretch=str(rng.choice(["", "0:.5,3:2"])), er_cloud_on=bool(rng.integers(0, 2)), er_taps=int(rng.choice([1, 16, 320, 2000])),
            er_max_ms=float(rng.choice([5.0, 45.0, 150.0])), space_ir_on=bool(rng.integers(0, 2)), space_ir_max_samps=int(rng.choice([4, 256, 12000])),
            stereo_on=bool(rng.integers(0, 2)), stereo_width=float(rng.uniform(-0.2, 1.2)), sat_drive=float(rng.choice([0.0, 1.0, 6.0])),
            env_a=float(rng.choice([0.0, 0.01, 20.0])), env_d=float(rng.choice([0.0, 250.0])), env_r=float(rng.choice([0.0, 1800.0])),
            env_s=float(rng.uniform(-0.1, 1.1)), env_curve=float(rng.choice([0.0, 1.0, 1.8])), _ir_audio=ir))
    ok = []
    for p in ps:
        assert hostplan.supported(p)
        try:
            P.plan_render(p) and T.pack_chunk([P.plan_render(p)])
            ok.append(p)
        except ValueError:                                   # attack longer than the output (main_v2.py:182): both planners raise
            with pytest.raises(ValueError):
                hostplan.plan_chunk([p])
    assert len(ok) > 100
    for a in range(0, len(ok), 16):
        assert_same_tables(ok[a:a + 16])
    # the whole family at once: several blocks of 32 renders, planned on three threads and appended in order (shared
    # envelopes, impulse responses in order of first use, odd lengths -- everything that spans blocks)
    want = T.pack_chunk([P.plan_render(p) for p in ok])
    for threads in (1, 3):
        got = hostplan.plan_chunk(ok, threads)
        for name in want.__dataclass_fields__:
            _same(getattr(want, name), getattr(got, name), "%s (threads=%d)" % (name, threads))


def test_streamed_planning_of_a_mixed_batch():
    """plan_stream sends the first slice to the native planner before it has looked at the rest; a batch whose later
    renders are outside the native family must still come out slice by slice with the tables the Python planner packs."""
    ps = [configs.c5_params(i, shared_ir=configs.synth_ir(0.1, 48000, 5)) for i in range(40)]
    ps[25] = configs.with_defaults(ps[25], cep_warp_on=True)
    ps[33] = configs.with_defaults(ps[33], gen_mode="Wavelet atoms")
    got = list(T.plan_stream(ps, [8, 16], workers=2, piece=8))
    T.shutdown_pool()
    assert [len(t.post) for t in got] == [8, 16, 16]
    for t, (a, b) in zip(got, ((0, 8), (8, 24), (24, 40))):
        want = T.pack_chunk([P.plan_render(p) for p in ps[a:b]])
        if b <= 24:
            for name in want.__dataclass_fields__:
                _same(getattr(want, name), getattr(t, name), "%s [%d:%d]" % (name, a, b))
        else:       # the mixed slice comes from the worker processes as merged pieces: same renders, its own pool layout
            assert np.array_equal(want.out_n, t.out_n) and np.array_equal(want.sy1["n"], t.sy1["n"])
            assert np.array_equal(want.ola_e["amp"], t.ola_e["amp"]) and np.array_equal(want.tap_gain, t.tap_gain)


def test_unsupported_renders_stay_with_the_python_planner():
    for kw in (dict(gen_mode="Wavelet atoms"), dict(event_process="Hawkes"), dict(cep_warp_on=True), dict(spectral_imprint_on=True),
               dict(res_bank_on=True), dict(gen_mode="no such generator")):
        assert not hostplan.supported(configs.with_defaults(kw))
        with pytest.raises(hostplan.Unsupported) as e:            # the conversion notices it too (no second pass over the batch)
            hostplan.plan_chunk([configs.with_defaults(), configs.with_defaults(kw)])
        assert e.value.args[0] == 1
    with pytest.raises(hostplan.Unsupported):
        hostplan.plan_chunk([configs.with_defaults(seed=-5)])
    assert hostplan.supported(configs.with_defaults())
