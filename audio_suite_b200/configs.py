"""Parameter schema and the canonical workloads.

FACTORY_DEFAULTS mirrors what `MicrosoundV2.get_params()` returns on a freshly built
window (reference main_v2.py:1166-1266 with the widget initial values of
main_v2.py:895-1135).  A headless caller builds a parameter set
the way `on_load_preset` does (main_v2.py:1286-1291): `{**FACTORY_DEFAULTS, **preset}`.

`canonical(name)` returns the five BASELINE.json configurations.
All inputs are synthetic and seeded; nothing is read from disk.
"""
from __future__ import annotations

import numpy as np

FACTORY_DEFAULTS = {
    "base_sr": 48000, "out_dur_s": 8.0, "time_unfold": 25.0, "peak": 0.98, "sat_drive": 1.0,
    "stereo_on": True, "stereo_width": 0.65,
    "gen_mode": "Gaussian click", "micro_ms": 1.25, "seed": 12345,
    "dust_density": 0.02, "noise_tilt": -3.0, "ring_hz": 4200.0, "ring_decay_ms": 12.0,
    "crackle_alpha": 1.4, "crackle_density": 180.0, "crackle_kernel": 64,
    "ss_threshold": 0.9, "ss_build": 0.06, "ss_decay": 0.75, "ss_noise": 0.08,
    "chaos_r": 3.92, "chaos_gate": 0.35,
    "wav_base_hz": 2400.0, "wav_count": 8, "wav_spread": 0.6,
    "unfold_mode": "Classic reinterpret", "partial_stretch": 1.0,
    "partial_lock_on": False, "pl_top_n": 24, "pl_neigh": 4,
    "nl_warp_on": False, "nl_warp_power": 1.25,
    "cep_warp_on": False, "cep_factor": 1.2,
    "mb_b1": 2000.0, "mb_b2": 8000.0, "mb_b3": 20000.0,
    "mb_u1": 35.0, "mb_u2": 20.0, "mb_u3": 12.0, "mb_roll": 2000.0,
    "bandlimit_on": True, "bandlimit_out_hz": 18000.0, "bandlimit_roll_hz": 2500.0,
    "event_process": "Single", "grains_per_sec": 18.0, "max_grains": 4000,
    "grain_amp_rand": 0.35, "grain_offset_on": True, "grain_offset_max_ms": 60.0,
    "cluster_size": 6, "cluster_spread_ms": 25.0, "hawkes_gain": 0.6, "hawkes_decay_s": 0.25,
    "bp_density": "0:18, 4:40, 8:14", "bp_unfold": "", "bp_cutoff": "", "bp_stretch": "",
    "res_bank_on": False, "res_modes": 24, "res_fmin": 120.0, "res_fmax": 12000.0, "res_decay_ms": 80.0,
    "wg_on": False, "wg_lines": 8, "wg_max_ms": 8.0, "wg_fb": 0.7,
    "event_feedback_on": False, "event_feedback_amt": 0.35,
    "spectral_imprint_on": False, "spectral_imprint_amt": 0.35, "spectral_imprint_smooth": 0.92,
    "er_cloud_on": True, "er_taps": 320, "er_max_ms": 45.0,
    "space_ir_on": False, "space_ir_max_samps": 12000,
    "env_a": 20.0, "env_d": 250.0, "env_s": 0.65, "env_r": 1800.0, "env_curve": 1.8,
}

BASIC_MODES = ("Gaussian click", "Dust impulses", "Noise burst", "Skewed transient", "Resonant strike")


def with_defaults(overrides=None, **kw):
    p = dict(FACTORY_DEFAULTS)
    p["_ir_audio"] = None
    p["_img_gray"] = None
    if overrides:
        p.update(overrides)
    p.update(kw)
    return p


def synth_ir(seconds, sr, seed, channels=2):
    """Synthetic decaying-noise impulse response: N(0,1) * exp(-6.9 t / T)."""
    n = int(round(seconds * sr))
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64) / sr
    body = rng.standard_normal((n, channels)) if channels > 1 else rng.standard_normal(n)
    env = np.exp(-t * 6.9 / seconds)
    return body * (env[:, None] if channels > 1 else env)


_C1 = dict(time_unfold=100.0, micro_ms=10.0, gen_mode="Resonant strike", out_dur_s=1.0,
           event_process="Single", er_cloud_on=False, space_ir_on=False, stereo_on=False,
           grain_offset_on=False, partial_stretch=1.0)

CANONICAL = ("C1", "C1b", "C2", "C3", "C4", "C5")


def canonical(name, index=0):
    """Parameter dict for one of the BASELINE.json configs.  `index` selects the render for C5."""
    if name == "C1":
        return with_defaults(_C1)
    if name == "C1b":
        return with_defaults(_C1, time_unfold=16.0, out_dur_s=0.16)
    if name == "C2":
        return with_defaults(_C1, partial_stretch=4.0)
    if name == "C3":
        return with_defaults(_C1, partial_stretch=4.0, space_ir_on=True, space_ir_max_samps=240000,
                             stereo_on=True, _ir_audio=synth_ir(5.0, 48000, 303))
    if name == "C4":
        return with_defaults(base_sr=96000, out_dur_s=600.0, time_unfold=500.0, micro_ms=10.0,
                             partial_stretch=2.5, gen_mode="Resonant strike", event_process="Poisson",
                             grains_per_sec=2.0, bp_density="", max_grains=50000, space_ir_on=True,
                             space_ir_max_samps=960000, er_cloud_on=True, stereo_on=True,
                             _ir_audio=synth_ir(10.0, 96000, 404))
    if name == "C5":
        return c5_params(index)
    raise KeyError(name)


_C5_IR = None


def c5_params(i, shared_ir=None):
    """Render i of the 4096-render preset sweep: C3 + 2 s output + ER cloud, with seed, unfold,
    stretch, generator mode, ring frequency and noise tilt drawn from default_rng([20260101, i])
    (the reference's own batch axes are seed x unfold x stretch, main_v2.py:1578-1584)."""
    global _C5_IR
    if shared_ir is None:
        if _C5_IR is None:
            _C5_IR = synth_ir(5.0, 48000, 303)
        shared_ir = _C5_IR
    rng = np.random.default_rng([20260101, int(i)])
    seed = int(rng.integers(0, 2_000_000_000))
    unfold = float(rng.integers(25, 201))
    stretch = round(float(rng.uniform(0.25, 4.0)), 2)
    mode = BASIC_MODES[int(rng.integers(0, 5))]
    ring = float(10.0 ** rng.uniform(2.0, 4.5))
    tilt = float(rng.uniform(-12.0, 12.0))
    return with_defaults(_C1, space_ir_on=True, space_ir_max_samps=240000, stereo_on=True,
                         out_dur_s=2.0, er_cloud_on=True, seed=seed, time_unfold=unfold,
                         partial_stretch=stretch, gen_mode=mode, ring_hz=ring, noise_tilt=tilt,
                         _ir_audio=shared_ir)
