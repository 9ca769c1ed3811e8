"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the Microsound offline render path.

A from-scratch float64 numpy restatement of what
`/root/reference/microsound_0.2.1/main_v2.py` (abbreviated `M:` below) computes on
the hot path named by BASELINE.json: `render()` (M:588-792) and the DSP helpers it
reaches for the five `gen_basic` generators, FFT band-limit, spectral stretch,
multiband unfold, overlap-add placement, ADSR, early-reflection cloud, short-IR
convolution, stereo diffusion, soft clip and normalise.  It is written against
the *behaviour* of those functions, stage by stage, not transcribed from them.

Parity status: PINNED.  The reference ships no tests or golden vectors
(SURVEY.md section 4), so this oracle is pinned by running the unmodified
reference file in the build container (oracle/ref_loader.py) on the same
parameters and seeds: tests/test_oracle_vs_reference.py requires <= 1e-12
max-abs agreement stage by stage and end to end, and tests/golden/*.npz holds
outputs of the reference itself (written by oracle/make_golden.py, numpy 2.3.5)
so the pin travels to machines where /root/reference does not exist.

The arithmetic that lives in third-party code is numpy's (version unpinned by the
reference, M README.txt:4; 2.3.5 here): pocketfft (`np.fft`), PCG64 +
ziggurat (`np.random.default_rng`), `np.interp`, `np.convolve`.  This file calls
the same numpy entry points, so those streams are identical by construction.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product (audio_suite_b200) never does.

Out of scope here ("next" rows): the non-basic generators and the
per-event extras (partial lock, power warp, cepstral warp, resonator bank,
waveguide, event feedback, spectral imprint).  `render` raises
NotImplementedError when a parameter set asks for one of them.
"""
from __future__ import annotations

import math

import numpy as np

BASIC_MODES = ("Gaussian click", "Dust impulses", "Noise burst", "Skewed transient", "Resonant strike")
DESIGN_SR_CAP = 30_000_000  # M:597, M:646
IR_TAP_CAP = 8192           # M:443


# --------------------------------------------------------------------------- post chain
def soft_saturate(x, drive):
    """M:31-34 -- tanh(x*d)/tanh(d); identity when d <= 0."""
    d = float(drive)
    if d <= 0:
        return x
    return np.tanh(x * d) / np.tanh(d)


def peak_normalize(x, peak):
    """M:26-29 -- scale so that max|x| == peak over the whole array (both channels)."""
    top = float(np.max(np.abs(x))) if x.size else 0.0
    if top <= 0:
        return x
    return x * (peak / top)


# --------------------------------------------------------------------------- spectral shaping
def _bin_freqs(n, sr):
    # M:36-37 -- np.fft.rfftfreq(n, d=1/sr); values are k * (1/(n*d)).
    return np.fft.rfftfreq(n, d=1.0 / sr)


def _fall_weights(f, f0, f1):
    # raised-cosine 1 -> 0 between f0 and f1 (M:56-57, M:98-99)
    return 0.5 * (1.0 + np.cos(np.pi * ((f - f0) / max(1e-12, f1 - f0))))


def _rise_weights(f, f0, f1):
    # raised-cosine 0 -> 1 between f0 and f1 (M:84-85)
    return 0.5 * (1.0 - np.cos(np.pi * ((f - f0) / max(1e-12, f1 - f0))))


def fft_lowpass(x, sr, cutoff, roll=0.0):
    """M:39-59 -- whole-grain rFFT mask: brick wall (roll<=0) or cosine edge of width roll."""
    n = len(x)
    if n < 8:
        return x
    nyq = 0.5 * sr
    fc = float(np.clip(cutoff, 1.0, nyq))
    width = float(max(0.0, roll))
    spec = np.fft.rfft(x)
    f = _bin_freqs(n, sr)
    if width <= 0:
        spec[f > fc] = 0.0
    else:
        top = min(nyq, fc + width)
        spec[f > top] = 0.0
        edge = (f >= fc) & (f <= top)
        if edge.any():
            spec[edge] *= _fall_weights(f[edge], fc, top)
    return np.fft.irfft(spec, n=n)


def fft_bandpass(x, sr, lo, hi, roll=0.0):
    """M:61-101 -- two-sided mask; cosine skirts lie outside [lo, hi]."""
    n = len(x)
    if n < 8:
        return x
    lo = max(0.0, float(lo))
    hi = max(lo, float(hi))
    spec = np.fft.rfft(x)          # (the reference transforms before the hi<=0 test too)
    f = _bin_freqs(n, sr)
    nyq = 0.5 * sr
    hi = min(hi, nyq)
    if hi <= 0:
        return np.zeros_like(x)
    width = float(max(0.0, roll))
    if lo > 0:
        if width <= 0:
            spec[f < lo] = 0.0
        else:
            a = max(0.0, lo - width)
            spec[f < a] = 0.0
            edge = (f >= a) & (f <= lo)
            if edge.any():
                spec[edge] *= _rise_weights(f[edge], a, lo)
    if hi < nyq:
        if width <= 0:
            spec[f > hi] = 0.0
        else:
            b = min(nyq, hi + width)
            spec[f > b] = 0.0
            edge = (f >= hi) & (f <= b)
            if edge.any():
                spec[edge] *= _fall_weights(f[edge], hi, b)
    return np.fft.irfft(spec, n=n)


def spectrum_stretch(x, factor):
    """M:117-128 -- Y[k] = lerp(X, k/factor) on Re and Im separately, zero outside; irfft."""
    n = len(x)
    if n < 16:
        return x
    factor = float(factor)
    if abs(factor - 1.0) < 1e-9:
        return x
    spec = np.fft.rfft(x)
    k = np.arange(spec.size, dtype=np.float64)
    src = k / max(1e-12, factor)
    y = np.interp(src, k, spec.real, left=0.0, right=0.0) + 1j * np.interp(src, k, spec.imag, left=0.0, right=0.0)
    return np.fft.irfft(y, n=n)


def spectrum_power_warp(x, power):
    """M:103-115 -- bin k of the output spectrum is the input spectrum interpolated (Re and Im separately, zeros
    outside) at (k / kmax) ** (1 / power) * kmax."""
    n = len(x)
    if n < 16:
        return x
    spec = np.fft.rfft(x)
    k = np.arange(spec.size, dtype=np.float64)
    kmax = max(1.0, k[-1])
    src = np.power(k / kmax, 1.0 / max(1e-6, float(power))) * kmax
    warped = np.interp(src, k, spec.real, left=0.0, right=0.0) + 1j * np.interp(src, k, spec.imag, left=0.0, right=0.0)
    return np.fft.irfft(warped, n=n)


def partial_lock(x, factor, top_n=24, neighborhood=4):
    """M:130-148 -- the top_n strongest rfft bins (DC excluded) are moved to round(k * factor) with a triangular
    spread over +-neighborhood bins, on top of 12 % of the original spectrum."""
    n = len(x)
    factor = float(factor)
    if n < 64 or abs(factor - 1.0) < 1e-9:
        return x
    spec = np.fft.rfft(x)
    strongest = np.argsort(np.abs(spec)[1:])[-top_n:] + 1
    moved = np.zeros_like(spec)
    for k in strongest:
        centre = int(round(k * factor))
        if 1 <= centre < moved.size:
            for d in range(-neighborhood, neighborhood + 1):
                if 1 <= centre + d < moved.size:
                    moved[centre + d] += spec[k] * (1.0 - abs(d) / (neighborhood + 1))
    return np.fft.irfft(moved + 0.12 * spec, n=n)


def cepstrum_warp(x, factor):
    """M:150-163 -- the real cepstrum of the grain (irfft of log(|X| + 1e-12)) is resampled on the quefrency axis
    (linear interpolation at t / factor, zeros outside), turned back into a log-magnitude (real part of its rfft)
    and recombined with the original phases."""
    n = len(x)
    if n < 64:
        return x
    spec = np.fft.rfft(x)
    mag = cepstrum_warp_magnitudes(spec, n, factor)
    return np.fft.irfft(mag * np.exp(1j * np.angle(spec)), n=n)


def cepstrum_warp_magnitudes(spec, n, factor):
    """M:154-161 on a given rfft spectrum: the new magnitudes exp(Re rfft(warped cepstrum)).  Split out so the
    stage-level parity tests can hand the oracle the SAME spectrum the device kernels saw: in bins the band-limit has
    emptied, |X| is rounding noise next to the +1e-12, so log|X| there differs between any two FFTs by ~1e-4."""
    cep = np.fft.irfft(np.log(np.abs(spec) + 1e-12), n=n)
    t = np.arange(n, dtype=np.float64)
    warped = np.interp(t / max(1e-12, float(factor)), t, cep, left=0.0, right=0.0)
    return np.exp(np.fft.rfft(warped).real)


def resonator_modes(seed, modes, f_min, f_max):
    """Scalar draws of resonator_bank (M:370, 377-380): per mode a log-spaced frequency detuned by 2**U(-.02, .02),
    a phase in [0, 2 pi) and the weight 1 / (1 + 0.35 k)."""
    rng = np.random.default_rng(int(seed) + 321)
    rows = []
    for k in range(int(max(1, modes))):
        f = float(f_min) * ((float(f_max) / max(1.0, float(f_min))) ** (k / max(1, modes - 1)))
        f *= 2.0 ** rng.uniform(-0.02, 0.02)
        rows.append((f, rng.uniform(0, 2 * np.pi), 1.0 / (1 + k * 0.35)))
    return rows


def resonator_bank(x, sr, modes, f_min, f_max, decay_ms, seed):
    """M:369-384 -- a bank of decaying sinusoids, peak-normalised, mixed in with the SIGN of the input:
    0.55 x + 0.45 bank sign(x).  (Where x is rounding noise its sign is too: see rounding_noise_floor.)"""
    n = len(x)
    if n < 32:
        return x
    t = np.arange(n, dtype=np.float64) / sr
    env = np.exp(-t / max(1e-6, decay_ms / 1000.0))
    bank = np.zeros_like(x)
    for f, ph, w in resonator_modes(seed, modes, f_min, f_max):
        bank += w * np.sin(2 * np.pi * f * t + ph) * env
    bank = bank / max(1e-12, np.max(np.abs(bank)))
    return 0.55 * x + 0.45 * (x * 0.0 + bank) * np.sign(x)


def waveguide_lines(seed, sr, lines, max_ms, feedback):
    """Scalar draws of waveguide_splinters (M:387, 391-396): per line the delay in samples, the loop gain and the mix."""
    rng = np.random.default_rng(int(seed) + 777)
    rows = []
    for _ in range(int(max(1, lines))):
        d = int(max(1, round((rng.uniform(0.4, max_ms) / 1000.0) * sr)))
        g = feedback * rng.uniform(0.6, 0.98)
        rows.append((d, g, rng.uniform(0.15, 0.45)))
    return rows


def waveguide(x, sr, lines, max_ms, feedback, seed):
    """M:386-402 -- a cascade of feedback comb filters: per line v[t] = y[t] + g v[t - d], y[t] <- (1 - mix) y[t] + mix v[t]."""
    n = len(x)
    if n < 64:
        return x
    y = x.copy()
    for d, g, mix in waveguide_lines(seed, sr, lines, max_ms, feedback):
        v = y.copy()
        for t in range(d, n):
            v[t] = y[t] + g * v[t - d]
        y = (1.0 - mix) * y + mix * v
    return y


def multiband_unfold(x, gen_sr, bands_out_hz, unfolds, roll_hz):
    """M:492-500 -- sum of band-passed copies, band edges scaled by each band's unfold."""
    acc = None
    for (lo, hi), u in zip(bands_out_hz, unfolds):
        part = fft_bandpass(x, gen_sr, lo * u, hi * u, roll=roll_hz)
        acc = part if acc is None else acc + part
    return x if acc is None else acc


# --------------------------------------------------------------------------- envelope / space
def adsr_envelope(n, sr, a_ms, d_ms, sustain, r_ms, curve):
    """M:172-195."""
    na = max(0, int(round(sr * a_ms / 1000.0)))
    nd = max(0, int(round(sr * d_ms / 1000.0)))
    nr = max(0, int(round(sr * r_ms / 1000.0)))
    s = float(np.clip(sustain, 0, 1))
    c = float(max(1e-6, curve))
    env = np.ones(n, dtype=np.float64)
    pos = 0
    if na > 0:
        env[:na] = np.linspace(0, 1, na, endpoint=False) ** c      # short n truncates the ramp
        pos = na
    dec_end = min(n, pos + nd)
    if nd > 0 and dec_end > pos:
        env[pos:dec_end] = 1.0 - (1.0 - s) * np.linspace(0, 1, dec_end - pos, endpoint=False) ** c
    rel_start = max(dec_end, n - nr)
    if rel_start > dec_end:
        env[dec_end:rel_start] = s
    if nr > 0 and n > rel_start:
        env[rel_start:] = s * (1.0 - np.linspace(0, 1, n - rel_start, endpoint=True) ** c)
    return env


def reflection_taps(sr, taps, max_ms, seed):
    """Delays (samples) and gains of M:409-421; kept as a separate function because the
    integer offsets are part of the bit-exact contract."""
    rng = np.random.default_rng(int(seed) + 202)
    delays = rng.uniform(0.3, max_ms, size=int(max(1, taps))) / 1000.0
    gains = rng.uniform(-1.0, 1.0, size=delays.size)
    gains *= np.exp(-delays * 42.0)
    offs = np.array([int(round(d * sr)) for d in delays], dtype=np.int64)
    return offs, gains


def reflection_cloud(x, sr, taps, max_ms, seed):
    """M:409-421 -- y = x + sum_t g_t * delay(x, off_t); taps with off<=0 or off>=n dropped."""
    n = len(x)
    offs, gains = reflection_taps(sr, taps, max_ms, seed)
    y = x.copy()
    for off, g in zip(offs.tolist(), gains.tolist()):
        if 0 < off < n:
            y[off:] += g * x[:-off]
    return y


def short_ir_convolve(x, ir):
    """M:438-445 -- mono-mix, first 8192 taps, causal direct convolution truncated to len(x)."""
    if ir is None or ir.size < 8:
        return x
    h = ir.astype(np.float64)
    if h.ndim > 1:
        h = h.mean(axis=1)
    h = h[:min(h.size, IR_TAP_CAP)]
    return np.convolve(x, h, mode="full")[:len(x)]


def stereo_shifts(sr, width):
    """Integer roll amounts of M:428-429."""
    w = float(np.clip(width, 0.0, 1.0))
    return int(round((1 + 7 * w) * 0.0005 * sr)), int(round((1 + 9 * w) * 0.0007 * sr)), w


def stereo_diffuse(x, sr, width):
    """M:423-436 -- L = roll(x, dl); R = irfft(rfft(roll(x,-dr)) * exp(i*0.9w*sin(2 pi k/kmax)))."""
    n = len(x)
    if n < 64:
        return np.column_stack([x, x])
    dl, dr, w = stereo_shifts(sr, width)
    left = np.roll(x, dl)
    spec = np.fft.rfft(np.roll(x, -dr))
    k = np.arange(spec.size, dtype=np.float64)
    spec = spec * np.exp(1j * (w * 0.9) * np.sin(2 * np.pi * k / max(1.0, k[-1])))
    return np.column_stack([left, np.fft.irfft(spec, n=n)])


# --------------------------------------------------------------------------- breakpoints / events
def parse_lane(text):
    """M:452-467 -- 't:v, t:v' -> sorted [(t, v)]; parts without ':' or with unparsable numbers are skipped silently,
    a part with more than one ':' raises ValueError (the unpacking at M:461 sits outside the reference's try)."""
    pts = []
    for part in (text or "").strip().split(","):
        part = part.strip()
        if not part or ":" not in part:
            continue
        t, v = part.split(":")
        try:
            pts.append((float(t.strip()), float(v.strip())))
        except Exception:
            continue
    pts.sort(key=lambda p: p[0])
    return pts


def lane_value(pts, t, default):
    """M:469-482 -- piecewise-linear, clamped at both ends."""
    if not pts:
        return default
    if t <= pts[0][0]:
        return pts[0][1]
    if t >= pts[-1][0]:
        return pts[-1][1]
    for (t0, v0), (t1, v1) in zip(pts[:-1], pts[1:]):
        if t0 <= t <= t1:
            a = (t - t0) / max(1e-12, t1 - t0)
            return (1 - a) * v0 + a * v1
    return default


def event_times(process, dur_s, rate, seed, cluster_size=6, cluster_spread_ms=25.0,
                hawkes_gain=0.6, hawkes_decay_s=0.25):
    """M:507-558 -- event onsets in seconds (python floats), stream default_rng(seed+9999)."""
    rng = np.random.default_rng(int(seed) + 9999)
    if process == "Single" or rate <= 0:
        return [0.0]
    out = []
    if process == "Poisson":
        t = 0.0
        while t < dur_s:
            t += rng.exponential(1.0 / rate)
            if t < dur_s:
                out.append(t)
        return out
    if process == "Clustered":
        parents = []
        t = 0.0
        prate = max(0.1, rate / max(1, cluster_size))
        while t < dur_s:
            t += rng.exponential(1.0 / prate)
            if t < dur_s:
                parents.append(t)
        spread = cluster_spread_ms / 1000.0
        for p in parents:
            kids = int(max(1, round(rng.uniform(0.6, 1.4) * cluster_size)))
            for _ in range(kids):
                tt = p + rng.normal(0.0, spread)
                if 0.0 <= tt < dur_s:
                    out.append(tt)
        out.sort()
        return out
    if process == "Hawkes":
        dt = 0.002
        act = 0.0
        for i in range(int(math.ceil(dur_s / dt))):
            act *= math.exp(-dt / max(1e-6, hawkes_decay_s))
            p = min(0.95, (rate + hawkes_gain * act * rate) * dt)
            if rng.random() < p:
                out.append(i * dt + rng.uniform(0, dt))
                act += 1.0
        return out
    return out


# --------------------------------------------------------------------------- generators
def grain_length(gen_sr, micro_ms, floor=16):
    """M:221 (floor 16 for gen_basic)."""
    return int(max(floor, round(gen_sr * micro_ms / 1000.0)))


def _tilted_noise(rng, n, gen_sr, tilt_db_per_oct):
    # M:224-233 -- white normal noise, spectrum scaled by (f/f[1])**alpha with f[0]:=f[1]
    w = rng.standard_normal(n)
    spec = np.fft.rfft(w)
    f = _bin_freqs(n, gen_sr)
    if f.size > 1:
        f[0] = f[1]
    alpha = math.log(10.0 ** (tilt_db_per_oct / 20.0), 2.0)
    spec *= (f / max(1e-12, f[1])) ** alpha
    return np.fft.irfft(spec, n=n)


def edge_fade(n):
    """M:265-268 -- linear fade in / out over max(8, int(0.01 n)) samples."""
    m = max(8, int(0.01 * n))
    w = np.ones(n, dtype=np.float64)
    w[:m] *= np.linspace(0, 1, m, endpoint=False)
    w[-m:] *= np.linspace(1, 0, m, endpoint=False)
    return w


def basic_transient(gen_sr, micro_ms, seed, mode, dust_density, tilt_db_per_oct, ring_hz, ring_decay_ms):
    """M:219-269 -- the five `gen_basic` modes (+ its fallback branch)."""
    rng = np.random.default_rng(int(seed))
    n = grain_length(gen_sr, micro_ms)
    t = np.arange(n, dtype=np.float64) / gen_sr
    micro_s = micro_ms / 1000.0
    if mode == "Gaussian click":
        sigma = max(1, int(0.0025 * n))
        bell = np.exp(-0.5 * ((np.arange(n) / sigma) ** 2))
        x = bell * (rng.standard_normal(n) * 0.12 + 1.0)
    elif mode == "Dust impulses":
        x = np.zeros(n, dtype=np.float64)
        k = int(max(1, round(dust_density * n)))
        where = rng.integers(0, n, size=k)
        x[where] = rng.uniform(-1, 1, size=k)          # duplicates: last write wins
        ker = np.exp(-np.linspace(0, 6, max(8, int(0.01 * n))))
        x = np.convolve(x, ker, mode="same")
    elif mode == "Noise burst":
        x = _tilted_noise(rng, n, gen_sr, tilt_db_per_oct) * np.exp(-t / max(1e-6, micro_s * 0.25))
    elif mode == "Skewed transient":
        w = np.maximum(0.0, _tilted_noise(rng, n, gen_sr, tilt_db_per_oct))
        x = np.diff(w, prepend=w[0]) * np.exp(-t / max(1e-6, micro_s * 0.2))
    elif mode == "Resonant strike":
        f = max(10.0, float(ring_hz))
        tau = max(1e-6, float(ring_decay_ms) / 1000.0)
        tone = np.sin(2 * np.pi * f * t) * np.exp(-t / tau)
        exc = rng.standard_normal(n) * np.exp(-t / max(1e-6, micro_s * 0.15))
        x = 0.9 * tone + 0.25 * exc
    else:
        x = rng.standard_normal(n) * 0.1
    return x * edge_fade(n)


def raised_cosine_window(n):
    """M:17-21 -- symmetric Hann window (ones for n <= 1)."""
    if n <= 1:
        return np.ones(n, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n, dtype=np.float64) / (n - 1))


WAVELET_FLOOR = 128          # M:319: wavelet grains are at least 128 samples (gen_basic: 16)


def wavelet_atom_draws(seed, micro_ms, n, base_hz, count, spread):
    """Scalar draws of gen_wavelet_atoms in stream order (M:322-327): per atom centre frequency, Gaussian width in
    ms, phase and circular shift.  Shared with the product's planner tests."""
    rng = np.random.default_rng(int(seed))
    rows = []
    for k in range(int(max(1, count))):
        f0 = base_hz * (2.0 ** rng.uniform(-spread, spread))
        sigma_ms = max(0.03, micro_ms * rng.uniform(0.04, 0.18))
        phase = rng.uniform(0, 2 * np.pi)
        shift = int(rng.integers(-n // 8, n // 8))
        rows.append((f0, sigma_ms, phase, shift, 1.0 / (1 + k * 0.6)))
    return rows


def wavelet_atoms(gen_sr, micro_ms, seed, base_hz, count, spread):
    """M:317-331 + morlet_atom M:165-170: a sum of circularly shifted Gaussian-windowed cosines under a Hann
    window.  The atoms are built with gen_basic's length rule (floor 16, M:166) while the grain uses floor 128
    (M:319): when they differ the reference's `x += atom[:n]` raises; so does this."""
    n = grain_length(gen_sr, micro_ms, floor=WAVELET_FLOOR)
    n_atom = grain_length(gen_sr, micro_ms, floor=16)
    x = np.zeros(n, dtype=np.float64)
    t = (np.arange(n_atom, dtype=np.float64) - (n_atom / 2)) / gen_sr
    for f0, sigma_ms, phase, shift, weight in wavelet_atom_draws(seed, micro_ms, n, base_hz, count, spread):
        sigma = max(1e-9, sigma_ms / 1000.0)
        atom = np.exp(-0.5 * (t / sigma) ** 2) * np.cos(2 * np.pi * f0 * t + phase)
        x += weight * np.roll(atom, shift)[:n]          # ValueError when n_atom < n, like the reference
    return x * raised_cosine_window(n)


def crackle_impulses(seed, n, alpha, density):
    """Scalar draws of gen_crackle (M:272-279): Pareto-distributed gaps accumulate into impulse times; every time
    below n gets one uniform amplitude, amplitudes landing on the same sample add up.  Returns (positions, sums)."""
    rng = np.random.default_rng(int(seed))
    times = np.cumsum(rng.pareto(alpha, int(max(8, density))))
    times = times[times < n].astype(int)
    x = np.zeros(n, dtype=np.float64)
    for ti in times:
        x[ti] += rng.uniform(-1, 1)
    pos = np.unique(times)
    return pos, x[pos]


def crackle(gen_sr, micro_ms, seed, alpha, density, kernel):
    """M:271-281 -- sparse impulses convolved ("same") with exp(-linspace(0, 6, max(8, kernel)))."""
    n = grain_length(gen_sr, micro_ms)
    x = np.zeros(n, dtype=np.float64)
    pos, val = crackle_impulses(seed, n, alpha, density)
    x[pos] = val
    return np.convolve(x, np.exp(-np.linspace(0, 6, max(8, int(kernel)))), mode="same")


def ir_fragment(ir_audio, gen_sr, micro_ms, seed):
    """M:333-348 -- 256 samples of the loaded IR from a random start, stretched to the grain length by linear
    interpolation, Hann-windowed, peak-normalised to 0.9; silence (gen_basic's length) when no IR is loaded."""
    rng = np.random.default_rng(int(seed))
    if ir_audio is None or ir_audio.size < 32:
        return np.zeros(grain_length(gen_sr, micro_ms))
    n = grain_length(gen_sr, micro_ms, floor=64)
    src = ir_audio.astype(np.float64)
    if src.ndim > 1:
        src = src.mean(axis=1)
    start = rng.integers(0, max(1, src.size - 256))
    piece = src[start:start + 256]
    x = np.interp(np.linspace(0, 1, n), np.linspace(0, 1, piece.size), piece) * raised_cosine_window(n)
    return peak_normalize(x, 0.9)


def image_scanline(img_gray, gen_sr, micro_ms, seed, with_note=False):
    """M:350-362 -- one random row of the loaded grey image, centred, stretched to the grain length, Hann-windowed
    and smoothed by exp(-linspace(0, 5, 48)); silence when no image is loaded.  `with_note`: also return the
    progress note the reference hands back (M:354, 362)."""
    rng = np.random.default_rng(int(seed))
    n = grain_length(gen_sr, micro_ms, floor=64)
    if img_gray is None:
        x = np.zeros(n, dtype=np.float64)
        return (x, "No image loaded") if with_note else x
    h, w = img_gray.shape
    y = int(rng.integers(0, h))
    row = img_gray[y, :].astype(np.float64) / 255.0
    row = (row - row.mean()) * 2.0
    x = np.interp(np.linspace(0, 1, n), np.linspace(0, 1, w), row) * raised_cosine_window(n)
    x = np.convolve(x, np.exp(-np.linspace(0, 5, 48)), mode="same")
    return (x, f"Image line y={y}") if with_note else x


def micro_chaos(gen_sr, micro_ms, seed, r, gate):
    """M:303-315 -- a logistic map sampled through a random gate (one uniform draw per sample), smoothed ("same") by
    exp(-linspace(0, 5, 48)) and Hann-windowed.  The map is iterated in Python floats: (r * y) * (1.0 - y)."""
    rng = np.random.default_rng(int(seed))
    n = grain_length(gen_sr, micro_ms, floor=64)
    x = np.zeros(n, dtype=np.float64)
    y = (int(seed) % 10000) / 10000.0
    for i in range(n):
        y = r * y * (1.0 - y)
        if rng.random() < gate:
            x[i] = y - 0.5
    return np.convolve(x, np.exp(-np.linspace(0, 5, 48)), mode="same") * raised_cosine_window(n)


def stick_slip(gen_sr, micro_ms, seed, threshold, build, decay, noise):
    """M:283-301 -- a two-state friction model driven by one normal draw per sample: while sticking the force builds
    up (output 0) until it exceeds the threshold; while slipping the output is the force plus noise and the force
    decays until it falls under 0.02.  Hann-windowed."""
    rng = np.random.default_rng(int(seed))
    n = grain_length(gen_sr, micro_ms, floor=64)
    x = np.zeros(n, dtype=np.float64)
    sticking, force = True, 0.0
    for i in range(n):
        z = rng.standard_normal()
        if sticking:
            force += build * (z * noise + 0.2)
            if abs(force) > threshold:
                sticking = False
        else:
            x[i] = force + 0.25 * z
            force *= decay
            if abs(force) < 0.02:
                sticking, force = True, 0.0
    return x * raised_cosine_window(n)


def generator_floor(mode, params):
    """Minimum grain length of each generator (M:221, 273, 319, 337-338, 352)."""
    if mode == "Wavelet atoms":
        return WAVELET_FLOOR
    if mode in ("Image scanline", "Micro-chaos", "Stick–slip friction"):
        return 64
    if mode == "IR fragment":
        ir = params.get("_ir_audio")
        return 16 if (ir is None or np.asarray(ir).size < 32) else 64
    return 16


class ImprintMemory:
    """SpectralImprint (M:565-581): an exponential moving average of the grains' magnitude spectra, carried from
    event to event of one render and restarted whenever the spectrum length changes."""

    def __init__(self):
        self.mem = None

    def apply(self, x, amount, smooth):
        n = len(x)
        if n < 64 or amount <= 0:
            return x
        spec = np.fft.rfft(x)
        blended = self.blend(np.abs(spec), amount, smooth)
        return np.fft.irfft(blended * np.exp(1j * np.angle(spec)), n=n)

    def blend(self, mag, amount, smooth):
        """M:575-579 on given magnitudes: advance the moving average by one grain and return the blended magnitudes."""
        if self.mem is None or self.mem.size != mag.size:
            self.mem = mag.copy()
        else:
            self.mem = smooth * self.mem + (1.0 - smooth) * mag
        return (1.0 - amount) * mag + amount * self.mem


def rounding_noise_floor(params, reference_audio=None):
    """max-abs change of the rendered audio under a 1e-15 relative perturbation of the grains entering the cepstral
    warp / the spectral imprint: the part of the reference's output that is decided by float64 rounding noise
    (0.0 when none is on).  resonator_bank multiplies by np.sign(x): where the grain has decayed below its own
    rounding noise the sign -- and with it 45 % of the output there -- is noise."""
    if not (params["spectral_imprint_on"] or params["cep_warp_on"] or params["res_bank_on"]):
        return 0.0
    a = reference_audio if reference_audio is not None else render(params)[0]
    b = render(params, jitter=1e-15)[0]
    return float(np.max(np.abs(a - b)))


imprint_noise_floor = rounding_noise_floor


# --------------------------------------------------------------------------- render
_UNSUPPORTED_FLAGS = ()


def design_rate(base_sr, unfold):
    """M:596-597 / M:645-646 -- round then clip to [base_sr, 30 MHz]."""
    return int(np.clip(int(round(base_sr * unfold)), base_sr, DESIGN_SR_CAP))


def plan_events(params):
    """Everything `render` decides with scalars before touching audio (M:589-646, 742-753):
    returns a list of dicts with the integer segment map and per-event scalars.  Used by the
    parity tests to compare the product's planner integer-for-integer."""
    base_sr = int(params["base_sr"])
    out_dur = float(params["out_dur_s"])
    out_n = int(max(1, round(out_dur * base_sr)))
    base_unfold = max(1.0, float(params["time_unfold"]))
    lanes = [parse_lane(params[k]) for k in ("bp_density", "bp_unfold", "bp_cutoff", "bp_stretch")]
    rate = float(params["grains_per_sec"])
    times = event_times(params["event_process"], out_dur, rate, int(params["seed"]),
                        int(params["cluster_size"]), float(params["cluster_spread_ms"]),
                        float(params["hawkes_gain"]), float(params["hawkes_decay_s"]))
    times = times[:int(params["max_grains"])]
    rng = np.random.default_rng(int(params["seed"]) + 123456)
    micro_ms = float(params["micro_ms"])
    spread = float(params["grain_amp_rand"])
    rows = []
    for i, t0 in enumerate(times):
        dens = lane_value(lanes[0], t0, rate)
        ufac = lane_value(lanes[1], t0, base_unfold)
        cutoff_out = lane_value(lanes[2], t0, float(params["bandlimit_out_hz"]))
        stretch = lane_value(lanes[3], t0, float(params["partial_stretch"]))
        amp = 1.0
        if rate > 0:
            amp *= np.clip(dens / max(1e-6, rate), 0.15, 4.0)
        amp *= rng.uniform(1.0 - spread, 1.0 + spread)
        ufac = max(1.0, float(ufac))
        sr_evt = design_rate(base_sr, ufac)
        n = grain_length(sr_evt, micro_ms, floor=generator_floor(params["gen_mode"], params))
        start = int(round(t0 * base_sr))
        row = dict(index=i, t0=t0, amp=float(amp), ufac=ufac, gen_sr=sr_evt, n=n,
                   cutoff_gen=cutoff_out * ufac, stretch=float(stretch), start=start,
                   offset=0, length=0, placed=False)
        if start < out_n:
            if params["grain_offset_on"]:
                max_off = int(round((float(params["grain_offset_max_ms"]) / 1000.0) * base_sr))
                if max_off > 0:
                    row["offset"] = int(rng.integers(0, max(1, min(max_off, n))))
            row["length"] = max(0, min(out_n - start, n - row["offset"]))
            row["placed"] = row["length"] > 0
        rows.append(row)
    return dict(base_sr=base_sr, out_n=out_n, design_sr_base=design_rate(base_sr, base_unfold), events=rows)


def render(params, progress=None, taps=None, jitter=None):
    """M:588-792.  `taps` (optional dict) receives intermediate buffers for stage-level tests.

    `jitter` (tests only): relative size of a seeded perturbation added to every grain right before the cepstral
    warp and before the spectral imprint.  Both keep the PHASE of every rfft bin and give it a new magnitude; in
    bins the band-limit has emptied, the phase is that of float64 rounding noise (|X| ~ 1e-16).  The imprint's
    moving average may still remember energy there (a falling `bp_cutoff` lane), and the warped cepstrum of a
    band-limited grain puts exp(-5..-16) there against an output of 1e-3: a 1e-17 relative perturbation of the
    input moves cepstral_warp's output by 2.4 %.  Rendering with and without a 1e-15 jitter measures how much of the
    reference's output is decided by rounding noise (rounding_noise_floor)."""
    for flag in _UNSUPPORTED_FLAGS:
        if params[flag]:
            raise NotImplementedError(f"oracle: '{flag}' is a SURVEY 8(f) 'next' row, not restated yet")
    mode = params["gen_mode"]
    plan = plan_events(params)
    base_sr, out_n = plan["base_sr"], plan["out_n"]
    if progress:
        progress(0, f"Output SR {base_sr} Hz | Design SR {plan['design_sr_base']} Hz")
    mix = np.zeros(out_n, dtype=np.float64)
    micro_last = grain_last = None
    seed = int(params["seed"])
    micro_ms = float(params["micro_ms"])
    n_evt = len(plan["events"])
    imprint = ImprintMemory() if params["spectral_imprint_on"] else None          # M:625
    previous = None                                                               # M:626: last event's grain, as placed
    for ev in plan["events"]:
        i = ev["index"]
        note = ""                                                                 # M:651
        if mode == "Wavelet atoms":
            g = wavelet_atoms(ev["gen_sr"], micro_ms, seed + i, float(params["wav_base_hz"]), int(params["wav_count"]),
                              float(params["wav_spread"]))
        elif mode == "Crackle / corona":
            g = crackle(ev["gen_sr"], micro_ms, seed + i, float(params["crackle_alpha"]), float(params["crackle_density"]),
                        int(params["crackle_kernel"]))
        elif mode == "Stick–slip friction":
            g = stick_slip(ev["gen_sr"], micro_ms, seed + i, float(params["ss_threshold"]), float(params["ss_build"]),
                           float(params["ss_decay"]), float(params["ss_noise"]))
        elif mode == "Micro-chaos":
            g = micro_chaos(ev["gen_sr"], micro_ms, seed + i, float(params["chaos_r"]), float(params["chaos_gate"]))
        elif mode == "IR fragment":
            ir_a = params.get("_ir_audio")
            g = ir_fragment(ir_a, ev["gen_sr"], micro_ms, seed + i)
            note = "No IR loaded" if (ir_a is None or ir_a.size < 32) else "IR fragment"       # M:336, 348
        elif mode == "Image scanline":
            g, note = image_scanline(params.get("_img_gray"), ev["gen_sr"], micro_ms, seed + i, with_note=True)
        elif mode in BASIC_MODES:
            g = basic_transient(ev["gen_sr"], micro_ms, seed + i, mode, float(params["dust_density"]),
                                float(params["noise_tilt"]), float(params["ring_hz"]),
                                float(params["ring_decay_ms"]))
        else:   # unknown mode string: M:686
            g = basic_transient(ev["gen_sr"], micro_ms, seed + i, "Noise burst", 0.01, -3.0, 4000, 12)
        micro_last = g.copy()
        if params["bandlimit_on"]:
            g = fft_lowpass(g, ev["gen_sr"], ev["cutoff_gen"], roll=float(params["bandlimit_roll_hz"]))
        if params["nl_warp_on"]:                                                  # M:694-695
            g = spectrum_power_warp(g, float(params["nl_warp_power"]))
        if params["cep_warp_on"]:                                                 # M:696-697
            if jitter:
                g = g + jitter * np.max(np.abs(g)) * np.random.default_rng(555 + i).standard_normal(g.size)
            g = cepstrum_warp(g, float(params["cep_factor"]))
        if params["partial_lock_on"]:                                             # M:699-702
            g = partial_lock(g, ev["stretch"], int(params["pl_top_n"]), int(params["pl_neigh"]))
        else:
            g = spectrum_stretch(g, ev["stretch"])
        if params["res_bank_on"]:                                                 # M:704-710
            if jitter:
                g = g + jitter * np.max(np.abs(g)) * np.random.default_rng(333 + i).standard_normal(g.size)
            g = resonator_bank(g, ev["gen_sr"], int(params["res_modes"]), float(params["res_fmin"]), float(params["res_fmax"]),
                               float(params["res_decay_ms"]), seed + i)
        if params["wg_on"]:                                                       # M:712-717
            g = waveguide(g, ev["gen_sr"], int(params["wg_lines"]), float(params["wg_max_ms"]), float(params["wg_fb"]), seed + i)
        if params["unfold_mode"] != "Classic reinterpret":
            b1, b2, b3 = float(params["mb_b1"]), float(params["mb_b2"]), float(params["mb_b3"])
            g = multiband_unfold(g, ev["gen_sr"], [(0, b1), (b1, b2), (b2, b3)],
                                 [float(params["mb_u1"]), float(params["mb_u2"]), float(params["mb_u3"])],
                                 float(params["mb_roll"]))
        grain_last = g.copy()
        if params["event_feedback_on"] and previous is not None:                  # M:731-734
            fb = float(params["event_feedback_amt"])
            m = min(len(g), len(previous))
            g[:m] = (1.0 - fb) * g[:m] + fb * previous[:m]
        if imprint is not None:                                                   # M:736-738 (after grain_last)
            if jitter:
                g = g + jitter * np.max(np.abs(g)) * np.random.default_rng(777 + i).standard_normal(g.size)
            g = imprint.apply(g, float(params["spectral_imprint_amt"]), float(params["spectral_imprint_smooth"]))
        previous = g.copy()                                                       # M:740 (before the placement test)
        if ev["placed"]:
            a, o, ln = ev["start"], ev["offset"], ev["length"]
            mix[a:a + ln] += ev["amp"] * g[o:o + ln]
        if progress and ev["placed"] and i % 50 == 0:       # M:757 sits after the `continue` at M:744
            progress(int(5 + 70 * (i / max(1, n_evt))), f"Events {i}/{n_evt}  {note}".strip())      # M:758
    mix *= adsr_envelope(out_n, base_sr, float(params["env_a"]), float(params["env_d"]),
                         float(params["env_s"]), float(params["env_r"]), float(params["env_curve"]))
    if taps is not None:
        taps["after_adsr"] = mix.copy()
    if params["er_cloud_on"]:
        mix = reflection_cloud(mix, base_sr, int(params["er_taps"]), float(params["er_max_ms"]), seed)
    if taps is not None:
        taps["after_er"] = mix.copy()
    ir = params.get("_ir_audio")
    if params["space_ir_on"] and ir is not None:
        mix = short_ir_convolve(mix, ir[:int(params["space_ir_max_samps"])])
    if taps is not None:
        taps["after_ir"] = mix.copy()
    if params["stereo_on"]:
        st = stereo_diffuse(mix, base_sr, float(params["stereo_width"]))
    else:
        st = np.column_stack([mix, mix])
    st = peak_normalize(soft_saturate(st, float(params["sat_drive"])), float(params["peak"]))
    if progress:
        progress(100, "Done.")
    meta = dict(out_sr=base_sr, design_sr_base=plan["design_sr_base"], micro_last=micro_last, grain_last=grain_last)
    return st.astype(np.float64), meta
