"""Job tables: RenderPlans -> flat, relocatable numpy tables -> one merged batch.

`pack_chunk(plans)` lays a list of render plans out in chunk-local coordinates (every offset starts at
0) and returns plain numpy arrays, so chunks can be built by worker processes and pickled cheaply;
`merge_chunks` concatenates them, shifts every offset field by the chunk's base in its buffer, and pairs
the spectral jobs across the whole batch (two signals of equal length share one complex transform).
A single in-process chunk goes through exactly the same code.

Buffers (offsets in elements): pool (real; per-event signals), mono (real; per chunk:
[OLA plane | FIR plane | right channels / odd-length scratch]), out (float32 frames), hpool (real;
combined FIR taps), taps (reflection delays/gains), irs (deduplicated IR taps), dust (impulses).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field

import numpy as np

from . import _abi, plan as P

_SPEC_OP_BYTES = np.dtype(_abi.SpecOp).itemsize
_SPEC_JOB = np.dtype(_abi.SpecJob)
_ZERO_OP = np.zeros(_SPEC_OP_BYTES, np.uint8)


def _recs(ctype, n):
    return np.zeros(n, dtype=np.dtype(ctype))


_BESSEL = {}


def bessel_coeffs(theta, K=_abi.POST_K):
    """J_m(theta), m = -K..K by the ascending series (theta <= 0.9 here).  exp(i theta sin(phi)) =
    sum_m J_m(theta) exp(i m phi): for even n the rotation of spectral_diffusion_stereo (main_v2.py:432-435)
    is exactly a (2K+1)-tap circular FIR with taps at even lags."""
    hit = _BESSEL.get(theta)
    if hit is not None:
        return hit
    out = np.zeros(2 * K + 1)
    for m in range(K + 1):
        s, k = 0.0, 0
        while True:
            term = (-1.0) ** k * (theta / 2.0) ** (2 * k + m) / (math.factorial(k) * math.factorial(k + m))
            s += term
            k += 1
            if abs(term) < 1e-22 or k > 40:
                break
        out[K + m] = s
        out[K - m] = s * (-1.0) ** m
    if len(_BESSEL) > 256:
        _BESSEL.clear()
    _BESSEL[theta] = out
    return out


class _Items:
    """Spectral work items (n, in, out, operator bytes) of one stage."""

    def __init__(self):
        self.n, self.src, self.dst, self.ops = [], [], [], []

    def add(self, n, src, dst, op):
        self.n.append(n)
        self.src.append(src)
        self.dst.append(dst)
        self.ops.append(bytes(op))

    def arrays(self):
        k = len(self.n)
        ops = np.frombuffer(b"".join(self.ops), dtype=np.uint8).reshape(k, _SPEC_OP_BYTES) if k else np.zeros((0, _SPEC_OP_BYTES), np.uint8)
        return (np.asarray(self.n, np.int64), np.asarray(self.src, np.int64), np.asarray(self.dst, np.int64), ops)


@dataclass
class Tables:
    sy1: np.ndarray = None
    sy2: np.ndarray = None
    ola_r: np.ndarray = None
    env_reps: np.ndarray = None     # one OlaRender per shared envelope table (its .env = table offset)
    env_n: int = 0
    ola_e: np.ndarray = None
    fir: np.ndarray = None
    post: np.ndarray = None
    tap_off: np.ndarray = None
    tap_gain: np.ndarray = None
    irs: np.ndarray = None
    dust_pos: np.ndarray = None
    dust_val: np.ndarray = None
    atoms: np.ndarray = None        # wavelet atoms, float64 [N, 4]
    atom_shift: np.ndarray = None   # int32 [N]
    seq: np.ndarray = None          # event-sequential tail (renders with event feedback), rows float64:
                                    # render, rank, cur, prev (-1: none), n_prev, fb_dst (-1), imp_dst (-1), n, fb, amount, smooth
    wg: tuple = None                # (rows int64 [N, 5] = src, dst, n, line_begin, line_count ; lines float64 [L, 3])
    res: tuple = None               # (rows int64 [N, 5] = src, dst, n, mode_begin, mode_count ; decay [N] ; modes float64 [M, 3])
    post_grain: tuple = None        # spectral items applied after the resonator bank (multiband): (n, src, dst, ops)
    cep: tuple = None               # (rows int64 [N, 3] = src, dst, n ; factor [N] ; pre ops [N, B] ; post ops [N, B])
    plock: tuple = None             # (rows int64 [N, 5] = src, dst, n, top_n, neigh ; factor [N] ; pre ops [N, B] ; post ops [N, B])
    imprint: np.ndarray = None      # rows per imprinted event, in event order per render: render, pool_in, pool_out, n
    imprint_par: np.ndarray = None  # per render: amount, smooth (nan when off)
    tilt: tuple = None          # (n, src, dst, ops)
    grain: tuple = None
    rot: tuple = None
    odd: np.ndarray = None      # rows: y_at, scratch, n, dr   (odd-length stereo)
    pool_n: int = 0
    mono_n: int = 0
    frames: int = 0
    h_total: int = 0
    max_h: int = 0
    max_out_n: int = 0
    out_at: np.ndarray = None   # per render: first frame
    out_n: np.ndarray = None
    y_at: np.ndarray = None
    last: np.ndarray = None     # per render: micro_at, grain_at, n of the last event (-1 if none)
    srs: np.ndarray = None      # per render: base_sr, design_sr_base
    alg: dict = field(default_factory=dict)   # algorithmic element counts per stage


def pack_chunk(plans) -> Tables:
    R = len(plans)
    n_evt = sum(len(rp.events) for rp in plans)
    sy1, sy2 = _recs(_abi.SynthEvt, n_evt), _recs(_abi.SynthEvt, n_evt)
    ola_r, post = _recs(_abi.OlaRender, R), _recs(_abi.PostRender, R)
    ola_e = []
    fir = []
    tilt, grain, rot = _Items(), _Items(), _Items()
    dust_pos, dust_val, n_dust = [], [], 0
    atoms, atom_shift, n_atoms = [], [], 0
    imprint_rows, imprint_par = [], np.full((R, 2), np.nan)
    pl_rows, pl_factor, pl_pre, pl_post = [], [], [], []
    cp_rows, cp_factor, cp_pre, cp_post = [], [], [], []
    rs_rows, rs_decay, rs_modes, n_modes = [], [], [], 0
    wg_rows, wg_lines, n_lines = [], [], 0
    seq_rows = []
    post_grain = _Items()
    tap_off, tap_gain, n_taps = [], [], 0
    irs, ir_index, n_ir = [], {}, 0
    pool_n = mono_n = h_total = max_h = 0
    mono_at, last = [], np.full((R, 3), -1, np.int64)
    fir_of = {}
    alg = dict(synth=0, tilt_spectral=0, grain_spectral=0, overlap_add=0, fir_in=0, fir_taps=0, post=0)
    env_seen = {}
    e = 0
    for r, rp in enumerate(plans):
        a, d, rel, S, curve = rp.adsr
        n = rp.out_n
        if a > n:
            raise ValueError(f"could not broadcast input array from shape ({a},) into shape ({n},)")   # main_v2.py:182
        d_end = min(n, a + d) if d > 0 else a
        sus_end = max(d_end, n - rel)
        ev_begin = len(ola_e)
        max_len = 0
        x_begin, x_end = n, 0
        prev_final, prev_n, rank = -1, 0, 0
        for ev in rp.events:
            st = np.random.PCG64(ev.seed).state["state"]
            s_hi, s_lo = st["state"] >> 64, st["state"] & 0xFFFFFFFFFFFFFFFF
            i_hi, i_lo = st["inc"] >> 64, st["inc"] & 0xFFFFFFFFFFFFFFFF
            micro = pool_n
            pool_n += ev.n
            out1, mode2, aux2, dust_b, dust_c = micro, -1, 0, 0, 0
            atom_b, atom_c = 0, 0
            alg["synth"] += ev.n
            if ev.mode == P.MODE_WAVELET:
                atom_b, atom_c = n_atoms, len(ev.atom_shift)
                atoms.append(ev.atoms)
                atom_shift.append(ev.atom_shift)
                n_atoms += atom_c
            if ev.mode == P.MODE_DUST:
                dust_b, dust_c = n_dust, len(ev.dust_pos)
                dust_pos.append(ev.dust_pos)
                dust_val.append(ev.dust_val)
                n_dust += dust_c
            elif ev.mode in (P.MODE_IRFRAG, P.MODE_SCANLINE):        # table rides in the impulse arrays (positions unused)
                dust_b, dust_c = n_dust, len(ev.table)
                dust_pos.append(np.zeros(dust_c, np.int32))
                dust_val.append(ev.table)
                n_dust += dust_c
                if ev.mode == P.MODE_SCANLINE:
                    aux2 = pool_n                                     # windowed, unsmoothed line
                    pool_n += ev.n
            elif ev.mode in (P.MODE_CHAOS, P.MODE_STICK):
                aux2 = pool_n                                         # gated logistic-map samples / the event's normals
                pool_n += ev.n
            elif ev.mode in (P.MODE_NOISE, P.MODE_SKEW):
                raw, tilted = pool_n, pool_n + ev.n
                pool_n += 2 * ev.n
                out1, mode2, aux2 = raw, ev.mode, tilted
                tilt.add(ev.n, raw, tilted, ev.tilt)
                alg["tilt_spectral"] += 2 * ev.n
            common = (s_hi, s_lo, i_hi, i_lo, ev.n)
            tail = (ev.fade, ev.sigma)
            inv_fade = 1.0 / ev.fade if ev.fade > 0 else (ev.stick_noise if ev.mode == P.MODE_STICK else 0.0)
            sy1[e] = common + (ev.mode,) + tail + (out1, ev.f_over_sr, inv_fade, ev.ring_decay, ev.env_decay,
                                                     dust_b, dust_c, ev.ker_len, aux2 if ev.mode in (P.MODE_SCANLINE, P.MODE_CHAOS, P.MODE_STICK) else 0,
                                                     atom_b, atom_c, 0)
            sy2[e] = common + (mode2,) + tail + (micro, ev.f_over_sr, inv_fade, ev.ring_decay, ev.env_decay,
                                                  0, 0, ev.ker_len, aux2, 0, 0, 0)
            g_at = micro
            if ev.cep is not None:
                g_at = pool_n
                pool_n += ev.n
                cp_rows.append((micro, g_at, ev.n))
                cp_factor.append(ev.cep[0])
                cp_pre.append(bytes(ev.cep[1]))
                cp_post.append(bytes(ev.cep[2]))
                alg["grain_spectral"] += 2 * ev.n * (1 + int(ev.cep[1].lp_on) + (1 if ev.cep[1].warp_exp else 0)
                                                     + int(ev.cep[2].stretch_on) + (1 if ev.cep[2].n_bands else 0))
            if ev.plock is not None:
                pl_src = g_at                      # the raw transient, or the cepstrally warped grain
                g_at = pool_n
                pool_n += ev.n
                pl_rows.append((pl_src, g_at, ev.n, ev.plock[1], ev.plock[2]))
                pl_factor.append(ev.plock[0])
                pl_pre.append(bytes(ev.plock[3]))
                pl_post.append(bytes(ev.spec))
                alg["grain_spectral"] += 2 * ev.n * (1 + int(ev.plock[3].lp_on) + (1 if ev.plock[3].warp_exp else 0) + (1 if ev.spec.n_bands else 0))
            elif ev.cep is None and ev.spec is not None:
                g_at = pool_n
                pool_n += ev.n
                grain.add(ev.n, micro, g_at, ev.spec)
                alg["grain_spectral"] += 2 * ev.n * (int(ev.spec.lp_on) + int(ev.spec.stretch_on) + (1 if ev.spec.n_bands else 0)
                                                     + (1 if ev.spec.warp_exp else 0))
            if ev.res is not None:
                r_at = pool_n                      # the resonator reads the grain so far and writes a new one
                pool_n += ev.n
                rs_rows.append((g_at, r_at, ev.n, n_modes, len(ev.res[0])))
                rs_decay.append(ev.res[1])
                rs_modes.append(ev.res[0])
                n_modes += len(ev.res[0])
                g_at = r_at
            if ev.wg is not None:
                w_at = pool_n                      # the combs run in place on a copy of the grain so far
                pool_n += ev.n
                wg_rows.append((g_at, w_at, ev.n, n_lines, len(ev.wg)))
                wg_lines.append(ev.wg)
                n_lines += len(ev.wg)
                g_at = w_at
            if ev.res is not None or ev.wg is not None:
                if ev.spec_b is not None:
                    b_at = pool_n
                    pool_n += ev.n
                    post_grain.add(ev.n, g_at, b_at, ev.spec_b)
                    alg["grain_spectral"] += 2 * ev.n
                    g_at = b_at
            last[r] = (micro, g_at, ev.n)
            if rp.feedback is not None:
                # M:731-740: feedback from the previous event's FINAL grain, then the imprint, rank by rank
                cur, fb_dst, imp_dst = g_at, -1, -1
                if prev_final >= 0:
                    fb_dst = pool_n
                    pool_n += ev.n
                    g_at = fb_dst
                imp_on = rp.imprint is not None and ev.n >= 64 and rp.imprint[0] > 0
                if imp_on:
                    imp_dst = pool_n
                    pool_n += ev.n
                    g_at = imp_dst
                seq_rows.append((r, rank, cur, prev_final, prev_n, fb_dst, imp_dst, ev.n, rp.feedback,
                                 rp.imprint[0] if imp_on else 0.0, rp.imprint[1] if imp_on else 0.0))
                prev_final, prev_n, rank = g_at, ev.n, rank + 1
            elif rp.imprint is not None and ev.n >= 64 and rp.imprint[0] > 0:         # M:570: short grains / amount <= 0 pass through
                src = g_at
                g_at = pool_n                      # imprinted grain (grain_last keeps the grain before it, M:729)
                pool_n += ev.n
                imprint_rows.append((r, src, g_at, ev.n))
                imprint_par[r] = rp.imprint
            if ev.placed:
                ola_e.append((g_at + ev.offset, ev.start, ev.length, ev.amp))
                max_len = max(max_len, ev.length)
                # (no "grain[0] == 0" shortcut here: that holds for a raw gen_basic transient only, not after the
                #  band-limit / stretch / multiband operators or for the generators without a fade-in)
                x_begin = min(x_begin, ev.start)
                x_end = max(x_end, ev.start + ev.length)
                alg["overlap_add"] += ev.length
            e += 1
        env_key = (n, a, d_end, sus_end, 1 if (rel > 0 and n > sus_end) else 0,
                   1.0 / a if a > 0 else 0.0, 1.0 / (d_end - a) if d_end > a else 0.0,
                   1.0 / (n - sus_end - 1) if n - sus_end > 1 else 0.0, S, curve)
        env_seen.setdefault(env_key, []).append(r)
        ola_r[r] = (mono_n, n, ev_begin, len(ola_e), max_len) + env_key[1:] + (-1,)
        alg["overlap_add"] += n
        # FIR: reflection cloud folded into the impulse response
        has_er = rp.er_offs is not None and rp.er_offs.size > 0
        if has_er or rp.ir is not None:
            if rp.ir is not None:
                key = rp.ir.ctypes.data if rp.ir.flags.owndata else rp.ir.tobytes()
                hit = ir_index.get(key)
                if hit is None or (not isinstance(key, bytes) and hit[2] is not rp.ir):
                    hit = ir_index[key] = (n_ir, rp.ir.size, rp.ir)
                    irs.append(rp.ir)
                    n_ir += rp.ir.size
                alg["fir_in"] += 2 * n
                alg["fir_taps"] += rp.ir.size
            else:
                hit = ir_index.get(b"delta")
                if hit is None:
                    hit = ir_index[b"delta"] = (n_ir, 1, None)
                    irs.append(np.ones(1))
                    n_ir += 1
            ir_at, ir_len = hit[0], hit[1]
            t0 = n_taps
            h_len = ir_len
            if has_er:
                if rp.er_offs.size > 4096:
                    raise ValueError("er_taps > 4096 is outside the accelerated path")
                tap_off.append(rp.er_offs)
                tap_gain.append(rp.er_gains)
                n_taps += rp.er_offs.size
                h_len = ir_len + int(rp.er_offs.max())
                alg["fir_in"] += 2 * n
            if x_begin == 0 and a > 0:
                x_begin = 1                      # env[0] = 0 ** curve = 0 (main_v2.py:181-182)
            fir_of[r] = len(fir)
            fir.append((ir_at, ir_len, h_len, h_total, t0, n_taps, mono_n, 0, n, min(x_begin, x_end), x_end, 0))
            h_total += h_len
            max_h = max(max_h, h_len)
        mono_at.append(mono_n)
        mono_n += n
    plane = mono_n
    extra = 0
    frames = 0
    odd = []
    out_at, out_n, y_at = np.zeros(R, np.int64), np.zeros(R, np.int64), np.zeros(R, np.int64)
    fir_arr = _recs(_abi.FirRender, len(fir))
    for i, f in enumerate(fir):
        fir_arr[i] = f
    for r, rp in enumerate(plans):
        n = rp.out_n
        y = mono_at[r]
        if r in fir_of:
            y = plane + mono_at[r]
            fir_arr[fir_of[r]]["y"] = y
        mode, dl, dr, rbuf = 0, 0, 0, 0
        coef = np.zeros(2 * _abi.POST_K + 1)
        if rp.stereo_on:
            dl, dr = rp.stereo_dl, rp.stereo_dr
            if n % 2 == 0:
                mode, coef, rbuf = 1, bessel_coeffs(rp.stereo_theta), 2 * plane + extra     # right channel, written by the max pass
                extra += n
            else:
                mode, rbuf = 2, 2 * plane + extra + n                                        # [rolled copy | right channel]
                odd.append((y, 2 * plane + extra, n, dr))
                op = _abi.SpecOp()
                op.kind, op.alpha = _abi.OP_ROT, rp.stereo_theta
                rot.add(n, 2 * plane + extra, 2 * plane + extra + n, op)
                extra += 2 * n
        post[r] = (y, frames, rbuf, n, mode, dl, dr, rp.drive, 1.0 / math.tanh(rp.drive) if rp.drive > 0 else 1.0, rp.peak, coef)
        out_at[r], out_n[r], y_at[r] = frames, n, y
        frames += n
        alg["post"] += n
    # envelopes shared by several renders are tabulated once (ms_adsr_tables) instead of one pow() per sample
    env_n = 0
    reps = []
    for key, members in env_seen.items():
        if len(members) >= 2:
            ola_r["env"][members] = env_n
            reps.append(ola_r[members[0]].copy())
            env_n += int(key[0])
    env_reps = np.array(reps, dtype=ola_r.dtype) if reps else np.zeros(0, dtype=ola_r.dtype)
    t = Tables()
    t.env_reps, t.env_n = env_reps, env_n
    t.sy1, t.sy2, t.ola_r, t.post, t.fir = sy1, sy2, ola_r, post, fir_arr
    t.ola_e = _recs(_abi.OlaEvt, len(ola_e))
    for i, rec in enumerate(ola_e):
        t.ola_e[i] = rec
    t.tap_off = np.concatenate(tap_off).astype(np.int32) if tap_off else np.zeros(0, np.int32)
    t.tap_gain = np.concatenate(tap_gain).astype(np.float64) if tap_gain else np.zeros(0)
    t.irs = np.concatenate(irs).astype(np.float64) if irs else np.zeros(0)
    t.dust_pos = np.concatenate(dust_pos).astype(np.int32) if dust_pos else np.zeros(0, np.int32)
    t.dust_val = np.concatenate(dust_val).astype(np.float64) if dust_val else np.zeros(0)
    t.atoms = np.concatenate(atoms).astype(np.float64) if atoms else np.zeros((0, 4))
    t.atom_shift = np.concatenate(atom_shift).astype(np.int32) if atom_shift else np.zeros(0, np.int32)
    npl = len(pl_rows)
    t.plock = (np.asarray(pl_rows, np.int64).reshape(-1, 5), np.asarray(pl_factor, np.float64),
               np.frombuffer(b"".join(pl_pre), np.uint8).reshape(npl, _SPEC_OP_BYTES) if npl else np.zeros((0, _SPEC_OP_BYTES), np.uint8),
               np.frombuffer(b"".join(pl_post), np.uint8).reshape(npl, _SPEC_OP_BYTES) if npl else np.zeros((0, _SPEC_OP_BYTES), np.uint8))
    t.res = (np.asarray(rs_rows, np.int64).reshape(-1, 5), np.asarray(rs_decay, np.float64),
             np.concatenate(rs_modes) if rs_modes else np.zeros((0, 3)))
    t.seq = np.asarray(seq_rows, np.float64).reshape(-1, 11)
    t.wg = (np.asarray(wg_rows, np.int64).reshape(-1, 5), np.concatenate(wg_lines) if wg_lines else np.zeros((0, 3)))
    t.post_grain = post_grain.arrays()
    ncp = len(cp_rows)
    t.cep = (np.asarray(cp_rows, np.int64).reshape(-1, 3), np.asarray(cp_factor, np.float64),
             np.frombuffer(b"".join(cp_pre), np.uint8).reshape(ncp, _SPEC_OP_BYTES) if ncp else np.zeros((0, _SPEC_OP_BYTES), np.uint8),
             np.frombuffer(b"".join(cp_post), np.uint8).reshape(ncp, _SPEC_OP_BYTES) if ncp else np.zeros((0, _SPEC_OP_BYTES), np.uint8))
    t.imprint = np.asarray(imprint_rows, np.int64).reshape(-1, 4)
    t.imprint_par = imprint_par
    t.tilt, t.grain, t.rot = tilt.arrays(), grain.arrays(), rot.arrays()
    t.odd = np.asarray(odd, np.int64).reshape(-1, 4)
    t.pool_n, t.mono_n, t.frames, t.h_total, t.max_h = pool_n, 2 * plane + extra, frames, h_total, max_h
    t.max_out_n = max((rp.out_n for rp in plans), default=0)
    t.out_at, t.out_n, t.y_at, t.last = out_at, out_n, y_at, last
    t.srs = np.asarray([(rp.base_sr, rp.design_sr_base) for rp in plans], np.int64).reshape(-1, 2)
    t.alg = alg
    return t


def _shift(arr, fields, base):
    if arr.size and base:
        for f in fields:
            arr[f] += base


def merge_chunks(chunks) -> Tables:
    if len(chunks) == 1:
        return chunks[0]
    m = Tables()
    pool_b = mono_b = frame_b = h_b = tap_b = ir_b = dust_b = olae_b = env_b = atom_b = render_b = 0
    parts = {k: [] for k in ("sy1", "sy2", "ola_r", "env_reps", "ola_e", "fir", "post", "tap_off", "tap_gain", "irs", "dust_pos", "dust_val",
                             "atoms", "atom_shift", "imprint", "imprint_par", "seq", "odd", "out_at", "out_n", "y_at", "last", "srs")}
    items = {k: [[], [], [], []] for k in ("tilt", "grain", "rot", "post_grain")}
    rs_parts, mode_b = [[], [], []], 0
    wg_parts, line_b = [[], []], 0
    pl_parts = [[], [], [], []]
    cp_parts = [[], [], [], []]
    alg = {}
    for c in chunks:
        _shift(c.sy1, ("out",), pool_b)
        if c.sy1.size and pool_b:
            c.sy1["aux"] += np.where((c.sy1["mode"] == P.MODE_SCANLINE) | (c.sy1["mode"] == P.MODE_CHAOS) | (c.sy1["mode"] == P.MODE_STICK), pool_b, 0)
        _shift(c.sy1, ("dust_begin",), dust_b)
        _shift(c.sy1, ("atom_begin",), atom_b)
        if c.imprint.size:
            c.imprint[:, 0] += render_b
            c.imprint[:, 1:3] += pool_b
        if c.seq.size:
            c.seq[:, 0] += render_b
            for col in (2, 3, 5, 6):
                c.seq[:, col] += np.where(c.seq[:, col] >= 0, pool_b, 0)
        if c.plock[0].size:
            c.plock[0][:, 0:2] += pool_b
        if c.cep[0].size:
            c.cep[0][:, 0:2] += pool_b
        for i in range(4):
            pl_parts[i].append(c.plock[i])
            cp_parts[i].append(c.cep[i])
        _shift(c.sy2, ("out", "aux"), pool_b)
        _shift(c.ola_r, ("out",), mono_b)
        _shift(c.ola_r, ("ev_begin", "ev_end"), olae_b)
        if env_b:
            c.ola_r["env"] += np.where(c.ola_r["env"] >= 0, env_b, 0)
            c.env_reps["env"] += env_b
        _shift(c.ola_e, ("grain",), pool_b)
        _shift(c.fir, ("ir",), ir_b)
        _shift(c.fir, ("h",), h_b)
        _shift(c.fir, ("tap_begin", "tap_end"), tap_b)
        _shift(c.fir, ("x", "y"), mono_b)
        _shift(c.post, ("y", "rbuf"), mono_b)
        _shift(c.post, ("out",), frame_b)
        if c.odd.size:
            c.odd[:, 0:2] += mono_b
        c.out_at += frame_b
        c.y_at += mono_b
        if c.last.size:
            c.last[:, 0:2] += np.where(c.last[:, 0:2] >= 0, pool_b, 0)
        for k in parts:
            parts[k].append(getattr(c, k))
        if c.res[0].size:
            c.res[0][:, 0:2] += pool_b
            c.res[0][:, 3] += mode_b
        for i in range(3):
            rs_parts[i].append(c.res[i])
        mode_b += len(c.res[2])
        if c.wg[0].size:
            c.wg[0][:, 0:2] += pool_b
            c.wg[0][:, 3] += line_b
        wg_parts[0].append(c.wg[0])
        wg_parts[1].append(c.wg[1])
        line_b += len(c.wg[1])
        for k, base in (("tilt", pool_b), ("grain", pool_b), ("rot", mono_b), ("post_grain", pool_b)):
            n, s, d, ops = getattr(c, k)
            items[k][0].append(n)
            items[k][1].append(s + base)
            items[k][2].append(d + base)
            items[k][3].append(ops)
        for k, v in c.alg.items():
            alg[k] = alg.get(k, 0) + v
        pool_b += c.pool_n
        mono_b += c.mono_n
        frame_b += c.frames
        h_b += c.h_total
        tap_b += c.tap_off.size
        ir_b += c.irs.size
        dust_b += c.dust_pos.size
        olae_b += c.ola_e.size
        env_b += c.env_n
        atom_b += c.atom_shift.size
        render_b += len(c.post)
        m.max_h = max(m.max_h, c.max_h)
        m.max_out_n = max(m.max_out_n, c.max_out_n)
    for k in parts:
        setattr(m, k, np.concatenate(parts[k]))
    for k in items:
        setattr(m, k, tuple(np.concatenate(x) for x in items[k]))
    m.plock = tuple(np.concatenate(x) for x in pl_parts)
    m.res = tuple(np.concatenate(x) for x in rs_parts)
    m.wg = tuple(np.concatenate(x) for x in wg_parts)
    m.cep = tuple(np.concatenate(x) for x in cp_parts)
    m.pool_n, m.mono_n, m.frames, m.h_total, m.alg, m.env_n = pool_b, mono_b, frame_b, h_b, alg, env_b
    return m


def pair_jobs(items):
    """(n, src, dst, ops) -> SpecJob record array.  Two signals of equal n share one complex transform;
    the leftover of an odd group rides alone."""
    n, src, dst, ops = items
    k = n.size
    if k == 0:
        return np.zeros(0, _SPEC_JOB)
    order = np.argsort(n, kind="stable")
    n, src, dst, ops = n[order], src[order], dst[order], ops[order]
    starts = np.flatnonzero(np.r_[True, n[1:] != n[:-1]])
    pos = np.arange(k) - np.repeat(starts, np.diff(np.r_[starts, k]))      # index inside its equal-n group
    first = pos % 2 == 0
    has_partner = np.zeros(k, bool)
    has_partner[:-1] = first[:-1] & (n[1:] == n[:-1]) & (pos[1:] == pos[:-1] + 1)
    a = np.flatnonzero(first)
    b = a + 1
    paired = has_partner[a]
    jobs = np.zeros(a.size, _SPEC_JOB)
    raw = jobs.view(np.uint8).reshape(a.size, _SPEC_JOB.itemsize)
    jobs["n"] = n[a]
    jobs["in_a"], jobs["out_a"] = src[a], dst[a]
    bb = np.where(paired, np.minimum(b, k - 1), 0)
    jobs["in_b"] = np.where(paired, src[bb], -1)
    jobs["out_b"] = np.where(paired, dst[bb], -1)
    op0 = _SPEC_JOB.fields["op"][1]
    raw[:, op0:op0 + _SPEC_OP_BYTES] = ops[a]
    raw[:, op0 + _SPEC_OP_BYTES:op0 + 2 * _SPEC_OP_BYTES] = np.where(paired[:, None], ops[bb], _ZERO_OP[None, :])
    return jobs


# --------------------------------------------------------------------------- worker pool
class _Pending:
    """Results of _WorkerPool.submit, retrievable in any order while the workers are still busy."""

    def __init__(self, n):
        import threading
        self.results, self.ready, self.error, self.cv, self.threads = [None] * n, [False] * n, None, threading.Condition(), []

    def put(self, i, value):
        with self.cv:
            self.results[i], self.ready[i] = value, True
            self.cv.notify_all()

    def fail(self, err):
        with self.cv:
            if self.error is None:
                self.error = err
            self.cv.notify_all()

    def get(self, i):
        with self.cv:
            while not self.ready[i] and self.error is None:
                self.cv.wait()
            if not self.ready[i]:
                raise self.error
            out, self.results[i] = self.results[i], None
            return out


class _WorkerPool:
    """Persistent `python -m audio_suite_b200.plan_worker` subprocesses, one feeder thread each."""

    def __init__(self, size):
        import os
        import subprocess
        import sys
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        env = dict(os.environ)
        env["PYTHONPATH"] = root + os.pathsep + env.get("PYTHONPATH", "")
        env["OMP_NUM_THREADS"] = env["OPENBLAS_NUM_THREADS"] = "1"
        self.size = size
        self.procs = [subprocess.Popen([sys.executable, "-m", "audio_suite_b200.plan_worker"], stdin=subprocess.PIPE,
                                       stdout=subprocess.PIPE, env=env) for _ in range(size)]
        # one worker per core, leaving the first cores of the affinity mask to the parent (kernel launches, table
        # uploads, the copy stream): measured on the 16-core B200 hosts, workers floating over all cores stall
        # the parent's CUDA calls for tens of ms at a time
        try:
            cores = sorted(os.sched_getaffinity(0))
            spare = int(os.environ.get("MS_PLAN_SPARE_CORES", "2")) if len(cores) >= 8 else 0
            usable = cores[spare:] or cores
            # (pinning only helps when this process owns the box; ranks of a torchrun job share it unpinned)
            if os.environ.get("MS_PLAN_PIN", "0") == "1":
                for i, p in enumerate(self.procs):
                    os.sched_setaffinity(p.pid, {usable[i % len(usable)]})
        except Exception:
            pass
        try:                                    # 1 MiB pipes: a request or a packed piece fits without blocking the writer
            import fcntl
            for p in self.procs:
                for f in (p.stdin, p.stdout):
                    fcntl.fcntl(f.fileno(), 1031, 1 << 20)          # F_SETPIPE_SZ
        except Exception:
            pass

    def submit(self, chunks, prep=None):
        """Start planning `chunks` (lists of parameter dicts); returns a _Pending whose get(i) blocks until
        chunk i is packed.  Chunks are handed out in order, so early chunks finish first.  `prep` (optional)
        is applied to a chunk by the feeder thread right before it is sent."""
        import pickle
        import struct
        import threading
        pend = _Pending(len(chunks))
        todo = list(enumerate(chunks))
        lock = threading.Lock()

        def serve(proc):
            # two requests in flight per worker: the worker finds its next piece already in the pipe when it
            # finishes one, instead of waiting for this thread to be scheduled on a box whose cores are all busy
            from collections import deque
            inflight = deque()

            def send_next():
                with lock:
                    if not todo or pend.error is not None:
                        return
                    i, chunk = todo.pop(0)
                if prep is not None:
                    chunk = prep(chunk)
                blob = pickle.dumps(chunk, protocol=pickle.HIGHEST_PROTOCOL)
                proc.stdin.write(struct.pack("<Q", len(blob)))
                proc.stdin.write(blob)
                proc.stdin.flush()
                inflight.append(i)
            try:
                send_next()
                send_next()
                while inflight:
                    i = inflight.popleft()
                    head = proc.stdout.read(8)
                    if len(head) < 8:
                        raise RuntimeError("planning worker died")
                    status, payload = pickle.loads(proc.stdout.read(struct.unpack("<Q", head)[0]))
                    if status != "ok":
                        raise payload
                    pend.put(i, payload)
                    send_next()
            except BaseException as e:
                pend.fail(e)
        pend.threads = [threading.Thread(target=serve, args=(p,), daemon=True) for p in self.procs]
        for t in pend.threads:
            t.start()
        return pend

    def map(self, chunks):
        pend = self.submit(chunks)
        try:
            return [pend.get(i) for i in range(len(chunks))]
        except BaseException:
            self.close()
            raise

    def close(self):
        for p in self.procs:
            try:
                p.stdin.close()
                p.terminate()
            except Exception:
                pass
        self.procs = []


_POOL = None


def shutdown_pool():
    global _POOL
    if _POOL is not None:
        _POOL.close()
        _POOL = None


def _pool(workers):
    import atexit
    global _POOL
    if _POOL is None or _POOL.size != workers or not _POOL.procs:
        shutdown_pool()
        _POOL = _WorkerPool(workers)
        atexit.register(shutdown_pool)
    return _POOL


def default_workers():
    """Planning worker processes per rank: the cores this process may run on, shared between the ranks of the box
    (torchrun sets LOCAL_WORLD_SIZE), minus two per rank for the parent (kernel launches, uploads, copy stream)."""
    import os
    forced = int(os.environ.get("MS_PLAN_WORKERS", "0"))
    if forced:
        return forced
    cores = len(os.sched_getaffinity(0))
    ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
    share = cores // ranks
    return min(32, max(1, share - 2 if share >= 8 else share - 1 if share >= 3 else share))


def plan_stream(params_list, chunk, workers=None, piece=32):
    """Generator of packed Tables for consecutive slices of `chunk` renders.  The whole list is handed to the
    worker pool at once (pieces of `piece` renders, in order); slice k is yielded as soon as its pieces are
    packed, while the workers carry on with the later ones -- the caller renders slice k meanwhile."""
    n = len(params_list)
    workers = default_workers() if workers is None else workers
    # slice boundaries; `chunk` may be a list of sizes (the last one repeats): a short first slice starts the GPU and
    # the device->host drain early
    sizes = list(chunk) if isinstance(chunk, (list, tuple)) else [int(chunk)]
    cuts, a = [], 0
    while a < n:
        c = max(1, int(sizes[min(len(cuts), len(sizes) - 1)]))
        cuts.append((a, min(n, a + c)))
        a += c
    # slices whose renders all belong to the native planner's family are planned in-process by libms_hostplan.so
    # (hostplan.plan_slice: no pickling, ~10x the Python planner's speed); the others go to the worker processes
    from . import hostplan
    have_native = hostplan.lib() is not None and not os.environ.get("MS_PLAN_PYTHON")

    def is_native(a, b):
        return have_native and all(hostplan.supported(p) for p in params_list[a:b])
    if have_native and len(cuts) > 1 and workers > 1 and is_native(*cuts[0]):
        # Optimistic native path: the first slice is planned before anything else is looked at; then two Python threads
        # alternate over the remaining slices (one converts the parameters of slice k+1 while the native call of slice k --
        # which releases the GIL and plans blocks of renders on its own threads -- runs).  Slices come out in order.  The
        # conversion itself notices a render outside the native family (hostplan.Unsupported): from that slice on the
        # batch goes through the general path below.
        from concurrent.futures import ThreadPoolExecutor
        nthr = max(1, min(hostplan.default_threads(), workers))
        yield hostplan.plan_chunk(params_list[cuts[0][0]:cuts[0][1]], nthr)
        rest = None
        with ThreadPoolExecutor(max_workers=2, thread_name_prefix="ms-hostplan") as ex:
            futs = [ex.submit(hostplan.plan_chunk, params_list[a:b], nthr) for a, b in cuts[1:]]
            for k, f in enumerate(futs):
                try:
                    tb = f.result()
                except hostplan.Unsupported:
                    for g in futs[k + 1:]:
                        g.cancel()
                    rest = k + 1
                    break
                yield tb
        if rest is not None:
            yield from plan_stream(params_list[cuts[rest][0]:], sizes[rest:] if len(sizes) > rest else sizes[-1:], workers=workers, piece=piece)
        return
    native = [is_native(a, b) for a, b in cuts]
    if workers <= 1 or all(native):
        for (a, b), nat in zip(cuts, native):
            yield hostplan.plan_slice(params_list[a:b]) if nat else pack_chunk([P.plan_render(p) for p in params_list[a:b]])
        return
    bounds = []
    for (a, b), nat in zip(cuts, native):
        bounds.append(None if nat else [(i, min(b, i + piece)) for i in range(a, b, piece)])
    flat = [ab for bs in bounds if bs is not None for ab in bs]
    # Static round-robin assignment: piece i belongs to worker i % W and is answered in order; the parent reads the
    # answers in piece order straight from the pipes (1 MiB pipes let a worker run a few pieces ahead).  Impulse
    # responses are digested before pickling (P._slim_params).
    import pickle
    import struct
    pool = _pool(workers)
    W = len(pool.procs)

    def send(w, obj):
        blob = pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
        f = pool.procs[w].stdin
        f.write(struct.pack("<Q", len(blob)))
        f.write(blob)
        f.flush()

    def slim(i):
        a, b = flat[i]
        return [P._slim_params(p) for p in params_list[a:b]]

    def recv(w):
        f = pool.procs[w].stdout
        head = f.read(8)
        if len(head) < 8:
            raise RuntimeError("planning worker died")
        status, payload = pickle.loads(f.read(struct.unpack("<Q", head)[0]))
        if status != "ok":
            raise payload
        return payload
    # Requests go out piece by piece, round-robin (piece i -> worker i % W), from a FEEDER THREAD: the thread that
    # reads the answers never blocks on a write.  (Writing from the reading thread can deadlock: a request larger than
    # the free space of a worker's stdin pipe blocks the parent while that worker is itself blocked writing an answer
    # nobody reads -- reproduced with 4 MB `_img_gray` arrays per piece.)  Every piece below the one being read has been
    # consumed, so the worker that owns it can always finish it: the pipeline cannot stall for good.
    import threading
    feed_err = []

    def feed():
        try:
            for i in range(len(flat)):
                send(i % W, slim(i))
        except BaseException as e:          # a closed pool (abandoned generator) ends the feeder quietly;
            feed_err.append(e)              # anything else (e.g. an unpicklable parameter) must not leave the reader waiting
            if not isinstance(e, (BrokenPipeError, ValueError, OSError)):
                shutdown_pool()
    feeder = threading.Thread(target=feed, daemon=True)
    feeder.start()
    try:
        k = 0
        for (a, b), bs in zip(cuts, bounds):
            if bs is None:
                yield hostplan.plan_slice(params_list[a:b])
                continue
            parts = [recv((k + t) % W) for t in range(len(bs))]
            k += len(bs)
            yield merge_chunks(parts)
        feeder.join()
    except BaseException as e:
        shutdown_pool()          # pipes may hold unread answers: start from fresh workers next time
        if feed_err and not isinstance(feed_err[0], (BrokenPipeError, ValueError, OSError)) and isinstance(e, RuntimeError):
            raise feed_err[0]
        raise


def plan_and_pack(params_list, workers=None):
    """Returns (Tables, plans or None).  Small batches are planned in-process (plans kept for progress /
    inspection); large ones by a persistent pool of worker processes that return packed chunks."""
    import os
    from . import hostplan
    n = len(params_list)
    if hostplan.lib() is not None and not os.environ.get("MS_PLAN_PYTHON") and all(hostplan.supported(p) for p in params_list):
        return hostplan.plan_slice(params_list), None
    if workers is None:
        workers = default_workers()
    min_batch = int(os.environ.get("MS_PLAN_MIN_BATCH", "256"))
    if n < min_batch or workers <= 1:
        plans = [P.plan_render(p) for p in params_list]
        return pack_chunk(plans), plans
    slim = [P._slim_params(p) for p in params_list]      # do not pickle multi-MB impulse responses per render
    per = max(min(32, max(1, n // workers)), (n + 2 * workers - 1) // (2 * workers))
    chunks = _pool(workers).map([slim[i:i + per] for i in range(0, n, per)])
    return merge_chunks(chunks), None
