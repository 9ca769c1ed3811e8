"""Render engine: host planning (plan.py) -> device-resident job tables -> the kernel sequence.

`render(params, progress=None)` is the drop-in for reference `render` (main_v2.py:588-792): same
arguments, same `(float64[out_n, 2], meta)` result, same progress call shapes, exceptions propagate.
`BatchRenderer` / `render_batch` run many independent renders (the reference's batch loop,
main_v2.py:1578-1593) as one batched launch sequence on one GPU; `parallel.py` shards a batch over
the GPUs of a box.

Stage order (one batched launch sequence for all renders):
  synth (PCG64 + ziggurat normals, closed-form modes)            ms_synth_normal / ms_synth_dust
  [tilted-noise modes: rfft -> power-law tilt -> irfft, finish]  ms_spectral_* / ms_synth_tilt_finish
  grain spectral op: low-pass -> stretch -> multiband            ms_spectral_*
  overlap-add placement + ADSR                                   ms_overlap_add
  reflection cloud (+) impulse response as one FIR, overlap-save ms_fir_build / ms_fir_*
  stereo diffusion, soft clip, normalise                         ms_post
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _abi, plan as P


# --------------------------------------------------------------------------- device abstraction
class CudaDevice:
    """torch-backed device memory on one GPU.  The only device the product ships."""

    def __init__(self, index=None):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("audio_suite_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.torch = torch
        self.index = torch.cuda.current_device() if index is None else int(index)
        self.dev = torch.device("cuda", self.index)
        self.lib = _abi.lib()
        self.uploaded = 0

    def stream_ptr(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.dev).cuda_stream)

    def empty(self, n, dtype):
        t = {np.float32: self.torch.float32, np.float64: self.torch.float64, np.uint8: self.torch.uint8,
             np.int32: self.torch.int32, np.uint64: self.torch.int64}[dtype]
        return self.torch.empty(max(1, int(n)), dtype=t, device=self.dev)

    def zeros(self, n, dtype):
        b = self.empty(n, dtype)
        b.zero_()
        return b

    def upload(self, arr):
        """numpy array (any dtype) -> device bytes."""
        a = np.array(arr, copy=True, order="C")
        self.uploaded += a.nbytes
        host = self.torch.from_numpy(a.view(np.uint8).reshape(-1) if a.size else np.zeros(1, np.uint8))
        return host.pin_memory().to(self.dev, non_blocking=True)

    def ptr(self, buf):
        return C.c_void_p(buf.data_ptr())

    def download(self, buf, offset, count):
        """slice of a device buffer -> numpy"""
        return buf[offset:offset + count].cpu().numpy()

    def synchronize(self):
        self.torch.cuda.synchronize(self.dev)


def _check(dev, rc):
    if rc != 0:
        raise RuntimeError("microsound_b200: " + (dev.lib.ms_last_error() or b"unknown error").decode())


def choose_precision(plans):
    """'auto' precision rule: float64.

    The kernels exist in float32 and float64 (B200 runs FP64 FMAs at half the FP32 rate, so f64 costs
    about 2x, not 30x).  float32 keeps every stage within ~3e-7 of that stage's *peak*, but render()
    ends with two amplifiers of small-signal error that the 1e-5 max-abs bar does not survive in
    general: (1) soft clip after the FIR stage -- with an impulse response or reflection cloud the
    pre-clip peak is routinely 1e2..1e3 and tanh has unit slope at every zero crossing; (2) peak
    normalisation -- when only the quiet start of a grain is placed (short outputs, slow ADSR attack:
    the factory default is a 1.25 ms grain under a 20 ms attack) the audible peak is <1 % of the grain
    peak and normalize() scales the float32 floor up with it.  float32 stays available as an explicit
    choice (`precision="f32"`); parity tests pin where it is within tolerance (C1, C1b, C2)."""
    return "f64"


def _recs(ctype, n):
    return np.zeros(n, dtype=np.dtype(ctype))


def _bessel_coeffs(theta, K=_abi.POST_K):
    """J_m(theta), m = -K..K by the ascending series (theta <= 0.9 here, converges in a few terms).
    exp(i theta sin(phi)) = sum_m J_m(theta) exp(i m phi): the stereo rotation of
    spectral_diffusion_stereo (main_v2.py:432-435) is, for even n, exactly a (2K+1)-tap circular FIR
    with taps at even lags."""
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
    return out


def _pair_jobs(items):
    """items: list of (n, in_off, out_off, SpecOp).  Two signals of equal n share one complex transform."""
    by_n = {}
    for it in items:
        by_n.setdefault(it[0], []).append(it)
    jobs = []
    for n, group in by_n.items():
        for i in range(0, len(group) - 1, 2):
            a, b = group[i], group[i + 1]
            j = _abi.SpecJob()
            j.n, j.in_a, j.in_b, j.out_a, j.out_b = n, a[1], b[1], a[2], b[2]
            j.op[0], j.op[1] = a[3], b[3]
            jobs.append(j)
        if len(group) % 2:
            a = group[-1]
            j = _abi.SpecJob()
            j.n, j.in_a, j.in_b, j.out_a, j.out_b = n, a[1], -1, a[2], -1
            j.op[0] = a[3]
            jobs.append(j)
    return jobs


class _SpectralStage:
    def __init__(self, dev, api, jobs, src, dst):
        self.dev, self.api, self.handle, self.njobs = dev, api, C.c_void_p(None), len(jobs)
        if not jobs:
            return
        arr = (_abi.SpecJob * len(jobs))(*jobs)
        need = api.ms_spectral_workspace_bytes(C.addressof(arr), len(jobs))
        if need == 0:
            _check(dev, -1)
        self.ws = dev.empty(need, np.uint8)
        _check(dev, api.ms_spectral_create(C.addressof(arr), len(jobs), dev.ptr(src), dev.ptr(dst),
                                           dev.ptr(self.ws), need, dev.stream_ptr(), C.byref(self.handle)))

    def run(self):
        if self.njobs:
            _check(self.dev, self.api.ms_spectral_run(self.handle, self.dev.stream_ptr()))

    def close(self):
        if self.handle:
            self.api.ms_spectral_destroy(self.handle)
            self.handle = C.c_void_p(None)


class BatchRenderer:
    """Plans a batch of independent renders once, keeps every table resident on the device and
    re-runs the kernel sequence on demand (`run()`), e.g. for benchmarking or repeated renders."""

    def __init__(self, params_list, device=None, precision="auto"):
        import time as _time
        self.dev = device or CudaDevice()
        t0 = _time.perf_counter()
        self.plans = P.plan_many(params_list)
        self.t_plan = _time.perf_counter() - t0
        self.precision = choose_precision(self.plans) if precision == "auto" else precision
        self.api = _abi.Api(self.dev.lib, self.precision)
        self.real = np.float32 if self.precision == "f32" else np.float64
        up0 = getattr(self.dev, "uploaded", 0) + self.dev.lib.ms_h2d_bytes()
        t0 = _time.perf_counter()
        self._pack()
        self.t_pack = _time.perf_counter() - t0
        self.h2d_bytes = getattr(self.dev, "uploaded", 0) + self.dev.lib.ms_h2d_bytes() - up0

    # ---- layout + tables ---------------------------------------------------------------------------
    def _pack(self):
        dev, plans, real = self.dev, self.plans, self.real
        n_evt = sum(len(rp.events) for rp in plans)
        R = len(plans)
        sy1 = _recs(_abi.SynthEvt, n_evt)            # stage 1 (normals / closed form / dust)
        sy2 = _recs(_abi.SynthEvt, n_evt)            # tilt finish
        ola_r = _recs(_abi.OlaRender, R)
        ola_e = []
        fir_r, post_r = [], _recs(_abi.PostRender, R)
        tilt_items, grain_items = [], []
        dust_pos, dust_val = [], []
        tap_off, tap_gain = [], []
        ir_chunks, ir_index = [], {}
        pool_n = 0
        mono_n = 0
        self.micro_at, self.grain_at, self.mono_at, self.out_at = [], [], [], []
        e = 0
        any_dust = any_tilt = False
        h_total = 0
        max_h = 0
        self.odd_stereo = []
        for r, rp in enumerate(plans):
            a, d, rel, S, curve = rp.adsr
            n = rp.out_n
            if a > n:
                raise ValueError(f"could not broadcast input array from shape ({a},) into shape ({n},)")   # M:182
            d_end = min(n, a + d) if d > 0 else a
            sus_end = max(d_end, n - rel)
            o = ola_r[r]
            o["out"], o["out_n"] = mono_n, n
            o["A"], o["D_end"], o["sus_end"] = a, d_end, sus_end
            o["has_release"] = 1 if (rel > 0 and n > sus_end) else 0
            o["inv_A"] = 1.0 / a if a > 0 else 0.0
            o["inv_D"] = 1.0 / (d_end - a) if d_end > a else 0.0
            o["inv_R"] = 1.0 / (n - sus_end - 1) if n - sus_end > 1 else 0.0
            o["S"], o["curve"] = S, curve
            o["ev_begin"] = len(ola_e)
            max_len = 0
            x_begin, x_end = n, 0
            last_micro = last_grain = None
            for ev in rp.events:
                rec1, rec2 = sy1[e], sy2[e]
                st = np.random.PCG64(ev.seed).state["state"]
                for rec in (rec1, rec2):
                    rec["s_hi"], rec["s_lo"] = st["state"] >> 64, st["state"] & 0xFFFFFFFFFFFFFFFF
                    rec["i_hi"], rec["i_lo"] = st["inc"] >> 64, st["inc"] & 0xFFFFFFFFFFFFFFFF
                    rec["n"], rec["mode"], rec["fade"], rec["sigma"] = ev.n, ev.mode, ev.fade, ev.sigma
                    rec["f_over_sr"], rec["inv_fade"] = ev.f_over_sr, 1.0 / ev.fade
                    rec["ring_decay"], rec["env_decay"] = ev.ring_decay, ev.env_decay
                    rec["ker_len"] = ev.ker_len
                micro = pool_n
                pool_n += ev.n
                rec1["out"] = micro
                rec2["out"] = micro
                rec2["mode"] = -1
                if ev.mode == P.MODE_DUST:
                    any_dust = True
                    rec1["dust_begin"], rec1["dust_count"] = sum(len(x) for x in dust_pos), len(ev.dust_pos)
                    dust_pos.append(ev.dust_pos)
                    dust_val.append(ev.dust_val)
                elif ev.mode in (P.MODE_NOISE, P.MODE_SKEW):
                    any_tilt = True
                    raw, tilted = pool_n, pool_n + ev.n
                    pool_n += 2 * ev.n
                    rec1["out"] = raw
                    rec2["mode"], rec2["aux"] = ev.mode, tilted
                    tilt_items.append((ev.n, raw, tilted, ev.tilt))
                grain = micro
                if ev.spec is not None:
                    grain = pool_n
                    pool_n += ev.n
                    grain_items.append((ev.n, micro, grain, ev.spec))
                last_micro, last_grain = (micro, ev.n), (grain, ev.n)
                if ev.placed:
                    ola_e.append((grain + ev.offset, ev.start, ev.length, ev.amp))
                    max_len = max(max_len, ev.length)
                    # grain[0] is exactly 0 (fade-in starts at 0, main_v2.py:267), so with no offset the
                    # first placed sample is an exact zero
                    x_begin = min(x_begin, ev.start + (1 if ev.offset == 0 else 0))
                    x_end = max(x_end, ev.start + ev.length)
                e += 1
            o["ev_end"], o["max_len"] = len(ola_e), max_len
            self.micro_at.append(last_micro)
            self.grain_at.append(last_grain)
            # FIR (reflection cloud folded into the impulse response)
            y_at = mono_n
            has_er = rp.er_offs is not None and rp.er_offs.size > 0
            if has_er or rp.ir is not None:
                if rp.ir is not None:
                    key = rp.ir.tobytes()
                    if key not in ir_index:
                        ir_index[key] = (sum(len(x) for x in ir_chunks), rp.ir.size)
                        ir_chunks.append(rp.ir.astype(real))
                    ir_at, ir_len = ir_index[key]
                else:
                    key = b"delta"
                    if key not in ir_index:
                        ir_index[key] = (sum(len(x) for x in ir_chunks), 1)
                        ir_chunks.append(np.ones(1, real))
                    ir_at, ir_len = ir_index[key]
                f = _abi.FirRender()
                f.ir, f.ir_len = ir_at, ir_len
                f.tap_begin = sum(len(x) for x in tap_off)
                if has_er:
                    if rp.er_offs.size > 4096:
                        raise ValueError("er_taps > 4096 is outside the accelerated path")
                    tap_off.append(rp.er_offs)
                    tap_gain.append(rp.er_gains.astype(real))
                    f.h_len = ir_len + int(rp.er_offs.max())
                else:
                    f.h_len = ir_len
                f.tap_end = sum(len(x) for x in tap_off)
                f.h = h_total
                h_total += f.h_len
                max_h = max(max_h, f.h_len)
                f.x, f.out_n = mono_n, n
                if x_begin == 0 and a > 0:
                    x_begin = 1                      # env[0] = 0 ** curve = 0 (main_v2.py:181-182)
                f.x_begin, f.x_end = min(x_begin, x_end), x_end
                fir_r.append((r, f))
            self.mono_at.append(mono_n)
            mono_n += n
        # second mono plane (FIR output) and the odd-length stereo scratch
        plane = mono_n
        extra = 0
        frames = 0
        for r, rp in enumerate(plans):
            pr = post_r[r]
            n = rp.out_n
            pr["n"], pr["out"] = n, frames
            pr["y"] = self.mono_at[r]
            pr["drive"] = rp.drive
            pr["inv_tanh_drive"] = 1.0 / math.tanh(rp.drive) if rp.drive > 0 else 1.0
            pr["peak"] = rp.peak
            if rp.stereo_on:
                pr["dl"], pr["dr"] = rp.stereo_dl, rp.stereo_dr
                if n % 2 == 0:
                    pr["stereo_mode"] = 1
                    pr["coef"] = _bessel_coeffs(rp.stereo_theta)
                    pr["rbuf"] = 2 * plane + extra               # right channel, written by the max pass
                    extra += n
                else:
                    pr["stereo_mode"] = 2
                    pr["rbuf"] = 2 * plane + extra + n          # [rolled copy | right channel]
                    self.odd_stereo.append((r, 2 * plane + extra, n, rp.stereo_dr, rp.stereo_theta))
                    extra += 2 * n
            self.out_at.append(frames)
            frames += n
        self.y_at = list(self.mono_at)
        for r, f in fir_r:
            f.y = plane + self.mono_at[r]
            post_r[r]["y"] = f.y
            self.y_at[r] = f.y
        self.n_renders, self.n_evt, self.frames = R, n_evt, frames
        self.max_out_n = max(rp.out_n for rp in plans)
        self.any_dust, self.any_tilt = any_dust, any_tilt
        self.pool_n, self.mono_n = pool_n, 2 * plane + extra

        # ---- device buffers
        self.pool = dev.empty(pool_n, real)
        self.mono = dev.zeros(self.mono_n, real)
        self.out = dev.empty(2 * frames, np.float32)
        self.maxbits = dev.zeros(R, np.uint64)
        self.d_sy1, self.d_sy2 = dev.upload(sy1), dev.upload(sy2)
        ola_e_arr = _recs(_abi.OlaEvt, len(ola_e))
        for i, (g, s, ln, amp) in enumerate(ola_e):
            ola_e_arr[i]["grain"], ola_e_arr[i]["start"], ola_e_arr[i]["len"], ola_e_arr[i]["amp"] = g, s, ln, amp
        self.d_ola_r, self.d_ola_e = dev.upload(ola_r), dev.upload(ola_e_arr)
        self.d_post = dev.upload(post_r)
        if any_dust:
            self.d_dpos = dev.upload(np.concatenate(dust_pos).astype(np.int32))
            self.d_dval = dev.upload(np.concatenate(dust_val).astype(real))
        # spectral stages
        self.tilt_stage = _SpectralStage(dev, self.api, _pair_jobs(tilt_items), self.pool, self.pool)
        self.grain_stage = _SpectralStage(dev, self.api, _pair_jobs(grain_items), self.pool, self.pool)
        # FIR
        self.fir_handle = C.c_void_p(None)
        self.n_fir = len(fir_r)
        if fir_r:
            arr = (_abi.FirRender * len(fir_r))(*[f for _, f in fir_r])
            self.fir_arr = arr
            self.d_fir = dev.upload(np.frombuffer(bytes(arr), dtype=np.uint8))
            self.d_tap_off = dev.upload(np.concatenate(tap_off).astype(np.int32) if tap_off else np.zeros(1, np.int32))
            self.d_tap_gain = dev.upload(np.concatenate(tap_gain).astype(real) if tap_gain else np.zeros(1, real))
            self.d_ir = dev.upload(np.concatenate(ir_chunks).astype(real))
            self.hpool = dev.empty(h_total, real)
            self.max_h = max_h
            need = self.api.ms_fir_workspace_bytes(C.addressof(arr), len(fir_r))
            if need == 0:
                _check(dev, -1)
            self.fir_ws = dev.empty(need, np.uint8)
            _check(dev, self.api.ms_fir_create(C.addressof(arr), len(fir_r), dev.ptr(self.hpool), dev.ptr(self.mono),
                                              dev.ptr(self.mono), dev.ptr(self.fir_ws), need, dev.stream_ptr(),
                                              C.byref(self.fir_handle)))
        # odd-length stereo: rolled copy -> rotation through the spectral engine -> right channel
        rot_items = []
        for (r, scratch, n, dr, theta) in self.odd_stereo:
            op = _abi.SpecOp()
            op.kind, op.alpha = _abi.OP_ROT, theta
            rot_items.append((n, scratch, scratch + n, op))
        self.rot_stage = _SpectralStage(dev, self.api, _pair_jobs(rot_items), self.mono, self.mono)

    # ---- execution -------------------------------------------------------------------------------------
    def run(self, mark=None):
        """Launch the whole kernel sequence on the current stream.  `mark(name)` (optional) is called
        after each stage has been enqueued (bench.py records a CUDA event there)."""
        dev, lib = self.dev, self.api
        st = dev.stream_ptr()
        mark = mark or (lambda name: None)
        if self.n_evt:
            _check(dev, lib.ms_synth_normal(dev.ptr(self.d_sy1), self.n_evt, dev.ptr(self.pool), st))
            if self.any_dust:
                _check(dev, lib.ms_synth_dust(dev.ptr(self.d_sy1), self.n_evt, dev.ptr(self.d_dpos), dev.ptr(self.d_dval),
                                              dev.ptr(self.pool), st))
            mark("synth")
            if self.any_tilt:
                self.tilt_stage.run()
                _check(dev, lib.ms_synth_tilt_finish(dev.ptr(self.d_sy2), self.n_evt, dev.ptr(self.pool), st))
                mark("tilt_spectral")
            self.grain_stage.run()
            mark("grain_spectral")
        _check(dev, lib.ms_overlap_add(dev.ptr(self.d_ola_r), self.n_renders, self.max_out_n, dev.ptr(self.d_ola_e),
                                       dev.ptr(self.pool), dev.ptr(self.mono), st))
        mark("overlap_add")
        if self.n_fir:
            _check(dev, lib.ms_fir_build(dev.ptr(self.d_fir), self.n_fir, self.max_h, dev.ptr(self.d_tap_off),
                                         dev.ptr(self.d_tap_gain), dev.ptr(self.d_ir), dev.ptr(self.hpool), st))
            mark("fir_build")
            _check(dev, lib.ms_fir_run(self.fir_handle, st))
            mark("fir_overlap_save")
        if self.odd_stereo:
            base, isz = dev.ptr(self.mono).value, np.dtype(self.real).itemsize
            for (r, scratch, n, dr, theta) in self.odd_stereo:
                _check(dev, lib.ms_roll(C.c_void_p(base + isz * self.y_at[r]), C.c_void_p(base + isz * scratch), n, dr, st))
            self.rot_stage.run()
        _check(dev, lib.ms_post(dev.ptr(self.d_post), self.n_renders, self.max_out_n, dev.ptr(self.mono),
                                dev.ptr(self.maxbits), dev.ptr(self.out), st))
        mark("post")

    # ---- results ---------------------------------------------------------------------------------------
    def output(self, r):
        """float32 [out_n, 2] of render r (host copy)."""
        n = self.plans[r].out_n
        return self.dev.download(self.out, 2 * self.out_at[r], 2 * n).reshape(n, 2)

    def outputs_device(self):
        return self.out

    def meta(self, r):
        rp = self.plans[r]
        m = dict(out_sr=rp.base_sr, design_sr_base=rp.design_sr_base, micro_last=None, grain_last=None)
        if self.micro_at[r] is not None:
            o, n = self.micro_at[r]
            m["micro_last"] = self.dev.download(self.pool, o, n).astype(np.float64)
            o, n = self.grain_at[r]
            m["grain_last"] = self.dev.download(self.pool, o, n).astype(np.float64)
        return m

    def close(self):
        for s in (self.tilt_stage, self.grain_stage, self.rot_stage):
            s.close()
        if self.fir_handle:
            self.api.ms_fir_destroy(self.fir_handle)
            self.fir_handle = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def render(params, progress=None, device=None, precision="auto"):
    """Drop-in for reference render() (main_v2.py:588-792)."""
    br = BatchRenderer([params], device=device, precision=precision)
    rp = br.plans[0]
    if progress:
        progress(0, f"Output SR {rp.base_sr} Hz | Design SR {rp.design_sr_base} Hz")
    br.run()
    if progress:
        n_evt = len(rp.events)
        for ev in rp.events:
            if ev.placed and ev.index % 50 == 0:
                progress(int(5 + 70 * (ev.index / max(1, n_evt))), f"Events {ev.index}/{n_evt}")
    audio = br.output(0).astype(np.float64)
    meta = br.meta(0)
    br.close()
    if progress:
        progress(100, "Done.")
    return audio, meta


def render_batch(params_list, device=None, precision="auto"):
    """Independent renders as one batched launch sequence.  Returns a list of float32 [out_n, 2]."""
    br = BatchRenderer(params_list, device=device, precision=precision)
    br.run()
    outs = [br.output(r) for r in range(br.n_renders)]
    br.close()
    return outs
