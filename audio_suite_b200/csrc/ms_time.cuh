// ms_time.cuh -- time-domain stages of render(): overlap-add placement + ADSR (main_v2.py:742-764,
// 172-195), the reflection-cloud / impulse-response FIR builder (main_v2.py:409-421, 438-445), and the
// stereo / soft-clip / normalise tail (main_v2.py:423-436, 31-34, 26-29, 775-781).

#define OLA_TILE 1024
#define OLA_NTHR 256
#define POST_K MS_POST_K          // Bessel taps -K..K of the stereo rotation
#define POST_NC (2 * POST_K + 1)

// ---- overlap-add ("unfold" placement) + ADSR --------------------------------------------------------
typedef ms_ola_render OlaRender;
typedef ms_ola_evt OlaEvt;


MS_DEV real adsr_gain(const OlaRender& R, int i) {
    // Evaluated in the working precision: an error relative to a loud sample is still an error relative
    // to the peak, and the FIR gain + soft clip downstream turn 1e-7 of that into 1e-5 (precision).
    const real S = (real)R.S, curve = (real)R.curve;
    if (i < R.A) return r_pow((real)((double)i * R.inv_A), curve);
    if (i < R.D_end) return (real)1.0 - ((real)1.0 - S) * r_pow((real)((double)(i - R.A) * R.inv_D), curve);
    if (i < R.sus_end || !R.has_release) return S;
    const real r = (i == R.out_n - 1 && R.out_n - R.sus_end > 1) ? (real)1.0 : (real)((double)(i - R.sus_end) * R.inv_R);
    return S * ((real)1.0 - r_pow(r, curve));
}

// grid = (ceil(max out_n / OLA_TILE), renders), block = OLA_NTHR.  Gather form: every output sample
// sums the events that cover it, in event order, so no atomics and each output is written once.
MS_DEV void ola_adsr_body(const OlaRender* MS_RESTRICT renders, const OlaEvt* MS_RESTRICT evts,
                          const real* MS_RESTRICT pool, real* MS_RESTRICT mono, const Ctx& c) {
    const OlaRender R = renders[c.by];
    const int t0 = c.bx * OLA_TILE;
    if (t0 >= R.out_n) return;
    const int t1 = (t0 + OLA_TILE) < R.out_n ? (t0 + OLA_TILE) : R.out_n;
    // candidates: start < t1 and start > t0 - max_len
    int lo = R.ev_begin, hi = R.ev_end;
    while (lo < hi) { const int m = (lo + hi) >> 1; if (evts[m].start >= t1) hi = m; else lo = m + 1; }
    const int eb = lo;
    lo = R.ev_begin; hi = eb;
    while (lo < hi) { const int m = (lo + hi) >> 1; if (evts[m].start > t0 - R.max_len) hi = m; else lo = m + 1; }
    const int ea = lo;
    real* out = mono + R.out;
    for (int i = t0 + c.tid; i < t1; i += c.nthr) {
        real acc = (real)0.;
        for (int e = ea; e < eb; ++e) {
            const int st = __ldg(&evts[e].start), ln = __ldg(&evts[e].len);
            const int k = i - st;
            if (k >= 0 && k < ln) acc += (real)__ldg(&evts[e].amp) * __ldg(&pool[evts[e].grain + k]);
        }
        out[i] = acc * adsr_gain(R, i);
    }
}

// ---- reflection cloud as a dense tap vector: e[off_t] += g_t  (early_reflection_cloud, main_v2.py:413-420) ----
// The cloud and the impulse response are both causal LTI filters, so they are applied as ONE filter whose
// spectrum is IRspec * (1 + FFT(e)); e is only max_off + 1 samples long.
typedef ms_fir_render FirRender;
struct ErJob { long long e; int elen, tap_begin, tap_end; };
MS_DEV void er_scatter_body(const ErJob* MS_RESTRICT jobs, const int* MS_RESTRICT tap_off, const real* MS_RESTRICT tap_gain,
                            real* ebase, const Ctx& c) {
    const ErJob J = jobs[c.by];
    real* e = ebase + J.e;
    for (int i = c.tid; i < J.elen; i += c.nthr) e[i] = (real)0.;
    c.sync();
    for (int t = J.tap_begin + c.tid; t < J.tap_end; t += c.nthr) {
        const int off = tap_off[t];
        if (off >= 0 && off < J.elen) {
#ifdef MS_HOST_EMUL
            e[off] += tap_gain[t];
#else
            atomicAdd(&e[off], tap_gain[t]);        // several taps may share a delay
#endif
        }
    }
}

// ---- stereo diffusion + soft clip + normalise -----------------------------------------------------------
typedef ms_post_render PostRender;
MS_DEV int wrap_idx(long long i, int n) { long long r = i % n; if (r < 0) r += n; return (int)r; }
// the clipped value is stored as float32, so tanh is evaluated in float32 (argument rounded once: 6e-8 relative)
MS_DEV real soft_clip(real v, real drive, real inv_t) { return drive > (real)0. ? (real)tanhf((float)(v * drive)) * inv_t : v; }

// pass 1: per-render max(|L|, |R|) before the clip (tanh is monotonic, so the clipped max follows).
// stereo_mode 1 also materialises the right channel at rbuf so that pass 2 does not redo the 25 taps.
// The tile's input window (1024 + 4K samples, circular) and the taps are staged in shared memory.
MS_DEV void post_max_body(const PostRender* MS_RESTRICT renders, real* mono, unsigned long long* MS_RESTRICT maxbits, const Ctx& c) {
    const PostRender& R = renders[c.by];
    const int n = R.n, mode = R.stereo_mode;
    const int t0 = c.bx * OLA_TILE;
    if (t0 >= n) return;
    const int t1 = (t0 + OLA_TILE) < n ? (t0 + OLA_TILE) : n;
    const real* y = mono + R.y;
    real* win = (real*)c.smem;                       // OLA_TILE + 4K
    real* coef = win + (OLA_TILE + 4 * POST_K);       // POST_NC
    real* red = coef + POST_NC + 1;                  // nthr
    if (mode == 1) {
        const int W = (t1 - t0) + 4 * POST_K;
        const long long w0 = (long long)t0 + R.dr - 2 * POST_K;
        for (int j = c.tid; j < W; j += c.nthr) win[j] = y[wrap_idx(w0 + j, n)];
        for (int j = c.tid; j < POST_NC; j += c.nthr) coef[j] = (real)R.coef[j];
        c.sync();
    }
    real m = (real)0.;
    for (int i = t0 + c.tid; i < t1; i += c.nthr) {
        m = r_max(m, r_abs(y[i]));
        if (mode == 1) {
            real acc = (real)0.;
            const real* w = win + (i - t0);
#pragma unroll
            for (int k = 0; k < POST_NC; ++k) acc += coef[k] * w[2 * k];
            mono[R.rbuf + i] = acc;
            m = r_max(m, r_abs(acc));
        } else if (mode == 2) {
            m = r_max(m, r_abs(mono[R.rbuf + i]));
        }
    }
    red[c.tid] = m;
    c.sync();
    for (int s = c.nthr >> 1; s > 0; s >>= 1) {
        if (c.tid < s) red[c.tid] = r_max(red[c.tid], red[c.tid + s]);
        c.sync();
    }
    if (c.tid == 0) {
        union { double f; unsigned long long u; } cv; cv.f = (double)red[0];
#ifdef MS_HOST_EMUL
        if (cv.u > maxbits[c.by]) maxbits[c.by] = cv.u;
#else
        atomicMax(&maxbits[c.by], cv.u);       // non-negative doubles order like their bit patterns
#endif
    }
}
// pass 2: write interleaved stereo, clipped and scaled to the requested peak
MS_DEV void post_write_body(const PostRender* MS_RESTRICT renders, const real* MS_RESTRICT mono, const unsigned long long* MS_RESTRICT maxbits,
                            float2* MS_RESTRICT out, const Ctx& c) {
    const PostRender R = renders[c.by];
    const int t0 = c.bx * OLA_TILE;
    if (t0 >= R.n) return;
    const int t1 = (t0 + OLA_TILE) < R.n ? (t0 + OLA_TILE) : R.n;
    const real* y = mono + R.y;
    union { double f; unsigned long long u; } cv; cv.u = maxbits[c.by];
    const real drive = (real)R.drive, inv_t = (real)R.inv_tanh_drive;
    const real top = soft_clip((real)cv.f, drive, inv_t);
    const real scale = top > (real)0. ? (real)R.peak / top : (real)1.0;
    float2* o = out + R.out;
    for (int i = t0 + c.tid; i < t1; i += c.nthr) {
        const real l = R.stereo_mode ? y[wrap_idx((long long)i - R.dl, R.n)] : y[i];
        const real r = R.stereo_mode ? mono[R.rbuf + i] : y[i];
        o[i] = make_float2((float)(soft_clip(l, drive, inv_t) * scale), (float)(soft_clip(r, drive, inv_t) * scale));
    }
}
// circular shift used by the odd-length stereo path: dst[i] = src[(i + shift) mod n]
MS_DEV void roll_body(const real* MS_RESTRICT src, real* MS_RESTRICT dst, int n, int shift, const Ctx& c) {
    for (int i = c.bx * c.nthr + c.tid; i < n; i += c.nthr * 64) dst[i] = src[wrap_idx((long long)i + shift, n)];
}
