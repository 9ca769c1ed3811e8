// ms_synth.cuh -- transient synthesis: the five gen_basic modes of the reference
// (main_v2.py:219-269) with numpy's exact random stream generated on the device.
//
// numpy's default_rng(seed).standard_normal(n) is PCG64 (128-bit LCG, XSL-RR output) feeding a
// 256-layer ziggurat.  About 1-2 % of the normals consume more than one 64-bit word (wedge test: one
// extra word; tail: two extra words per attempt), so "output j" is not "word j".  One CTA renders one
// event: every thread jumps the LCG to its own run of C consecutive words, classifies them, and the
// block resolves who-starts-in-which-state with a prefix scan over 6-state transition maps
// (S0 = fresh word, W = wedge test pending, T1+-/T2+- = tail loop).  Accepted normals go to a
// shared staging buffer in stream order; the mode's closed-form terms (sine with float64 phase,
// exponentials, fades) are applied on the way out with coalesced stores.

enum { SY_GAUSS = 0, SY_DUST = 1, SY_NOISE = 2, SY_SKEW = 3, SY_RES = 4, SY_PLAIN = 5, SY_WAVELET = 6, SY_IRFRAG = 7, SY_SCANLINE = 8, SY_SILENT = 9, SY_CHAOS = 10, SY_STICK = 11 };

#define SY_C 8            // words per thread per round
#define SY_NTHR 256
#define SY_W (SY_C * SY_NTHR)

typedef ms_synth_evt SynthEvt;

struct u128 { unsigned long long hi, lo; };

MS_DEV unsigned long long ms_mulhi64(unsigned long long a, unsigned long long b) {
#ifdef MS_HOST_EMUL
    return (unsigned long long)(((unsigned __int128)a * b) >> 64);
#else
    return __umul64hi(a, b);
#endif
}
MS_DEV u128 u128_mul(u128 a, u128 b) {
    u128 r;
    r.lo = a.lo * b.lo;
    r.hi = ms_mulhi64(a.lo, b.lo) + a.lo * b.hi + a.hi * b.lo;
    return r;
}
MS_DEV u128 u128_add(u128 a, u128 b) {
    u128 r;
    r.lo = a.lo + b.lo;
    r.hi = a.hi + b.hi + (r.lo < a.lo ? 1ull : 0ull);
    return r;
}
#define PCG_MULT_HI 0x2360ED051FC65DA4ull
#define PCG_MULT_LO 0x4385DF649FCCF645ull
MS_DEV u128 pcg_step(u128 s, u128 inc) {
    u128 m; m.hi = PCG_MULT_HI; m.lo = PCG_MULT_LO;
    return u128_add(u128_mul(s, m), inc);
}
MS_DEV unsigned long long pcg_output(u128 s) {
    const unsigned long long x = s.hi ^ s.lo;
    const unsigned rot = (unsigned)(s.hi >> 58);
    return (x >> rot) | (x << ((64u - rot) & 63u));
}
// (mult, plus) such that state_{k+delta} = mult * state_k + plus
MS_DEV void pcg_jump_consts(u128 inc, unsigned long long delta, u128* mult, u128* plus) {
    u128 am, ap, cm, cp;
    am.hi = 0; am.lo = 1; ap.hi = 0; ap.lo = 0;
    cm.hi = PCG_MULT_HI; cm.lo = PCG_MULT_LO; cp = inc;
    while (delta) {
        if (delta & 1ull) { am = u128_mul(am, cm); ap = u128_add(u128_mul(ap, cm), cp); }
        u128 one; one.hi = 0; one.lo = 1;
        cp = u128_mul(u128_add(cm, one), cp);
        cm = u128_mul(cm, cm);
        delta >>= 1;
    }
    *mult = am; *plus = ap;
}

// ---- ziggurat state machine ------------------------------------------------------------------------
enum { ZS_S0 = 0, ZS_W = 1, ZS_T1P = 2, ZS_T1N = 3, ZS_T2P = 4, ZS_T2N = 5 };
#define ZIG_R 3.6541528853610087963519472518
#define ZIG_INV_R 0.27366123732975827203338247596

struct ZigTables { const unsigned long long* ki; const double* wi; const double* fi; };

MS_DEV double word_to_double(unsigned long long w) { return (double)(w >> 11) * (1.0 / 9007199254740992.0); }

// One step of the machine on word `cur` (with `prev` = the word before it).  Returns the next state;
// *emit = 1 and *val = the normal when this word completes one.
MS_DEV int zig_step(int s, unsigned long long prev, unsigned long long cur, const ZigTables& T, int* emit, double* val) {
    *emit = 0;
    if (s == ZS_S0) {
        unsigned long long r = cur;
        const int idx = (int)(r & 0xff);
        r >>= 8;
        const int sign = (int)(r & 1);
        const unsigned long long rabs = (r >> 1) & 0x000fffffffffffffull;
        if (rabs < T.ki[idx]) {
            double x = (double)rabs * T.wi[idx];
            *val = sign ? -x : x; *emit = 1;
            return ZS_S0;
        }
        if (idx == 0) return ((rabs >> 8) & 1) ? ZS_T1N : ZS_T1P;
        return ZS_W;
    }
    if (s == ZS_W) {
        unsigned long long r = prev;
        const int idx = (int)(r & 0xff);
        r >>= 8;
        const int sign = (int)(r & 1);
        const unsigned long long rabs = (r >> 1) & 0x000fffffffffffffull;
        double x = (double)rabs * T.wi[idx];
        if (sign) x = -x;
        const double u = word_to_double(cur);
        if ((T.fi[idx - 1] - T.fi[idx]) * u + T.fi[idx] < exp(-0.5 * x * x)) { *val = x; *emit = 1; }
        return ZS_S0;
    }
    if (s == ZS_T1P) return ZS_T2P;
    if (s == ZS_T1N) return ZS_T2N;
    // T2: prev word gave xx, this one gives yy
    const double xx = -ZIG_INV_R * log1p(-word_to_double(prev));
    const double yy = -log1p(-word_to_double(cur));
    if (yy + yy > xx * xx) {
        *val = (s == ZS_T2N) ? -(ZIG_R + xx) : (ZIG_R + xx);
        *emit = 1;
        return ZS_S0;
    }
    return s == ZS_T2N ? ZS_T1N : ZS_T1P;
}

MS_DEV real fade_gain(int j, int n, int fade, double inv_fade) {
    real w = (real)1.;
    if (j < fade) w *= (real)((double)j * inv_fade);
    const int t = j - (n - fade);
    if (t >= 0) w *= (real)(1.0 - (double)t * inv_fade);
    return w;
}

// shared-memory carve-up of the synthesis kernel
struct SynthSmem {
    unsigned long long ki[256];
    double wi[256];
    double fi[256];
    unsigned long long words[SY_W + 1];     // [0] = word before the round, then i-major: 1 + i*NTHR + t
    real stage[SY_W + 8];
    int endst[SY_NTHR];                     // state each thread's run ends in
    int scan[2][SY_NTHR];                   // prefix sums of the per-thread output counts
    int flag[3];                            // "some thread changed its start state" (rotating, see below)
    int carry_state, _pad;
    int xend, xtotal;                       // cluster form: end state of this CTA's last run / its count of normals, read by the peers
};
MS_DEV unsigned long long sy_word(const SynthSmem* S, int p) {   // p in [-1, SY_W)
    if (p < 0) return S->words[0];
    const int t = p / SY_C, i = p - t * SY_C;
    return S->words[1 + i * SY_NTHR + t];
}
// One thread's run of SY_C consecutive words from state `s`: vals[i] = the normal completed by word i when bit i
// of *mask is set (static indices: the array stays in registers).  Returns the end state.
MS_DEV int zig_run(int s, const SynthSmem* S, int p0, const ZigTables& T, double* vals, unsigned* mask) {
    unsigned m = 0; int em;
#pragma unroll
    for (int i = 0; i < SY_C; ++i) {
        double v = 0.0;
        s = zig_step(s, sy_word(S, p0 + i - 1), sy_word(S, p0 + i), T, &em, &v);
        vals[i] = v;
        m |= (unsigned)em << i;
    }
    *mask = m;
    return s;
}

// mode-specific sample from normal z at index j
MS_DEV real synth_sample(const SynthEvt& E, int j, real z) {
    const real fj = (real)j;
    real x;
    if (E.mode == SY_GAUSS) {
        const real q = fj / (real)E.sigma;
        x = r_exp(-(real)0.5 * q * q) * (z * (real)0.12 + (real)1.0);
    } else if (E.mode == SY_RES) {
        double cyc = (double)j * E.f_over_sr;
        cyc -= floor(cyc);
        const real tone = r_sinpi((real)2.0 * (real)cyc) * r_exp(-fj * (real)E.ring_decay);
        x = (real)0.9 * tone + (real)0.25 * z * r_exp(-(real)j * (real)E.env_decay);
    } else if (E.mode == SY_PLAIN) {
        x = z * (real)0.1;
    } else {
        return z;           // SY_NOISE / SY_SKEW: raw normals, shaped later
    }
    return x * fade_gain(j, E.n, E.fade, E.inv_fade);
}

// One CTA per event.  blockDim.x must be SY_NTHR.
MS_DEV void synth_normal_body(const SynthEvt* MS_RESTRICT evts, real* MS_RESTRICT pool, const Ctx& c) {
    const SynthEvt E = evts[c.bx];
    if (E.mode == SY_DUST) return;
    SynthSmem* S = (SynthSmem*)c.smem;
    for (int i = c.tid; i < 256; i += c.nthr) {
        S->ki[i] = MS_ZIG_KI[i];
        union { unsigned long long u; double d; } cv;
        cv.u = MS_ZIG_WI_BITS[i]; S->wi[i] = cv.d;
        cv.u = MS_ZIG_FI_BITS[i]; S->fi[i] = cv.d;
    }
    if (c.tid == 0) S->carry_state = ZS_S0;
    ZigTables T; T.ki = S->ki; T.wi = S->wi; T.fi = S->fi;
    u128 inc; inc.hi = E.i_hi; inc.lo = E.i_lo;
    u128 st; st.hi = E.s_hi; st.lo = E.s_lo;
    u128 jm, jp;
    pcg_jump_consts(inc, (unsigned long long)(c.tid * SY_C), &jm, &jp);
    st = u128_add(u128_mul(st, jm), jp);
    pcg_jump_consts(inc, (unsigned long long)(SY_W - SY_C), &jm, &jp);
    real* out = pool + E.out;
    int out_base = 0;
    c.sync();
    const int max_rounds = (int)((2ll * E.n) / SY_W) + 64;
    for (int round = 0; round < max_rounds && out_base < E.n; ++round) {
        // ---- generate this thread's words
        if (c.tid == 0) { S->words[0] = pcg_output(st); S->flag[0] = 0; }
        for (int i = 0; i < SY_C; ++i) { st = pcg_step(st, inc); S->words[1 + i * SY_NTHR + c.tid] = pcg_output(st); }
        st = u128_add(u128_mul(st, jm), jp);
        c.sync();
        // ---- every run is first walked as if it began on a fresh word (true for ~99 % of them: a run inherits a
        //      pending wedge / tail state only when the previous run ended inside a multi-word draw).  Then the
        //      end states are published and any thread whose neighbour says otherwise re-walks its run from the
        //      state it really starts in; repeat until nobody changes (normally one check, rarely two).
        const int p0 = c.tid * SY_C;
        const int first = c.tid == 0 ? S->carry_state : ZS_S0;
        int start = first;
        unsigned mask;
        double vals[SY_C];
        int end = zig_run(start, S, p0, T, vals, &mask);
        for (int it = 0;; ++it) {
            S->endst[c.tid] = end;
            if (c.tid == 0) S->flag[(it + 1) % 3] = 0;
            c.sync();
            const int want = c.tid == 0 ? first : S->endst[c.tid - 1];
            if (want != start) {
                start = want;
                end = zig_run(start, S, p0, T, vals, &mask);
                S->flag[it % 3] = 1;
            }
            c.sync();
            if (!S->flag[it % 3]) break;
        }
        // ---- where each run's normals go: exclusive prefix sum of the counts
        const int cnt = MS_POPC(mask);
#ifdef MS_HOST_EMUL
        int cur = 0;                                       // (emulator: Hillis-Steele over the block)
        S->scan[0][c.tid] = cnt;
        c.sync();
        for (int d = 1; d < c.nthr; d <<= 1) {
            int v = S->scan[cur][c.tid];
            if (c.tid >= d) v += S->scan[cur][c.tid - d];
            S->scan[cur ^ 1][c.tid] = v;
            cur ^= 1;
            c.sync();
        }
        const int total = S->scan[cur][c.nthr - 1];
        int my_off = S->scan[cur][c.tid] - cnt;
#else
        // warp shuffles for the 32 lanes, one value per warp through shared memory: two block barriers instead of nine
        // (ncu: a fifth of this kernel's stall samples sat at the barriers of the eight-step block scan)
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, incl, d); if ((c.tid & 31) >= d) incl += o; }
        if ((c.tid & 31) == 31) S->scan[0][c.tid >> 5] = incl;
        c.sync();
        int wbase = 0, total = 0;
#pragma unroll
        for (int w = 0; w < SY_NTHR / 32; ++w) { const int t = S->scan[0][w]; if (w < (c.tid >> 5)) wbase += t; total += t; }
        int my_off = wbase + incl - cnt;                   // (scan[0] is rewritten a round later, behind several barriers)
#endif
#pragma unroll
        for (int i = 0; i < SY_C; ++i) if ((mask >> i) & 1u) S->stage[my_off++] = (real)vals[i];
        if (c.tid == c.nthr - 1) S->carry_state = end;
        c.sync();
        const int take = (E.n - out_base) < total ? (E.n - out_base) : total;
        for (int i = c.tid; i < take; i += c.nthr) out[out_base + i] = synth_sample(E, out_base + i, S->stage[i]);
        out_base += total;
        c.sync();
    }
}

#ifndef MS_HOST_EMUL
// ---- cluster form: CL CTAs share one event ---------------------------------------------------------------------------
// For small batches (a rank's share of a multi-GPU sweep, a single render) one CTA per event leaves most SMs idle and
// the longest event sets the kernel's duration.  Here a thread-block cluster walks an event together: in every round CTA
// `rank` takes the words [rank * SY_W, (rank + 1) * SY_W) of the cluster's CL * SY_W, and the two things a CTA needs from
// its left neighbour -- the ziggurat state its first run starts in, and how many normals the CTAs before it produced --
// are read from the neighbour's shared memory (DSMEM: mapa + ld.shared::cluster) between hardware cluster barriers.
// Same words, same state machine, same output order as synth_normal_body: bit-identical output.
MS_DEV void sy_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
MS_DEV unsigned sy_cluster_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
MS_DEV int sy_ld_peer(const int* p, unsigned rank) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    unsigned ra; int v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
    asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(v) : "r"(ra) : "memory");
    return v;
}
template <int CL>
__global__ void __launch_bounds__(SY_NTHR, 4) synth_normal_cluster_kernel(const SynthEvt* __restrict__ evts, real* __restrict__ pool) {
    extern __shared__ float4 ms_dyn_smem[];
    const int tid = threadIdx.x;
    const unsigned rank = sy_cluster_rank();
    const SynthEvt E = evts[blockIdx.x / CL];
    if (E.mode == SY_DUST) return;                          // (the whole cluster leaves)
    SynthSmem* S = (SynthSmem*)ms_dyn_smem;
    for (int i = tid; i < 256; i += SY_NTHR) {
        S->ki[i] = MS_ZIG_KI[i];
        union { unsigned long long u; double d; } cv;
        cv.u = MS_ZIG_WI_BITS[i]; S->wi[i] = cv.d;
        cv.u = MS_ZIG_FI_BITS[i]; S->fi[i] = cv.d;
    }
    ZigTables T; T.ki = S->ki; T.wi = S->wi; T.fi = S->fi;
    u128 inc; inc.hi = E.i_hi; inc.lo = E.i_lo;
    u128 st; st.hi = E.s_hi; st.lo = E.s_lo;
    u128 jm, jp;
    pcg_jump_consts(inc, (unsigned long long)(rank * SY_W + tid * SY_C), &jm, &jp);
    st = u128_add(u128_mul(st, jm), jp);
    pcg_jump_consts(inc, (unsigned long long)(CL * SY_W - SY_C), &jm, &jp);
    real* out = pool + E.out;
    int out_base = 0;
    int carry = ZS_S0;                                      // state the cluster's round starts in (used by rank 0, thread 0)
    __syncthreads();
    const int max_rounds = (int)((2ll * E.n) / (CL * SY_W)) + 64;
    for (int round = 0; round < max_rounds && out_base < E.n; ++round) {
        if (tid == 0) { S->words[0] = pcg_output(st); S->flag[0] = 0; }
        for (int i = 0; i < SY_C; ++i) { st = pcg_step(st, inc); S->words[1 + i * SY_NTHR + tid] = pcg_output(st); }
        st = u128_add(u128_mul(st, jm), jp);
        __syncthreads();
        const int p0 = tid * SY_C;
        const int first = (tid == 0 && rank == 0) ? carry : ZS_S0;
        int start = first;
        unsigned mask;
        double vals[SY_C];
        int end = zig_run(start, S, p0, T, vals, &mask);
        for (int it = 0;; ++it) {
            S->endst[tid] = end;
            if (tid == SY_NTHR - 1) S->xend = end;
            if (tid == 0) S->flag[(it + 1) % 3] = 0;
            sy_cluster_sync();
            int want;
            if (tid == 0) want = rank == 0 ? first : sy_ld_peer(&S->xend, rank - 1);
            else want = S->endst[tid - 1];
            if (want != start) {
                start = want;
                end = zig_run(start, S, p0, T, vals, &mask);
                S->flag[it % 3] = 1;
            }
            sy_cluster_sync();
            int any = 0;
#pragma unroll
            for (int r = 0; r < CL; ++r) any |= sy_ld_peer(&S->flag[it % 3], (unsigned)r);
            if (!any) break;
        }
        const int cnt = MS_POPC(mask);
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, incl, d); if ((tid & 31) >= d) incl += o; }
        if ((tid & 31) == 31) S->scan[0][tid >> 5] = incl;
        __syncthreads();
        int wbase = 0, total = 0;
#pragma unroll
        for (int w = 0; w < SY_NTHR / 32; ++w) { const int t = S->scan[0][w]; if (w < (tid >> 5)) wbase += t; total += t; }
        int my_off = wbase + incl - cnt;
#pragma unroll
        for (int i = 0; i < SY_C; ++i) if ((mask >> i) & 1u) S->stage[my_off++] = (real)vals[i];
        if (tid == 0) S->xtotal = total;
        sy_cluster_sync();                                  // (xend holds the final end state of the last run since the last iteration)
        int before = 0, all = 0;
#pragma unroll
        for (int r = 0; r < CL; ++r) { const int t = sy_ld_peer(&S->xtotal, (unsigned)r); if ((unsigned)r < rank) before += t; all += t; }
        if (tid == 0 && rank == 0) carry = sy_ld_peer(&S->xend, CL - 1);
        const int at = out_base + before;
        int take = E.n - at;
        take = take < 0 ? 0 : (take < total ? take : total);
        for (int i = tid; i < take; i += SY_NTHR) out[at + i] = synth_sample(E, at + i, S->stage[i]);
        out_base += all;
        sy_cluster_sync();                                  // peers have read xend / xtotal; stage and words may be rewritten
    }
}
#endif

// Finalize for the tilted-noise modes (main_v2.py:246-255): one thread per sample.
MS_DEV void synth_tilt_finish_body(const SynthEvt* MS_RESTRICT evts, real* MS_RESTRICT pool, const Ctx& c) {
    const SynthEvt E = evts[c.by];
    if (E.mode != SY_NOISE && E.mode != SY_SKEW) return;
    const real* t = pool + E.aux;
    real* out = pool + E.out;
    for (int j = c.bx * c.nthr + c.tid; j < E.n; j += c.nthr * 64) {   // gridDim.x == 64
        real v = t[j];
        if (E.mode == SY_SKEW) {
            const real a = v > (real)0. ? v : (real)0.;
            const real pv = j > 0 ? t[j - 1] : v;
            const real b = pv > (real)0. ? pv : (real)0.;
            v = a - b;
        }
        out[j] = v * r_exp(-(real)j * (real)E.env_decay) * fade_gain(j, E.n, E.fade, E.inv_fade);
    }
}

// Dust impulses (main_v2.py:239-245): sparse impulses convolved ("same") with ker = exp(-linspace(0,6,K)).
// ker is tabulated once per CTA in shared memory (K <= 0.01 n entries), so an output costs one table lookup
// per impulse in reach instead of one exp().  Events longer than the table fall back to exp().
#define DUST_KER_MAX 4096
#define DUST_STAGE_MAX 1024
#define DUST_CTAS 16
// grid = (DUST_CTAS, events): CTA bx owns the contiguous outputs [bx * span, (bx + 1) * span), span = ceil(n / DUST_CTAS)
// (16 CTAs per event: every CTA tabulates the kernel once, so fewer, longer chunks amortise those K exponentials).  The impulses
// that can reach them (positions in (first + ctr - K, last + ctr]) are found with two binary searches and staged in
// shared memory together with the kernel table, so an output costs a short search in shared memory plus one
// multiply-add per impulse in reach.  (More than DUST_STAGE_MAX impulses in reach: read them from global memory.)
MS_DEV void synth_dust_body(const SynthEvt* MS_RESTRICT evts, const int* MS_RESTRICT dpos, const real* MS_RESTRICT dval,
                            real* MS_RESTRICT pool, const Ctx& c) {
    const SynthEvt E = evts[c.by];
    if (E.mode != SY_DUST) return;
    const int* pos = dpos + E.dust_begin;
    const real* val = dval + E.dust_begin;
    real* out = pool + E.out;
    const int K = E.ker_len, ctr = (K - 1) / 2;
    const real rate = (real)6.0 / (real)(K - 1);
    const int span = (E.n + DUST_CTAS - 1) / DUST_CTAS;
    const int j0 = c.bx * span, j1 = (j0 + span) < E.n ? (j0 + span) : E.n;
    if (j0 >= E.n) return;
    real* ker = (real*)c.smem;
    real* sval = ker + DUST_KER_MAX;
    int* spos = (int*)(sval + DUST_STAGE_MAX);
    const int tab = K <= DUST_KER_MAX;
    if (tab) for (int d = c.tid; d < K; d += c.nthr) ker[d] = r_exp(-rate * (real)d);
    // impulses with  j0 + ctr - K < p <= j1 - 1 + ctr
    int lo = 0, hi_i = E.dust_count;
    while (lo < hi_i) { const int mid = (lo + hi_i) >> 1; if (__ldg(&pos[mid]) > j0 + ctr - K) hi_i = mid; else lo = mid + 1; }
    const int q0 = lo;
    lo = q0; hi_i = E.dust_count;
    while (lo < hi_i) { const int mid = (lo + hi_i) >> 1; if (__ldg(&pos[mid]) > j1 - 1 + ctr) hi_i = mid; else lo = mid + 1; }
    const int q1 = lo, cnt = q1 - q0;
    const int staged = cnt <= DUST_STAGE_MAX;
    if (staged) for (int q = c.tid; q < cnt; q += c.nthr) { spos[q] = __ldg(&pos[q0 + q]); sval[q] = __ldg(&val[q0 + q]); }
    c.sync();
    for (int j = j0 + c.tid; j < j1; j += c.nthr) {
        const int hi = j + ctr;            // impulses p with hi-K < p <= hi contribute ker[hi-p]
        real acc = (real)0.;
        if (staged) {
            int a = 0, b = cnt;            // first staged impulse with pos > hi - K
            while (a < b) { const int mid = (a + b) >> 1; if (spos[mid] > hi - K) b = mid; else a = mid + 1; }
            for (int q = a; q < cnt; ++q) {
                const int p = spos[q];
                if (p > hi) break;
                acc += sval[q] * (tab ? ker[hi - p] : r_exp(-rate * (real)(hi - p)));
            }
        } else {
            int a = q0, b = q1;
            while (a < b) { const int mid = (a + b) >> 1; if (__ldg(&pos[mid]) > hi - K) b = mid; else a = mid + 1; }
            for (int q = a; q < q1; ++q) {
                const int p = __ldg(&pos[q]);
                if (p > hi) break;
                acc += __ldg(&val[q]) * (tab ? ker[hi - p] : r_exp(-rate * (real)(hi - p)));
            }
        }
        out[j] = acc * fade_gain(j, E.n, E.fade, E.inv_fade);
    }
}

// Wavelet atoms (main_v2.py:317-331, morlet_atom :165-170): x[j] = hann(n)[j] * sum_k w_k * atom_k[(j - shift_k) mod n],
// atom_k[m] = exp(-0.5 (t/sigma)^2) cos(2 pi f0 t + phase), t = (m - n/2) / gen_sr.  The scalar draws (f0, sigma, phase,
// shift) come from the host planner; phase arithmetic is float64 (f0 t reaches 1e4 cycles).  One thread per sample.
typedef ms_wavelet_atom WaveletAtom;
MS_DEV void synth_wavelet_body(const SynthEvt* MS_RESTRICT evts, const WaveletAtom* MS_RESTRICT atoms, const int* MS_RESTRICT shifts,
                               real* MS_RESTRICT pool, const Ctx& c) {
    const SynthEvt E = evts[c.by];
    if (E.mode != SY_WAVELET) return;
    const WaveletAtom* A = atoms + E.atom_begin;
    const int* S = shifts + E.atom_begin;
    real* out = pool + E.out;
    const int n = E.n;
    const double half = 0.5 * (double)n;
    for (int j = c.bx * c.nthr + c.tid; j < n; j += c.nthr * 64) {
        double acc = 0.0;
        for (int k = 0; k < E.atom_count; ++k) {
            int m = j - S[k];                      // np.roll(atom, shift)[j] = atom[(j - shift) mod n]
            if (m < 0) m += n; else if (m >= n) m -= n;
            const double d = (double)m - half;
            const double q = d * A[k].inv_sigma;
            double cyc = d * A[k].f0_over_sr;
            cyc -= floor(cyc);
            acc += A[k].weight * exp(-0.5 * q * q) * cos(6.283185307179586476925286766559 * cyc + A[k].phase);
        }
        const double hann = n > 1 ? 0.5 - 0.5 * cospi(2.0 * (double)j / (double)(n - 1)) : 1.0;
        out[j] = (real)(acc * hann);
    }
}

// Table generators (gen_ir_fragment main_v2.py:333-348, gen_image_scanline :350-362): a host-chosen table of M values is
// stretched to n samples (np.interp over linspace(0,1,M) -> linspace(0,1,n)) under a Hann window; the IR fragment is then
// peak-normalised to 0.9, the scan line smoothed ("same") by exp(-linspace(0, 5, K)).  One CTA per event.
MS_DEV real table_sample(const real* MS_RESTRICT tab, int M, int n, int j) {
    double v;
    if (M < 2 || n < 2) v = (double)tab[0];
    else {
        const double pos = ((double)j / (double)(n - 1)) * (double)(M - 1);
        int m = (int)pos;
        if (m >= M - 1) m = M - 2;
        const double fr = pos - (double)m;
        v = (double)tab[m] + ((double)tab[m + 1] - (double)tab[m]) * fr;
    }
    const double hann = n > 1 ? 0.5 - 0.5 * cospi(2.0 * (double)j / (double)(n - 1)) : 1.0;
    return (real)(v * hann);
}
MS_DEV void synth_table_body(const SynthEvt* MS_RESTRICT evts, const real* MS_RESTRICT dval, real* pool, const Ctx& c) {
    const SynthEvt E = evts[c.bx];
    real* out = pool + E.out;
    const int n = E.n;
    if (E.mode == SY_SILENT) { for (int j = c.tid; j < n; j += c.nthr) out[j] = (real)0.; return; }
    const real* tab = dval + E.dust_begin;
    const int M = E.dust_count;
    if (E.mode == SY_IRFRAG) {
        real* red = (real*)c.smem;
        real mx = (real)0.;
        for (int j = c.tid; j < n; j += c.nthr) { const real v = table_sample(tab, M, n, j); out[j] = v; mx = r_max(mx, r_abs(v)); }
        red[c.tid] = mx;
        c.sync();
        for (int s = c.nthr >> 1; s > 0; s >>= 1) { if (c.tid < s) red[c.tid] = r_max(red[c.tid], red[c.tid + s]); c.sync(); }
        const real peak = red[0];
        if (peak > (real)0.) {                               // normalize(x, 0.9), main_v2.py:26-29
            const real g = (real)0.9 / peak;
            for (int j = c.tid; j < n; j += c.nthr) out[j] *= g;
        }
        return;
    }
    real* tmp = pool + E.aux;
    if (E.mode == SY_STICK) {
        // gen_stick_slip (main_v2.py:283-301): tmp holds the event's normals (one per sample, both states draw one).
        // The state machine is walked by one thread in the reference's float64 operation order, no fused multiply-add.
        if (c.tid == 0) {
            const double threshold = E.f_over_sr, build = E.ring_decay, decay = E.env_decay, noise = E.inv_fade;
            double force = 0.0;
            bool sticking = true;
            for (int j = 0; j < n; ++j) {
                const double z = (double)tmp[j];
                double x = 0.0;
                if (sticking) {
#ifdef MS_HOST_EMUL
                    volatile double t1 = z * noise; volatile double t2 = t1 + 0.2; volatile double t3 = build * t2; force = force + t3;
#else
                    force = __dadd_rn(force, __dmul_rn(build, __dadd_rn(__dmul_rn(z, noise), 0.2)));
#endif
                    if (fabs(force) > threshold) sticking = false;
                } else {
#ifdef MS_HOST_EMUL
                    volatile double t1 = 0.25 * z; x = force + t1; volatile double t2 = force * decay; force = t2;
#else
                    x = __dadd_rn(force, __dmul_rn(0.25, z));
                    force = __dmul_rn(force, decay);
#endif
                    if (fabs(force) < 0.02) { sticking = true; force = 0.0; }
                }
                out[j] = (real)x;
            }
        }
        c.sync();
        for (int j = c.tid; j < n; j += c.nthr)
            out[j] *= (real)(n > 1 ? 0.5 - 0.5 * cospi(2.0 * (double)j / (double)(n - 1)) : 1.0);
        return;
    }
    if (E.mode == SY_CHAOS) {
        // gen_micro_chaos (main_v2.py:303-315).  The gate draws (one uniform per sample: (word >> 11) * 2^-53 < gate) are
        // position-independent, so every thread jumps the 128-bit LCG to its own samples; the logistic map itself is a
        // chaotic recurrence and is iterated by ONE thread with the reference's exact float64 operation order
        // ((r * y) * (1.0 - y), no fused multiply-add), so it stays bit-identical for any grain length.
        u128 inc; inc.hi = E.i_hi; inc.lo = E.i_lo;
        u128 st; st.hi = E.s_hi; st.lo = E.s_lo;
        u128 jm, jp;
        pcg_jump_consts(inc, (unsigned long long)c.tid + 1ull, &jm, &jp);      // sample j uses word j (state after j + 1 steps)
        st = u128_add(u128_mul(st, jm), jp);
        pcg_jump_consts(inc, (unsigned long long)c.nthr, &jm, &jp);
        const double gate = E.ring_decay;
        for (int j = c.tid; j < n; j += c.nthr) {
            tmp[j] = word_to_double(pcg_output(st)) < gate ? (real)1. : (real)0.;
            st = u128_add(u128_mul(st, jm), jp);
        }
        c.sync();
        if (c.tid == 0) {
            const double r = E.f_over_sr;
            double y = E.env_decay;
            for (int j = 0; j < n; ++j) {
#ifdef MS_HOST_EMUL
                volatile double ry = r * y; volatile double om = 1.0 - y; y = ry * om;
#else
                y = __dmul_rn(__dmul_rn(r, y), __dsub_rn(1.0, y));
#endif
                tmp[j] = tmp[j] != (real)0. ? (real)(y - 0.5) : (real)0.;
            }
        }
        c.sync();
    } else {            // SY_SCANLINE
        for (int j = c.tid; j < n; j += c.nthr) tmp[j] = table_sample(tab, M, n, j);
        c.sync();
    }
    const int K = E.ker_len, ctr = (K - 1) / 2;
    const real rate = (real)5.0 / (real)(K - 1);
    for (int j = c.tid; j < n; j += c.nthr) {
        real acc = (real)0.;
        for (int m = 0; m < K; ++m) {                          // np.convolve(x, ker, "same")[j] = sum_m ker[m] x[j + ctr - m]
            const int i = j + ctr - m;
            if (i >= 0 && i < n) acc += r_exp(-rate * (real)m) * tmp[i];
        }
        if (E.mode == SY_CHAOS) acc *= (real)(n > 1 ? 0.5 - 0.5 * cospi(2.0 * (double)j / (double)(n - 1)) : 1.0);   // hann after the smoothing
        out[j] = acc;
    }
}
