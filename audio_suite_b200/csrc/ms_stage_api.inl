// ms_stage_api.inl -- C-ABI entry points of the non-spectral stages (include/microsound_b200.h).

struct SynthNormalK { static constexpr int MAXT = SY_NTHR;
    static constexpr int MINB = 4;
    static MS_DEV void run(const SynthEvt* e, real* pool, const Ctx& c) { synth_normal_body(e, pool, c); } };
struct SynthTiltK { static constexpr int MAXT = 256;
    static constexpr int MINB = 1;
    static MS_DEV void run(const SynthEvt* e, real* pool, const Ctx& c) { synth_tilt_finish_body(e, pool, c); } };
struct SynthDustK { static constexpr int MAXT = 256;
    static constexpr int MINB = 1;
    static MS_DEV void run(const SynthEvt* e, const int* dp, const real* dv, real* pool, const Ctx& c) { synth_dust_body(e, dp, dv, pool, c); } };
struct SynthTableK { static constexpr int MAXT = 256;
    static constexpr int MINB = 1;
    static MS_DEV void run(const SynthEvt* e, const real* dv, real* pool, const Ctx& c) { synth_table_body(e, dv, pool, c); } };
struct SynthWaveletK { static constexpr int MAXT = 256;
    static constexpr int MINB = 1;
    static MS_DEV void run(const SynthEvt* e, const WaveletAtom* a, const int* sh, real* pool, const Ctx& c) { synth_wavelet_body(e, a, sh, pool, c); } };
struct FeedbackK { static constexpr int MAXT = 256;
    static constexpr int MINB = 1;
    static MS_DEV void run(const FeedbackEvt* e, real* pool, const Ctx& c) { feedback_body(e, pool, c); } };
struct ImprintStepK { static constexpr int MAXT = 256;
    static constexpr int MINB = 1;
    static MS_DEV void run(const ImprintStepEvt* e, cpx* z, real* mem, const int* pb, int mb, const Ctx& c) { imprint_step_body(e, z, mem, pb, mb, c); } };
struct ImprintCommitK { static constexpr int MAXT = 256;
    static constexpr int MINB = 1;
    static MS_DEV void run(const ImprintStepEvt* e, int n, int* pb, const Ctx& c) { imprint_commit_body(e, n, pb, c); } };
struct WaveguideK { static constexpr int MAXT = 256;
    static constexpr int MINB = 1;
    static MS_DEV void run(const WgEvt* e, const WgLine* l, real* pool, const Ctx& c) { waveguide_body(e, l, pool, c); } };
struct ResonatorK { static constexpr int MAXT = 256;
    static constexpr int MINB = 1;
    static MS_DEV void run(const ResEvt* e, const ResMode* m, real* pool, const Ctx& c) { resonator_body(e, m, pool, c); } };
struct PartialLockK { static constexpr int MAXT = PLOCK_NTHR;
    static constexpr int MINB = 1;
    static MS_DEV void run(const PlockEvt* e, cpx* z, real* scratch, const Ctx& c) { partial_lock_body(e, z, scratch, c); } };
template <int STEP> struct CepstralK { static constexpr int MAXT = 256;
    static constexpr int MINB = 1;
    static MS_DEV void run(const CepEvt* e, cpx* z1, cpx* z2, const cpx* z3, real* scratch, const Ctx& c) { cepstral_body<STEP>(e, z1, z2, z3, scratch, c); } };
struct ImprintK { static constexpr int MAXT = 256;
    static constexpr int MINB = 1;
    static MS_DEV void run(const ImprintEvt* e, const ImprintRender* r, cpx* z, const Ctx& c) { imprint_body(e, r, z, c); } };
#ifndef MS_OLA_MINB
#define MS_OLA_MINB 1
#endif
struct OlaK { static constexpr int MAXT = OLA_NTHR;
    static constexpr int MINB = MS_OLA_MINB;
    static MS_DEV void run(const OlaRender* r, const OlaEvt* e, const real* pool, const real* env, real* mono, const Ctx& c) { ola_adsr_body(r, e, pool, env, mono, c); } };
struct AdsrTableK { static constexpr int MAXT = OLA_NTHR;
    static constexpr int MINB = 1;
    static MS_DEV void run(const OlaRender* r, real* env, const Ctx& c) { adsr_table_body(r, env, c); } };
struct ErScatterK { static constexpr int MAXT = 256;
    static constexpr int MINB = 1;
    static MS_DEV void run(const ErJob* j, const int* to, const real* tg, real* e, const Ctx& c) { er_scatter_body(j, to, tg, e, c); } };
#ifndef MS_POSTMAX_MINB
#define MS_POSTMAX_MINB 6          // <= 42 registers: six 256-thread CTAs per SM (measured 3.7 -> 3.0 ms on the C5 sweep)
#endif
struct PostMaxK { static constexpr int MAXT = OLA_NTHR;
    static constexpr int MINB = MS_POSTMAX_MINB;
    static MS_DEV void run(const PostRender* r, real* mono, unsigned long long* mb, const Ctx& c) { post_max_body(r, mono, mb, c); } };
struct PostWriteK { static constexpr int MAXT = OLA_NTHR;
    static constexpr int MINB = 1;
    static MS_DEV void run(const PostRender* r, const real* mono, const unsigned long long* mb, float2* out, const Ctx& c) { post_write_body(r, mono, mb, out, c); } };
struct RollK { static constexpr int MAXT = 256;
    static constexpr int MINB = 1;
    static MS_DEV void run(const real* src, real* dst, int n, int shift, const Ctx& c) { roll_body(src, dst, n, shift, c); } };

static inline MsDim mk_dim(unsigned x, unsigned y) { MsDim d; d.x = x; d.y = y; return d; }
#define MS_FOR_Y_CHUNKS(total, body) for (int _y0 = 0; _y0 < (total); _y0 += 32768) { const int _yc = std::min(32768, (total) - _y0); body }

extern "C" int MS_API(ms_synth_normal)(const ms_synth_evt* evts, int n, real* pool, void* stream) {
    for (int x0 = 0; x0 < n; x0 += 1 << 20) {
        const int cnt = std::min(1 << 20, n - x0);
        if (ms_launch<SynthNormalK>(mk_dim((unsigned)cnt, 1), SY_NTHR, sizeof(SynthSmem), (ms_stream_t)stream, evts + x0, pool)) return -1;
    }
    return 0;
}
extern "C" int MS_API(ms_synth_tilt_finish)(const ms_synth_evt* evts, int n, real* pool, void* stream) {
    MS_FOR_Y_CHUNKS(n, { if (ms_launch<SynthTiltK>(mk_dim(64, (unsigned)_yc), 256, 0, (ms_stream_t)stream, evts + _y0, pool)) return -1; })
    return 0;
}
extern "C" int MS_API(ms_synth_dust)(const ms_synth_evt* evts, int n, const int32_t* dpos, const real* dval, real* pool, void* stream) {
    MS_FOR_Y_CHUNKS(n, { if (ms_launch<SynthDustK>(mk_dim(DUST_CTAS, (unsigned)_yc), 256, DUST_KER_MAX * sizeof(real) + DUST_STAGE_MAX * (sizeof(real) + sizeof(int)), (ms_stream_t)stream, evts + _y0, (const int*)dpos, dval, pool)) return -1; })
    return 0;
}
extern "C" int MS_API(ms_adsr_tables)(const ms_ola_render* reps, int n_tables, int max_out_n, real* envpool, void* stream) {
    const unsigned gx = (unsigned)((max_out_n + OLA_TILE - 1) / OLA_TILE);
    MS_FOR_Y_CHUNKS(n_tables, { if (ms_launch<AdsrTableK>(mk_dim(gx, (unsigned)_yc), OLA_NTHR, 0, (ms_stream_t)stream, reps + _y0, envpool)) return -1; })
    return 0;
}
extern "C" int MS_API(ms_synth_table)(const ms_synth_evt* evts, int n, const real* dval, real* pool, void* stream) {
    for (int x0 = 0; x0 < n; x0 += 1 << 20) {
        const int cnt = std::min(1 << 20, n - x0);
        if (ms_launch<SynthTableK>(mk_dim((unsigned)cnt, 1), 256, 256 * sizeof(real), (ms_stream_t)stream, evts + x0, dval, pool)) return -1;
    }
    return 0;
}
extern "C" int MS_API(ms_synth_wavelet)(const ms_synth_evt* evts, int n, const ms_wavelet_atom* atoms, const int32_t* shifts,
                                real* pool, void* stream) {
    MS_FOR_Y_CHUNKS(n, { if (ms_launch<SynthWaveletK>(mk_dim(64, (unsigned)_yc), 256, 0, (ms_stream_t)stream, evts + _y0, atoms, (const int*)shifts, pool)) return -1; })
    return 0;
}
extern "C" int MS_API(ms_feedback)(const ms_feedback_evt* evts, int n, int max_n, real* pool, void* stream) {
    const unsigned gx = (unsigned)((max_n + 255) / 256);
    MS_FOR_Y_CHUNKS(n, { if (ms_launch<FeedbackK>(mk_dim(gx, (unsigned)_yc), 256, 0, (ms_stream_t)stream, evts + _y0, pool)) return -1; })
    return 0;
}
extern "C" int MS_API(ms_imprint_step)(const ms_imprint_step_evt* evts, int n, int max_bins, real* z_base, real* mem, int32_t* prev_bins,
                               void* stream) {
    const unsigned gx = (unsigned)((max_bins + 255) / 256);
    MS_FOR_Y_CHUNKS(n, { if (ms_launch<ImprintStepK>(mk_dim(gx, (unsigned)_yc), 256, 0, (ms_stream_t)stream, evts + _y0, (cpx*)z_base, mem,
                                                     (const int*)prev_bins, max_bins)) return -1; })
    return ms_launch<ImprintCommitK>(mk_dim(64, 1), 256, 0, (ms_stream_t)stream, evts, n, (int*)prev_bins);
}
extern "C" int MS_API(ms_waveguide)(const ms_wg_evt* evts, int n, const ms_wg_line* lines, real* pool, void* stream) {
    for (int x0 = 0; x0 < n; x0 += 1 << 20) {
        const int cnt = std::min(1 << 20, n - x0);
        if (ms_launch<WaveguideK>(mk_dim((unsigned)cnt, 1), 256, 0, (ms_stream_t)stream, evts + x0, lines, pool)) return -1;
    }
    return 0;
}
extern "C" int MS_API(ms_resonator)(const ms_res_evt* evts, int n, const ms_res_mode* modes, real* pool, void* stream) {
    for (int x0 = 0; x0 < n; x0 += 1 << 20) {
        const int cnt = std::min(1 << 20, n - x0);
        if (ms_launch<ResonatorK>(mk_dim((unsigned)cnt, 1), 256, 256 * sizeof(real), (ms_stream_t)stream, evts + x0, modes, pool)) return -1;
    }
    return 0;
}
extern "C" int MS_API(ms_partial_lock)(const ms_plock_evt* evts, int n, real* z_base, real* scratch, void* stream) {
    for (int x0 = 0; x0 < n; x0 += 1 << 20) {
        const int cnt = std::min(1 << 20, n - x0);
        if (ms_launch<PartialLockK>(mk_dim((unsigned)cnt, 1), PLOCK_NTHR, PLOCK_NTHR * sizeof(int), (ms_stream_t)stream,
                                    evts + x0, (cpx*)z_base, scratch)) return -1;
    }
    return 0;
}
extern "C" int MS_API(ms_cepstral)(int step, const ms_cep_evt* evts, int n, int max_n, real* z1, real* z2, real* z3, real* scratch,
                           void* stream) {
    const unsigned gx = (unsigned)(((step == 1 ? max_n : max_n / 2 + 1) + 255) / 256);
    MS_FOR_Y_CHUNKS(n, {
        int rc;
        if (step == 0) rc = ms_launch<CepstralK<0>>(mk_dim(gx, (unsigned)_yc), 256, 0, (ms_stream_t)stream, evts + _y0, (cpx*)z1, (cpx*)z2, (const cpx*)z3, scratch);
        else if (step == 1) rc = ms_launch<CepstralK<1>>(mk_dim(gx, (unsigned)_yc), 256, 0, (ms_stream_t)stream, evts + _y0, (cpx*)z1, (cpx*)z2, (const cpx*)z3, scratch);
        else rc = ms_launch<CepstralK<2>>(mk_dim(gx, (unsigned)_yc), 256, 0, (ms_stream_t)stream, evts + _y0, (cpx*)z1, (cpx*)z2, (const cpx*)z3, scratch);
        if (rc) return -1;
    })
    return 0;
}
extern "C" int MS_API(ms_imprint)(const ms_imprint_evt* evts, const ms_imprint_render* renders, int n_renders, int max_bins,
                          real* z_base, void* stream) {
    const unsigned gx = (unsigned)((max_bins + 255) / 256);
    MS_FOR_Y_CHUNKS(n_renders, { if (ms_launch<ImprintK>(mk_dim(gx, (unsigned)_yc), 256, 0, (ms_stream_t)stream, evts, renders + _y0, (cpx*)z_base)) return -1; })
    return 0;
}
extern "C" int MS_API(ms_overlap_add)(const ms_ola_render* renders, int n_renders, int max_out_n, const ms_ola_evt* evts,
                              const real* pool, const real* envpool, real* mono, void* stream) {
    const unsigned gx = (unsigned)((max_out_n + OLA_TILE - 1) / OLA_TILE);
    MS_FOR_Y_CHUNKS(n_renders, { if (ms_launch<OlaK>(mk_dim(gx, (unsigned)_yc), OLA_NTHR, 0, (ms_stream_t)stream, renders + _y0, evts, pool, envpool, mono)) return -1; })
    return 0;
}
extern "C" int MS_API(ms_post)(const ms_post_render* renders, int n_renders, int max_n, real* mono, uint64_t* maxbits,
                       float* out, void* stream) {
    const unsigned gx = (unsigned)((max_n + OLA_TILE - 1) / OLA_TILE);
    if (ms_memset(maxbits, 0, sizeof(uint64_t) * (size_t)n_renders, (ms_stream_t)stream)) return -1;
    MS_FOR_Y_CHUNKS(n_renders, { if (ms_launch<PostMaxK>(mk_dim(gx, (unsigned)_yc), OLA_NTHR, (2 * POST_PAR + POST_NC + 1 + OLA_NTHR + OLA_TILE + OLA_TILE / 8 + 8) * sizeof(real), (ms_stream_t)stream,
                                   renders + _y0, mono, (unsigned long long*)maxbits + _y0)) return -1; })
    MS_FOR_Y_CHUNKS(n_renders, { if (ms_launch<PostWriteK>(mk_dim(gx, (unsigned)_yc), OLA_NTHR, 0, (ms_stream_t)stream,
                                   renders + _y0, (const real*)mono, (const unsigned long long*)maxbits + _y0, (float2*)out)) return -1; })
    return 0;
}
extern "C" int MS_API(ms_roll)(const real* src, real* dst, int n, int shift, void* stream) {
    return ms_launch<RollK>(mk_dim(64, 1), 256, 0, (ms_stream_t)stream, src, dst, n, shift);
}

// ---- FIR by overlap-save ----------------------------------------------------------------------------------
static int ols_block_len(int h_len) {
    if (h_len <= 3072) return 8192;
    int B = 32768;
    while (B < 4 * h_len && B < (1 << 20)) B <<= 1;
    return B;
}
// Plan: (1) at create time, the spectrum of every distinct impulse response (FFT_B(ir)/B, scrambled [k1][k2]
// layout); (2) per run and per render with reflection taps: scatter the taps into a dense vector, transform it
// and compose the render's filter spectrum IRspec * (1 + FFT(e)); (3) overlap-save with two blocks per transform.
struct FirPlan {
    std::vector<FftJob> ijobs, hjobs, cjobs;
    FftJob *ijobs_dev, *hjobs_dev, *cjobs_dev;
    std::vector<ErJob> ejobs; ErJob* ejobs_dev;
    const int* tap_off; const real* tap_gain; real* ebase;
};
struct FirLayout {
    size_t ijobs_off, hjobs_off, cjobs_off, ejobs_off, ebuf_off, ispec_off, hspec_off, work_off, total;
    std::vector<size_t> hspec_at, e_at; std::vector<int> B, ispec_of; std::vector<std::pair<long long, std::pair<int, int>>> irs;
    int ncj, nh;
};

static int fir_layout(const ms_fir_render* r, int n, FirLayout& L) {
    L.hspec_at.assign(n, 0); L.e_at.assign(n, 0); L.B.resize(n); L.ispec_of.resize(n); L.irs.clear();
    size_t hs = 0, wk = 0, is = 0, eb = 0; int ncj = 0, nh = 0;
    for (int i = 0; i < n; ++i) {
        const int B = ols_block_len(r[i].h_len);
        if (r[i].h_len < 1 || r[i].h_len > B / 2 || r[i].ir_len < 1) MS_FAIL("fir: %d taps unsupported", r[i].h_len);
        L.B[i] = B;
        const std::pair<long long, std::pair<int, int>> key(r[i].ir, std::make_pair(r[i].ir_len, B));
        int found = -1;
        for (size_t k = 0; k < L.irs.size(); ++k) if (L.irs[k] == key) { found = (int)k; break; }
        if (found < 0) { found = (int)L.irs.size(); L.irs.push_back(key); is += (size_t)B; }
        L.ispec_of[i] = found;
        if (r[i].tap_end > r[i].tap_begin) {
            L.hspec_at[i] = hs; hs += (size_t)B;
            L.e_at[i] = eb; eb += (size_t)(r[i].h_len - r[i].ir_len + 1);
            ++nh;
        }
        const int hop = B - r[i].h_len + 1;
        const int nblk = (r[i].out_n + hop - 1) / hop;
        const int nj = (nblk + 1) / 2;
        ncj += nj;
        if (B > MS_SMALL_MAX) wk += (size_t)B * (size_t)nj;
    }
    L.ncj = ncj; L.nh = nh;
    L.ijobs_off = 0;
    L.hjobs_off = ms_align256(sizeof(FftJob) * L.irs.size());
    L.cjobs_off = L.hjobs_off + ms_align256(sizeof(FftJob) * (size_t)nh);
    L.ejobs_off = L.cjobs_off + ms_align256(sizeof(FftJob) * (size_t)ncj);
    L.ebuf_off = L.ejobs_off + ms_align256(sizeof(ErJob) * (size_t)nh);
    L.ispec_off = L.ebuf_off + ms_align256(sizeof(real) * eb);
    L.hspec_off = L.ispec_off + ms_align256(sizeof(cpx) * is);
    L.work_off = L.hspec_off + ms_align256(sizeof(cpx) * hs);
    L.total = L.work_off + ms_align256(sizeof(cpx) * wk);
    return 0;
}
extern "C" size_t MS_API(ms_fir_workspace_bytes)(const ms_fir_render* r, int n) {
    FirLayout L;
    if (n <= 0) return 256;
    if (fir_layout(r, n, L)) return 0;
    return L.total;
}
extern "C" int MS_API(ms_fir_create)(const ms_fir_render* r, int n, const real* irpool, const int32_t* tap_off, const real* tap_gain,
                             const real* mono_in, real* mono_out, void* ws, size_t ws_bytes, void* stream, void** handle) {
    *handle = nullptr;
    ms_stream_t st = (ms_stream_t)stream;
    FirPlan* P = new FirPlan();
    P->ijobs_dev = P->hjobs_dev = P->cjobs_dev = nullptr; P->ejobs_dev = nullptr;
    if (n <= 0) { *handle = P; return 0; }
    FirLayout L;
    if (fir_layout(r, n, L)) { delete P; return -1; }
    if (ws_bytes < L.total) { delete P; MS_FAIL("ms_fir_create: workspace %zu < required %zu", ws_bytes, L.total); }
    char* base = (char*)ws;
    cpx* ispec = (cpx*)(base + L.ispec_off);
    cpx* hspec = (cpx*)(base + L.hspec_off);
    cpx* work = (cpx*)(base + L.work_off);
    P->ebase = (real*)(base + L.ebuf_off);
    P->tap_off = (const int*)tap_off; P->tap_gain = tap_gain;
    // (1) impulse-response spectra
    std::vector<size_t> ispec_at(L.irs.size());
    size_t is = 0;
    for (size_t k = 0; k < L.irs.size(); ++k) {
        const int B = L.irs[k].second.second;
        FftJob I; memset(&I, 0, sizeof I);
        if (FftEngine::get().prepare(I, B, st)) { delete P; return -1; }
        I.n = L.irs[k].second.first; I.in_a = irpool + L.irs[k].first; I.work = ispec + is; I.out_scale = (real)1.0 / (real)B;
        ispec_at[k] = is; is += (size_t)B;
        P->ijobs.push_back(I);
    }
    // (2)+(3) per render, in order of launch class
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return (L.B[a] > MS_SMALL_MAX) < (L.B[b] > MS_SMALL_MAX); });
    size_t wk = 0;
    for (int oi = 0; oi < n; ++oi) {
        const int i = order[oi];
        FftJob G; memset(&G, 0, sizeof G);
        if (FftEngine::get().prepare(G, L.B[i], st)) { delete P; return -1; }
        const cpx* filt = ispec + ispec_at[L.ispec_of[i]];
        if (r[i].tap_end > r[i].tap_begin) {
            FftJob H = G;
            ErJob E; E.e = (long long)L.e_at[i]; E.elen = r[i].h_len - r[i].ir_len + 1; E.tap_begin = r[i].tap_begin; E.tap_end = r[i].tap_end;
            P->ejobs.push_back(E);
            H.n = E.elen; H.in_a = P->ebase + E.e; H.cin = filt; H.work = hspec + L.hspec_at[i]; H.out_scale = (real)1.0;
            P->hjobs.push_back(H);
            filt = hspec + L.hspec_at[i];
        }
        const int hop = L.B[i] - r[i].h_len + 1;
        const int nblk = (r[i].out_n + hop - 1) / hop;
        const int nj = (nblk + 1) / 2;
        for (int j = 0; j < nj; ++j) {
            FftJob J = G;
            const int ba = j, bb = j + nj;                 // pair block j with block j + nj of the same render
            J.in_a = mono_in + r[i].x; J.out_a = mono_out + r[i].y;
            J.p0_a = (long long)ba * hop - (r[i].h_len - 1);
            if (bb < nblk) { J.in_b = J.in_a; J.out_b = J.out_a; J.p0_b = (long long)bb * hop - (r[i].h_len - 1); }
            J.ols_n = r[i].out_n; J.ols_skip = r[i].h_len - 1;
            J.live_lo = r[i].x_begin;
            J.live_hi = (int)std::min<long long>(r[i].out_n, (long long)r[i].x_end + r[i].h_len - 1);
            J.bspec = filt;
            if (L.B[i] > MS_SMALL_MAX) { J.work = work + wk; wk += (size_t)L.B[i]; }
            P->cjobs.push_back(J);
        }
    }
    auto by_class = [](const FftJob& a, const FftJob& b) { return FftEngine::job_class(a) < FftEngine::job_class(b); };
    std::stable_sort(P->ijobs.begin(), P->ijobs.end(), by_class);
    P->ijobs_dev = (FftJob*)(base + L.ijobs_off);
    P->hjobs_dev = (FftJob*)(base + L.hjobs_off);
    P->cjobs_dev = (FftJob*)(base + L.cjobs_off);
    P->ejobs_dev = (ErJob*)(base + L.ejobs_off);
    if (ms_h2d(P->ijobs_dev, P->ijobs.data(), sizeof(FftJob) * P->ijobs.size(), st)) { delete P; return -1; }
    if (!P->hjobs.empty()) {
        if (ms_h2d(P->hjobs_dev, P->hjobs.data(), sizeof(FftJob) * P->hjobs.size(), st)) { delete P; return -1; }
        if (ms_h2d(P->ejobs_dev, P->ejobs.data(), sizeof(ErJob) * P->ejobs.size(), st)) { delete P; return -1; }
    }
    if (ms_h2d(P->cjobs_dev, P->cjobs.data(), sizeof(FftJob) * P->cjobs.size(), st)) { delete P; return -1; }
    if (FftEngine::get().filter_spectrum(P->ijobs, P->ijobs_dev, st)) { delete P; return -1; }      // once per plan
    *handle = P;
    return 0;
}
extern "C" int MS_API(ms_fir_run)(void* handle, void* stream) {
    FirPlan* P = (FirPlan*)handle;
    if (!P) MS_FAIL("ms_fir_run: null handle");
    if (P->cjobs.empty()) return 0;
    ms_stream_t st = (ms_stream_t)stream;
    if (!P->hjobs.empty()) {
        for (size_t y0 = 0; y0 < P->ejobs.size(); y0 += 32768) {
            const unsigned yc = (unsigned)std::min<size_t>(32768, P->ejobs.size() - y0);
            if (ms_launch<ErScatterK>(mk_dim(1, yc), 256, 0, st, (const ErJob*)(P->ejobs_dev + y0), P->tap_off, P->tap_gain, P->ebase)) return -1;
        }
        if (FftEngine::get().filter_compose(P->hjobs, P->hjobs_dev, st)) return -1;
    }
    return FftEngine::get().overlap_save(P->cjobs, P->cjobs_dev, st);
}
extern "C" void MS_API(ms_fir_destroy)(void* handle) { delete (FirPlan*)handle; }
