// ms_stage_api.inl -- C-ABI entry points of the non-spectral stages (include/microsound_b200.h).

struct SynthNormalK { static constexpr int MAXT = SY_NTHR;
    static MS_DEV void run(const SynthEvt* e, real* pool, const Ctx& c) { synth_normal_body(e, pool, c); } };
struct SynthTiltK { static constexpr int MAXT = 256;
    static MS_DEV void run(const SynthEvt* e, real* pool, const Ctx& c) { synth_tilt_finish_body(e, pool, c); } };
struct SynthDustK { static constexpr int MAXT = 256;
    static MS_DEV void run(const SynthEvt* e, const int* dp, const real* dv, real* pool, const Ctx& c) { synth_dust_body(e, dp, dv, pool, c); } };
struct OlaK { static constexpr int MAXT = OLA_NTHR;
    static MS_DEV void run(const OlaRender* r, const OlaEvt* e, const real* pool, real* mono, const Ctx& c) { ola_adsr_body(r, e, pool, mono, c); } };
struct FirBuildK { static constexpr int MAXT = OLA_NTHR;
    static MS_DEV void run(const FirRender* r, const int* to, const real* tg, const real* ir, real* h, const Ctx& c) { fir_build_body(r, to, tg, ir, h, c); } };
struct PostMaxK { static constexpr int MAXT = OLA_NTHR;
    static MS_DEV void run(const PostRender* r, real* mono, unsigned long long* mb, const Ctx& c) { post_max_body(r, mono, mb, c); } };
struct PostWriteK { static constexpr int MAXT = OLA_NTHR;
    static MS_DEV void run(const PostRender* r, const real* mono, const unsigned long long* mb, float2* out, const Ctx& c) { post_write_body(r, mono, mb, out, c); } };
struct RollK { static constexpr int MAXT = 256;
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
    MS_FOR_Y_CHUNKS(n, { if (ms_launch<SynthDustK>(mk_dim(64, (unsigned)_yc), 256, 0, (ms_stream_t)stream, evts + _y0, (const int*)dpos, dval, pool)) return -1; })
    return 0;
}
extern "C" int MS_API(ms_overlap_add)(const ms_ola_render* renders, int n_renders, int max_out_n, const ms_ola_evt* evts,
                              const real* pool, real* mono, void* stream) {
    const unsigned gx = (unsigned)((max_out_n + OLA_TILE - 1) / OLA_TILE);
    MS_FOR_Y_CHUNKS(n_renders, { if (ms_launch<OlaK>(mk_dim(gx, (unsigned)_yc), OLA_NTHR, 0, (ms_stream_t)stream, renders + _y0, evts, pool, mono)) return -1; })
    return 0;
}
extern "C" int MS_API(ms_fir_build)(const ms_fir_render* renders, int n_renders, int max_h_len, const int32_t* tap_off,
                            const real* tap_gain, const real* irpool, real* hpool, void* stream) {
    const unsigned gx = (unsigned)((max_h_len + OLA_TILE - 1) / OLA_TILE);
    MS_FOR_Y_CHUNKS(n_renders, { if (ms_launch<FirBuildK>(mk_dim(gx, (unsigned)_yc), OLA_NTHR, FIR_MAX_TAPS * 8, (ms_stream_t)stream,
                                   renders + _y0, (const int*)tap_off, tap_gain, irpool, hpool)) return -1; })
    return 0;
}
extern "C" int MS_API(ms_post)(const ms_post_render* renders, int n_renders, int max_n, real* mono, uint64_t* maxbits,
                       float* out, void* stream) {
    const unsigned gx = (unsigned)((max_n + OLA_TILE - 1) / OLA_TILE);
    if (ms_memset(maxbits, 0, sizeof(uint64_t) * (size_t)n_renders, (ms_stream_t)stream)) return -1;
    MS_FOR_Y_CHUNKS(n_renders, { if (ms_launch<PostMaxK>(mk_dim(gx, (unsigned)_yc), OLA_NTHR, OLA_NTHR * sizeof(real), (ms_stream_t)stream,
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
struct FirGroup { size_t h0, h1, c0, c1; };
struct FirPlan { std::vector<FftJob> hjobs, cjobs; FftJob *hjobs_dev, *cjobs_dev; std::vector<FirGroup> groups; };
struct FirLayout { size_t hjobs_off, cjobs_off, hspec_off, work_off, total; std::vector<size_t> hspec_at; std::vector<int> B; int ncj; };

static int fir_layout(const ms_fir_render* r, int n, FirLayout& L) {
    L.hspec_at.resize(n); L.B.resize(n);
    size_t hs = 0, wk = 0; int ncj = 0;
    for (int i = 0; i < n; ++i) {
        const int B = ols_block_len(r[i].h_len);
        if (r[i].h_len < 1 || r[i].h_len > B / 2) MS_FAIL("fir: %d taps unsupported", r[i].h_len);
        L.B[i] = B; L.hspec_at[i] = hs; hs += (size_t)B;
        const int hop = B - r[i].h_len + 1;
        const int nblk = (r[i].out_n + hop - 1) / hop;
        const int nj = (nblk + 1) / 2;
        ncj += nj;
        if (B > MS_SMALL_MAX) wk += (size_t)B * (size_t)nj;
    }
    L.ncj = ncj;
    L.hjobs_off = 0;
    L.cjobs_off = ms_align256(sizeof(FftJob) * (size_t)n);
    L.hspec_off = L.cjobs_off + ms_align256(sizeof(FftJob) * (size_t)ncj);
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
extern "C" int MS_API(ms_fir_create)(const ms_fir_render* r, int n, const real* hpool, const real* mono_in, real* mono_out,
                             void* ws, size_t ws_bytes, void* stream, void** handle) {
    *handle = nullptr;
    ms_stream_t st = (ms_stream_t)stream;
    FirPlan* P = new FirPlan();
    P->hjobs_dev = P->cjobs_dev = nullptr;
    if (n <= 0) { *handle = P; return 0; }
    FirLayout L;
    if (fir_layout(r, n, L)) { delete P; return -1; }
    if (ws_bytes < L.total) { delete P; MS_FAIL("ms_fir_create: workspace %zu < required %zu", ws_bytes, L.total); }
    char* base = (char*)ws;
    cpx* hspec = (cpx*)(base + L.hspec_off);
    cpx* work = (cpx*)(base + L.work_off);
    size_t wk = 0;
    // renders in order of their block length so both job lists come out sorted by launch class and aligned
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return (L.B[a] > MS_SMALL_MAX) < (L.B[b] > MS_SMALL_MAX); });
    size_t acc = 0;
    FirGroup cur; cur.h0 = cur.c0 = 0;
    for (int oi = 0; oi < n; ++oi) {
        const int i = order[oi];
        {
            const int hop_i = L.B[i] - r[i].h_len + 1;
            const int nj_i = ((r[i].out_n + hop_i - 1) / hop_i + 1) / 2;
            const size_t bytes = sizeof(cpx) * (size_t)L.B[i] * (size_t)(1 + nj_i) + 2 * sizeof(real) * (size_t)r[i].out_n;
            if (acc && (acc + bytes > MS_L2_GROUP_BYTES || (oi > 0 && (L.B[order[oi - 1]] > MS_SMALL_MAX) != (L.B[i] > MS_SMALL_MAX)))) {
                cur.h1 = P->hjobs.size(); cur.c1 = P->cjobs.size();
                P->groups.push_back(cur);
                cur.h0 = cur.h1; cur.c0 = cur.c1; acc = 0;
            }
            acc += bytes;
        }
        FftJob H; memset(&H, 0, sizeof H);
        if (FftEngine::get().prepare(H, L.B[i], st)) { delete P; return -1; }
        FftJob Cj = H;
        H.n = r[i].h_len; H.in_a = hpool + r[i].h; H.work = hspec + L.hspec_at[i]; H.out_scale = (real)1.0 / (real)L.B[i];
        P->hjobs.push_back(H);
        const int hop = L.B[i] - r[i].h_len + 1;
        const int nblk = (r[i].out_n + hop - 1) / hop;
        const int nj = (nblk + 1) / 2;
        for (int j = 0; j < nj; ++j) {
            FftJob J = Cj;
            const int ba = j, bb = j + nj;                 // pair block j with block j + nj of the same render
            J.in_a = mono_in + r[i].x; J.out_a = mono_out + r[i].y;
            J.p0_a = (long long)ba * hop - (r[i].h_len - 1);
            if (bb < nblk) { J.in_b = J.in_a; J.out_b = J.out_a; J.p0_b = (long long)bb * hop - (r[i].h_len - 1); }
            J.ols_n = r[i].out_n; J.ols_skip = r[i].h_len - 1;
            J.live_lo = r[i].x_begin;
            J.live_hi = (int)std::min<long long>(r[i].out_n, (long long)r[i].x_end + r[i].h_len - 1);
            J.bspec = hspec + L.hspec_at[i];
            if (L.B[i] > MS_SMALL_MAX) { J.work = work + wk; wk += (size_t)L.B[i]; }
            P->cjobs.push_back(J);
        }
    }
    cur.h1 = P->hjobs.size(); cur.c1 = P->cjobs.size();
    P->groups.push_back(cur);
    P->hjobs_dev = (FftJob*)(base + L.hjobs_off);
    P->cjobs_dev = (FftJob*)(base + L.cjobs_off);
    if (ms_h2d(P->hjobs_dev, P->hjobs.data(), sizeof(FftJob) * P->hjobs.size(), st)) { delete P; return -1; }
    if (ms_h2d(P->cjobs_dev, P->cjobs.data(), sizeof(FftJob) * P->cjobs.size(), st)) { delete P; return -1; }
    *handle = P;
    return 0;
}
extern "C" int MS_API(ms_fir_run)(void* handle, void* stream) {
    FirPlan* P = (FirPlan*)handle;
    if (!P) MS_FAIL("ms_fir_run: null handle");
    if (P->hjobs.empty()) return 0;
    // filter spectrum and overlap-save of one group back to back: the per-render spectrum (B complex) and the
    // two-pass scratch are consumed out of L2
    for (const FirGroup& g : P->groups) {
        if (FftEngine::get().filter_spectrum(P->hjobs, P->hjobs_dev, (ms_stream_t)stream, g.h0, g.h1)) return -1;
        if (FftEngine::get().overlap_save(P->cjobs, P->cjobs_dev, (ms_stream_t)stream, g.c0, g.c1)) return -1;
    }
    return 0;
}
extern "C" void MS_API(ms_fir_destroy)(void* handle) { delete (FirPlan*)handle; }
