// ms_stage_api.inl -- C-ABI entry points of the non-spectral stages (include/microsound_b200.h).

struct SynthNormalK { static constexpr int MAXT = SY_NTHR;
    static constexpr int MINB = 4;
    static MS_DEV void run(const SynthEvt* e, real* pool, const Ctx& c) { synth_normal_body(e, pool, c); } };
// (latency-bound elementwise kernels gain from residency, measured on the C5 sweep: SynthTiltK 0.59 -> 0.39 ms at eight CTAs
//  per SM, SynthDustK 0.76 -> 0.54 at five; six and eight for the dust kernel were no better)
#ifndef MS_TILT_MINB
#define MS_TILT_MINB 8
#endif
#ifndef MS_DUST_MINB
#define MS_DUST_MINB 5
#endif
struct SynthTiltK { static constexpr int MAXT = 256;
    static constexpr int MINB = MS_TILT_MINB;
    static MS_DEV void run(const SynthEvt* e, real* pool, const Ctx& c) { synth_tilt_finish_body(e, pool, c); } };
struct SynthDustK { static constexpr int MAXT = 256;
    static constexpr int MINB = MS_DUST_MINB;
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
#define MS_OLA_MINB 8          // 32 registers, eight CTAs per SM: the kernel waits on dependent global loads (measured 1.54 -> 1.26 ms)
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

#ifndef MS_HOST_EMUL
template <int CL>
static int synth_launch_cluster(const ms_synth_evt* evts, int n, real* pool, ms_stream_t st) {
    auto kern = synth_normal_cluster_kernel<CL>;
    static bool configured = false;
    if (!configured) {
        MS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SynthSmem)));
        configured = true;
    }
    cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3((unsigned)(CL * n), 1, 1); cfg.blockDim = dim3(SY_NTHR, 1, 1); cfg.dynamicSmemBytes = sizeof(SynthSmem); cfg.stream = st;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    MS_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, (const SynthEvt*)evts, pool));
    ++ms_launch_counter();
    if (ms_launch_hook()) ms_launch_hook()(CL == 2 ? "K = synth_normal_cluster_kernel<2>;" : CL == 4 ? "K = synth_normal_cluster_kernel<4>;" : "K = synth_normal_cluster_kernel<8>;", (void*)st);
    return 0;
}
// CTAs per event: small batches spread every event over a thread-block cluster (development switch MS_SYNTH_CLUSTER=1|2|4|8
// forces a size; 1 = always one CTA per event)
static int synth_cluster_size(int n_events) {
    const char* e = getenv("MS_SYNTH_CLUSTER");            // (read per call: the tests switch it between launches)
    const int forced = e ? atoi(e) : 0;
    if (forced == 1 || forced == 2 || forced == 4 || forced == 8) return forced;
    static int slots = 0;                                   // CTAs of this kernel the device holds at once
    if (!slots) { int dev = 0, sms = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); slots = 4 * sms; }
    if (n_events * 8 <= slots) return 8;
    if (n_events * 4 <= 3 * slots) return 4;
    if (n_events * 2 <= 3 * slots) return 2;
    return 1;
}
#endif
extern "C" int MS_API(ms_synth_normal)(const ms_synth_evt* evts, int n, real* pool, void* stream) {
#ifndef MS_HOST_EMUL
    if (n > 0 && n < (1 << 20)) {
        const int cl = synth_cluster_size(n);
        if (cl == 8) return synth_launch_cluster<8>(evts, n, pool, (ms_stream_t)stream);
        if (cl == 4) return synth_launch_cluster<4>(evts, n, pool, (ms_stream_t)stream);
        if (cl == 2) return synth_launch_cluster<2>(evts, n, pool, (ms_stream_t)stream);
    }
#endif
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
        if (ms_launch<PartialLockK>(mk_dim((unsigned)cnt, 1), PLOCK_NTHR, (PLOCK_NTHR + 1 + 2 * PLOCK_SEL_MAX) * sizeof(int), (ms_stream_t)stream,
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
    const unsigned gx = (unsigned)((max_out_n + OLA_ATILE - 1) / OLA_ATILE);
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

// ---- EXTENSION: polyphase decimation (see ms_time.cuh; not on the reference's path) -----------------------------------
struct DecimateK { static constexpr int MAXT = DEC_NTHR; static constexpr int MINB = 1;
    static MS_DEV void run(const real* x, long long xs, int n, const real* h, int taps, int q, real* y, long long ys, int n_out, const Ctx& c) {
        decimate_body(x, xs, n, h, taps, q, y, ys, n_out, c); } };
extern "C" int MS_API(ms_polyphase_decimate)(const real* x, int64_t x_stride, int n, int n_signals, const real* h, int taps, int q,
                                             real* y, int64_t y_stride, void* stream) {
    if (taps < 1 || taps > DEC_TAPS_MAX || q < 1 || n < 0) MS_FAIL("ms_polyphase_decimate: unsupported taps %d / factor %d", taps, q);
    const int n_out = n + taps - 1 <= 0 ? 0 : (n + taps - 1 + q - 1) / q;
    if (n_out == 0 || n_signals <= 0) return 0;
    const size_t smem = sizeof(real) * ((size_t)taps + (size_t)(DEC_OUT - 1) * q + taps + DEC_NTHR);
    if (smem > 200 * 1024) MS_FAIL("ms_polyphase_decimate: window of %zu bytes does not fit shared memory (factor %d, %d taps)", smem, q, taps);
    MS_FOR_Y_CHUNKS(n_signals, { if (ms_launch<DecimateK>(mk_dim((unsigned)((n_out + DEC_OUT - 1) / DEC_OUT), (unsigned)_yc), DEC_NTHR, smem, (ms_stream_t)stream,
                                    x + (long long)_y0 * x_stride, (long long)x_stride, n, h, taps, q, y + (long long)_y0 * y_stride, (long long)y_stride, n_out)) return -1; })
    return 0;
}

// ---- FIR by overlap-save ----------------------------------------------------------------------------------
static int ols_block_len(int h_len, int out_n) {
    if (h_len <= 3072) return 8192;
    int B = 32768;
    while (B < 4 * h_len && B < (1 << 20)) B <<= 1;
    // 65536-point blocks have the fused kernel (ms_fir_fused.cuh) and waste less on overlap; 32768 stays only where
    // the whole output fits in one transform of two blocks
    if (B == 32768 && (out_n + (B - h_len)) / (B - h_len + 1) > 2) B = 65536;
    return B;
}
// How the 65536-point blocks run (ms_fir_fused.cuh):  0 = legacy five-kernel sequence through the spectral engine,
// 1 = fused phases, one launch per phase (scratch per unit), 4 / 8 / 16 = fused phases inside one persistent kernel of
// thread-block clusters of that many CTAs (scratch per cluster, L2-resident).  Development switch: MS_FIR_MODE.
static int fir_mode() {
#ifdef MS_HOST_EMUL
    const char* e = getenv("MS_FIR_MODE");
    return (e && atoi(e) == 0) ? 0 : 1;                     // the block emulator has no clusters
#else
    static int mode = -1;
    // default 1: measured on B200 (C5 sweep, f64) the one-launch-per-phase form wins -- 10.2 ms against 12.9 / 13.7 / 14.9 ms
    // for clusters of 4 / 8 / 16 (legacy 15.3 ms): the phases are latency-bound, not DRAM-bound, and a cluster keeps all
    // its CTAs in the same phase at three CTAs per SM, while separate launches run P1 / P3 at four
    if (mode < 0) { const char* e = getenv("MS_FIR_MODE"); mode = e ? atoi(e) : 1; if (mode != 0 && mode != 4 && mode != 8 && mode != 16) mode = 1; }
    return mode;
#endif
}
struct FirP1K { static constexpr int MAXT = FF_NTHR; static constexpr int MINB = 4;
    static MS_DEV void run(const FirUnit* u, FirTables T, cpx* s, const Ctx& c) { fir_phase_body(1, u, T, s, c); } };
struct FirP2K { static constexpr int MAXT = FF_NTHR; static constexpr int MINB = 3;
    static MS_DEV void run(const FirUnit* u, FirTables T, cpx* s, const Ctx& c) { fir_phase_body(2, u, T, s, c); } };
struct FirP3K { static constexpr int MAXT = FF_NTHR; static constexpr int MINB = 4;
    static MS_DEV void run(const FirUnit* u, FirTables T, cpx* s, const Ctx& c) { fir_phase_body(3, u, T, s, c); } };
struct FirSortK { static constexpr int MAXT = 256; static constexpr int MINB = 1;
    static MS_DEV void run(const FirSortJob* j, const int* to, const real* tg, int* rp, int* so, real* sg, const Ctx& c) { fir_sort_taps_body(j, to, tg, rp, so, sg, c); } };
static const size_t FF_SMEM1 = sizeof(cpx) * (size_t)FF_TILE * FF_RS;       // one transposing tile
static const size_t FF_SMEM2 = sizeof(cpx) * (size_t)(FF_TILE + FF_EROWS) * FF_RS;   // + the reflection-spectrum rows of phase 2

// Plan: (1) at create time, the spectrum of every distinct impulse response (FFT_B(ir)/B, [k1][k2] layout) and the
// reflection taps of the fused renders sorted by residue; (2) legacy renders, per run: scatter the taps into a dense
// vector, transform it and compose the render's filter spectrum IRspec * (1 + FFT(e)); (3) overlap-save with two blocks
// per transform: fused units (65536-point blocks) or legacy jobs.
struct FirPlan {
    std::vector<FftJob> ijobs, hjobs, cjobs;
    FftJob *ijobs_dev, *hjobs_dev, *cjobs_dev;
    std::vector<ErJob> ejobs; ErJob* ejobs_dev;
    const int* tap_off; const real* tap_gain; real* ebase;
    std::vector<FirUnit> units; FirUnit* units_dev;
    FirTables ft; cpx* scratch; int n_scratch;           // n_scratch: 65536-point scratch blocks available
};
struct FirLayout {
    size_t ijobs_off, hjobs_off, cjobs_off, ejobs_off, ebuf_off, ispec_off, hspec_off, work_off, total;
    size_t units_off, sort_off, res_off, soff_off, sgain_off, scratch_off;
    std::vector<size_t> hspec_at, e_at; std::vector<int> B, ispec_of, fused; std::vector<std::pair<long long, std::pair<int, int>>> irs;
    int ncj, nh, n_units, n_sort, n_scratch, n_taps;
};
static int fir_max_clusters(int cl) { return std::max(1, (148 * 3) / cl); }      // resident clusters at three CTAs per SM

static int fir_layout(const ms_fir_render* r, int n, FirLayout& L) {
    L.hspec_at.assign(n, 0); L.e_at.assign(n, 0); L.B.resize(n); L.ispec_of.resize(n); L.fused.assign(n, 0); L.irs.clear();
    size_t hs = 0, wk = 0, is = 0, eb = 0; int ncj = 0, nh = 0, nu = 0, ns = 0, nt = 0;
    const int mode = fir_mode();
    for (int i = 0; i < n; ++i) {
        const int B = ols_block_len(r[i].h_len, r[i].out_n);
        if (r[i].h_len < 1 || r[i].h_len > B / 2 || r[i].ir_len < 1) MS_FAIL("fir: %d taps unsupported", r[i].h_len);
        L.B[i] = B;
        L.fused[i] = (mode != 0 && B == FF_N * FF_N) ? 1 : 0;
        const std::pair<long long, std::pair<int, int>> key(r[i].ir, std::make_pair(r[i].ir_len, B));
        int found = -1;
        for (size_t k = 0; k < L.irs.size(); ++k) if (L.irs[k] == key) { found = (int)k; break; }
        if (found < 0) { found = (int)L.irs.size(); L.irs.push_back(key); is += (size_t)B; }
        L.ispec_of[i] = found;
        const bool has_taps = r[i].tap_end > r[i].tap_begin;
        nt = std::max(nt, r[i].tap_end);
        const int hop = B - r[i].h_len + 1;
        const int nblk = (r[i].out_n + hop - 1) / hop;
        const int nj = (nblk + 1) / 2;
        if (L.fused[i]) { nu += nj; if (has_taps) ++ns; continue; }
        if (has_taps) {
            L.hspec_at[i] = hs; hs += (size_t)B;
            L.e_at[i] = eb; eb += (size_t)(r[i].h_len - r[i].ir_len + 1);
            ++nh;
        }
        ncj += nj;
        if (B > MS_SMALL_MAX) wk += (size_t)B * (size_t)nj;
    }
    L.ncj = ncj; L.nh = nh; L.n_units = nu; L.n_sort = ns; L.n_taps = nt;
    L.n_scratch = nu == 0 ? 0 : (mode == 1 ? nu : std::min(nu, fir_max_clusters(mode)));
    L.ijobs_off = 0;
    L.hjobs_off = ms_align256(sizeof(FftJob) * L.irs.size());
    L.cjobs_off = L.hjobs_off + ms_align256(sizeof(FftJob) * (size_t)nh);
    L.ejobs_off = L.cjobs_off + ms_align256(sizeof(FftJob) * (size_t)ncj);
    L.ebuf_off = L.ejobs_off + ms_align256(sizeof(ErJob) * (size_t)nh);
    L.ispec_off = L.ebuf_off + ms_align256(sizeof(real) * eb);
    L.hspec_off = L.ispec_off + ms_align256(sizeof(cpx) * is);
    L.work_off = L.hspec_off + ms_align256(sizeof(cpx) * hs);
    L.units_off = L.work_off + ms_align256(sizeof(cpx) * wk);
    L.sort_off = L.units_off + ms_align256(sizeof(FirUnit) * (size_t)nu);
    L.res_off = L.sort_off + ms_align256(sizeof(FirSortJob) * (size_t)ns);
    L.soff_off = L.res_off + ms_align256(sizeof(int) * FF_RES_STRIDE * (size_t)ns);
    L.sgain_off = L.soff_off + ms_align256(sizeof(int) * (size_t)(ns ? nt : 0));
    L.scratch_off = L.sgain_off + ms_align256(sizeof(real) * (size_t)(ns ? nt : 0));
    L.total = L.scratch_off + ms_align256(sizeof(cpx) * (size_t)L.n_scratch * (FF_N * FF_N));
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
    P->ijobs_dev = P->hjobs_dev = P->cjobs_dev = nullptr; P->ejobs_dev = nullptr; P->units_dev = nullptr; P->scratch = nullptr; P->n_scratch = 0;
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
    std::vector<FirSortJob> sorts;
    for (int oi = 0; oi < n; ++oi) {
        const int i = order[oi];
        FftJob G; memset(&G, 0, sizeof G);
        if (FftEngine::get().prepare(G, L.B[i], st)) { delete P; return -1; }
        const cpx* filt = ispec + ispec_at[L.ispec_of[i]];
        const int hop = L.B[i] - r[i].h_len + 1;
        const int nblk = (r[i].out_n + hop - 1) / hop;
        const int nj = (nblk + 1) / 2;
        const int live_hi = (int)std::min<long long>(r[i].out_n, (long long)r[i].x_end + r[i].h_len - 1);
        if (L.fused[i]) {
            if (G.F1 != FF_N || G.F2 != FF_N) { delete P; MS_FAIL("fir: internal: 65536 is not planned as 256 x 256"); }
            P->ft.tw = G.tw1; P->ft.twM_hi = G.twM_hi; P->ft.twM_lo = G.twM_lo;
            int res_at = -1;
            if (r[i].tap_end > r[i].tap_begin) {
                FirSortJob sj; sj.tap_begin = r[i].tap_begin; sj.tap_end = r[i].tap_end; sj.res_at = FF_RES_STRIDE * (int)sorts.size(); sj._pad = 0;
                res_at = sj.res_at;
                sorts.push_back(sj);
            }
            for (int j = 0; j < nj; ++j) {
                FirUnit U; memset(&U, 0, sizeof U);
                const int bb = j + nj;                         // pair block j with block j + nj of the same render
                U.in = mono_in + r[i].x; U.out = mono_out + r[i].y; U.filt = filt;
                U.p0_a = (long long)j * hop - (r[i].h_len - 1);
                U.has_b = bb < nblk;
                U.p0_b = U.has_b ? (long long)bb * hop - (r[i].h_len - 1) : 0;
                U.ols_n = r[i].out_n; U.ols_skip = r[i].h_len - 1;
                U.live_lo = r[i].x_begin; U.live_hi = live_hi;
                U.tap_res = res_at;
                P->units.push_back(U);
            }
            continue;
        }
        if (r[i].tap_end > r[i].tap_begin) {
            FftJob H = G;
            ErJob E; E.e = (long long)L.e_at[i]; E.elen = r[i].h_len - r[i].ir_len + 1; E.tap_begin = r[i].tap_begin; E.tap_end = r[i].tap_end;
            P->ejobs.push_back(E);
            H.n = E.elen; H.in_a = P->ebase + E.e; H.cin = filt; H.work = hspec + L.hspec_at[i]; H.out_scale = (real)1.0;
            P->hjobs.push_back(H);
            filt = hspec + L.hspec_at[i];
        }
        for (int j = 0; j < nj; ++j) {
            FftJob J = G;
            const int ba = j, bb = j + nj;                 // pair block j with block j + nj of the same render
            J.in_a = mono_in + r[i].x; J.out_a = mono_out + r[i].y;
            J.p0_a = (long long)ba * hop - (r[i].h_len - 1);
            if (bb < nblk) { J.in_b = J.in_a; J.out_b = J.out_a; J.p0_b = (long long)bb * hop - (r[i].h_len - 1); }
            J.ols_n = r[i].out_n; J.ols_skip = r[i].h_len - 1;
            J.live_lo = r[i].x_begin;
            J.live_hi = live_hi;
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
    if (!P->cjobs.empty() && ms_h2d(P->cjobs_dev, P->cjobs.data(), sizeof(FftJob) * P->cjobs.size(), st)) { delete P; return -1; }
    if (FftEngine::get().filter_spectrum(P->ijobs, P->ijobs_dev, st)) { delete P; return -1; }      // once per plan
    if (!P->units.empty()) {
        P->units_dev = (FirUnit*)(base + L.units_off);
        P->scratch = (cpx*)(base + L.scratch_off);
        P->n_scratch = L.n_scratch;
        int* res_ptr = (int*)(base + L.res_off);
        int* soff = (int*)(base + L.soff_off);
        real* sgain = (real*)(base + L.sgain_off);
        P->ft.res_ptr = res_ptr; P->ft.tap_off = soff; P->ft.tap_gain = sgain;
        if (ms_h2d(P->units_dev, P->units.data(), sizeof(FirUnit) * P->units.size(), st)) { delete P; return -1; }
        if (!sorts.empty()) {                                  // residue-sorted copies of the taps: table building, once per plan
            FirSortJob* sj = (FirSortJob*)(base + L.sort_off);
            if (ms_h2d(sj, sorts.data(), sizeof(FirSortJob) * sorts.size(), st)) { delete P; return -1; }
            for (size_t y0 = 0; y0 < sorts.size(); y0 += 32768) {
                const unsigned yc = (unsigned)std::min<size_t>(32768, sorts.size() - y0);
                if (ms_launch<FirSortK>(mk_dim(1, yc), 256, (257 + 256) * sizeof(int), st, (const FirSortJob*)(sj + y0), P->tap_off, P->tap_gain,
                                        res_ptr, soff, sgain)) { delete P; return -1; }
            }
        }
    }
    *handle = P;
    return 0;
}
#ifndef MS_HOST_EMUL
template <int CL>
static int fir_launch_cluster(FirPlan* P, ms_stream_t st) {
    static int max_clusters = -1;
    auto kern = fir_cluster_kernel<CL>;
    if (max_clusters < 0) {
        MS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FF_SMEM2));
        if (CL > 8) MS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cudaLaunchConfig_t q; memset(&q, 0, sizeof q);
        q.gridDim = dim3(CL * fir_max_clusters(CL), 1, 1); q.blockDim = dim3(FF_NTHR, 1, 1); q.dynamicSmemBytes = FF_SMEM2;
        cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = CL; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
        q.attrs = a; q.numAttrs = 1;
        int nc = 0;
        MS_CUDA_OK(cudaOccupancyMaxActiveClusters(&nc, kern, &q));
        if (nc < 1) MS_FAIL("fir: clusters of %d CTAs cannot be scheduled on this device", CL);
        max_clusters = nc;
    }
    const int ncl = std::min(std::min(max_clusters, P->n_scratch), (int)P->units.size());
    cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3((unsigned)(CL * ncl), 1, 1); cfg.blockDim = dim3(FF_NTHR, 1, 1); cfg.dynamicSmemBytes = FF_SMEM2; cfg.stream = st;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    MS_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, (const FirUnit*)P->units_dev, (int)P->units.size(), P->ft, P->scratch));
    ++ms_launch_counter();
    if (ms_launch_hook()) ms_launch_hook()(CL == 4 ? "K = fir_cluster_kernel<4>;" : CL == 8 ? "K = fir_cluster_kernel<8>;" : "K = fir_cluster_kernel<16>;", (void*)st);
    return 0;
}
#endif
static int fir_run_fused(FirPlan* P, ms_stream_t st) {
    const int mode = fir_mode();
#ifndef MS_HOST_EMUL
    if (mode == 4) return fir_launch_cluster<4>(P, st);
    if (mode == 8) return fir_launch_cluster<8>(P, st);
    if (mode == 16) return fir_launch_cluster<16>(P, st);
#endif
    (void)mode;
    for (size_t y0 = 0; y0 < P->units.size(); y0 += 32768) {
        const unsigned yc = (unsigned)std::min<size_t>(32768, P->units.size() - y0);
        const FirUnit* u = P->units_dev + y0;
        cpx* sc = P->scratch + y0 * (size_t)(FF_N * FF_N);
        if (ms_launch<FirP1K>(mk_dim(FF_TILES, yc), FF_NTHR, FF_SMEM1, st, u, P->ft, sc)) return -1;
        if (ms_launch<FirP2K>(mk_dim(FF_TILES2, yc), FF_NTHR, FF_SMEM2, st, u, P->ft, sc)) return -1;
        if (ms_launch<FirP3K>(mk_dim(FF_TILES, yc), FF_NTHR, FF_SMEM1, st, u, P->ft, sc)) return -1;
    }
    return 0;
}
extern "C" int MS_API(ms_fir_run)(void* handle, void* stream) {
    FirPlan* P = (FirPlan*)handle;
    if (!P) MS_FAIL("ms_fir_run: null handle");
    ms_stream_t st = (ms_stream_t)stream;
    if (!P->units.empty() && fir_run_fused(P, st)) return -1;
    if (P->cjobs.empty()) return 0;
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
