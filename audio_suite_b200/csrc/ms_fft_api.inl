// ms_fft_api.inl -- C-ABI entry points of the spectral stage (include/microsound_b200.h).
// Included by ms_lib.cu (nvcc, product) and by tests/host_emul (g++, block emulator).

static_assert(sizeof(ms_band_edge) == sizeof(BandEdge), "ABI mirror of BandEdge");
static_assert(sizeof(ms_spec_op) == sizeof(SpecOp) && offsetof(ms_spec_op, lp) == offsetof(SpecOp, lp) &&
              offsetof(ms_spec_op, warp_exp) == offsetof(SpecOp, warp_exp), "ABI mirror of SpecOp");

static inline size_t ms_align256(size_t x) { return (x + 255) & ~(size_t)255; }

static void copy_edge(BandEdge& d, const ms_band_edge& s) {
    d.lo_f0 = s.lo_f0; d.lo_f1 = s.lo_f1; d.hi_f0 = s.hi_f0; d.hi_f1 = s.hi_f1;
    d.lo_mode = s.lo_mode; d.hi_mode = s.hi_mode; d.zero = s.zero; d._pad = 0;
}

struct SpecLayout { size_t jobs_off, z_off, work_off, total; std::vector<size_t> z_at, work_at; };

// geometry only (no buffers); fills `out` in input order
static int spec_prepare(const ms_spec_job* in, int njobs, std::vector<FftJob>& out, SpecLayout& lay, ms_stream_t st) {
    out.resize(njobs);
    lay.z_at.resize(njobs); lay.work_at.resize(njobs);
    size_t z = 0, w = 0;
    for (int i = 0; i < njobs; ++i) {
        FftJob& J = out[i];
        memset(&J, 0, sizeof J);
        if (FftEngine::get().prepare(J, in[i].n, st)) return -1;
        lay.z_at[i] = z; z += (size_t)J.n;
        lay.work_at[i] = w; if (J.F1 > 1) w += (size_t)J.M;
    }
    lay.jobs_off = 0;
    lay.z_off = ms_align256(sizeof(FftJob) * (size_t)njobs);
    lay.work_off = lay.z_off + ms_align256(sizeof(cpx) * z);
    lay.total = lay.work_off + ms_align256(sizeof(cpx) * w);
    return 0;
}

extern "C" size_t MS_API(ms_spectral_workspace_bytes)(const ms_spec_job* jobs, int njobs) {
    std::vector<FftJob> tmp; SpecLayout lay;
    if (njobs <= 0) return 256;
    if (spec_prepare(jobs, njobs, tmp, lay, (ms_stream_t)0)) return 0;
    return lay.total;
}

// groups: job index boundaries such that one group's working set (signals, spectrum, scratch) stays in L2
struct SpectralPlan { std::vector<FftJob> jobs; FftJob* jobs_dev; std::vector<size_t> groups; std::vector<size_t> z_at; size_t z_off; };
// Measured on B200 (C5 sweep): walking the batch in 56 MB groups was 14 % SLOWER (2023 small launches, tail
// effects) than one launch per pass over the whole batch -- these passes are latency-bound, not DRAM-bound --
// so grouping is off.
static const size_t MS_L2_GROUP_BYTES = (size_t)1 << 62;

extern "C" int MS_API(ms_spectral_create)(const ms_spec_job* in, int njobs, const real* src, real* dst,
                                  void* ws, size_t ws_bytes, void* stream, void** handle) {
    *handle = nullptr;
    ms_stream_t st = (ms_stream_t)stream;
    SpectralPlan* P = new SpectralPlan();
    P->jobs_dev = nullptr;
    if (njobs <= 0) { *handle = P; return 0; }
    std::vector<FftJob>& jobs = P->jobs; SpecLayout lay;
    if (spec_prepare(in, njobs, jobs, lay, st)) { delete P; return -1; }
    if (ws_bytes < lay.total) { delete P; MS_FAIL("ms_spectral_create: workspace %zu < required %zu", ws_bytes, lay.total); }
    P->z_at = lay.z_at; P->z_off = lay.z_off;
    char* base = (char*)ws;
    for (int i = 0; i < njobs; ++i) {
        FftJob& J = jobs[i];
        const ms_spec_job& s = in[i];
        J.in_a = src + s.in_a; J.in_b = s.in_b >= 0 ? src + s.in_b : nullptr;
        J.out_a = dst + s.out_a; J.out_b = s.out_b >= 0 ? dst + s.out_b : nullptr;
        J.Z = (cpx*)(base + lay.z_off) + lay.z_at[i];
        J.work = J.F1 > 1 ? (cpx*)(base + lay.work_off) + lay.work_at[i] : nullptr;
        J.out_scale = (real)1.0 / (real)J.n;
        for (int w = 0; w < 2; ++w) {
            SpecOp& op = J.op[w];
            const ms_spec_op& so = s.op[w];
            op.kind = so.kind; op.n_bands = so.n_bands; op.lp_on = so.lp_on; op.stretch_on = so.stretch_on;
            op.df = so.df; op.factor = so.factor; op.alpha = so.alpha; op.warp_exp = so.warp_exp;
            copy_edge(op.lp, so.lp);
            for (int b = 0; b < 3; ++b) copy_edge(op.mb[b], so.mb[b]);
        }
    }
    std::stable_sort(jobs.begin(), jobs.end(), [](const FftJob& a, const FftJob& b) {
        return FftEngine::job_class(a) < FftEngine::job_class(b); });
    P->jobs_dev = (FftJob*)(base + lay.jobs_off);
    if (ms_h2d(P->jobs_dev, jobs.data(), sizeof(FftJob) * jobs.size(), st)) { delete P; return -1; }
    // forward and inverse of one group run back to back so the spectrum and the two-pass scratch are
    // still in the 126 MB L2 when the next kernel wants them
    size_t acc = 0;
    P->groups.push_back(0);
    for (size_t i = 0; i < jobs.size(); ++i) {
        const FftJob& J = jobs[i];
        const size_t bytes = sizeof(cpx) * ((size_t)J.n + (J.F1 > 1 ? (size_t)J.M : 0)) + 4 * sizeof(real) * (size_t)J.n;
        if (acc && acc + bytes > MS_L2_GROUP_BYTES) { P->groups.push_back(i); acc = 0; }
        acc += bytes;
    }
    P->groups.push_back(jobs.size());
    *handle = P;
    return 0;
}
#ifndef MS_HOST_EMUL
// Side streams for the length classes of a spectral stage: the classes are independent launch chains (columns -> rows ->
// operators -> columns -> rows), and run side by side their tails and small grids fill each other's idle SMs.  Fork / join
// with events, so the caller's stream sees one ordered operation (and a stream capture records a forked graph).
struct SpecStreams {
    enum { N = 5 };
    cudaStream_t s[N]; cudaEvent_t fork, join[N]; bool ok;
    SpecStreams() : ok(true) {
        for (int i = 0; i < N; ++i) ok = ok && cudaStreamCreateWithFlags(&s[i], cudaStreamNonBlocking) == cudaSuccess
                                             && cudaEventCreateWithFlags(&join[i], cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&fork, cudaEventDisableTiming) == cudaSuccess;
    }
    static SpecStreams& get() { static thread_local SpecStreams v; return v; }     // (per host thread: the events are re-recorded by every call)
};
// development switch MS_SPEC_STREAMS=0: one stream, class after class
static bool spec_streams_on() { static int v = -1; if (v < 0) { const char* e = getenv("MS_SPEC_STREAMS"); v = (e && e[0] == '0') ? 0 : 1; } return v != 0; }
#endif

extern "C" int MS_API(ms_spectral_run)(void* handle, void* stream) {
    SpectralPlan* P = (SpectralPlan*)handle;
    if (!P) MS_FAIL("ms_spectral_run: null handle");
    if (P->jobs.empty()) return 0;
    ms_stream_t st = (ms_stream_t)stream;
#ifndef MS_HOST_EMUL
    if (spec_streams_on() && !ms_launch_hook() && P->groups.size() == 2) {
        std::vector<std::pair<size_t, size_t>> rg = FftEngine::class_ranges(P->jobs);
        if (rg.size() > 1 && SpecStreams::get().ok) {
            // heaviest class first, on the caller's stream; the others on side streams
            auto weight = [&](const std::pair<size_t, size_t>& r) { double w = 0; for (size_t i = r.first; i < r.second; ++i) w += (double)P->jobs[i].n; return w; };
            std::sort(rg.begin(), rg.end(), [&](const std::pair<size_t, size_t>& a, const std::pair<size_t, size_t>& b) { return weight(a) > weight(b); });
            SpecStreams& S = SpecStreams::get();
            int used = 0, rc = cudaEventRecord(S.fork, st) == cudaSuccess ? 0 : -1;
            for (size_t k = 0; k < rg.size() && !rc; ++k) {
                // chain k -> stream (k - 1) mod N for k >= 1 (more classes than side streams: they queue up behind each other)
                ms_stream_t q = st;
                if (k > 0) {
                    const int i = (int)((k - 1) % SpecStreams::N);
                    q = S.s[i];
                    if ((int)k - 1 < SpecStreams::N) { rc = cudaStreamWaitEvent(q, S.fork, 0) == cudaSuccess ? 0 : -1; used = i + 1; }
                    if (rc) break;
                }
                rc = FftEngine::get().forward(P->jobs, P->jobs_dev, q, rg[k].first, rg[k].second);
                if (!rc) rc = FftEngine::get().inverse(P->jobs, P->jobs_dev, q, rg[k].first, rg[k].second);
            }
            for (int i = 0; i < used; ++i) {                    // always join, also after an error: a capture must not be left forked
                cudaEventRecord(S.join[i], S.s[i]);
                cudaStreamWaitEvent(st, S.join[i], 0);
            }
            if (rc) MS_FAIL("ms_spectral_run: a launch chain of the stage failed");
            return 0;
        }
    }
#endif
    for (size_t g = 0; g + 1 < P->groups.size(); ++g) {
        if (FftEngine::get().forward(P->jobs, P->jobs_dev, st, P->groups[g], P->groups[g + 1])) return -1;
        if (FftEngine::get().inverse(P->jobs, P->jobs_dev, st, P->groups[g], P->groups[g + 1])) return -1;
    }
    return 0;
}
extern "C" int MS_API(ms_spectral_forward)(void* handle, void* stream) {
    SpectralPlan* P = (SpectralPlan*)handle;
    if (!P) MS_FAIL("ms_spectral_forward: null handle");
    if (P->jobs.empty()) return 0;
    return FftEngine::get().forward(P->jobs, P->jobs_dev, (ms_stream_t)stream);
}
extern "C" int MS_API(ms_spectral_inverse)(void* handle, void* stream) {
    SpectralPlan* P = (SpectralPlan*)handle;
    if (!P) MS_FAIL("ms_spectral_inverse: null handle");
    if (P->jobs.empty()) return 0;
    return FftEngine::get().inverse(P->jobs, P->jobs_dev, (ms_stream_t)stream);
}
extern "C" int MS_API(ms_spectral_z_table)(void* handle, int64_t* z_offsets, size_t* z_base_bytes) {
    SpectralPlan* P = (SpectralPlan*)handle;
    if (!P) MS_FAIL("ms_spectral_z_table: null handle");
    for (size_t i = 0; i < P->z_at.size(); ++i) z_offsets[i] = (int64_t)P->z_at[i];
    *z_base_bytes = P->z_off;
    return 0;
}
extern "C" void MS_API(ms_spectral_destroy)(void* handle) { delete (SpectralPlan*)handle; }

extern "C" int MS_API(ms_spectral_apply)(const ms_spec_job* in, int njobs, const real* src, real* dst,
                                 void* ws, size_t ws_bytes, void* stream) {
    if (njobs <= 0) return 0;
    void* h = nullptr;
    if (MS_API(ms_spectral_create)(in, njobs, src, dst, ws, ws_bytes, stream, &h)) return -1;
    const int rc = MS_API(ms_spectral_run)(h, stream);
    MS_API(ms_spectral_destroy)(h);
    return rc;
}

extern "C" size_t MS_API(ms_fft_pair_workspace_bytes)(int n) {
    ms_spec_job j; memset(&j, 0, sizeof j); j.n = n;
    return MS_API(ms_spectral_workspace_bytes)(&j, 1);
}

extern "C" int MS_API(ms_fft_pair_forward)(const real* a, const real* b, int n, real* z_out,
                                   void* ws, size_t ws_bytes, void* stream) {
    ms_stream_t st = (ms_stream_t)stream;
    ms_spec_job s; memset(&s, 0, sizeof s); s.n = n;
    std::vector<FftJob> jobs; SpecLayout lay;
    if (spec_prepare(&s, 1, jobs, lay, st)) return -1;
    if (ws_bytes < lay.total) MS_FAIL("ms_fft_pair_forward: workspace %zu < required %zu", ws_bytes, lay.total);
    char* base = (char*)ws;
    FftJob& J = jobs[0];
    J.in_a = a; J.in_b = b; J.Z = (cpx*)z_out;
    J.work = J.F1 > 1 ? (cpx*)(base + lay.work_off) : nullptr;
    FftJob* jd = (FftJob*)(base + lay.jobs_off);
    if (ms_h2d(jd, jobs.data(), sizeof(FftJob), st)) return -1;
    return FftEngine::get().forward(jobs, jd, st);
}
