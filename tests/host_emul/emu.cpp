// emu.cpp -- cooperative-fibre CUDA block emulator (TEST INFRASTRUCTURE; see csrc/ms_rt.cuh).
// Each CUDA thread of a block is a ucontext fibre; Ctx::sync() yields to the next fibre, so a
// barrier is "everyone has reached it" exactly as on the device.  Blocks run one after another.
#define MS_HOST_EMUL 1
#include "../../audio_suite_b200/csrc/ms_launch.cuh"
#include <ucontext.h>
#include <vector>
#include <stdexcept>
#include <cstdlib>
#include <utility>

namespace msemu {
static ucontext_t g_main;
static std::vector<ucontext_t> g_ctx;
static std::vector<char> g_done;
static int g_cur = -1, g_n = 0;
static const std::function<void(const Ctx&)>* g_body = nullptr;
static Ctx g_tmpl;
static const size_t STACK = 256 * 1024;

static void trampoline() {
    Ctx c = g_tmpl; c.tid = g_cur;
    (*g_body)(c);
    g_done[g_cur] = 1;
    swapcontext(&g_ctx[g_cur], &g_main);
}
// Counting barriers: a fibre that arrives waits (yielding) until its barrier's generation changes.  Block barriers
// expect every fibre that has not exited; warp barriers the live fibres of the warp (32 consecutive thread ids), so
// warps may execute different numbers of warp-level barriers between two block barriers (warp-specialised kernels).
static int g_blk_cnt = 0; static unsigned g_blk_gen = 0;
static std::vector<int> g_warp_cnt; static std::vector<unsigned> g_warp_gen;
static int live_in(int lo, int hi) { int n = 0; for (int t = lo; t < hi && t < g_n; ++t) n += g_done[t] ? 0 : 1; return n; }
void yield_barrier() {
    const unsigned gen = g_blk_gen;
    ++g_blk_cnt;
    for (;;) {
        if (g_blk_gen != gen) return;
        if (g_blk_cnt >= live_in(0, g_n)) { g_blk_cnt = 0; ++g_blk_gen; return; }
        swapcontext(&g_ctx[g_cur], &g_main);
    }
}
void yield_warp_barrier() {
    const int w = g_cur >> 5;
    const unsigned gen = g_warp_gen[w];
    ++g_warp_cnt[w];
    for (;;) {
        if (g_warp_gen[w] != gen) return;
        if (g_warp_cnt[w] >= live_in(32 * w, 32 * w + 32)) { g_warp_cnt[w] = 0; ++g_warp_gen[w]; return; }
        swapcontext(&g_ctx[g_cur], &g_main);
    }
}

void run(MsDim grid, int block, size_t smem, const std::function<void(const Ctx&)>& body) {
    std::vector<char> shared(smem + 64);
    std::vector<std::vector<char>> stacks(block, std::vector<char>(STACK));
    g_body = &body; g_n = block;
    for (unsigned by = 0; by < grid.y; ++by) for (unsigned bx = 0; bx < grid.x; ++bx) {
        g_ctx.assign(block, ucontext_t());
        g_done.assign(block, 0);
        g_blk_cnt = 0; g_warp_cnt.assign((block + 31) / 32, 0); g_warp_gen.assign((block + 31) / 32, 0);
        g_tmpl.nthr = block; g_tmpl.bx = (int)bx; g_tmpl.by = (int)by; g_tmpl.smem = shared.data();
        for (int t = 0; t < block; ++t) {
            getcontext(&g_ctx[t]);
            g_ctx[t].uc_stack.ss_sp = stacks[t].data();
            g_ctx[t].uc_stack.ss_size = STACK;
            g_ctx[t].uc_link = &g_main;
            makecontext(&g_ctx[t], trampoline, 0);
        }
        // MS_EMUL_SHUFFLE=<seed>: every sweep visits the fibres in a fresh pseudo-random order.  A kernel whose result depends
        // on which thread runs first between two barriers (a shared-memory race) then gives different results under
        // different seeds -- the stand-in for compute-sanitizer's racecheck, which is closed on the GPU pool
        // (tests/test_emul_kernels.py::test_results_do_not_depend_on_the_thread_schedule).
        static const char* shuf = getenv("MS_EMUL_SHUFFLE");
        static unsigned long long rs = shuf ? strtoull(shuf, nullptr, 10) * 2654435761ull + 88172645463325252ull : 0ull;
        std::vector<int> order(block);
        for (int t = 0; t < block; ++t) order[t] = t;
        for (;;) {      // one sweep = one barrier interval
            int alive = 0, finished = 0;
            if (shuf) for (int t = block - 1; t > 0; --t) { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; std::swap(order[t], order[(int)(rs % (unsigned long long)(t + 1))]); }
            for (int ti = 0; ti < block; ++ti) {
                const int t = order[ti];
                if (g_done[t]) { ++finished; continue; }
                g_cur = t;
                swapcontext(&g_main, &g_ctx[t]);
                if (g_done[t]) ++finished; else ++alive;
            }
            if (alive == 0) break;
            if (finished != 0 && alive != 0) {
                // some threads exited while others wait at a barrier: legal in CUDA only if the
                // exited threads never reach it again; keep sweeping the live ones.
            }
        }
    }
}
}  // namespace msemu

std::string& ms_err_slot() { static thread_local std::string s; return s; }
unsigned long long& ms_launch_counter() { static unsigned long long n = 0; return n; }
extern "C" unsigned long long ms_launch_count(void) { return ms_launch_counter(); }
unsigned long long& ms_h2d_counter() { static unsigned long long n = 0; return n; }
ms_launch_hook_t& ms_launch_hook() { static ms_launch_hook_t h = nullptr; return h; }
extern "C" void ms_set_launch_hook(ms_launch_hook_t h) { ms_launch_hook() = h; }
extern "C" unsigned long long ms_h2d_bytes(void) { return ms_h2d_counter(); }
