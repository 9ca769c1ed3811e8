// emu.cpp -- cooperative-fibre CUDA block emulator (TEST INFRASTRUCTURE; see csrc/ms_rt.cuh).
// Each CUDA thread of a block is a ucontext fibre; Ctx::sync() yields to the next fibre, so a
// barrier is "everyone has reached it" exactly as on the device.  Blocks run one after another.
#define MS_HOST_EMUL 1
#include "../../audio_suite_b200/csrc/ms_launch.cuh"
#include <ucontext.h>
#include <vector>
#include <stdexcept>

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
void yield_barrier() { swapcontext(&g_ctx[g_cur], &g_main); }

void run(MsDim grid, int block, size_t smem, const std::function<void(const Ctx&)>& body) {
    std::vector<char> shared(smem + 64);
    std::vector<std::vector<char>> stacks(block, std::vector<char>(STACK));
    g_body = &body; g_n = block;
    for (unsigned by = 0; by < grid.y; ++by) for (unsigned bx = 0; bx < grid.x; ++bx) {
        g_ctx.assign(block, ucontext_t());
        g_done.assign(block, 0);
        g_tmpl.nthr = block; g_tmpl.bx = (int)bx; g_tmpl.by = (int)by; g_tmpl.smem = shared.data();
        for (int t = 0; t < block; ++t) {
            getcontext(&g_ctx[t]);
            g_ctx[t].uc_stack.ss_sp = stacks[t].data();
            g_ctx[t].uc_stack.ss_size = STACK;
            g_ctx[t].uc_link = &g_main;
            makecontext(&g_ctx[t], trampoline, 0);
        }
        for (;;) {      // one sweep = one barrier interval
            int alive = 0, finished = 0;
            for (int t = 0; t < block; ++t) {
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
