// ms_launch.cuh -- one launcher for both builds (see ms_rt.cuh).  A kernel is a struct K with
//   static constexpr int MAXT;                    // __launch_bounds__
//   static MS_DEV void run(args..., const Ctx&);   // the body
#pragma once
#include "ms_rt.cuh"
#include <stddef.h>
#include <stdio.h>
#include <string>
#include <string.h>
#include <mutex>

struct MsDim { unsigned x, y; };

// thread-local last error (C-ABI: ms_last_error)
std::string& ms_err_slot();
// kernels launched by this library since it was loaded (C-ABI: ms_launch_count)
unsigned long long& ms_launch_counter();
// bytes of job tables this library has copied host->device itself (C-ABI: ms_h2d_bytes)
unsigned long long& ms_h2d_counter();
// optional observer called after every kernel launch with the kernel's type name and the stream (C-ABI:
// ms_set_launch_hook; bench.py records a CUDA event there to time individual kernels inside a step)
typedef void (*ms_launch_hook_t)(const char* kernel, void* stream);
ms_launch_hook_t& ms_launch_hook();
#define MS_FAIL(...) do { char _b[512]; snprintf(_b, sizeof _b, __VA_ARGS__); ms_err_slot() = _b; return -1; } while (0)

#ifdef MS_HOST_EMUL
#include <functional>
#include <cstdlib>
typedef void* ms_stream_t;
namespace msemu {
    void run(MsDim grid, int block, size_t smem, const std::function<void(const Ctx&)>& body);
}
template <class K, class... Args>
int ms_launch(MsDim grid, int block, size_t smem, ms_stream_t, Args... args) {
    msemu::run(grid, block, smem, [&](const Ctx& c) { K::run(args..., c); });
    ++ms_launch_counter();
    if (ms_launch_hook()) ms_launch_hook()(__PRETTY_FUNCTION__, nullptr);
    return 0;
}
static inline void* ms_dev_alloc(size_t bytes) { return calloc(1, bytes ? bytes : 1); }
static inline void ms_dev_free(void* p) { free(p); }
static inline int ms_h2d(void* dst, const void* src, size_t bytes, ms_stream_t) { ms_h2d_counter() += bytes; memcpy(dst, src, bytes); return 0; }
static inline int ms_memset(void* dst, int v, size_t bytes, ms_stream_t) { memset(dst, v, bytes); return 0; }
#else
typedef cudaStream_t ms_stream_t;
template <class K, class... Args>
__global__ void __launch_bounds__(K::MAXT, K::MINB) ms_kernel(Args... args) {
    extern __shared__ float4 ms_dyn_smem[];
    Ctx c;
    c.tid = threadIdx.x; c.nthr = blockDim.x; c.bx = blockIdx.x; c.by = blockIdx.y;
    c.smem = (char*)ms_dyn_smem;
    K::run(args..., c);
}
#define MS_CUDA_OK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) MS_FAIL("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); } while (0)
template <class K, class... Args>
int ms_launch(MsDim grid, int block, size_t smem, ms_stream_t st, Args... args) {
    static size_t configured = 0;          // per instantiation
    if (block > K::MAXT) MS_FAIL("launch: block %d exceeds kernel bound %d", block, K::MAXT);
    if (smem > 48 * 1024 && smem > configured) {
        MS_CUDA_OK(cudaFuncSetAttribute(ms_kernel<K, Args...>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    if (grid.x == 0 || grid.y == 0) return 0;
    ms_kernel<K, Args...><<<dim3(grid.x, grid.y, 1), block, smem, st>>>(args...);
    ++ms_launch_counter();
    MS_CUDA_OK(cudaGetLastError());
    if (ms_launch_hook()) ms_launch_hook()(__PRETTY_FUNCTION__, (void*)st);
    return 0;
}
static inline void* ms_dev_alloc(size_t bytes) { void* p = nullptr; if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr; return p; }
static inline void ms_dev_free(void* p) { cudaFree(p); }
// Job tables are built in pageable host memory (std::vector).  A cudaMemcpyAsync from pageable memory makes
// the host wait for the stream, which would serialise "build the tables of slice k+1" behind "render slice k".
// So every upload is staged through a small ring of pinned blocks owned by the library: memcpy into the ring,
// async DMA from there; a block is reused only after the event recorded behind its last copy has completed.
struct MsStageRing {
    static const size_t BLK = (size_t)8 << 20;
    static const int NB = 6;
    char* base[NB]; cudaEvent_t ev[NB]; bool used[NB]; cudaStream_t last[NB];
    int cur; size_t off; std::mutex mu; bool ok;
    MsStageRing() : cur(0), off(0), ok(true) { for (int i = 0; i < NB; ++i) { base[i] = nullptr; used[i] = false; last[i] = 0; } }
    int open(int b) {
        if (!base[b]) {
            // all blocks at once, on first use: pinning memory later, while planning workers keep the cores busy,
            // stalls the caller for tens of ms (measured on the B200 hosts)
            for (int i = 0; i < NB; ++i) {
                if (base[i]) continue;
                if (cudaHostAlloc((void**)&base[i], BLK, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); base[i] = nullptr; return -1; }
                if (cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess) return -1;
            }
        }
        if (used[b]) { if (cudaEventSynchronize(ev[b]) != cudaSuccess) return -1; used[b] = false; }
        return 0;
    }
    int copy(void* dst, const void* src, size_t bytes, cudaStream_t st) {
        std::lock_guard<std::mutex> lk(mu);
        const char* s = (const char*)src; char* d = (char*)dst;
        if (!base[cur] && open(cur)) return -1;
        while (bytes) {
            if (off >= BLK || (off > 0 && last[cur] != st)) {          // next block (also when the stream changes)
                cur = (cur + 1) % NB; off = 0;
                if (open(cur)) return -1;
            }
            const size_t c = bytes < BLK - off ? bytes : BLK - off;
            memcpy(base[cur] + off, s, c);
            if (cudaMemcpyAsync(d, base[cur] + off, c, cudaMemcpyHostToDevice, st) != cudaSuccess) return -1;
            if (cudaEventRecord(ev[cur], st) != cudaSuccess) return -1;
            used[cur] = true; last[cur] = st;
            off += (c + 255) & ~(size_t)255; s += c; d += c; bytes -= c;
        }
        return 0;
    }
};
MsStageRing& ms_stage_ring();
static inline int ms_h2d(void* dst, const void* src, size_t bytes, ms_stream_t st) {
    ms_h2d_counter() += bytes;
    if (bytes == 0) return 0;
    if (ms_stage_ring().copy(dst, src, bytes, st)) {
        cudaGetLastError();
        MS_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));       // pageable fallback (slower, still correct)
    }
    return 0;
}
static inline int ms_memset(void* dst, int v, size_t bytes, ms_stream_t st) {
    MS_CUDA_OK(cudaMemsetAsync(dst, v, bytes, st)); return 0;
}
#endif
