// ms_all.inl -- stamps the whole kernel library once per precision.
//   #define MS_REAL float  / MS_CPX float2  / MS_SFX _f32 / MS_NS msf     (then include this file)
//   #define MS_REAL double / MS_CPX double2 / MS_SFX _f64 / MS_NS msd     (and again)
// Everything precision-dependent lives in namespace MS_NS with `real` / `cpx` typedefs; the extern "C"
// entry points get the suffix (ms_post_f32, ms_post_f64, ...).
#define MS_CAT2(a, b) a##b
#define MS_CAT(a, b) MS_CAT2(a, b)
#define MS_API(name) MS_CAT(name, MS_SFX)
namespace MS_NS {
typedef MS_REAL real;
typedef MS_CPX cpx;
MS_DEV cpx mk(real x, real y) { cpx r; r.x = x; r.y = y; return r; }
MS_DEV cpx c_zero() { return mk((real)0, (real)0); }
#include "ms_fft_core.cuh"
#include "ms_fft_kernels.cuh"
#include "ms_fft_host.h"
#include "ms_fft_api.inl"
#include "ms_synth.cuh"
#include "ms_time.cuh"
#include "ms_fir_fused.cuh"
#include "ms_stage_api.inl"
}  // namespace MS_NS
#undef MS_API
#undef MS_CAT
#undef MS_CAT2
