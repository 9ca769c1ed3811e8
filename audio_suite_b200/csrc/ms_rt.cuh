// ms_rt.cuh -- thin portability layer so every kernel body in this library is written once
// and can be compiled two ways:
//   * nvcc, sm_100a: the product.  Ctx::sync() is __syncthreads(), smem is dynamic shared memory.
//   * g++ with -DMS_HOST_EMUL: a cooperative-fibre block emulator (tests/host_emul/) that runs the
//     same kernel bodies on the CPU, one fibre per CUDA thread, so index arithmetic is unit-tested
//     in the build container where there is no GPU.  This is test infrastructure, never a fallback:
//     the shipped shared library contains no host compute path.
#pragma once
#include <stdint.h>
#include <math.h>

#ifdef MS_HOST_EMUL
  #include <cmath>
  #include <cstring>
  #include <algorithm>
  #define MS_DEV inline
  #define MS_HD inline
  #define MS_RESTRICT
  struct float2 { float x, y; };
  struct float4 { float x, y, z, w; };
  struct double2 { double x, y; };
  static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
  static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
  static inline float __ldg(const float* p) { return *p; }
  static inline double __ldg(const double* p) { return *p; }
  static inline float2 __ldg(const float2* p) { return *p; }
  static inline int __ldg(const int* p) { return *p; }
  static inline void sincospi(double a, double* s, double* c) { *s = std::sin(M_PI * a); *c = std::cos(M_PI * a); }
  static inline void sincospif(float a, float* s, float* c) { *s = (float)std::sin(M_PI * (double)a); *c = (float)std::cos(M_PI * (double)a); }
  static inline float cospif(float a) { return (float)std::cos(M_PI * (double)a); }
  static inline float sinpif(float a) { return (float)std::sin(M_PI * (double)a); }
  static inline double cospi(double a) { return std::cos(M_PI * a); }
  static inline float __fmaf_rn(float a, float b, float c) { return std::fma(a, b, c); }
  namespace msemu { void yield_barrier(); }
  struct Ctx {
      int tid, nthr, bx, by;
      char* smem;
      void sync() const { msemu::yield_barrier(); }
  };
#else
  #include <cuda_runtime.h>
  #define MS_DEV __device__ __forceinline__
  #define MS_HD __host__ __device__ __forceinline__
  #define MS_RESTRICT __restrict__
  struct Ctx {
      int tid, nthr, bx, by;
      char* smem;
      __device__ __forceinline__ void sync() const { __syncthreads(); }
  };
#endif

// ---- complex helpers (float2 = re, im) -------------------------------------------------------
MS_DEV float2 c_mul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
MS_DEV float2 c_mulc(float2 a, float2 b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }  // a * conj(b)
MS_DEV float2 c_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
MS_DEV float2 c_sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
MS_DEV float2 c_scale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
MS_DEV float2 c_swap(float2 a) { return make_float2(a.y, a.x); }
MS_DEV float2 c_conj(float2 a) { return make_float2(a.x, -a.y); }
MS_DEV float2 c_zero() { return make_float2(0.f, 0.f); }
// multiply by -i (forward-FFT quarter turn): (x+iy)(-i) = y - ix
MS_DEV float2 c_mul_mi(float2 a) { return make_float2(a.y, -a.x); }
