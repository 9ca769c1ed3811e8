// ms_lib.cu -- the one translation unit of libmicrosound_b200.so (nvcc, sm_100a).
#include "ms_launch.cuh"
std::string& ms_err_slot() { static thread_local std::string s; return s; }
#include "ms_fft_api.inl"
#include "ms_stage_api.inl"

extern "C" int ms_version(void) { return MS_ABI_VERSION; }
extern "C" const char* ms_last_error(void) { return ms_err_slot().c_str(); }
extern "C" int ms_is_cuda_build(void) { return 1; }
