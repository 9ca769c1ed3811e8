// ms_lib.cu -- precision-independent entry points of libmicrosound_b200.so (nvcc, sm_100a).
#include "ms_prelude.h"
std::string& ms_err_slot() { static thread_local std::string s; return s; }
extern "C" int ms_version(void) { return MS_ABI_VERSION; }
extern "C" const char* ms_last_error(void) { return ms_err_slot().c_str(); }
extern "C" int ms_is_cuda_build(void) { return 1; }
