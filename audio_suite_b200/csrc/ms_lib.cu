// ms_lib.cu -- precision-independent entry points of libmicrosound_b200.so (nvcc, sm_100a).
#include "ms_prelude.h"
std::string& ms_err_slot() { static thread_local std::string s; return s; }
unsigned long long& ms_launch_counter() { static unsigned long long n = 0; return n; }
extern "C" unsigned long long ms_launch_count(void) { return ms_launch_counter(); }
unsigned long long& ms_h2d_counter() { static unsigned long long n = 0; return n; }
ms_launch_hook_t& ms_launch_hook() { static ms_launch_hook_t h = nullptr; return h; }
extern "C" void ms_set_launch_hook(ms_launch_hook_t h) { ms_launch_hook() = h; }
extern "C" unsigned long long ms_h2d_bytes(void) { return ms_h2d_counter(); }
MsStageRing& ms_stage_ring() { static MsStageRing r; return r; }
extern "C" int ms_version(void) { return MS_ABI_VERSION; }
extern "C" const char* ms_last_error(void) { return ms_err_slot().c_str(); }
extern "C" int ms_is_cuda_build(void) { return 1; }
