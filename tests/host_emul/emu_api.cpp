// emu_api.cpp -- builds the library's C-ABI against the block emulator (TEST INFRASTRUCTURE).
#define MS_HOST_EMUL 1
#include "../../audio_suite_b200/csrc/ms_prelude.h"
#define MS_REAL float
#define MS_CPX float2
#define MS_SFX _f32
#define MS_NS msf
#include "../../audio_suite_b200/csrc/ms_all.inl"
#undef MS_REAL
#undef MS_CPX
#undef MS_SFX
#undef MS_NS
#define MS_REAL double
#define MS_CPX double2
#define MS_SFX _f64
#define MS_NS msd
#include "../../audio_suite_b200/csrc/ms_all.inl"
extern "C" int ms_version(void) { return MS_ABI_VERSION; }
extern "C" const char* ms_last_error(void) { return ms_err_slot().c_str(); }
extern "C" int ms_is_cuda_build(void) { return 0; }
