// emu_api.cpp -- builds the library's C-ABI against the block emulator (TEST INFRASTRUCTURE).
#define MS_HOST_EMUL 1
#include "../../audio_suite_b200/csrc/ms_fft_api.inl"
#include "../../audio_suite_b200/csrc/ms_stage_api.inl"
extern "C" int ms_version(void) { return MS_ABI_VERSION; }
extern "C" const char* ms_last_error(void) { return ms_err_slot().c_str(); }
extern "C" int ms_is_cuda_build(void) { return 0; }
