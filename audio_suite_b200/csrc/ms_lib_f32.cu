// float build of every stage (entry points *_f32)
#include "ms_prelude.h"
#define MS_REAL float
#define MS_CPX float2
#define MS_SFX _f32
#define MS_NS msf
#include "ms_all.inl"
