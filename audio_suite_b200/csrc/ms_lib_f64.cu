// double build of every stage (entry points *_f64)
#include "ms_prelude.h"
#define MS_REAL double
#define MS_CPX double2
#define MS_SFX _f64
#define MS_NS msd
#include "ms_all.inl"
