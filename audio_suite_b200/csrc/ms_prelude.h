// ms_prelude.h -- everything that must be seen once, outside the per-precision namespaces.
#pragma once
#include "ms_rt.cuh"
#include "ms_launch.cuh"
#include "ms_zig_tables.h"
#include "../../include/microsound_b200.h"
#include <map>
#include <mutex>
#include <vector>
#include <algorithm>
#include <utility>
#include <string.h>
#include <stdlib.h>
#ifdef MS_HOST_EMUL
#define MS_POPC(x) __builtin_popcount(x)
#else
#define MS_POPC(x) __popc(x)
#endif
