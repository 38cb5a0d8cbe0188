// goliath_b200/csrc/capi.cu — library-level entry points of the C ABI (include/goliath_b200.h).
#include "common.cuh"

namespace gb {
unsigned long long g_launch_count = 0;
}

GB_API int gb_version(void) { return 1000; }

// kernels launched by this library since load (or since the last reset); host-side counter
GB_API unsigned long long gb_launch_count(void) { return __atomic_load_n(&gb::g_launch_count, __ATOMIC_RELAXED); }
GB_API void gb_launch_count_reset(void) { __atomic_store_n(&gb::g_launch_count, 0ull, __ATOMIC_RELAXED); }
