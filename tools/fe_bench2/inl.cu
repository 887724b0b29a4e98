#define IMPL inl
#define BP_FE_KARATSUBA 0
#define BP_FE_CALL 0
#include "kern.cuh"
