#define IMPL new
#define BP_FE_KARATSUBA 0
#include "kern.cuh"
