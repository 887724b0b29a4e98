#define IMPL kara
#define BP_FE_KARATSUBA 1
#include "kern.cuh"
