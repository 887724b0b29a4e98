#define IMPL new
#include "kern.cuh"
