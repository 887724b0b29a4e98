#define IMPL old
#include "kern.cuh"
