// explicit instantiation of the MSM kernel group (see kernel_groups.h)
#define KGROUP_DEFINING
#include "kernel_groups.h"
KGROUP_MSM(KDEFINE)
