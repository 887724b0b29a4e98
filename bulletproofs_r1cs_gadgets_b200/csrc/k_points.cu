// explicit instantiation of the POINTS kernel group (see kernel_groups.h)
#define KGROUP_DEFINING
#include "kernel_groups.h"
KGROUP_POINTS(KDEFINE)
