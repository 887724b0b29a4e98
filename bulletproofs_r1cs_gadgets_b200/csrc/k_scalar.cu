// explicit instantiation of the SCALAR kernel group (see kernel_groups.h)
#define KGROUP_DEFINING
#include "kernel_groups.h"
KGROUP_SCALAR(KDEFINE)
