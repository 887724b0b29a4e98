// explicit instantiation of the FOLD kernel group (see kernel_groups.h)
#define KGROUP_DEFINING
#include "kernel_groups.h"
KGROUP_FOLD(KDEFINE)
