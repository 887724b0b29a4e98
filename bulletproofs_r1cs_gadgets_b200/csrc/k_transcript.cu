// explicit instantiation of the TRANSCRIPT kernel group (see kernel_groups.h)
#define KGROUP_DEFINING
#include "kernel_groups.h"
KGROUP_TRANSCRIPT(KDEFINE)
