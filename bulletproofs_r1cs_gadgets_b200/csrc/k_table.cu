// explicit instantiation of the TABLE kernel group (see kernel_groups.h)
#define KGROUP_DEFINING
#include "kernel_groups.h"
KGROUP_TABLE(KDEFINE)
