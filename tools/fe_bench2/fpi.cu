#define FPN(x) fpi_##x
#define BP_FE_CALL 0
#include "fp.cu"
