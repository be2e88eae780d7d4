#define FDW_ORDER 6
#include "fdw_kernels_inst.inc"
