#define FDW_ORDER 2
#include "fdw_kernels_inst.inc"
