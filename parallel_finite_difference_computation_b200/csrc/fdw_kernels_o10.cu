#define FDW_ORDER 10
#include "fdw_kernels_inst.inc"
