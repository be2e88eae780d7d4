#define FDW_ORDER 14
#include "fdw_kernels_inst.inc"
