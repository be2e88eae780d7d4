#define FDW_ORDER 8
#include "fdw_kernels_inst.inc"
