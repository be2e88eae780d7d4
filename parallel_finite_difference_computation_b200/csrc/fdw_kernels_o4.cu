#define FDW_ORDER 4
#include "fdw_kernels_inst.inc"
