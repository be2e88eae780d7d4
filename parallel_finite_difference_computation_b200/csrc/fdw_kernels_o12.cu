#define FDW_ORDER 12
#include "fdw_kernels_inst.inc"
