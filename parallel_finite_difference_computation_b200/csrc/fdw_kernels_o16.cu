#define FDW_ORDER 16
#include "fdw_kernels_inst.inc"
