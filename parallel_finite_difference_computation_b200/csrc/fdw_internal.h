/* fdw_internal.h -- private declarations shared by the libfdwave sources. */
#ifndef FDW_INTERNAL_H
#define FDW_INTERNAL_H

#ifdef __cplusplus
extern "C" {
#endif

void fdw_set_error(const char *fmt, ...);

/* per-order kernel tables (fdw_kernels_o{2,4,6,8,10,12,14,16}.cu) */
const void *fdw_step_kernel_o2(int recipe, int epi, int sponge);
const void *fdw_step_kernel_o4(int recipe, int epi, int sponge);
const void *fdw_step_kernel_o6(int recipe, int epi, int sponge);
const void *fdw_step_kernel_o8(int recipe, int epi, int sponge);
const void *fdw_step_kernel_o10(int recipe, int epi, int sponge);
const void *fdw_step_kernel_o12(int recipe, int epi, int sponge);
const void *fdw_step_kernel_o14(int recipe, int epi, int sponge);
const void *fdw_step_kernel_o16(int recipe, int epi, int sponge);
const void *fdw_persist_kernel_o2(int recipe, int epi);
const void *fdw_persist_kernel_o4(int recipe, int epi);
const void *fdw_persist_kernel_o6(int recipe, int epi);
const void *fdw_persist_kernel_o8(int recipe, int epi);
const void *fdw_persist_kernel_o10(int recipe, int epi);
const void *fdw_persist_kernel_o12(int recipe, int epi);
const void *fdw_persist_kernel_o14(int recipe, int epi);
const void *fdw_persist_kernel_o16(int recipe, int epi);
const void *fdw_tile_kernel_o2(int recipe, int epi); /* epi < 0: GPU-family backward */
const void *fdw_tile_kernel_o4(int recipe, int epi);
const void *fdw_tile_kernel_o6(int recipe, int epi);
const void *fdw_tile_kernel_o8(int recipe, int epi);
const void *fdw_tile_kernel_o10(int recipe, int epi);
const void *fdw_tile_kernel_o12(int recipe, int epi);
const void *fdw_tile_kernel_o14(int recipe, int epi);
const void *fdw_tile_kernel_o16(int recipe, int epi);
const void *fdw_pslab_kernel_o2(int recipe, int epi);
const void *fdw_pslab_kernel_o4(int recipe, int epi);
const void *fdw_pslab_kernel_o6(int recipe, int epi);
const void *fdw_pslab_kernel_o8(int recipe, int epi);
const void *fdw_pslab_kernel_o10(int recipe, int epi);
const void *fdw_pslab_kernel_o12(int recipe, int epi);
const void *fdw_pslab_kernel_o14(int recipe, int epi);
const void *fdw_pslab_kernel_o16(int recipe, int epi);
const void *fdw_lap_kernel_o2(void);
const void *fdw_lap_kernel_o4(void);
const void *fdw_lap_kernel_o6(void);
const void *fdw_lap_kernel_o8(void);
const void *fdw_lap_kernel_o10(void);
const void *fdw_lap_kernel_o12(void);
const void *fdw_lap_kernel_o14(void);
const void *fdw_lap_kernel_o16(void);

#ifdef __cplusplus
}
#endif
#endif
