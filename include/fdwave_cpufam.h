/*
 * fdwave_cpufam.h -- the CPU family's function-level API with the reference's
 * exact names and signatures, served by libfdwave_cpufam.so on top of
 * libfdwave.so.  Drop-in for
 *   dpct_gpu_rtm_domain_division/include/timestep/fd.h:4-7
 *   dpct_gpu_rtm_domain_division/include/boundary/taper.h:4-8
 *   dpct_gpu_rtm_domain_division/include/source/ptsrc.h:4-6
 * The reference builds these sources as C++ (CC = g++), so its callers
 * (mod_main.cpp, rtm_main.cpp) reference C++-mangled symbols; the shim is C++
 * for the same reason and this header deliberately has no extern "C".
 * float** arguments are CWP alloc2float tables: p[0] is the base of one
 * contiguous [nx][nz] block (lib/cwp/src/cwp/lib/alloc.c).
 *
 * fd_step runs on the GPU (one fused Laplacian+leapfrog launch, bit-exact
 * recipe C); because this API keeps the fields in caller-owned host arrays the
 * shim moves p, pp, v2 over PCIe on every call -- it is the compatibility
 * path.  The resident-in-HBM path is fdwave.h (fdw_model_shot, fdw_rtm_shot_cpu).
 * taper_* / ptsrc / ricker_wavelet / extendvel act on the caller's host arrays
 * exactly like the reference (0.1 % of its run time, gprof in SURVEY 8a).
 */
#ifndef FDWAVE_CPUFAM_H
#define FDWAVE_CPUFAM_H

void fd_init(int order, int nx, int nz, float dx, float dz, float dt);              /* fd.c:11-22 */
void fd_step(int order, float **p, float **pp, float **v2, int nz, int nx);         /* fd.c:24-46 */
void fd_destroy();                                                                  /* fd.c:48-52 */
float *calc_coefs(int order);                                                       /* fd.c:54-97 */
void extendvel(int nx, int nz, int nxb, int nzb, float *vel);                       /* taper.c:7-23 */
void taper_init(int nxb, int nzb, float F);                                         /* taper.c:25-45 */
void taper_apply(float **pp, int nx, int nz, int nxb, int nzb);                     /* taper.c:47-67 */
void taper_apply2(float **pp, int nx, int nz, int nxb, int nzb);                    /* taper.c:69-84 */
void taper_destroy();                                                               /* taper.c:87-91 */
void ptsrc(int xs, int zs, int nx, int nz, float ts, float **s);                    /* ptsrc.c:12-58 */
void ricker_wavelet(int ns, float dt, float peak, float *s);                        /* ptsrc.c:88-99 */
float ricker(float t, float fpeak);                                                 /* ptsrc.c:60-86 */

#endif
