/*
 * fdwave_gpufam.h -- the GPU family's function-level API with the reference's
 * exact names and signatures, served by libfdwave_gpufam.so on top of
 * libfdwave.so.  Drop-in for the functions cuda_reference_RTM/src/fd-code.cu
 * defines around its kernels:
 *   fd_init        fd-code.cu:200-224   (C linkage via lib/include/functions.h:15)
 *   fd_init_cuda   fd-code.cu:146-198   (C linkage via functions.h:16)
 *   write_buffers  fd-code.cu:226-245
 *   fd_forward     fd-code.cu:247-288
 *   fd_back        fd-code.cu:290-341
 * Callers pass nz = nze, nx = nxe (fd-code.cu:499,518).  One instance per
 * process, like the reference's file-static state.  The truncated launch
 * extents of the reference (quirk Q1) are reproduced unless FDW_COMPAT=0.
 */
#ifndef FDWAVE_GPUFAM_H
#define FDWAVE_GPUFAM_H

#ifdef __cplusplus
extern "C" {
#endif
void fd_init(int order, int nx, int nz, int nxb, int nzb, int nt, int ns, float fac, float dx, float dz, float dt);
void fd_init_cuda(int order, int nxe, int nze, int nxb, int nzb, int nt, int ns, float fac);
#ifdef __cplusplus
}
void write_buffers(float **p, float **pp, float **v2, float *taperx, float *taperz, float **d_obs, float **imloc,
                   int is, int flag);
void fd_forward(int order, float **p, float **pp, float **v2, int nz, int nx, int nt, int is, int sz, int *sx,
                float *srce, int propag);
void fd_back(int order, float **p, float **pp, float **pr, float **ppr, float **v2, int nz, int nx, int nt, int is,
             int sz, int gz, float ***snaps, float **imloc, float **d_obs);
#endif
#endif
