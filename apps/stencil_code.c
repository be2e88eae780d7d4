/*
 * stencil_code -- drop-in for cuda_reference_stencil_computation/fd-source-code.cu
 * (main(): :277-352) and its DPC++ migrations: one Laplacian sweep of the
 * snapshot named by `tmpdir=` in input.dat.
 *
 *   ./stencil_code ./input.dat [output.bin]
 *
 * Output: ../bin/output_cuda.bin like the CUDA reference (fd-source-code.cu:337)
 * unless a second argument names another file (the DPC++ builds write
 * output_teste.bin, fd-source-code.dp.cpp:386).
 */
#include "common.h"

int main(int argc, char **argv)
{
    if (argc < 2) DIE("usage: %s input.dat [output.bin]", argv[0]);
    fdw_input in;
    FDW(fdw_read_input_stencil(argv[1], &in));
    printf("Local do arquivo: %s\n", in.tmpdir);
    printf("nzb = %i\n", in.nzb);
    printf("nzb = %i\n", in.nxb);
    printf("nz = %i\n", in.nz);
    printf("nx = %i\n", in.nx);
    printf("dz = %f\n", in.dz);
    printf("dx = %f\n", in.dx);
    printf("order = %i\n", in.order);
    const int nxe = in.nx + 2 * in.nxb, nze = in.nz + 2 * in.nzb;
    const size_t n = (size_t)nxe * nze;
    float *p = xalloc(n), *lap = xalloc(n);
    if (read_floats(in.tmpdir, p, n, 1) != n)
        printf("Input reading error!\n");
    else
        printf("Input reading was successful.\n");
    double t0 = now_s();
    FDW(fdw_stencil(in.order, nxe, nze, in.dx, in.dz, p, lap, env_int("FDW_DEVICE", 0)));
    fprintf(stderr, "[fdwave] laplacian %dx%d order %d: %.3f ms (incl. context + copies)\n", nxe, nze, in.order,
            1e3 * (now_s() - t0));
    const char *out = argc > 2 ? argv[2] : "../bin/output_cuda.bin";
    FILE *f = fopen(out, "wb");
    if (!f) {
        printf("Unable to open file!\n");
        return 1;
    }
    printf("Output successfully opened for writing.\n");
    if (fwrite(lap, sizeof(float), n, f) != n)
        printf("Output writing error!\n");
    else
        printf("Output writing was successful.\n");
    fclose(f);
    free(p);
    free(lap);
    return 0;
}
