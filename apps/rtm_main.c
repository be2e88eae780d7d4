/*
 * rtm_main -- drop-in for dpct_gpu_rtm_domain_division/src/rtm_main.cpp
 * (:45-282): RTM with the full forward history.  The history stays resident in
 * HBM and the imaging condition is accumulated during the backward pass in the
 * reference's summation order, so dir.img / dir.image match bit for bit while
 * the receiver history is never stored.
 *
 *   ./rtm_main par=input.dat      (writes dir.img and dir.image in the CWD)
 */
#include "cpu_family_args.h"

int main(int argc, char **argv)
{
    fdw_input in;
    cpu_family_input(argc, argv, &in);
    cpu_family_banner(&in);
    const int nx = in.nx, nz = in.nz, nt = in.nt, ns = in.ns;
    const size_t ni = (size_t)nx * nz;
    float *srce = xalloc(nt);
    FDW(fdw_ricker_wavelet(nt, in.dt, in.fpeak, FDW_FAMILY_CPU, srce));
    float *vel2 = cpu_family_vel2(&in);
    float *dobs = xalloc((size_t)ns * nx * nt);
    read_floats(in.datfile, dobs, (size_t)ns * nx * nt, 1);
    fdw_params prm;
    memset(&prm, 0, sizeof prm);
    prm.nx = nx; prm.nz = nz; prm.nxb = in.nxb; prm.nzb = in.nzb; prm.order = in.order;
    prm.dx = in.dx; prm.dz = in.dz; prm.dt = in.dt; prm.fac = in.fac;
    prm.family = FDW_FAMILY_CPU; prm.recipe = FDW_RECIPE_C; prm.taper = FDW_TAPER_TOP;
    prm.device = env_int("FDW_DEVICE", 0);
    prm.nt = nt;
    prm.history = 1;
    fdw_ctx *ctx = NULL;
    FDW(fdw_create(&prm, &ctx));
    FDW(fdw_set_v2(ctx, vel2));
    FDW(fdw_set_wavelet(ctx, srce, nt));
    float *imloc = xalloc(ni), *img = xalloc(ni);
    FILE *flim = fopen("dir.img", "w+");
    if (!flim) DIE("cannot open dir.img");
    const int sz = in.sz + in.nzb, gz = in.gz + in.nzb;
    double t0 = now_s();
    for (int is = 0; is < ns; is++) {
        const int sx = in.fsx + is * in.ds + in.nxb;
        fprintf(stdout, "** source %d, at (%d,%d) \n", is + 1, sx - in.nxb, sz - in.nzb);
        fprintf(stdout, "** backward propagation %d, at (%d,%d) \n", is + 1, sx - in.nxb, sz - in.nzb);
        FDW(fdw_rtm_shot_cpu(ctx, sx, sz, gz, dobs, ns, is, imloc));
        fwrite(imloc, sizeof(float), ni, flim);
        for (size_t k = 0; k < ni; k++) img[k] += imloc[k];
    }
    printf("Execution Time: %.2f seconds", now_s() - t0);
    fclose(flim);
    write_floats("dir.image", img, ni, "w+");
    fdw_destroy(ctx);
    free(srce); free(vel2); free(dobs); free(imloc); free(img);
    return 0;
}
