/*
 * mod_main -- drop-in for dpct_gpu_rtm_domain_division/src/mod_main.cpp
 * (:42-208): acoustic modelling, writes the seismograms [ns][nx][nt] to datfile.
 *
 *   ./mod_main par=input.dat
 */
#include "cpu_family_args.h"

int main(int argc, char **argv)
{
    fdw_input in;
    cpu_family_input(argc, argv, &in);
    cpu_family_banner(&in);
    const int nx = in.nx, nt = in.nt, ns = in.ns;
    float *srce = xalloc(nt);
    FDW(fdw_ricker_wavelet(nt, in.dt, in.fpeak, FDW_FAMILY_CPU, srce));
    float *vel2 = cpu_family_vel2(&in);
    fdw_params prm;
    memset(&prm, 0, sizeof prm);
    prm.nx = nx; prm.nz = in.nz; prm.nxb = in.nxb; prm.nzb = in.nzb; prm.order = in.order;
    prm.dx = in.dx; prm.dz = in.dz; prm.dt = in.dt; prm.fac = in.fac;
    prm.family = FDW_FAMILY_CPU; prm.recipe = FDW_RECIPE_C; prm.taper = FDW_TAPER_FOUR;
    prm.device = env_int("FDW_DEVICE", 0);
    prm.nt = nt;
    fdw_ctx *ctx = NULL;
    FDW(fdw_create(&prm, &ctx));
    FDW(fdw_set_v2(ctx, vel2));
    FDW(fdw_set_wavelet(ctx, srce, nt));
    float *data = xalloc((size_t)ns * nx * nt);
    const int sz = in.sz + in.nzb, gz = in.gz + in.nzb;
    double t0 = now_s();
    for (int is = 0; is < ns; is++) {
        const int sx = in.fsx + is * in.ds + in.nxb;
        fprintf(stdout, "** source %d, at (%d,%d) \n", is + 1, sx - in.nxb, sz - in.nzb);
        FDW(fdw_model_shot(ctx, sx, sz, gz, data + (size_t)is * nx * nt));
    }
    double dt = now_s() - t0;
    fprintf(stderr, "[fdwave] modelling: %.3f s (%.2f Gpts/s)\n", dt,
            (double)ns * nt * (nx + 2.0 * in.nxb) * (in.nz + 2.0 * in.nzb) / dt / 1e9);
    write_floats(in.datfile, data, (size_t)ns * nx * nt, "w+");
    fdw_destroy(ctx);
    free(srce); free(vel2); free(data);
    return 0;
}
