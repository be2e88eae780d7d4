/*
 * rtm_code -- drop-in for cuda_reference_RTM/src/fd-code.cu main() (:380-584):
 * reverse-time migration with random-velocity borders, two saved forward
 * levels and time-reversed source reconstruction.
 *
 *   ./rtm_code ./models/<m>/input.dat
 *
 * Same input.dat dialect and raw float32 files; writes <tmpdir>/dir.image,
 * <tmpdir>/dir.image_lap (zeros, as the reference), the empty dir.snaps*,
 * and image.num in the working directory.  The two saved levels never leave
 * the device.  FDW_COMPAT=0 switches the reference's truncated launch extents
 * (quirk Q1, fd-code.cu:185-195) off; the default reproduces them so that
 * images match the reference bit for bit.
 */
#include "common.h"

int main(int argc, char **argv)
{
    if (argc < 2) DIE("usage: %s input.dat", argv[0]);
    const double t_start = now_s();
    fdw_input in;
    FDW(fdw_read_input_gpu(argv[1], 1, &in));
    const int vel_ext_flag = in.has_vel_ext_file;
    printf("## vp = %s, d_obs = %s, vel_ext_file = %s, vel_ext_flag = %d \n", in.vpfile, in.datfile,
           vel_ext_flag ? in.vel_ext_file : "(null)", vel_ext_flag);
    printf("## nz = %d, nx = %d, nt = %d \n", in.nz, in.nx, in.nt);
    printf("## dz = %f, dx = %f, dt = %f \n", in.dz, in.dx, in.dt);
    printf("## ns = %d, sz = %d, fsx = %d, ds = %d, gz = %d \n", in.ns, in.sz, in.fsx, in.ds, in.gz);
    printf("## order = %d, nzb = %d, nxb = %d, F = %f, rnd = %d \n", in.order, in.nzb, in.nxb, in.fac, in.rnd);
    if (!in.has_datfile) DIE("input.dat names no datfile (the reference crashes here, fd-code.cu:422)");
    const int nx = in.nx, nz = in.nz, nt = in.nt, ns = in.ns, nxb = in.nxb, nzb = in.nzb;
    const int nxe = nx + 2 * nxb, nze = nz + 2 * nzb;
    const int sz = in.sz + nzb, gz = in.gz + nzb;
    const size_t ne = (size_t)nxe * nze, ni = (size_t)nx * nz, ntr = (size_t)nx * nt;

    float *srce = xalloc(nt);
    FDW(fdw_ricker_wavelet(nt, in.dt, in.fpeak, FDW_FAMILY_GPU, srce));
    float *vel_ext = NULL;
    if (vel_ext_flag) {
        vel_ext = xalloc(ne * ns);
        read_floats(in.vel_ext_file, vel_ext, ne * ns, 1);
    }
    float *d_obs = xalloc(ntr * ns);
    read_floats(in.datfile, d_obs, ntr * ns, 1);
    float *vp = xalloc(ni), *vpe = xalloc(ne), *vel2 = xalloc(ne);
    read_floats(in.vpfile, vp, ni, 1);
    for (int ix = 0; ix < nx; ix++)
        memcpy(vpe + (size_t)(ix + nxb) * nze + nzb, vp + (size_t)ix * nz, nz * sizeof(float));

    fdw_params prm;
    memset(&prm, 0, sizeof prm);
    prm.nx = nx; prm.nz = nz; prm.nxb = nxb; prm.nzb = nzb; prm.order = in.order;
    prm.dx = in.dx; prm.dz = in.dz; prm.dt = in.dt; prm.fac = in.fac;
    prm.family = FDW_FAMILY_GPU; prm.recipe = FDW_RECIPE_G; prm.taper = FDW_TAPER_TOP;
    prm.compat_extents = env_int("FDW_COMPAT", 1);
    prm.device = env_int("FDW_DEVICE", 0);
    prm.nt = nt;
    fdw_ctx *ctx = NULL;
    const double t_io = now_s() - t_start;
    FDW(fdw_create(&prm, &ctx));
    const double t_ctx = now_s() - t_start - t_io;
    FDW(fdw_set_wavelet(ctx, srce, nt));

    char path[1024];
    const char *empties[] = {"dir.snaps", "dir.snaps_rec", "dir.snapr"};
    for (int k = 0; k < 3; k++) {
        snprintf(path, sizeof path, "%s/%s", in.tmpdir, empties[k]);
        FILE *f = fopen(path, "w");
        if (f) fclose(f);
    }
    float *imloc = xalloc(ni), *img = xalloc(ni), *img_lap = xalloc(ni);
    double t_dev = 0;
    for (int is = 0; is < ns; is++) {
        const int sx = in.fsx + is * in.ds + nxb;
        fprintf(stdout, "** source %d, at (%d,%d) \n", is + 1, sx - nxb, sz - nzb);
        const float *v = vpe;
        if (vel_ext_flag)
            v = vel_ext + (size_t)is * ne;
        else
            FDW(fdw_extendvel_linear(nx, nz, nxb, nzb, vpe)); /* libc rand(), continuing across shots like the reference */
        for (size_t k = 0; k < ne; k++) vel2[k] = v[k] * v[k];
        double t0 = now_s();
        FDW(fdw_set_v2(ctx, vel2));
        FDW(fdw_forward(ctx, sx, sz, NULL, NULL));
        fprintf(stdout, "\n** backward propagation %d, at (%d,%d) \n", is + 1, sx - nxb, sz - nzb);
        FDW(fdw_backward(ctx, NULL, NULL, d_obs + (size_t)is * ntr, gz, imloc));
        t_dev += now_s() - t0;
        fprintf(stdout, "\n");
        FDW(fdw_image_stack_shot("image.num", is, nx, nz, img, imloc)); /* img += imloc + the text dump, fd-code.cu:521-528 */
    }
    printf("> Exec time = %.2f (s)\n", (double)(long)(now_s() - t_start)); /* whole seconds, like fd-code.cu:536 */
    fprintf(stderr, "[fdwave] %d shot(s), %d steps each: %.3f s in forward+backward (%.2f Gpts/s, 3 updates/pt/step); "
                    "input %.3f s, context %.3f s, total %.3f s\n",
            ns, nt, t_dev, 3.0 * ns * nt * (double)ne / t_dev / 1e9, t_io, t_ctx, now_s() - t_start);
    snprintf(path, sizeof path, "%s/dir.image", in.tmpdir);
    write_floats(path, img, ni, "w");
    /* dir.image_lap: the reference never computes it and writes zeros (fd-code.cu:477,542); FDW_IMAGE_LAP=1 fills it
     * with the filter of models/3lay_mod/laplace.f90 on the GPU */
    if (getenv("FDW_IMAGE_LAP") && atoi(getenv("FDW_IMAGE_LAP")) != 0)
        FDW(fdw_image_laplacian(nx, nz, in.dx, in.dz, img, img_lap, prm.device));
    snprintf(path, sizeof path, "%s/dir.image_lap", in.tmpdir);
    write_floats(path, img_lap, ni, "w");
    fdw_destroy(ctx);
    free(srce); free(vel_ext); free(d_obs); free(vp); free(vpe); free(vel2); free(imloc); free(img); free(img_lap);
    return 0;
}
