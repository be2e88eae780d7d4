/* apps/cpu_family_args.h -- `par=<file>` + name=value command-line handling of the
 * CPU-family drivers (CWP initargs/getpar: later definitions win,
 * lib/cwp/src/par/lib/getpars.c:447-453). */
#ifndef FDW_CPU_FAMILY_ARGS_H
#define FDW_CPU_FAMILY_ARGS_H
#include "common.h"

static void cpu_family_input(int argc, char **argv, fdw_input *in)
{
    /* concatenate the par file (if any) and the other arguments into one token file, in order */
    char tmpl[] = "/tmp/fdwparXXXXXX";
    int fd = mkstemp(tmpl);
    if (fd < 0) DIE("mkstemp failed");
    FILE *out = fdopen(fd, "w");
    for (int i = 1; i < argc; i++) {
        if (!strncmp(argv[i], "par=", 4)) {
            FILE *f = fopen(argv[i] + 4, "r");
            if (!f) DIE("cannot open par file %s", argv[i] + 4);
            int ch;
            while ((ch = fgetc(f)) != EOF) fputc(ch, out);
            fputc('\n', out);
            fclose(f);
        } else {
            fprintf(out, "%s\n", argv[i]);
        }
    }
    fclose(out);
    FDW(fdw_read_input_cpu(tmpl, 1, in));
    remove(tmpl);
    /* MUSTGETPAR* keys, mod_main.cpp:65-74 */
    if (!in->tmpdir[0] || !in->vpfile[0] || !in->has_datfile || in->nz < 0 || in->nx < 0 || in->nt < 0 ||
        in->dz < 0 || in->dx < 0 || in->dt < 0 || in->fpeak < 0)
        DIE("must specify tmpdir vpfile datfile nz nx nt dz dx dt fpeak");
}

static void cpu_family_banner(const fdw_input *in)
{
    fprintf(stdout, "## vp = %s \n", in->vpfile);
    fprintf(stdout, "## nz = %d, nx = %d, nt = %d \n", in->nz, in->nx, in->nt);
    fprintf(stdout, "## dz = %f, dx = %f, dt = %f \n", in->dz, in->dx, in->dt);
    fprintf(stdout, "## ns = %d, sz = %d, fsx = %d, ds = %d, gz = %d \n", in->ns, in->sz, in->fsx, in->ds, in->gz);
    fprintf(stdout, "## order = %d, nzb = %d, nxb = %d, F = %f \n", in->order, in->nzb, in->nxb, in->fac);
}

/* vel2 = vp^2 embedded + constant extension (mod_main.cpp:113-126) */
static float *cpu_family_vel2(const fdw_input *in)
{
    const int nx = in->nx, nz = in->nz, nxb = in->nxb, nzb = in->nzb;
    const int nze = nz + 2 * nzb, nxe = nx + 2 * nxb;
    float *vp = xalloc((size_t)nx * nz), *vel2 = xalloc((size_t)nxe * nze);
    read_floats(in->vpfile, vp, (size_t)nx * nz, 1);
    for (int ix = 0; ix < nx; ix++)
        for (int iz = 0; iz < nz; iz++) {
            float v = vp[(size_t)ix * nz + iz];
            vel2[(size_t)(ix + nxb) * nze + iz + nzb] = v * v;
        }
    FDW(fdw_extendvel(nx, nz, nxb, nzb, vel2));
    free(vp);
    return vel2;
}
#endif
