/*
 * fdw_host.c -- host-side tables and parameter-file readers of libfdwave.
 * Pure C, no CUDA: callable on a machine without a GPU.
 *
 * Each table reproduces the reference's value bit for bit, which means
 * reproducing which libm entry point (float or double) the reference's
 * compiler picked:
 *   GPU family host code  cuda_reference_RTM/lib/src/functions.c is C (gcc):
 *       exp/pow/cos in double;  but the sponge table is built inside
 *       fd-code.cu:159-166, C++ under nvcc: log/sqrt of a float are the float
 *       overloads, pow(float,int) promotes to double.
 *   CPU family  dpct_gpu_rtm_domain_division/src/... .c are built as C++ by
 *       g++ (src/timestep/Makefile:17): exp(float) -> expf, cos/pow on floats
 *       -> cosf/powf, pow(float,int) -> double.
 * Build without FMA contraction and without -ffast-math.
 */
#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fdwave.h"

#define FDW_PI (3.141592653589793) /* functions.h:7, cwp.h:126 */

void fdw_set_error(const char *fmt, ...);

/* ---- second-derivative weights ------------------------------------------ */
static void windowed_sinc(float *coef, int order, int family)
{
    /* makeo2: functions.c:125-157 (double libm) / fd.c:99-130 (float libm) */
    float alpha = .54, beta = 6.;
    float h_beta = 0.5 * beta, alpha1 = 2. * alpha - 1.0, alpha2 = 2. * (1.0 - alpha);
    float central = 0.0;
    int half = order / 2, sgn = -1;
    for (int k = 1; k <= half; k++) {
        float filt, arg, wind;
        sgn = -sgn;
        filt = (2. * sgn) / (k * k);
        arg = FDW_PI * k / (2. * (half + 2));
        if (family == FDW_FAMILY_GPU)
            wind = pow((alpha1 + alpha2 * cos(arg) * cos(arg)), h_beta);
        else
            wind = powf(alpha1 + alpha2 * cosf(arg) * cosf(arg), h_beta);
        coef[half + k] = filt * wind;
        coef[half - k] = coef[half + k];
        central = central + coef[half + k];
    }
    coef[half] = -2. * central;
}

int fdw_calc_coefs(int order, int family, float *c)
{
    if (!c || order < 2 || (order & 1) || order > 62) {
        fdw_set_error("fdw_calc_coefs: order must be even, 2..62 (got %d)", order);
        return FDW_ERR_ARG;
    }
    memset(c, 0, (size_t)(order + 1) * sizeof(float));
    switch (order) {
    case 2:
        c[0] = 1.; c[1] = -2.; c[2] = 1.;
        break;
    case 4:
        c[0] = c[4] = -1. / 12.; c[1] = c[3] = 4. / 3.; c[2] = -5. / 2.;
        break;
    case 6:
        c[0] = c[6] = 1. / 90.; c[1] = c[5] = -3. / 20.; c[2] = c[4] = 3. / 2.; c[3] = -49. / 18.;
        break;
    case 8:
        c[0] = c[8] = -1. / 560.; c[1] = c[7] = 8. / 315.; c[2] = c[6] = -1. / 5.; c[3] = c[5] = 8. / 5.;
        c[4] = -205. / 72.;
        break;
    default:
        windowed_sinc(c, order, family);
    }
    return FDW_OK;
}

/* ---- Ricker -------------------------------------------------------------- */
int fdw_ricker_wavelet(int nt, float dt, float fpeak, int family, float *s)
{
    if (!s || nt < 0) {
        fdw_set_error("fdw_ricker_wavelet: bad arguments");
        return FDW_ERR_ARG;
    }
    for (int it = 0; it < nt; it++) {
        if (family == FDW_FAMILY_CPU && it * dt > 2.0 / fpeak) { /* ptsrc.c:92-93 */
            s[it] = 0.0;
            continue;
        }
        float t = it * dt - 1.0 / fpeak; /* functions.c:297 / ptsrc.c:96 */
        float x = FDW_PI * fpeak * t;
        float xx = x * x;
        if (family == FDW_FAMILY_GPU)
            s[it] = exp(-xx) * (1.0 - 2.0 * xx); /* functions.c:290, double exp */
        else
            s[it] = expf(-xx) * (1.0 - 2.0 * xx); /* ptsrc.c:85 built as C++ */
    }
    return FDW_OK;
}

/* ---- sponge -------------------------------------------------------------- */
int fdw_taper_table(int nb, float fac, int family, float *tab)
{
    if (!tab || nb < 0) {
        fdw_set_error("fdw_taper_table: bad arguments");
        return FDW_ERR_ARG;
    }
    float slope = fac; /* taper.c:35: F used directly */
    if (family == FDW_FAMILY_GPU) slope = sqrtf(-logf(fac)) / (1. * nb); /* fd-code.cu:159,164 */
    for (int i = 0; i < nb; i++) {
        double a = (double)(slope * (nb - i));
        tab[i] = exp(-(a * a));
    }
    return FDW_OK;
}

/* ---- velocity extension --------------------------------------------------- */
int fdw_extendvel(int nx, int nz, int nxb, int nzb, float *vel)
{
    if (!vel || nx < 1 || nz < 1 || nxb < 0 || nzb < 0) {
        fdw_set_error("fdw_extendvel: bad arguments");
        return FDW_ERR_ARG;
    }
    const size_t nze = (size_t)nz + 2 * (size_t)nzb;
    const int nxe = nx + 2 * nxb;
    for (int ix = nxb; ix < nxb + nx; ix++) { /* z direction first, taper.c:11-16 */
        float *row = vel + ix * nze;
        const float top = row[nzb], bot = row[nzb + nz - 1];
        for (int iz = 0; iz < nzb; iz++) row[iz] = top;
        for (size_t iz = (size_t)nzb + nz; iz < nze; iz++) row[iz] = bot;
    }
    for (int ix = 0; ix < nxb; ix++) memcpy(vel + ix * nze, vel + nxb * nze, nze * sizeof(float));
    for (int ix = nxb + nx; ix < nxe; ix++)
        memcpy(vel + ix * nze, vel + (size_t)(nxb + nx - 1) * nze, nze * sizeof(float));
    return FDW_OK;
}

static float border_draw(float v, float frac_num, float frac_den)
{
    /* functions.c:313-314: a linear ramp from v down to 300 m/s, then a uniform
     * integer draw in [ramp-200, v+200] from libc rand(). */
    const float floor_v = 300., jitter = 200.;
    float ramp = v - (v - floor_v) * frac_num / frac_den;
    return rand() % (int)(v + jitter - (ramp - jitter) + 1) + ramp - jitter;
}

int fdw_extendvel_linear(int nx, int nz, int nxb, int nzb, float *vel)
{
    if (!vel || nx < 1 || nz < 1 || nxb < 2 || nzb < 2) {
        fdw_set_error("fdw_extendvel_linear: bad arguments");
        return FDW_ERR_ARG;
    }
    const size_t nze = (size_t)nz + 2 * (size_t)nzb;
    const int xl = nxb, xr = nxb + nx - 1, zt = nzb, zb = nzb + nz - 1;
    const int nxe = nx + 2 * nxb;
#define AT(ix, iz) vel[(size_t)(ix) * nze + (size_t)(iz)]
    for (int ix = xl; ix <= xr; ix++)
        for (int k = 0; k < nzb; k++) {
            AT(ix, k) = AT(ix, zt);
            AT(ix, zb + 1 + k) = border_draw(AT(ix, zb), k, nzb - 1);
        }
    for (int iz = zt; iz <= zb; iz++)
        for (int k = 0; k < nxb; k++) {
            AT(xl - 1 - k, iz) = border_draw(AT(xl, iz), k, nxb - 1);
            AT(xr + 1 + k, iz) = border_draw(AT(xr, iz), k, nxb - 1);
        }
    for (int k = 0; k < nzb; k++)
        for (int m = 0; m < nxb; m++) {
            AT(m, k) = AT(xl, k);
            AT(xr + 1 + m, k) = AT(xr, k);
        }
    for (int k = 0; k < nzb; k++) /* bottom-left triangle pair, functions.c:340-348 */
        for (int m = 0; m <= k; m++) {
            float v = AT(xl, zb);
            AT(m, nze - 1 - k) = border_draw(v, nxb - 1 - m, nzb - 1);
            AT(k, nze - 1 - m) = border_draw(v, nxb - 1 - m, nzb - 1);
        }
    for (int k = 0; k < nzb; k++) /* bottom-right, functions.c:350-358 */
        for (int m = 0; m <= k; m++) {
            float v = AT(xr, zb);
            AT(nxe - 1 - m, nze - 1 - k) = border_draw(v, nxb - 1 - m, nzb - 1);
            AT(nxe - 1 - k, nze - 1 - m) = border_draw(v, nxb - 1 - m, nzb - 1);
        }
#undef AT
    return FDW_OK;
}

int fdw_ptsrc_weights(float *w49)
{
    if (!w49) return FDW_ERR_ARG;
    for (int di = -3; di <= 3; di++)
        for (int dj = -3; dj <= 3; dj++) {
            float xn = di, zn = dj;
            w49[(di + 3) * 7 + (dj + 3)] = expf(-xn * xn - zn * zn); /* ptsrc.c:55 as C++ */
        }
    return FDW_OK;
}

/* ---- input.dat ------------------------------------------------------------ */
static void input_clear(fdw_input *in)
{
    memset(in, 0, sizeof(*in));
    in->nz = in->nx = in->nt = in->ns = in->sz = in->fsx = in->ds = in->gz = -1;
    in->order = in->nzb = in->nxb = in->iss = in->rnd = -1;
    in->dz = in->dx = in->dt = in->fpeak = in->fac = -1.0f;
}

static void input_defaults(fdw_input *in)
{
    /* fd-code.cu:367-377 == mod_main.cpp:76-85 */
    if (in->iss == -1) in->iss = 0;
    if (in->ns == -1) in->ns = 1;
    if (in->sz == -1) in->sz = 0;
    if (in->fsx == -1) in->fsx = 0;
    if (in->ds == -1) in->ds = 1;
    if (in->gz == -1) in->gz = 0;
    if (in->order == -1) in->order = 8;
    if (in->nzb == -1) in->nzb = 40;
    if (in->nxb == -1) in->nxb = 40;
    if (in->fac == -1.0f) in->fac = 0.7f;
}

/* value text of the first line containing `key` as a substring: the text
 * between the first and second '=' (strtok semantics, functions.c:21-25).
 * Returns 0 when no line matches or the line has no value. */
static int gpu_dialect_find(const char *path, const char *key, char *val, size_t cap)
{
    FILE *fp = fopen(path, "r");
    if (!fp) return -1;
    char *line = NULL;
    size_t len = 0;
    int found = 0;
    while (getline(&line, &len, fp) != -1) {
        if (!strstr(line, key)) continue;
        char *s = line;
        while (*s == '=') s++; /* strtok skips leading delimiters */
        char *eq = strchr(s, '=');
        if (eq) {
            eq++;
            while (*eq == '=') eq++;
            size_t n = strcspn(eq, "=");
            if (n > 0) {
                if (n >= cap) n = cap - 1;
                memcpy(val, eq, n);
                val[n] = 0;
                found = 1;
            }
        }
        break; /* first matching line wins, whatever it holds */
    }
    free(line);
    fclose(fp);
    return found;
}

static int gpu_int(const char *path, const char *key)
{
    char v[1024];
    return gpu_dialect_find(path, key, v, sizeof v) == 1 ? atoi(v) : -1;
}

static float gpu_float(const char *path, const char *key)
{
    char v[1024];
    return gpu_dialect_find(path, key, v, sizeof v) == 1 ? (float)atof(v) : -1.0f;
}

static int gpu_str(const char *path, const char *key, char *dst, size_t cap)
{
    char v[1024];
    dst[0] = 0;
    if (gpu_dialect_find(path, key, v, sizeof v) != 1) return 0;
    size_t n = strlen(v);
    if (n) v[n - 1] = 0; /* the reference drops the last character (the newline), functions.c:70 */
    snprintf(dst, cap, "%s", v);
    return 1;
}

int fdw_read_input_gpu(const char *path, int apply_defaults, fdw_input *in)
{
    if (!path || !in) return FDW_ERR_ARG;
    FILE *fp = fopen(path, "r");
    if (!fp) {
        fdw_set_error("cannot open %s", path);
        return FDW_ERR_IO;
    }
    fclose(fp);
    input_clear(in);
    gpu_str(path, "tmpdir", in->tmpdir, sizeof in->tmpdir);
    gpu_str(path, "vpfile", in->vpfile, sizeof in->vpfile);
    in->has_datfile = gpu_str(path, "datfile", in->datfile, sizeof in->datfile);
    in->has_vel_ext_file = gpu_str(path, "vel_ext_file", in->vel_ext_file, sizeof in->vel_ext_file);
    in->nz = gpu_int(path, "nz");
    in->nx = gpu_int(path, "nx");
    in->nt = gpu_int(path, "nt");
    in->ns = gpu_int(path, "ns");
    in->sz = gpu_int(path, "sz");
    in->fsx = gpu_int(path, "fsx");
    in->ds = gpu_int(path, "ds");
    in->gz = gpu_int(path, "gz");
    in->order = gpu_int(path, "order");
    in->nzb = gpu_int(path, "nzb");
    in->nxb = gpu_int(path, "nxb");
    in->iss = gpu_int(path, "iss");
    in->rnd = gpu_int(path, "rnd");
    in->dz = gpu_float(path, "dz");
    in->dx = gpu_float(path, "dx");
    in->dt = gpu_float(path, "dt");
    in->fpeak = gpu_float(path, "fpeak");
    in->fac = gpu_float(path, "fac");
    if (apply_defaults) input_defaults(in);
    return FDW_OK;
}

int fdw_read_input_stencil(const char *path, fdw_input *in)
{
    /* fd-source-code.cu:34-108: one pass, every line is tested against every key
     * (so a later line overrides an earlier one); nz / nx only accept lines
     * whose name part is at most 2 characters long. */
    if (!path || !in) return FDW_ERR_ARG;
    FILE *fp = fopen(path, "r");
    if (!fp) {
        fdw_set_error("cannot open %s", path);
        return FDW_ERR_IO;
    }
    input_clear(in);
    char *line = NULL;
    size_t len = 0;
    while (getline(&line, &len, fp) != -1) {
        char *eq = strchr(line, '=');
        size_t name_len = eq ? (size_t)(eq - line) : strlen(line);
        const char *val = eq ? eq + 1 : NULL;
        if (!val) continue;
        if (strstr(line, "tmpdir")) {
            snprintf(in->tmpdir, sizeof in->tmpdir, "%s", val);
            size_t n = strlen(in->tmpdir);
            if (n) in->tmpdir[n - 1] = 0;
        }
        if (strstr(line, "nzb")) in->nzb = atoi(val);
        if (strstr(line, "nxb")) in->nxb = atoi(val);
        if (strstr(line, "nz") && name_len <= 2) in->nz = atoi(val);
        if (strstr(line, "nx") && name_len <= 2) in->nx = atoi(val);
        if (strstr(line, "dz")) in->dz = (float)atof(val);
        if (strstr(line, "dx")) in->dx = (float)atof(val);
        if (strstr(line, "order")) in->order = atoi(val);
    }
    free(line);
    fclose(fp);
    return FDW_OK;
}

int fdw_read_input_cpu(const char *path, int apply_defaults, fdw_input *in)
{
    if (!path || !in) return FDW_ERR_ARG;
    FILE *fp = fopen(path, "r");
    if (!fp) {
        fdw_set_error("cannot open %s", path);
        return FDW_ERR_IO;
    }
    input_clear(in);
    char tok[2048];
    while (fscanf(fp, "%2047s", tok) == 1) { /* whitespace-separated name=value tokens */
        char *eq = strchr(tok, '=');
        if (!eq) continue;
        *eq = 0;
        const char *k = tok, *v = eq + 1;
#define STR(name, field)                                         \
    if (!strcmp(k, name)) {                                      \
        snprintf(in->field, sizeof in->field, "%s", v);          \
        continue;                                                \
    }
#define INT(name)                 \
    if (!strcmp(k, #name)) {      \
        in->name = atoi(v);       \
        continue;                 \
    }
#define FLT(name)                    \
    if (!strcmp(k, #name)) {         \
        in->name = (float)atof(v);   \
        continue;                    \
    }
        STR("tmpdir", tmpdir) STR("vpfile", vpfile) STR("vel_ext_file", vel_ext_file)
        if (!strcmp(k, "datfile")) {
            snprintf(in->datfile, sizeof in->datfile, "%s", v);
            in->has_datfile = 1;
            continue;
        }
        INT(nz) INT(nx) INT(nt) INT(ns) INT(sz) INT(fsx) INT(ds) INT(gz) INT(order) INT(nzb) INT(nxb) INT(iss) INT(rnd)
        FLT(dz) FLT(dx) FLT(dt) FLT(fpeak) FLT(fac)
#undef STR
#undef INT
#undef FLT
    }
    fclose(fp);
    in->has_vel_ext_file = in->vel_ext_file[0] != 0;
    if (apply_defaults) input_defaults(in);
    return FDW_OK;
}

/* ---- raw float32 files and the image.num text dump of the drop-in surface */
long long fdw_read_floats(const char *path, float *dst, long long n)
{
    if (!path || !dst || n < 0) return -1;
    FILE *f = fopen(path, "rb");
    if (!f) return -1;
    size_t got = fread(dst, sizeof(float), (size_t)n, f);
    fclose(f);
    return (long long)got;
}

int fdw_write_floats(const char *path, const float *src, long long n, int append)
{
    if (!path || !src || n < 0) return FDW_ERR_ARG;
    FILE *f = fopen(path, append ? "ab" : "wb");
    if (!f) {
        fdw_set_error("cannot open %s for writing", path);
        return FDW_ERR_IO;
    }
    const size_t put = fwrite(src, sizeof(float), (size_t)n, f);
    if (fclose(f) != 0 || put != (size_t)n) {
        fdw_set_error("short write to %s", path);
        return FDW_ERR_IO;
    }
    return FDW_OK;
}

int fdw_image_stack_shot(const char *path, int is, int nx, int nz, float *img, const float *imloc)
{
    if (!img || !imloc || nx < 1 || nz < 1) return FDW_ERR_ARG;
    FILE *f = NULL;
    if (path) {
        f = fopen(path, is == 0 ? "w" : "a");
        if (!f) {
            fdw_set_error("cannot open %s for writing", path);
            return FDW_ERR_IO;
        }
        fprintf(f, "======== %i ========\n", is);
    }
    for (int iz = 0; iz < nz; iz++)
        for (int ix = 0; ix < nx; ix++) {
            img[(size_t)ix * nz + iz] += imloc[(size_t)ix * nz + iz];
            if (f) fprintf(f, " %f \n", img[(size_t)ix * nz + iz]);
        }
    if (f && fclose(f) != 0) {
        fdw_set_error("short write to %s", path);
        return FDW_ERR_IO;
    }
    return FDW_OK;
}
