/* shim_cpufam.cpp -- libfdwave_cpufam.so: the reference CPU family's function names
 * (include/fdwave_cpufam.h) on top of the C ABI.  C++ on purpose: the reference's
 * callers are built by g++ and bind the mangled names. */
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "fdwave.h"
#include "fdwave_cpufam.h"

namespace {
struct State {
    fdw_ctx *ctx = nullptr;
    int order = 0, nxe = 0, nze = 0;
    float dx = 0, dz = 0, dt = 0;
    std::vector<float> tx, tz;
    const float *v2_ptr = nullptr; /* velocity array of the last upload ... */
    unsigned long long v2_hash = 0; /* ... and its content hash */
} S;

/* 64-bit multiply-xor hash over the array's bytes (one pass at memory speed; an upload costs a PCIe
 * transfer, a device kernel and a synchronisation) */
unsigned long long hash_floats(const float *a, size_t n)
{
    unsigned long long h = 0x9E3779B97F4A7C15ull;
    const unsigned *w = (const unsigned *)a;
    for (size_t i = 0; i < n; i++) {
        h ^= w[i];
        h *= 0x100000001B3ull;
        h ^= h >> 29;
    }
    return h;
}

void die(const char *what)
{
    fprintf(stderr, "libfdwave_cpufam: %s: %s\n", what, fdw_last_error());
    exit(EXIT_FAILURE);
}
} // namespace

void fd_init(int order, int nx, int nz, float dx, float dz, float dt)
{
    if (S.ctx) fdw_destroy(S.ctx);
    fdw_params p;
    memset(&p, 0, sizeof p);
    p.nx = nx; p.nz = nz; /* already the extended grid: no border known at this level */
    p.order = order; p.dx = dx; p.dz = dz; p.dt = dt; p.fac = 0.5f;
    p.family = FDW_FAMILY_CPU; p.recipe = FDW_RECIPE_C; p.taper = FDW_TAPER_NONE;
    p.device = getenv("FDW_DEVICE") ? atoi(getenv("FDW_DEVICE")) : 0;
    if (fdw_create(&p, &S.ctx) != FDW_OK) die("fd_init");
    S.v2_ptr = nullptr;
    S.order = order; S.nxe = nx; S.nze = nz; S.dx = dx; S.dz = dz; S.dt = dt;
}

void fd_step(int order, float **p, float **pp, float **v2, int nz, int nx)
{
    if (!S.ctx || order != S.order || nx != S.nxe || nz != S.nze) {
        fprintf(stderr, "libfdwave_cpufam: fd_step(order=%d,%dx%d) does not match fd_init(order=%d,%dx%d)\n", order, nx,
                nz, S.order, S.nxe, S.nze);
        exit(EXIT_FAILURE);
    }
    /* the velocity is uploaded (and premultiplied by dt^2 on the device) only when it changed: same
     * array and same content hash as at the previous call = the device copy is current */
    const size_t n = (size_t)nx * nz;
    const unsigned long long h = hash_floats(v2[0], n);
    if (v2[0] != S.v2_ptr || h != S.v2_hash) {
        if (fdw_set_v2(S.ctx, v2[0]) != FDW_OK) die("fd_step/set_v2");
        S.v2_ptr = v2[0];
        S.v2_hash = h;
    }
    /* p = stencil input (newest), pp = older level in / new level out: the new level is downloaded
     * straight into pp, p stays untouched on the host (fd.c:39-43) */
    if (fdw_fields_upload(S.ctx, 0, p[0], pp[0]) != FDW_OK) die("fd_step/upload");
    if (fdw_advance(S.ctx, 0, 1) != FDW_OK) die("fd_step/advance");
    if (fdw_fields_download(S.ctx, 0, pp[0], nullptr) != FDW_OK) die("fd_step/download");
}

void fd_destroy()
{
    if (S.ctx) fdw_destroy(S.ctx);
    S.ctx = nullptr;
}

float *calc_coefs(int order)
{
    float *c = (float *)calloc(order + 1, sizeof(float));
    fdw_calc_coefs(order, FDW_FAMILY_CPU, c);
    return c;
}

void extendvel(int nx, int nz, int nxb, int nzb, float *vel) { fdw_extendvel(nx, nz, nxb, nzb, vel); }

void taper_init(int nxb, int nzb, float F)
{
    S.tx.assign(nxb > 0 ? nxb : 1, 1.0f);
    S.tz.assign(nzb > 0 ? nzb : 1, 1.0f);
    fdw_taper_table(nxb, F, FDW_FAMILY_CPU, S.tx.data());
    fdw_taper_table(nzb, F, FDW_FAMILY_CPU, S.tz.data());
}

void taper_apply(float **a, int nx, int nz, int nxb, int nzb)
{
    const int nxe = nx + 2 * nxb, nze = nz + 2 * nzb;
    for (int i = 0; i < nxe; i++)
        for (int j = 0; j < nzb; j++) {
            a[i][j] *= S.tz[j];
            a[i][nze - 1 - j] *= S.tz[j];
        }
    for (int j = 0; j < nze; j++)
        for (int i = 0; i < nxb; i++) {
            a[i][j] *= S.tx[i];
            a[nxe - 1 - i][j] *= S.tx[i];
        }
}

void taper_apply2(float **a, int nx, int nz, int nxb, int nzb)
{
    const int nxe = nx + 2 * nxb;
    (void)nz;
    for (int i = 0; i < nxe; i++)
        for (int j = 0; j < nzb; j++) a[i][j] *= S.tz[j];
    for (int i = 0; i < nxb; i++)
        for (int j = 0; j < nzb; j++) {
            a[i][j] *= S.tx[i];
            a[nxe - 1 - i][j] *= S.tx[i];
        }
}

void taper_destroy()
{
    S.tx.clear();
    S.tz.clear();
}

void ptsrc(int xs, int zs, int nx, int nz, float ts, float **s)
{
    float w[49];
    fdw_ptsrc_weights(w);
    for (int di = -3; di <= 3; di++)
        for (int dj = -3; dj <= 3; dj++) {
            int i = xs + di, j = zs + dj;
            if (i < 0 || i > nx - 1 || j < 0 || j > nz - 1) continue;
            s[i][j] += ts * w[(di + 3) * 7 + (dj + 3)];
        }
}

void ricker_wavelet(int ns, float dt, float peak, float *s) { fdw_ricker_wavelet(ns, dt, peak, FDW_FAMILY_CPU, s); }

float ricker(float t, float fpeak)
{
    float x = 3.141592653589793 * fpeak * t;
    float xx = x * x;
    return expf(-xx) * (1.0 - 2.0 * xx);
}
