/* shim_gpufam.cpp -- libfdwave_gpufam.so: the reference GPU family's function names
 * (include/fdwave_gpufam.h) on top of the C ABI. */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "fdwave.h"
#include "fdwave_gpufam.h"

namespace {
struct State {
    fdw_ctx *ctx = nullptr;
    int order = 8, nxe = 0, nze = 0, nxb = 0, nzb = 0, nt = 0, ns = 1;
    float fac = 0.7f, dx = 0, dz = 0, dt = 0;
} S;

void die(const char *what)
{
    fprintf(stderr, "libfdwave_gpufam: %s: %s\n", what, fdw_last_error());
    exit(EXIT_FAILURE);
}

void ensure_ctx()
{
    if (S.ctx) return;
    fdw_params p;
    memset(&p, 0, sizeof p);
    p.nx = S.nxe - 2 * S.nxb; p.nz = S.nze - 2 * S.nzb; p.nxb = S.nxb; p.nzb = S.nzb;
    p.order = S.order; p.dx = S.dx; p.dz = S.dz; p.dt = S.dt; p.fac = S.fac; p.nt = S.nt;
    p.family = FDW_FAMILY_GPU; p.recipe = FDW_RECIPE_G; p.taper = FDW_TAPER_TOP;
    p.compat_extents = getenv("FDW_COMPAT") ? atoi(getenv("FDW_COMPAT")) : 1;
    p.device = getenv("FDW_DEVICE") ? atoi(getenv("FDW_DEVICE")) : 0;
    if (fdw_create(&p, &S.ctx) != FDW_OK) die("fd_init");
}
} // namespace

extern "C" void fd_init_cuda(int order, int nxe, int nze, int nxb, int nzb, int nt, int ns, float fac)
{
    S.order = order; S.nxe = nxe; S.nze = nze; S.nxb = nxb; S.nzb = nzb; S.nt = nt; S.ns = ns; S.fac = fac;
}

extern "C" void fd_init(int order, int nx, int nz, int nxb, int nzb, int nt, int ns, float fac, float dx, float dz,
                        float dt)
{
    if (S.ctx) { fdw_destroy(S.ctx); S.ctx = nullptr; }
    S.dx = dx; S.dz = dz; S.dt = dt;
    fd_init_cuda(order, nx, nz, nxb, nzb, nt, ns, fac);
    ensure_ctx();
}

void write_buffers(float **, float **, float **, float *, float *, float **, float **, int, int)
{
    /* device buffers are owned by the library context; fd_forward / fd_back upload what they need */
}

void fd_forward(int order, float **p, float **pp, float **v2, int nz, int nx, int nt, int is, int sz, int *sx,
                float *srce, int)
{
    if (!S.ctx || order != S.order || nx != S.nxe || nz != S.nze || nt != S.nt) {
        fprintf(stderr, "libfdwave_gpufam: fd_forward arguments do not match fd_init\n");
        exit(EXIT_FAILURE);
    }
    if (fdw_set_v2(S.ctx, v2[0]) != FDW_OK) die("fd_forward/set_v2");
    if (fdw_set_wavelet(S.ctx, srce, nt) != FDW_OK) die("fd_forward/set_wavelet");
    if (fdw_set_source(S.ctx, sx[is], sz, FDW_SRC_POINT) != FDW_OK) die("fd_forward/set_source");
    /* the reference swaps first (fd-code.cu:260-262): its stencil input is pp */
    if (fdw_propagate(S.ctx, pp[0], p[0], 0, nt) != FDW_OK) die("fd_forward");
}

void fd_back(int order, float **, float **, float **, float **, float **v2, int nz, int nx, int nt, int is, int,
             int gz, float ***snaps, float **imloc, float **d_obs)
{
    if (!S.ctx || order != S.order || nx != S.nxe || nz != S.nze || nt != S.nt) {
        fprintf(stderr, "libfdwave_gpufam: fd_back arguments do not match fd_init\n");
        exit(EXIT_FAILURE);
    }
    if (fdw_set_v2(S.ctx, v2[0]) != FDW_OK) die("fd_back/set_v2");
    /* p, pp, pr, ppr arrive zeroed in the reference (fd-code.cu:510-515); imloc accumulates from 0 */
    if (fdw_backward(S.ctx, snaps[0][0], snaps[1][0], d_obs[is], gz, imloc[0]) != FDW_OK) die("fd_back");
}
