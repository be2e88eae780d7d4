/* apps/common.h -- small helpers shared by the C drivers. */
#ifndef FDW_APPS_COMMON_H
#define FDW_APPS_COMMON_H
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#include "fdwave.h"

#define DIE(...)                                  \
    do {                                          \
        fprintf(stderr, __VA_ARGS__);             \
        fprintf(stderr, "\n");                    \
        exit(EXIT_FAILURE);                       \
    } while (0)

#define FDW(call)                                                                  \
    do {                                                                           \
        int rc_ = (call);                                                          \
        if (rc_ != FDW_OK) DIE("%s failed (%d): %s", #call, rc_, fdw_last_error()); \
    } while (0)

static float *xalloc(size_t n)
{
    float *p = (float *)calloc(n ? n : 1, sizeof(float));
    if (!p) DIE("out of host memory (%zu floats)", n);
    return p;
}

/* like the reference: a short read leaves the rest of the (zeroed) buffer untouched */
static size_t read_floats(const char *path, float *dst, size_t n, int must_exist)
{
    FILE *f = fopen(path, "rb");
    if (!f) {
        if (must_exist) DIE("cannot open %s", path);
        return 0;
    }
    size_t got = fread(dst, sizeof(float), n, f);
    fclose(f);
    return got;
}

static void write_floats(const char *path, const float *src, size_t n, const char *mode)
{
    FILE *f = fopen(path, mode);
    if (!f) DIE("cannot open %s for writing", path);
    if (fwrite(src, sizeof(float), n, f) != n) DIE("short write to %s", path);
    fclose(f);
}

static double now_s(void)
{
    struct timeval t;
    gettimeofday(&t, NULL);
    return t.tv_sec + 1e-6 * t.tv_usec;
}

static int env_int(const char *name, int dflt)
{
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}
#endif
