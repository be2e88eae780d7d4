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

/* raw float32 files: the library's readers / writers (fdw_read_floats / fdw_write_floats, csrc/fdw_host.c).
 * Like the reference, a short read leaves the rest of the (zeroed) buffer untouched. */
static size_t read_floats(const char *path, float *dst, size_t n, int must_exist)
{
    const long long got = fdw_read_floats(path, dst, (long long)n);
    if (got < 0) {
        if (must_exist) DIE("cannot open %s", path);
        return 0;
    }
    return (size_t)got;
}

static void write_floats(const char *path, const float *src, size_t n, const char *mode)
{
    if (fdw_write_floats(path, src, (long long)n, mode[0] == 'a') != FDW_OK) DIE("%s", fdw_last_error());
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
