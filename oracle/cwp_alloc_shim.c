/*
 * cwp_alloc_shim.c -- TEST INFRASTRUCTURE ONLY.
 *
 * The reference's CPU-family objects (fd.c, taper.c) call four CWP/SU
 * allocator functions (cwp/lib/alloc.c: alloc1float, alloc2float,
 * free1float, free2float).  To load those objects as a shared library for
 * function-level parity tests without linking the vendored non-PIC CWP
 * archives, this file provides the same four entry points with the CWP
 * contract: alloc2float(n1,n2) returns n2 row pointers into one contiguous
 * n1*n2 block, p[0] is the block base.
 */
#include <stdlib.h>

float *alloc1float(size_t n1) { return (float *)malloc(n1 * sizeof(float)); }

float **alloc2float(size_t n1, size_t n2)
{
    float **rows = (float **)malloc(n2 * sizeof(float *));
    if (!rows) return NULL;
    rows[0] = (float *)malloc(n1 * n2 * sizeof(float));
    if (!rows[0]) {
        free(rows);
        return NULL;
    }
    for (size_t i = 1; i < n2; i++) rows[i] = rows[0] + i * n1;
    return rows;
}

void free1float(float *p) { free(p); }

void free2float(float **p)
{
    free(p[0]);
    free(p);
}
