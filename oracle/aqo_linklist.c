/* aqo_linklist.c -- CPU ORACLE (test infrastructure, not product code).
 * Restates the LinkList / RadixSort / Sort / UnSort tools of AQUAgpusph 5.0.4.
 * See aqo.h for conventions. */
#include "aqo.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* CalcServer.cpp:245-257: an evaluated <Define> is solved as float, printed with
 * "%#G" (6 significant digits) and suffixed with "f"; the OpenCL compiler then
 * parses that decimal literal back to the nearest float. */
float aqo_define_round6(float v)
{
    char s[128];
    snprintf(s, sizeof(s), "%#G", (double)v);
    return strtof(s, NULL);
}

/* basic.xml:119-123: H = h, CONW = 1/(h^dims), CONF = 1/(h^(dims+2)), evaluated
 * by the tokenizer in double/long double (Tokenizer_exprtk.hpp:38,88-143) on the
 * float variable h, narrowed to float (CalcServer.cpp:250-252) then 6 digits. */
void aqo_make_defs(aqo_defs* d, int dims, float h)
{
    d->dims = dims;
    d->H = aqo_define_round6(h);
    d->CONW = aqo_define_round6((float)(1.0 / pow((double)h, (double)dims)));
    d->CONF = aqo_define_round6((float)(1.0 / pow((double)h, (double)(dims + 2))));
    d->SUPPORT = 2.f;
}

/* LinkList.cpp:80-91 + Reduction.hcl.in:73,90: component-wise min/max with
 * identities VEC_INFINITY / -VEC_INFINITY (w identity is 0 in 3D, so the w
 * component reduces against 0). min/max are order independent => exact. */
void aqo_minmax(const float* r, aqo_usize N, int dims, float* rmin, float* rmax)
{
    const int vs = (dims == 3) ? 4 : 2;
    for (int c = 0; c < vs; c++) {
        const int is_w = (dims == 3 && c == 3);
        rmin[c] = is_w ? 0.f : INFINITY;
        rmax[c] = is_w ? -0.f : -INFINITY;
    }
    for (aqo_usize i = 0; i < N; i++)
        for (int c = 0; c < vs; c++) {
            rmin[c] = fminf(rmin[c], r[(size_t)i * vs + c]);
            rmax[c] = fmaxf(rmax[c], r[(size_t)i * vs + c]);
        }
}

/* LinkList.cpp:153-156 (_cell_length = support * h, float product) and
 * LinkList.cpp:185-232 (n_a = (ulong)((max_a - min_a) / cell_length) + 6). */
int aqo_ncells(const float* rmin, const float* rmax, int dims, float support,
               float h, aqo_usize ncells[4])
{
    const float cell_length = support * h;
    if (!cell_length)
        return -1;
    uint64_t n[3] = { 1, 1, 1 };
    for (int c = 0; c < dims; c++)
        n[c] = (uint64_t)((rmax[c] - rmin[c]) / cell_length) + 6;
    ncells[0] = (aqo_usize)n[0];
    ncells[1] = (aqo_usize)n[1];
    ncells[2] = (aqo_usize)n[2];
    ncells[3] = (aqo_usize)(n[0] * n[1] * n[2]);
    return 0;
}

/* LinkList.cl.in:54-85 */
void aqo_icell(aqo_usize* icell, const float* r, aqo_usize N, int dims,
               const float* rmin, float support, float h,
               const aqo_usize ncells[4])
{
    const int vs = (dims == 3) ? 4 : 2;
    const float idist = 1.f / (support * h);
    for (aqo_usize i = 0; i < N; i++) {
        const float* ri = r + (size_t)i * vs;
        const aqo_usize cx = (aqo_usize)((ri[0] - rmin[0]) * idist) + 3u;
        const aqo_usize cy = (aqo_usize)((ri[1] - rmin[1]) * idist) + 3u;
        aqo_usize id = cx - 1u + (cy - 1u) * ncells[0];
        if (dims == 3) {
            const aqo_usize cz = (aqo_usize)((ri[2] - rmin[2]) * idist) + 3u;
            id += (cz - 1u) * ncells[0] * ncells[1];
        }
        icell[i] = id;
    }
}

/* RadixSort.cpp:129-303 + RadixSort.cl.in:35-323.  Only the observable
 * semantics are restated: a STABLE ascending sort of the keys (ties keep the
 * input order because every work-item owns a contiguous chunk and histograms
 * are laid out radix-major, RadixSort.cl.in:113-123,289-301); perm[k] = input
 * index of the k-th output (:295-296); inv_perm[perm[k]] = k (:313-323).  The
 * UINT_MAX padding up to n_radix (:35-52) never mixes with real keys.  A
 * bottom-up stable merge sort gives the identical permutation. */
void aqo_radix_sort(aqo_usize* keys, aqo_usize n, aqo_usize* perm,
                    aqo_usize* inv_perm)
{
    if (!n)
        return;
    aqo_usize* a = (aqo_usize*)malloc(sizeof(aqo_usize) * n);
    aqo_usize* b = (aqo_usize*)malloc(sizeof(aqo_usize) * n);
    for (aqo_usize i = 0; i < n; i++)
        a[i] = i;
    for (aqo_usize w = 1; w < n; w *= 2) {
        for (aqo_usize lo = 0; lo < n; lo += 2 * w) {
            aqo_usize mid = lo + w < n ? lo + w : n;
            aqo_usize hi = lo + 2 * w < n ? lo + 2 * w : n;
            aqo_usize i = lo, j = mid, k = lo;
            while (i < mid && j < hi)
                b[k++] = (keys[a[j]] < keys[a[i]]) ? a[j++] : a[i++];
            while (i < mid)
                b[k++] = a[i++];
            while (j < hi)
                b[k++] = a[j++];
        }
        aqo_usize* t = a;
        a = b;
        b = t;
    }
    for (aqo_usize k = 0; k < n; k++)
        b[k] = keys[a[k]];
    memcpy(keys, b, sizeof(aqo_usize) * n);
    if (perm)
        memcpy(perm, a, sizeof(aqo_usize) * n);
    if (inv_perm)
        for (aqo_usize k = 0; k < n; k++)
            inv_perm[a[k]] = k;
    free(a);
    free(b);
}

/* LinkList.cl.in:32-42 (iHoc: every cell = N) and :92-113 (linkList heads).
 * The reference launches linkList on N-1 items, so with N == 1 no head is
 * written at all (ihoc stays N everywhere); restated verbatim. */
void aqo_ihoc(const aqo_usize* icell, aqo_usize N, aqo_usize* ihoc,
              aqo_usize n_cells_w)
{
    for (aqo_usize c = 0; c < n_cells_w; c++)
        ihoc[c] = N;
    if (N < 2)
        return;
    for (aqo_usize i = 0; i + 1 < N; i++) {
        const aqo_usize c = icell[i], c2 = icell[i + 1];
        if (i == 0)
            ihoc[c] = 0;
        if (c2 != c)
            ihoc[c2] = i + 1;
    }
}

/* LinkList.cpp:326-494 */
int aqo_linklist(const float* r, aqo_usize N, int dims, float support, float h,
                 int recompute_grid, float* rmin, float* rmax,
                 aqo_usize ncells[4], aqo_usize* icell, aqo_usize* ihoc,
                 size_t ihoc_capacity, aqo_usize* perm, aqo_usize* inv_perm)
{
    if (recompute_grid)
        aqo_minmax(r, N, dims, rmin, rmax);
    if (aqo_ncells(rmin, rmax, dims, support, h, ncells))
        return -2;
    if ((size_t)ncells[3] > ihoc_capacity)
        return -1;
    aqo_icell(icell, r, N, dims, rmin, support, h, ncells);
    aqo_radix_sort(icell, N, perm, inv_perm);
    aqo_ihoc(icell, N, ihoc, ncells[3]);
    return 0;
}

/* basic/Sort.cl:57-78,102-124 (X[id_sorted[i]] = X_in[i]) and
 * UnSort.cl.in:30-42 (out[id[i]] = in[i]) are the same scatter. */
void aqo_scatter(void* out, const void* in, const aqo_usize* idx, aqo_usize N,
                 size_t elem_bytes)
{
    for (aqo_usize i = 0; i < N; i++)
        memcpy((char*)out + (size_t)idx[i] * elem_bytes,
               (const char*)in + (size_t)i * elem_bytes, elem_bytes);
}
