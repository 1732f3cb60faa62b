/* mc33_internal.h -- shared between the plain-C files of the drop-in library
 * (mc33_api.c, mc33_io.c).  Not installed; nothing outside csrc/ includes it. */
#ifndef MC33_INTERNAL_H
#define MC33_INTERNAL_H
#include <stddef.h>
#include "../../include/marching_cubes_33.h"

#if defined(__GNUC__)
#define MC33_HIDDEN __attribute__((visibility("hidden")))
#else
#define MC33_HIDDEN
#endif

/* result arrays of a surface: page-locked pool memory (include/mc33cu.h
 * mc33cu_host_alloc) with plain malloc as the fallback; either kind goes back
 * through mc33_result_free */
MC33_HIDDEN void *mc33_result_alloc(size_t bytes);
MC33_HIDDEN void mc33_result_free(void *p);

/* _GRD.internal_data values: 0 caller-owned samples (grid_from_data_pointer),
 * 1 one malloc per x-row (alloc_F, as in the reference MC33_util_grd.c:147-169),
 * 2 ONE contiguous block (readers): F[0][0] is its base, rows point into it */
#define MC33_GRD_ROWS 1
#define MC33_GRD_BLOCK 2

/* allocate Z->F for Z->N as row-pointer tables over one contiguous x-fastest
 * block taken from the page-locked pool (so that the upload of the whole grid is
 * a single DMA at link speed); 0 on success */
MC33_HIDDEN int mc33_alloc_F_block(_GRD *Z);

#endif
