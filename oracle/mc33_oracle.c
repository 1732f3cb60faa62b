/* mc33_oracle.c -- TEST INFRASTRUCTURE ONLY (see mc33_oracle.h).
 * Instantiates the type-generic restatement for the five element types the
 * reference can be compiled for (include/marching_cubes_33.h:66-88).
 * Parity status: pinned against oracle/_ref (the compiled reference) by
 * tests/test_oracle_vs_reference.py. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "mc33_oracle.h"
#include "../mc33_c_library_b200/csrc/mc33_tables.h"

void mc33o_free(mc33o_mesh *m)
{
	if (!m) return;
	free(m->V); free(m->N); free(m->T); free(m->vkey); free(m->tcell); free(m->tpat);
	memset(m, 0, sizeof(*m));
}

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

#define REAL float
#define SAMPLE float
#define DIFF_T float
#define FN(x) CAT(x, _f32)
#include "mc33_oracle_impl.h"
#undef REAL
#undef SAMPLE
#undef DIFF_T
#undef FN

#define REAL double
#define SAMPLE double
#define DIFF_T double
#define FN(x) CAT(x, _f64)
#include "mc33_oracle_impl.h"
#undef REAL
#undef SAMPLE
#undef DIFF_T
#undef FN

#define REAL float
#define SAMPLE uint8_t
#define DIFF_T int
#define FN(x) CAT(x, _u8)
#include "mc33_oracle_impl.h"
#undef REAL
#undef SAMPLE
#undef DIFF_T
#undef FN

#define REAL float
#define SAMPLE uint16_t
#define DIFF_T int
#define FN(x) CAT(x, _u16)
#include "mc33_oracle_impl.h"
#undef REAL
#undef SAMPLE
#undef DIFF_T
#undef FN

#define REAL float
#define SAMPLE uint32_t
#define DIFF_T unsigned
#define FN(x) CAT(x, _u32)
#include "mc33_oracle_impl.h"
#undef REAL
#undef SAMPLE
#undef DIFF_T
#undef FN

static int dispatch(int dtype, const void *data, uint32_t nx, uint32_t ny, uint32_t nz, double iso,
                    const mc33o_geom *g, int count_only, mc33o_mesh *out, uint16_t *pat)
{
	if (!data || !nx || !ny || !nz) return -1;
	switch (dtype) {
	case MC33O_F32: return extract_f32((const float *)data, nx, ny, nz, iso, g, count_only, out, pat);
	case MC33O_F64: return extract_f64((const double *)data, nx, ny, nz, iso, g, count_only, out, pat);
	case MC33O_U8:  return extract_u8((const uint8_t *)data, nx, ny, nz, iso, g, count_only, out, pat);
	case MC33O_U16: return extract_u16((const uint16_t *)data, nx, ny, nz, iso, g, count_only, out, pat);
	case MC33O_U32: return extract_u32((const uint32_t *)data, nx, ny, nz, iso, g, count_only, out, pat);
	}
	return -1;
}

int mc33o_extract(int dtype, const void *data, uint32_t nx, uint32_t ny, uint32_t nz,
                  double iso, const mc33o_geom *g, int count_only, mc33o_mesh *out)
{
	mc33o_geom ident;
	if (!out) return -1;
	if (!g) {
		memset(&ident, 0, sizeof ident);
		ident.D[0] = ident.D[1] = ident.D[2] = 1.0;
		ident.ca = ident.cb = 1.0;
		ident.A[0] = ident.A[4] = ident.A[8] = 1.0;
		ident.Ai[0] = ident.Ai[4] = ident.Ai[8] = 1.0;
		g = &ident;
	}
	return dispatch(dtype, data, nx, ny, nz, iso, g, count_only, out, 0);
}

int mc33o_cell_patterns(int dtype, const void *data, uint32_t nx, uint32_t ny, uint32_t nz,
                        double iso, uint16_t *pat_out)
{
	mc33o_mesh tmp;
	mc33o_geom ident;
	if (!pat_out) return -1;
	memset(&ident, 0, sizeof ident);
	return dispatch(dtype, data, nx, ny, nz, iso, &ident, 1, &tmp, pat_out);
}
