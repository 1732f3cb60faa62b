/* mc33_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, single thread) of the reference's
 * calculate_isosurface path (reference: source/marching_cubes_33.c:329-1889,
 * source/MC33_util_grd.c:87-114), written order-free: instead of the reference's
 * slice-to-slice vertex reuse tables (Dx,Dy,Ux,Uy,Lz) every vertex has exactly
 * one owner and the mesh comes out in this project's
 * canonical order.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may use it; nothing under mc33_c_library_b200/ links it.
 *
 * Parity status: PINNED -- tests/test_oracle_vs_reference.py checks this
 * restatement against the compiled, unmodified reference (the shared objects under oracle/_ref/) on
 * every element type and store variant, and against the known-answer counts in
 * BASELINE.md (K1..K7) committed under tests/golden/.
 *
 * Canonical order produced here (and by the CUDA path):
 *   vertices : first all SHARED vertices, by point row in (z,y) order; within
 *              a row first the X plane by x (X edge p -> p+ex, or the POINT
 *              vertex of a sample exactly on the isovalue), then the Y plane
 *              (edges p -> p+ey), then the Z plane (p -> p+ez);
 *              then all CENTRE vertices by owning cell in (z,y,x) order.
 *   triangles: by cell in (z,y,x) order -- the reference's sweep order
 *              (marching_cubes_33.c:1832-1865) -- then in table order, with the
 *              reference's winding (marching_cubes_33.c:1246-1250).
 * Consequently T[] lists the same triangles in the same order as the
 * reference; only the vertex numbering differs.
 */
#ifndef MC33_ORACLE_H
#define MC33_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { MC33O_F32 = 0, MC33O_F64 = 1, MC33O_U8 = 2, MC33O_U16 = 3, MC33O_U32 = 4 };

/* store variants of the reference: marching_cubes_33.c:485-621 */
enum { MC33O_SPN0 = 0, MC33O_SPNA = 1, MC33O_SPNB = 2, MC33O_SPNC = 3 };

/* Geometry exactly as create_MC33 leaves it in the MC33 struct
 * (marching_cubes_33.c:1762-1782): O, D, ca, cb already narrowed to MC33_real
 * (passed here as doubles holding those values), A = G->_A[j][i]*d[i],
 * Ai = G->A_[j][i]/d[j]. */
typedef struct {
	int store;          /* MC33O_SPN*                                   */
	int normal_neg;     /* MC33_NORMAL_NEG build option                 */
	int tsa;            /* 1: mult_Abf == _multTSA_bf (upper triangular)*/
	int pad_;
	double O[3], D[3], ca, cb;
	double A[9], Ai[9]; /* row major 3x3                                */
} mc33o_geom;

typedef struct {
	uint64_t nV, nT;        /* totals                                       */
	uint64_t nShared;       /* vertices [0,nShared) are EDGE/POINT vertices */
	uint64_t nCentre;       /* vertices [nShared,nV) are cell centres       */
	uint64_t nPoint;        /* how many of the shared ones are POINT        */
	uint64_t nActive;       /* cells with index != 0, 0xFF                  */
	void *V;                /* nV x 3 MC33_real (float, or double for F64)  */
	float *N;               /* nV x 3                                       */
	uint32_t *T;            /* nT x 3                                       */
	uint64_t *vkey;         /* nV: (linear point or cell id)*4 + slot (0 X/POINT,1 Y,2 Z,3 CENTRE) */
	uint64_t *tcell;        /* nT: linear cell id (z*ny + y)*nx + x         */
	uint16_t *tpat;         /* nT: pattern start in MC33_TRI of the owning cell */
} mc33o_mesh;

/* data: contiguous samples, x fastest, (nx+1)*(ny+1)*(nz+1) of them.
 * nx,ny,nz are INTERVAL counts as in _GRD.N (include/marching_cubes_33.h).
 * count_only != 0: fill only the counters.  Returns 0, or -1 on bad args / OOM. */
int mc33o_extract(int dtype, const void *data, uint32_t nx, uint32_t ny, uint32_t nz,
                  double iso, const mc33o_geom *g, int count_only, mc33o_mesh *out);
void mc33o_free(mc33o_mesh *m);

/* per-cell pattern only (for debugging / per-cell parity): writes, for every
 * cell, the pattern start (0xFFFF for inactive cells). */
int mc33o_cell_patterns(int dtype, const void *data, uint32_t nx, uint32_t ny, uint32_t nz,
                        double iso, uint16_t *pat_out);

#ifdef __cplusplus
}
#endif
#endif
