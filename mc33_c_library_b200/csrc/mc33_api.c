/*
 * mc33_api.c -- the marching_cubes_33.h API in plain C on top of the CUDA C-ABI
 * (include/mc33cu.h).  Compiled once per element-type variant, exactly like the
 * reference library is (reference source/libMC33.c:17-34 and
 * include/marching_cubes_33.h:57-88):
 *
 *   default                                   float grid,   MC33_real float
 *   -DGRD_TYPE_SIZE=8                         double grid,  MC33_real double
 *   -DINTEGER_GRD -DGRD_TYPE_SIZE={1,2,4}     u8/u16/u32,   MC33_real float
 *   -DGRD_ORTHOGONAL                          no inclined-grid members
 *   -DMC33_NORMAL_NEG=1  -DDEFAULT_SURFACE_COLOR=0x..   as in the reference
 *
 * What each function replaces is cited at its definition.  Nothing here does
 * marching cubes arithmetic: the host side wraps grids, snapshots geometry,
 * moves bytes and owns the malloc'ed result arrays.
 */
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "mc33_internal.h"
#include "../../include/mc33cu.h"

#ifndef DEFAULT_SURFACE_COLOR
#define DEFAULT_SURFACE_COLOR 0xff5c5c5c
#endif
#ifndef MC33_NORMAL_NEG
#define MC33_NORMAL_NEG 0
#endif

#if defined(INTEGER_GRD)
#  if GRD_TYPE_SIZE == 1
#    define MC33_DTYPE MC33CU_U8
#  elif GRD_TYPE_SIZE == 2
#    define MC33_DTYPE MC33CU_U16
#  else
#    define MC33_DTYPE MC33CU_U32
#  endif
#elif GRD_TYPE_SIZE == 8
#  define MC33_DTYPE MC33CU_F64
#else
#  define MC33_DTYPE MC33CU_F32
#endif

/* reference marching_cubes_33.c:76-80 */
int DefaultColorMC = (int)DEFAULT_SURFACE_COLOR;

/* The MC33 handed to the caller is the head of this private record.  The grid is cut into
 * nslab contiguous z-slabs of cell layers, one CUDA context (= one GPU) each: the same
 * decomposition bench.py drives over NCCL ranks, here inside one process, with
 * the per-slab counts exchanged through the host -- they are needed there anyway to size the
 * surface.  Every GPU pulls its slab over its own PCIe link and pushes its part of the mesh
 * straight into the final host arrays. */
#define MC33_MAX_SLABS 64
typedef struct {
	MC33 pub;
	int nslab;
	mc33cu_ctx *ctx[MC33_MAX_SLABS];
	int dev[MC33_MAX_SLABS];     /* device of every slab */
	int ngpu;                    /* distinct devices in use */
	int grid_uploaded;
	const void *registered;      /* the grid's sample block, page-locked by us at the first upload */
	int reg_tried;
	/* largest mesh this extractor has produced so far: sizes the result arrays of the next call before
	 * its counts are complete (calculate_isosurface) */
	unsigned long long hist_nV, hist_nT;
} mc33_private;

static int g_last_ngpu = 0, g_last_nslab = 0;
/* how many GPUs / z-slabs the most recently created MC33 uses (reporting hooks of bench.py) */
int mc33_dropin_gpus_last(void) { return g_last_ngpu; }
int mc33_dropin_slabs_last(void) { return g_last_nslab; }

/* ------------------------------------------------------------------------- */
/* 3x3 helpers: reference MC33_util_grd.c:86-121                              */
/* ------------------------------------------------------------------------- */
#ifndef GRD_ORTHOGONAL
/* upper-triangular A: c = A b (t == 0) keeps the terms j >= i, c = A^T b (t != 0) the
 * terms j <= i; c may alias b only in the order the reference evaluates them
 * (row 2 first for the transpose, row 0 first otherwise: MC33_util_grd.c:87-98) */
void _multTSA_bf(const double (*A)[3], MC33_real *b, MC33_real *c, int t)
{
	for (int n = 0; n < 3; n++) {
		const int i = t ? 2 - n : n;
		double acc = 0.0;
		int first = 1;
		for (int j = t ? 0 : i; j <= (t ? i : 2); j++) {
			const double term = (t ? A[j][i] : A[i][j]) * b[j];
			acc = first ? term : acc + term;
			first = 0;
		}
		c[i] = (MC33_real)acc;
	}
}

void _multA_bf(const double (*A)[3], MC33_real *b, MC33_real *c, int t)
{
	double r[3];
	for (int i = 0; i < 3; i++)
		r[i] = t ? A[0][i] * b[0] + A[1][i] * b[1] + A[2][i] * b[2]
		         : A[i][0] * b[0] + A[i][1] * b[1] + A[i][2] * b[2];
	for (int i = 0; i < 3; i++) c[i] = (MC33_real)r[i];
}

void (*mult_Abf)(const double (*)[3], MC33_real *, MC33_real *, int) = _multA_bf;

void setIdentMat3x3d(double (*A)[3])
{
	for (int j = 0; j < 3; j++)
		for (int i = 0; i < 3; i++) A[j][i] = (i == j) ? 1.0 : 0.0;
}
#endif

/* Markers for MC33.store.  In the reference these are the four vertex store routines
 * (marching_cubes_33.c:485-621), called by MC33_findCase for every new vertex; here the
 * store runs inside the CUDA kernels (mc33_core.cuh store_vertex) and the pointer only
 * records which variant create_MC33 selected.  There is no host-side vertex stream to
 * append to, so a direct call cannot do what the caller expects: it stops the program
 * with a message instead of returning a made-up index. */
static unsigned int store_marker_called(const char *name)
{
	fprintf(stderr, "libMC33_b200: %s() is a marker for MC33.store; vertices are stored on the GPU by "
	                "calculate_isosurface and the routine cannot be called directly\n", name);
	abort();
	return 0;
}
unsigned int MC33_spn0(void *m, MC33_real *r) { (void)m; (void)r; return store_marker_called("MC33_spn0"); }
unsigned int MC33_spnA(void *m, MC33_real *r) { (void)m; (void)r; return store_marker_called("MC33_spnA"); }
unsigned int MC33_spnB(void *m, MC33_real *r) { (void)m; (void)r; return store_marker_called("MC33_spnB"); }
#ifndef GRD_ORTHOGONAL
unsigned int MC33_spnC(void *m, MC33_real *r) { (void)m; (void)r; return store_marker_called("MC33_spnC"); }
#endif

/* ------------------------------------------------------------------------- */
/* grids: reference MC33_util_grd.c:125-169, :585-686                         */
/* ------------------------------------------------------------------------- */
void free_memory_grd(_GRD *Z)
{
	if (!Z) return;
	GRD_data_type ***F = Z->F;
	if (F) {
		const unsigned int nzp = Z->N[2] + 1, nyp = Z->N[1] + 1;
		if (Z->internal_data == MC33_GRD_BLOCK && F[0]) mc33_result_free(F[0][0]);
		for (unsigned int k = 0; k < nzp; k++) {
			GRD_data_type **rows = F[k];
			if (Z->internal_data == MC33_GRD_ROWS) {
				if (!rows) break;           /* alloc_F stopped here */
				for (unsigned int j = 0; j < nyp; j++) free(rows[j]);
			}
			free(rows);
		}
		free(F);
	}
	free(Z);
}

/* Row tables of Z->N[2]+1 slices x Z->N[1]+1 rows; rows == NULL: every x-row gets its
 * own malloc (the reference's layout, MC33_util_grd.c:147-169: callers may replace or
 * free single rows), else row (k,j) points at rows + (k*ny+j)*nx samples.
 * On failure the slice that could not be completed is left NULL (free_memory_grd stops there). */
static int build_F(_GRD *Z, GRD_data_type *rows)
{
	const size_t nx = (size_t)Z->N[0] + 1, ny = (size_t)Z->N[1] + 1, nz = (size_t)Z->N[2] + 1;
	Z->F = (GRD_data_type ***)calloc(nz, sizeof(GRD_data_type **));
	if (!Z->F) return -1;
	for (size_t k = 0; k < nz; k++) {
		GRD_data_type **tab = (GRD_data_type **)calloc(ny, sizeof(GRD_data_type *));
		if (!tab) return -1;
		size_t j = 0;
		for (; j < ny; j++) {
			tab[j] = rows ? rows + (k * ny + j) * nx : (GRD_data_type *)malloc(nx * sizeof(GRD_data_type));
			if (!tab[j]) break;
		}
		if (j < ny) {
			while (j) free(tab[--j]);
			free(tab);
			return -1;
		}
		Z->F[k] = tab;
	}
	return 0;
}

int alloc_F(_GRD *Z)
{
	Z->internal_data = MC33_GRD_ROWS;
	return build_F(Z, 0);
}

int mc33_alloc_F_block(_GRD *Z)
{
	const size_t n = ((size_t)Z->N[0] + 1) * ((size_t)Z->N[1] + 1) * ((size_t)Z->N[2] + 1);
	GRD_data_type *blk = (GRD_data_type *)mc33_result_alloc(n * sizeof(GRD_data_type));
	Z->F = 0;
	if (!blk) return -1;
	Z->internal_data = MC33_GRD_BLOCK;
	if (build_F(Z, blk)) {
		/* free_memory_grd releases the block through F[0][0]: make sure it can, or do it here */
		if (!Z->F || !Z->F[0]) { mc33_result_free(blk); Z->internal_data = 0; }
		return -1;
	}
	return 0;
}

static void grd_defaults(_GRD *Z)
{
#ifndef GRD_ORTHOGONAL
	for (int i = 0; i < 3; i++) Z->Ang[i] = 90.0f;
	Z->nonortho = 0;
	setIdentMat3x3d(Z->_A);
	setIdentMat3x3d(Z->A_);
#else
	(void)Z;
#endif
}

_GRD *grid_from_data_pointer(unsigned int Nx, unsigned int Ny, unsigned int Nz, GRD_data_type *data)
{
	if (!data || !Nx || !Ny || !Nz) return 0;
	_GRD *Z = (_GRD *)malloc(sizeof(_GRD));
	if (!Z) return 0;
	Z->internal_data = 0;
	Z->F = (GRD_data_type ***)malloc(Nz * sizeof(void *));
	if (!Z->F) { free(Z); return 0; }
	for (unsigned int k = 0; k < Nz; k++) {
		GRD_data_type **rows = (GRD_data_type **)malloc(Ny * sizeof(void *));
		if (!rows) {
			while (k) free(Z->F[--k]);
			free(Z->F); free(Z);
			return 0;
		}
		for (unsigned int j = 0; j < Ny; j++) rows[j] = data + ((size_t)k * Ny + j) * Nx;
		Z->F[k] = rows;
	}
	Z->N[0] = Nx - 1; Z->N[1] = Ny - 1; Z->N[2] = Nz - 1;
	for (int i = 0; i < 3; i++) {
		Z->L[i] = (float)Z->N[i];
		Z->d[i] = 1.0;
		Z->r0[i] = 0.0;
	}
	grd_defaults(Z);
	return Z;
}

_GRD *generate_grid_from_fn(double xi, double yi, double zi, double xf, double yf, double zf,
                            double dx, double dy, double dz, double (*fn)(double x, double y, double z))
{
	double lo[3] = {xi, yi, zi}, hi[3] = {xf, yf, zf}, st[3] = {dx, dy, dz};
	for (int i = 0; i < 3; i++) {
		if (st[i] <= 0 || lo[i] == hi[i]) return 0;
		if (lo[i] > hi[i]) { double t = lo[i]; lo[i] = hi[i]; hi[i] = t; }
		if (hi[i] - lo[i] < st[i]) st[i] = hi[i] - lo[i];
	}
	_GRD *Z = (_GRD *)malloc(sizeof(_GRD));
	if (!Z) return 0;
	for (int i = 0; i < 3; i++) Z->N[i] = (unsigned int)(int)((hi[i] - lo[i]) / st[i] + 0.5);
	if (alloc_F(Z)) { free_memory_grd(Z); return 0; }
	for (int i = 0; i < 3; i++) { Z->d[i] = st[i]; Z->r0[i] = lo[i]; }
	if (fn) {
		/* coordinates are accumulated (x += dx) in double, as the reference does
		 * (MC33_util_grd.c:661-672): this decides the last bits of the samples */
		double z = lo[2];
		for (unsigned int k = 0; k <= Z->N[2]; k++, z += st[2]) {
			double y = lo[1];
			for (unsigned int j = 0; j <= Z->N[1]; j++, y += st[1]) {
				GRD_data_type *row = Z->F[k][j];
				double x = lo[0];
				for (unsigned int i = 0; i <= Z->N[0]; i++, x += st[0]) row[i] = (GRD_data_type)fn(x, y, z);
			}
		}
	}
	for (int i = 0; i < 3; i++) Z->L[i] = (float)(Z->N[i] * Z->d[i]);
	grd_defaults(Z);
	return Z;
}

/* ------------------------------------------------------------------------- */
/* surfaces: reference marching_cubes_33.c:84-127                             */
/* ------------------------------------------------------------------------- */
/* Result arrays come from the C-ABI's pool of page-locked memory (so that the
 * device-to-host copies run at PCIe speed) with plain malloc as the fallback;
 * either kind is released here.  The reference's contract is kept: the arrays are
 * ordinary writable host memory owned by the library and released through
 * free_surface_memory. */
void *mc33_result_alloc(size_t bytes)
{
	void *p = 0;
	if (!getenv("MC33_B200_NO_PIN") && mc33cu_host_alloc(bytes, &p) == MC33CU_OK && p) return p;
	return malloc(bytes ? bytes : 1);
}

void mc33_result_free(void *p)
{
	if (p && mc33cu_host_free(p) != MC33CU_OK) free(p);
}

void free_surface_memory(surface *S)
{
	if (!S) return;
	mc33_result_free(S->T); mc33_result_free(S->V); mc33_result_free(S->N); mc33_result_free(S->color);
	free(S);
}

static int shrink(void **p, size_t bytes)
{
	void *q = malloc(bytes ? bytes : 1);
	if (!q) return -1;
	memcpy(q, *p, bytes);
	mc33_result_free(*p);
	*p = q;
	return 0;
}

void adjustvectorlenght_s(surface *S)
{
	if (!S) return;
	if (S->capv > S->nV) {
		if (shrink((void **)&S->color, sizeof(int) * S->nV)) return;
		if (shrink((void **)&S->N, 3 * sizeof(float) * S->nV)) return;
		if (shrink((void **)&S->V, 3 * sizeof(MC33_real) * S->nV)) return;
		S->capv = S->nV;
	}
	if (S->capt > S->nT) {
		if (shrink((void **)&S->T, 3 * sizeof(int) * S->nT)) return;
		S->capt = S->nT;
	}
}

/* ------------------------------------------------------------------------- */
/* extractor: reference marching_cubes_33.c:1727-1940                         */
/* ------------------------------------------------------------------------- */
void free_MC33(MC33 *M)
{
	if (!M) return;
	mc33_private *p = (mc33_private *)M;
	for (int i = 0; i < p->nslab; i++) mc33cu_destroy(p->ctx[i]);
	if (p->registered) mc33cu_host_unregister(p->registered);
	free(p);
}

/* cell layers [*z0, *z1) of slab i of n (the first slabs take the remainder), and the sample slices
 * it needs: one below (normals / on-iso neighbours of its first slice), two above (the seam slice is
 * numbered by the next slab, whose on-iso points look one slice further) */
static void slab_range(unsigned int nz, int i, int n, unsigned int *z0, unsigned int *z1, unsigned int *lo, unsigned int *hi)
{
	const unsigned int base = nz / (unsigned int)n, rem = nz % (unsigned int)n, ui = (unsigned int)i;
	*z0 = ui * base + (ui < rem ? ui : rem);
	*z1 = *z0 + base + (ui < rem ? 1u : 0u);
	*lo = (*z0 > 0 ? *z0 : 1u) - 1u;
	*hi = *z1 + 2 < nz + 1 ? *z1 + 2 : nz + 1;
}

static void describe(const MC33 *M, mc33cu_desc *d, int slab, int nslab)
{
	memset(d, 0, sizeof *d);
	d->dtype = MC33_DTYPE;
	d->nx = M->nx; d->ny = M->ny; d->nz = M->nz;
	slab_range(M->nz, slab, nslab, &d->cell_z0, &d->cell_z1, &d->z_lo, &d->z_hi);
	d->is_last = slab == nslab - 1;
	d->normal_neg = MC33_NORMAL_NEG;
	d->store = M->store == MC33_spn0 ? MC33CU_SPN0 : M->store == MC33_spnA ? MC33CU_SPNA
	         : M->store == MC33_spnB ? MC33CU_SPNB : -1;
	for (int i = 0; i < 3; i++) { d->O[i] = M->O[i]; d->D[i] = M->D[i]; }
	d->ca = M->ca; d->cb = M->cb;
	d->A[0] = d->A[4] = d->A[8] = 1.0;
	d->Ai[0] = d->Ai[4] = d->Ai[8] = 1.0;
#ifndef GRD_ORTHOGONAL
	if (M->store == MC33_spnC) {
		d->store = MC33CU_SPNC;
		for (int j = 0; j < 3; j++)
			for (int i = 0; i < 3; i++) { d->A[3 * j + i] = M->_A[j][i]; d->Ai[3 * j + i] = M->A_[j][i]; }
		/* a user-supplied mult_Abf cannot run on the GPU: only the two library
		 * routines are recognised */
		d->tsa = mult_Abf == _multTSA_bf ? 1 : (mult_Abf == _multA_bf ? 0 : -1);
	}
#endif
}

MC33 *create_MC33(_GRD *G)
{
	if (!G || !G->F) return 0;
	mc33_private *p = (mc33_private *)calloc(1, sizeof(mc33_private));
	if (!p) return 0;
	MC33 *M = &p->pub;
	M->nx = G->N[0]; M->ny = G->N[1]; M->nz = G->N[2];
	M->ca = M->cb = 1;   /* the reference leaves these unset unless spnB is chosen */
	/* choice of store variant and geometry snapshot: reference c:1762-1782 */
#ifndef GRD_ORTHOGONAL
	if (G->nonortho) {
		M->store = MC33_spnC;
		for (int j = 0; j < 3; j++)
			for (int i = 0; i < 3; i++) {
				M->_A[j][i] = G->_A[j][i] * G->d[i];
				M->A_[j][i] = G->A_[j][i] / G->d[j];
			}
	} else
#endif
	if (G->d[0] != G->d[1] || G->d[1] != G->d[2]) {
		M->ca = (MC33_real)(G->d[2] / G->d[0]);
		M->cb = (MC33_real)(G->d[2] / G->d[1]);
		M->store = MC33_spnB;
	} else {
		M->store = (G->d[0] == 1 && G->r0[0] == 0 && G->r0[1] == 0 && G->r0[2] == 0) ? MC33_spn0 : MC33_spnA;
	}
	for (int j = 0; j < 3; j++) {
		M->O[j] = (MC33_real)G->r0[j];
		M->D[j] = (MC33_real)G->d[j];
	}
	M->F = (const GRD_data_type ***)G->F;
	if (!M->nx || !M->ny || !M->nz) { free(p); return 0; }
	/* how many GPUs: MC33_B200_GPUS (default: all visible), never more than cell layers; without an
	 * explicit request a GPU's share is not made smaller than 64 MB of samples (below that the per-call
	 * overheads of one more device outweigh its share of the copy) */
	const double bytes = ((double)M->nx + 1) * ((double)M->ny + 1) * ((double)M->nz + 1) * sizeof(GRD_data_type);
	int ndev = mc33cu_device_count(), want = ndev;
	const char *e = getenv("MC33_B200_GPUS");
	if (e && atoi(e) > 0) want = atoi(e);
	else {
		const int by_size = (int)(bytes / (64.0 * 1024 * 1024));
		if (want > by_size) want = by_size;
	}
	if (want > ndev) want = ndev;
	if (want > MC33_MAX_SLABS) want = MC33_MAX_SLABS;
	if (want < 1) want = 1;
	/* z-chunks per GPU (MC33_B200_CHUNKS, default: about 64 MB of samples each, at most 8): the link is the
	 * bottleneck of a call (the samples go up, the mesh comes down), so a GPU's share is cut further and its
	 * chunks are uploaded one after the other -- chunk k is counted, emitted and its part of the mesh
	 * downloaded while chunk k+1 is still arriving (PCIe is full duplex).  Chunks are dealt round-robin:
	 * slab i sits on GPU i % ngpu, so that all GPUs advance through z together and the vertex base of a slab
	 * (the counts of every slab below it) is known as early as possible. */
	int chunks = 1;
	const char *ce = getenv("MC33_B200_CHUNKS");
	if (ce && atoi(ce) > 0) chunks = atoi(ce);
	else {
		chunks = (int)(bytes / want / (64.0 * 1024 * 1024));
		if (chunks > 8) chunks = 8;
	}
	if (chunks < 1) chunks = 1;
	if (want * chunks > MC33_MAX_SLABS) chunks = MC33_MAX_SLABS / want;
	int nslab = want * chunks;
	int devs[MC33_MAX_SLABS];
	for (int i = 0; i < MC33_MAX_SLABS; i++) devs[i] = i % want;
	/* MC33_B200_DEVICES="2,3" names the device of every slab explicitly (a device may appear more than
	 * once: slabs are independent contexts) and overrides the two counts above */
	const char *dl = getenv("MC33_B200_DEVICES");
	if (dl && *dl) {
		int n = 0;
		for (const char *q = dl; *q && n < MC33_MAX_SLABS;) {
			char *end = 0;
			const long v = strtol(q, &end, 10);
			if (end == q) break;
			devs[n++] = (int)v;
			q = *end == ',' ? end + 1 : end;
		}
		if (n > 0) nslab = n;
	}
	if ((unsigned int)nslab > M->nz) nslab = (int)M->nz;
	if (nslab < 1) nslab = 1;
	for (int i = 0; i < nslab; i++) {
		mc33cu_desc d;
		describe(M, &d, i, nslab);
		if (d.tsa < 0) d.tsa = 0;
		if (mc33cu_create(&d, devs[i], &p->ctx[i]) != MC33CU_OK) {
			if (getenv("MC33_B200_VERBOSE")) fprintf(stderr, "create_MC33: %s\n", mc33cu_last_error());
			for (int j = 0; j < i; j++) mc33cu_destroy(p->ctx[j]);
			free(p);
			return 0;
		}
		p->dev[i] = devs[i];
		p->nslab = i + 1;
	}
	p->ngpu = 0;
	for (int i = 0; i < p->nslab; i++) {
		int seen = 0;
		for (int j = 0; j < i; j++) seen |= p->dev[j] == p->dev[i];
		p->ngpu += !seen;
	}
	g_last_ngpu = p->ngpu;
	g_last_nslab = p->nslab;
	return M;
}

/* is the whole grid one x-fastest block (grid_from_data_pointer, the block readers)? */
static const void *whole_block(const MC33 *M, size_t *bytes)
{
	const size_t rowb = ((size_t)M->nx + 1) * sizeof(GRD_data_type), NY = (size_t)M->ny + 1, NZ = (size_t)M->nz + 1;
	const char *first = (const char *)M->F[0][0];
	for (size_t z = 0; z < NZ; z++)
		for (size_t y = 0; y < NY; y++)
			if ((const char *)M->F[z][y] != first + (z * NY + y) * rowb) return 0;
	*bytes = NZ * NY * rowb;
	return first;
}

/* shared front half of calculate_isosurface / size_of_isosurface: start the upload of every slab and its
 * count behind it; nothing waits here.  Slabs that share a device upload one after the other (each
 * slab's copy is chained behind the previous one's), so slab k is counted while slab k+1 arrives. */
static int start_counts(MC33 *M, MC33_real iso)
{
	mc33_private *p = (mc33_private *)M;
	int rc;
	for (int i = 0; i < p->nslab; i++) {
		mc33cu_desc d;
		describe(M, &d, i, p->nslab);
		if (d.store < 0 || d.tsa < 0) return MC33CU_ERR_ARG;
		rc = mc33cu_set_geometry(p->ctx[i], &d);
		if (rc) return rc;
	}
	/* The reference reads the samples at calculate time (c:1792, c:1820), so they are uploaded on
	 * every call; MC33_B200_CACHE_GRID=1 promises they do not change between calls on the same MC33
	 * and uploads them once.  A grid that is one block is page-locked at its first upload (DMA at
	 * link speed, all GPUs in flight together) and released in free_MC33: G and its samples must
	 * outlive M, as for the reference (M borrows G->F, c:1792).  MC33_B200_NO_PIN=1 turns that off. */
	const char *cache = getenv("MC33_B200_CACHE_GRID");
	const int upload = !(p->grid_uploaded && cache && cache[0] == '1');
	if (upload && !p->reg_tried) {
		p->reg_tried = 1;
		size_t bytes = 0;
		const void *blk = getenv("MC33_B200_NO_PIN") ? 0 : whole_block(M, &bytes);
		if (blk && mc33cu_host_register(blk, bytes) == MC33CU_OK) p->registered = blk;
	}
	M->iso = iso;
	for (int i = 0; i < p->nslab; i++) {
		if (upload) {
			rc = mc33cu_grid_upload_rows_async(p->ctx[i], (const void *const *const *)M->F);
			if (rc) return rc;
			for (int j = i + 1; j < p->nslab; j++)
				if (p->dev[j] == p->dev[i]) {          /* the next slab of this device: its copy follows this one */
					rc = mc33cu_stream_wait(p->ctx[j], p->ctx[i]);
					if (rc) return rc;
					break;
				}
		}
		rc = mc33cu_count_async(p->ctx[i], (double)iso, 0);
		if (rc) return rc;
	}
	if (upload) p->grid_uploaded = 1;
	return MC33CU_OK;
}

/* wait for the count of slab i and fetch it */
static int slab_counts(mc33_private *p, int i, mc33cu_counts *k)
{
	int rc = mc33cu_sync(p->ctx[i]);
	if (rc) return rc;
	return mc33cu_get_counts(p->ctx[i], k);
}

static void drain(mc33_private *p)
{
	for (int i = 0; i < p->nslab; i++) mc33cu_sync(p->ctx[i]);
}

unsigned long long size_of_isosurface(MC33 *M, MC33_real iso, unsigned int *nV, unsigned int *nT)
{
	mc33_private *p = (mc33_private *)M;
	mc33cu_counts k, ki;
	memset(&k, 0, sizeof k);
	int rc = M ? start_counts(M, iso) : MC33CU_ERR_ARG;
	for (int i = 0; rc == MC33CU_OK && i < p->nslab; i++) {
		rc = slab_counts(p, i, &ki);
		k.nV += ki.nV; k.nT += ki.nT;
	}
	if (rc == MC33CU_OK && (k.nV >= 0xFFFFFFFFull || k.nT >= 0xFFFFFFFFull)) rc = MC33CU_ERR_RANGE;   /* unsigned int indices, h:140 */
	if (rc != MC33CU_OK) {
		if (M) drain(p);
		if (nV) *nV = 0;
		if (nT) *nT = 0;
		return 0;
	}
	M->nV = (unsigned int)k.nV; M->nT = (unsigned int)k.nT;
	if (nV) *nV = (unsigned int)k.nV;
	if (nT) *nT = (unsigned int)k.nT;
	/* same formula as the reference, including its 6*sizeof(MC33_real) (c:1939) */
	return k.nV * (6 * sizeof(MC33_real) + sizeof(int)) + k.nT * (3 * sizeof(int)) + sizeof(surface);
}

/* estimated capacities are rounded up to eight steps per octave, so that the calls of an iso sweep ask the
 * page-locked pool for the same few block sizes again and again */
static unsigned long long round_cap(unsigned long long n)
{
	unsigned long long step = 1;
	while (step * 16 <= n) step *= 2;              /* step = 2^(floor(log2 n) - 3) */
	return (n + step - 1) / step * step;
}

static int alloc_result(surface *S, unsigned long long capv, unsigned long long capt)
{
	if (capv > 0xFFFFFFFFull) capv = 0xFFFFFFFFull;
	if (capt > 0xFFFFFFFFull) capt = 0xFFFFFFFFull;
	if (!capt) capt = 1;
	S->capv = (unsigned int)capv; S->capt = (unsigned int)capt;
	S->T = (unsigned int (*)[3])mc33_result_alloc((size_t)capt * 3 * sizeof(int));
	S->V = (MC33_real (*)[3])mc33_result_alloc((size_t)capv * 3 * sizeof(MC33_real));
	S->N = (float (*)[3])mc33_result_alloc((size_t)capv * 3 * sizeof(float));
	S->color = (int *)mc33_result_alloc((size_t)capv * sizeof(int));
	return S->T && S->V && S->N && S->color ? 0 : -1;
}

/* how much room the result arrays get when they must be sized before every slab has been counted: what the
 * extractor has seen so far on this grid (an iso sweep), else the slabs counted so far scaled to the whole grid,
 * with head room (capv / capt > nV / nT, as after the reference's block-wise growth, c:489-510): factor 1.5 on the
 * extrapolation (MC33_B200_SPECULATE=<factor> changes it; 0 waits for all the counts instead), 0.8 of that on the
 * history (the largest mesh so far is a better guide than the first slabs of this one). */
/* MC33_B200_TRACE=1: host-clock time line of a call on stderr (when each slab's count was in, when the call ended) */
static double now_ms(void)
{
	struct timespec t;
	clock_gettime(CLOCK_MONOTONIC, &t);
	return t.tv_sec * 1e3 + t.tv_nsec * 1e-6;
}

static double speculate_factor(void)
{
	const char *e = getenv("MC33_B200_SPECULATE");
	return e && *e ? atof(e) : 1.5;
}

surface *calculate_isosurface(MC33 *M, MC33_real iso)
{
	if (!M) return 0;
	mc33_private *p = (mc33_private *)M;
	surface *S = (surface *)calloc(1, sizeof(surface));
	if (!S) return 0;
	mc33cu_counts ks[MC33_MAX_SLABS];
	M->memoryfault = 0;
	const int trace = getenv("MC33_B200_TRACE") != 0;
	const double t_begin = trace ? now_ms() : 0;
	int rc = start_counts(M, iso);
	if (trace) fprintf(stderr, "[mc33 trace] iso %g: %d slab(s) started at +%.3f ms\n", (double)iso, p->nslab, now_ms() - t_begin);
	/* Slab i's vertices are [vb, vb + nV_i) and its triangles [tb, tb + nT_i) of the result (the running nV / nT of
	 * the reference's sweep, c:487, c:1245, across slabs); every device writes its range straight into the final host
	 * arrays.  A slab's place only depends on the slabs below it, so it is emitted as soon as it has been counted
	 * -- while the slabs above it are still being uploaded -- into arrays sized from an estimate; slabs that do
	 * not fit the estimate wait until every count is in, the arrays are replaced by exact ones (the part already
	 * there is copied over) and the rest follows. */
	const double fac = speculate_factor();
	unsigned long long vb = 0, tb = 0;             /* totals of the slabs counted so far */
	unsigned long long dvb = 0, dtb = 0;           /* ... of the slabs emitted so far */
	int first_deferred = -1, allocated = 0;
	for (int i = 0; rc == MC33CU_OK && i < p->nslab; i++) {
		rc = slab_counts(p, i, &ks[i]);
		if (rc != MC33CU_OK) break;
		if (trace) fprintf(stderr, "[mc33 trace]   slab %d counted at +%.3f ms (nV %llu nT %llu)\n", i, now_ms() - t_begin,
		                   (unsigned long long)ks[i].nV, (unsigned long long)ks[i].nT);
		const unsigned long long nvb = vb + ks[i].nV, ntb = tb + ks[i].nT;
		if (nvb >= 0xFFFFFFFFull || ntb >= 0xFFFFFFFFull) { rc = MC33CU_ERR_RANGE; break; }   /* unsigned int indices, h:140 */
		if (first_deferred < 0 && ks[i].nV) {
			if (!allocated) {
				unsigned long long cv, ct;
				if (i == p->nslab - 1) { cv = nvb; ct = ntb; }                /* everything is known: exact */
				else if (fac <= 0) cv = ct = 0;
				else {
					const double up = (double)p->nslab / (i + 1);
					const double ev = (double)nvb * up, et = (double)ntb * up;
					const double hv = 0.8 * (double)p->hist_nV, ht = 0.8 * (double)p->hist_nT;
					cv = round_cap((unsigned long long)(fac * (hv > ev ? hv : ev)) + 1024);
					ct = round_cap((unsigned long long)(fac * (ht > et ? ht : et)) + 1024);
				}
				if (cv >= nvb && ct >= ntb) {
					if (alloc_result(S, cv, ct)) { rc = MC33CU_ERR_NOMEM; break; }
					allocated = 1;
				}
			}
			if (allocated && nvb <= S->capv && ntb <= S->capt) {
				rc = mc33cu_emit_host_async(p->ctx[i], &S->V[vb], (float *)&S->N[vb], S->color + vb, (unsigned int *)&S->T[tb],
				                            (uint32_t)vb, (uint32_t)nvb, DefaultColorMC);
				if (rc != MC33CU_OK) break;
				if (trace) fprintf(stderr, "[mc33 trace]   slab %d emit issued at +%.3f ms\n", i, now_ms() - t_begin);
				dvb = nvb; dtb = ntb;
			} else {
				first_deferred = i;
			}
		}
		vb = nvb; tb = ntb;
	}
	if (rc != MC33CU_OK) goto fail;
	if (vb == 0) {
		/* empty isosurface: zero-filled struct, iso included (reference c:1880-1883) */
		drain(p);
		M->nV = M->nT = 0;
		return S;
	}
	if (first_deferred >= 0) {
		/* the estimate was too small (or there was none): exact arrays now, the part already downloaded moves over */
		surface old = *S;
		if (alloc_result(S, vb, tb)) {
			rc = MC33CU_ERR_NOMEM;
			drain(p);
			if (allocated) { mc33_result_free(old.T); mc33_result_free(old.V); mc33_result_free(old.N); mc33_result_free(old.color); }
			goto fail;
		}
		if (allocated) {
			for (int i = 0; i < first_deferred; i++) {
				const int r2 = mc33cu_sync(p->ctx[i]);
				if (r2 != MC33CU_OK) rc = r2;
			}
			memcpy(S->V, old.V, (size_t)dvb * 3 * sizeof(MC33_real));
			memcpy(S->N, old.N, (size_t)dvb * 3 * sizeof(float));
			memcpy(S->color, old.color, (size_t)dvb * sizeof(int));
			memcpy(S->T, old.T, (size_t)dtb * 3 * sizeof(int));
			mc33_result_free(old.T); mc33_result_free(old.V); mc33_result_free(old.N); mc33_result_free(old.color);
			if (rc != MC33CU_OK) goto fail;
		}
		unsigned long long v0 = 0, t0 = 0;
		for (int i = 0; i < p->nslab; i++) {
			if (i >= first_deferred && ks[i].nV) {
				rc = mc33cu_emit_host_async(p->ctx[i], &S->V[v0], (float *)&S->N[v0], S->color + v0, (unsigned int *)&S->T[t0],
				                            (uint32_t)v0, (uint32_t)(v0 + ks[i].nV), DefaultColorMC);
				if (rc != MC33CU_OK) goto fail;
			}
			v0 += ks[i].nV; t0 += ks[i].nT;
		}
	}
	for (int i = 0; i < p->nslab; i++) {
		const int r2 = mc33cu_sync(p->ctx[i]);
		if (r2 != MC33CU_OK) rc = r2;
	}
	if (rc != MC33CU_OK) goto fail;
	if (trace) fprintf(stderr, "[mc33 trace]   done at +%.3f ms (capv %u capt %u, first deferred slab %d)\n", now_ms() - t_begin,
	                   S->capv, S->capt, first_deferred);
	S->nV = (unsigned int)vb; S->nT = (unsigned int)tb;
	S->iso = iso;
	if (vb > p->hist_nV) p->hist_nV = vb;
	if (tb > p->hist_nT) p->hist_nT = tb;
	/* the MC33 mirrors the surface head, as in the reference (c:1873) */
	memcpy(M, S, offsetof(MC33, memoryfault));
	return S;
fail:
	if (getenv("MC33_B200_VERBOSE")) fprintf(stderr, "calculate_isosurface: %s\n", mc33cu_last_error());
	drain(p);                    /* nothing may still be writing into the arrays that are about to be released */
	M->memoryfault = 1;
	free_surface_memory(S);
	return 0;
}
