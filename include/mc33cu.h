/* mc33cu.h -- C-ABI of the B200 (sm_100a) Marching Cubes 33 extractor.
 *
 * Plain C, plain pointers and sizes, no torch / C++ types.  This is the
 * boundary a binding of the reference's hot path attaches to: it replaces the
 * body of
 *     create_MC33            /root/reference source/marching_cubes_33.c:1750-1811
 *     calculate_isosurface   source/marching_cubes_33.c:1816-1889
 *     size_of_isosurface     source/marching_cubes_33.c:1892-1940
 *     free_MC33              source/marching_cubes_33.c:1733-1744
 * i.e. the sweep, case selection (MC33_findCase, :673-779), vertex creation
 * (:780-1253, MC33_surfint :628-649) and vertex store (MC33_spn0/A/B/C :485-621).
 * The drop-in implementation of include/marching_cubes_33.h on top of it is
 * mc33_c_library_b200/csrc/mc33_api.c; INTEGRATION.md shows the binding.
 *
 * There is no CPU fallback: every entry point fails with MC33CU_ERR_CUDA when no
 * CUDA device / kernel image is usable.
 */
#ifndef MC33CU_H
#define MC33CU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* element types = the reference's compile-time GRD_data_type choices
 * (include/marching_cubes_33.h:66-88) as a run-time tag */
enum { MC33CU_F32 = 0, MC33CU_F64 = 1, MC33CU_U8 = 2, MC33CU_U16 = 3, MC33CU_U32 = 4 };
/* vertex store variants = MC33_spn0 / spnA / spnB / spnC (marching_cubes_33.c:485-621) */
enum { MC33CU_SPN0 = 0, MC33CU_SPNA = 1, MC33CU_SPNB = 2, MC33CU_SPNC = 3 };

enum {
	MC33CU_OK = 0,
	MC33CU_ERR_ARG = -1,      /* NULL / zero-sized / inconsistent description    */
	MC33CU_ERR_CUDA = -2,     /* CUDA runtime error, see mc33cu_last_error()     */
	MC33CU_ERR_NOMEM = -3,    /* host or device allocation failed                */
	MC33CU_ERR_CAPACITY = -4, /* output buffers too small for this isosurface    */
	MC33CU_ERR_RANGE = -5,    /* more than 2^32-1 vertices or triangles (the API's
	                             indices are unsigned int, marching_cubes_33.h:140) */
	MC33CU_ERR_STATE = -6     /* call order (no grid bound, emit before count)   */
};

typedef struct mc33cu_ctx mc33cu_ctx;

/* Grid + extractor description: what create_MC33 copies out of _GRD
 * (marching_cubes_33.c:1758-1782), plus the z-slab this context works on. */
typedef struct {
	int32_t  dtype;               /* MC33CU_F32 ...                                  */
	uint32_t nx, ny, nz;          /* GLOBAL interval counts, _GRD.N                  */
	/* z-slab (single GPU: z_lo = 0, z_hi = nz+1, cell_z0 = 0, cell_z1 = nz, is_last = 1):
	 * the sample array bound to the context holds global slices [z_lo, z_hi); the
	 * context emits the cells of layers [cell_z0, cell_z1), the shared vertices of
	 * the sample slices [cell_z0, cell_z1) (+ slice nz when is_last).  Required halo:
	 * z_lo <= max(cell_z0,1)-1 and z_hi >= min(cell_z1+2, nz+1).                    */
	uint32_t z_lo, z_hi;
	uint32_t cell_z0, cell_z1;
	int32_t  is_last;
	/* geometry exactly as create_MC33 leaves it in the MC33 struct */
	int32_t  store;               /* MC33CU_SPN*                                     */
	int32_t  normal_neg;          /* MC33_NORMAL_NEG build option                    */
	int32_t  tsa;                 /* mult_Abf == _multTSA_bf (MC33_util_grd.c:86-98)  */
	double   O[3], D[3], ca, cb;  /* MC33.O, .D, .ca, .cb (values of MC33_real)      */
	double   A[9], Ai[9];         /* MC33._A, .A_ row major (scaled by d)            */
} mc33cu_desc;

typedef struct {
	uint64_t nV, nT;              /* vertices / triangles this context emits         */
	uint64_t nShared;             /* edge + on-iso-point vertices (numbered first)   */
	uint64_t nCentre;             /* cell-centre vertices (numbered after them)      */
	uint64_t nSharedHalo;         /* shared vertices of sample slice cell_z1 (owned by
	                                 the next slab; 0 when is_last)                  */
} mc33cu_counts;

/* Device output buffers of ONE context (slab).  The arrays are indexed locally:
 * entry i of V/N/color is the slab's i-th vertex, whose global id is vbase + i;
 * T holds the slab's triangles as GLOBAL vertex ids.  Concatenating the arrays
 * of all slabs in rank order gives the whole mesh.
 * V: 3 MC33_real per vertex (double for MC33CU_F64, else float); N: 3 float;
 * color: int; T: 3 uint32. */
typedef struct {
	void     *V;
	float    *N;
	int32_t  *color;
	uint32_t *T;
	uint64_t *vkey;               /* optional (may be NULL): canonical vertex key:
	                                 (global linear point id)*4 + plane 0/1/2, or
	                                 (global linear cell id)*4 + 3 for centres      */
	uint64_t *tcell;              /* optional: global linear cell id per triangle   */
	uint32_t  capV, capT;         /* capacities in vertices / triangles             */
	uint32_t  vbase;              /* global id of this slab's first vertex          */
	uint32_t  vbase_next;         /* global id of the next slab's first vertex      */
	int32_t   color_value;        /* DefaultColorMC (marching_cubes_33.c:80)        */
	int32_t   pad_;
	const uint32_t *dev_bases;    /* optional DEVICE pointer to {vbase, vbase_next}: when
	                                 set it overrides the two host values, so the bases
	                                 can come straight from an on-device all-gather + scan
	                                 of the per-slab counts with no host round trip    */
} mc33cu_out;

const char *mc33cu_last_error(void);
int  mc33cu_device_count(void);

int  mc33cu_create(const mc33cu_desc *desc, int device, mc33cu_ctx **out);
void mc33cu_destroy(mc33cu_ctx *ctx);
/* replace the geometry part of the description (store .. Ai); the reference reads
 * MC33.O/D/ca/cb/_A/A_ and mult_Abf at store time, so the drop-in re-sends them
 * before every extraction */
int  mc33cu_set_geometry(mc33cu_ctx *ctx, const mc33cu_desc *desc);
/* cudaStream_t as void*; NULL = the context's own stream.  The stream may be changed between calls: the sets of an
 * iso sweep (mc33cu_classify_sweep) are independent of each other, so mc33cu_extract_set_device for different sets
 * may be issued on different streams and run side by side (the caller orders them after the sweep classify and keeps
 * one set of output arrays per stream; the context keeps its per-extraction scratch per stream).  mc33cu_sync waits
 * for the current stream and for every extraction issued on another one. */
int  mc33cu_set_stream(mc33cu_ctx *ctx, void *cuda_stream);

/* bind the samples.  *_device borrows a device pointer (no copy); the two upload
 * calls copy host memory into context-owned device memory: contiguous x-fastest
 * data of the slab, or the reference's triple pointer F[z][y] (_GRD.F,
 * marching_cubes_33.h:112; rows may be separate mallocs, MC33_util_grd.c:147-169)
 * indexed by GLOBAL z. */
int  mc33cu_grid_device(mc33cu_ctx *ctx, const void *dev_samples);
int  mc33cu_grid_upload(mc33cu_ctx *ctx, const void *host_samples);
int  mc33cu_grid_upload_rows(mc33cu_ctx *ctx, const void *const *const *F);
/* the same without waiting for the copy (it is ordered on the context's stream in front of whatever is
 * launched next): several contexts on several devices upload their slabs side by side, each over its own
 * link.  Only page-locked host memory is copied asynchronously by the driver; mc33cu_grid_upload_rows_async
 * falls back to the staged, synchronous row copy when the rows are separate allocations. */
int  mc33cu_grid_upload_async(mc33cu_ctx *ctx, const void *host_samples);
int  mc33cu_grid_upload_rows_async(mc33cu_ctx *ctx, const void *const *const *F);
/* *block = the contiguous host block the slab's rows form (grid_from_data_pointer layout,
 * MC33_util_grd.c:600-612) and its size, or NULL / 0 when the rows are separate allocations */
int  mc33cu_grid_rows_block(mc33cu_ctx *ctx, const void *const *const *F, const void **block, size_t *bytes);
/* Explicit page-locking of caller-owned memory (cudaHostRegisterPortable, reference counted per
 * block, process wide).  The upload calls never register memory on their own; the drop-in
 * registers the grid's sample block at its first upload and releases it in free_MC33.  The
 * block must stay mapped until it is unregistered. */
/* order what is issued on ctx's stream from now on behind everything issued so far on `after`'s stream (a CUDA
 * event, no host wait; the two contexts may sit on different devices).  The drop-in chains the uploads of the
 * z-chunks that share a device so that they cross the link one after the other. */
int  mc33cu_stream_wait(mc33cu_ctx *ctx, mc33cu_ctx *after);
int  mc33cu_host_register(const void *p, size_t bytes);
int  mc33cu_host_unregister(const void *p);

/* classify + count + scan for isovalue iso; synchronises and returns the counts
 * (the GPU form of size_of_isosurface).  iso is converted to MC33_real. */
int  mc33cu_count(mc33cu_ctx *ctx, double iso, mc33cu_counts *counts);
/* the same without host synchronisation: the counts {nV, nT, nShared, nCentre} are
 * left in the caller's DEVICE buffer dev_counts4 (4 x uint32), ready for an
 * all-gather across slabs; follow with mc33cu_emit_device */
int  mc33cu_count_async(mc33cu_ctx *ctx, double iso, uint32_t *dev_counts4);
/* z-slabs: after the all-gather of every slab's dev_counts4 into the DEVICE
 * array dev_counts_all [world][4], leave {vbase, vbase_next} of slab `rank` in the DEVICE
 * buffer dev_bases2 (the exclusive sum of the vertex counts; asynchronous, context stream):
 * the global running nV of the reference (marching_cubes_33.c:487) across slabs.  Pass
 * dev_bases2 as mc33cu_out.dev_bases. */
int  mc33cu_slab_bases(mc33cu_ctx *ctx, const uint32_t *dev_counts_all, int rank, int world, uint32_t *dev_bases2);
/* the same with slab r's counts at dev_counts_all + r * stride_words (e.g. one all-gather of [sets][4] per slab) */
int  mc33cu_slab_bases_strided(mc33cu_ctx *ctx, const uint32_t *dev_counts_all, uint32_t stride_words, int rank, int world,
                               uint32_t *dev_bases2);
/* emit the mesh of the last mc33cu_count into device buffers (asynchronous on the
 * context's stream; mc33cu_sync reports a capacity overflow). */
int  mc33cu_emit_device(mc33cu_ctx *ctx, const mc33cu_out *out);
/* classify + count + scan + emit back to back without any host synchronisation;
 * counts (optional) are read after mc33cu_sync via mc33cu_get_counts. */
int  mc33cu_extract_device(mc33cu_ctx *ctx, double iso, const mc33cu_out *out);
/* Iso sweep over one grid (BASELINE config 2): the reference calls calculate_isosurface once
 * per isovalue and re-reads every sample each time (marching_cubes_33.c:1832-1859).  Here the
 * samples are streamed ONCE for up to 8 isovalues: mc33cu_classify_sweep leaves one set of
 * sign / on-iso bitmaps per isovalue; mc33cu_count_set_async and mc33cu_extract_set_device are
 * mc33cu_count_async / mc33cu_extract_device for pre-classified set `set` (0 .. n-1), in any
 * order, until the next mc33cu_classify_sweep on the context (single-isovalue calls in between
 * do not disturb the sets).  The samples must not change between the sweep classify and the
 * last use of its sets.  Results are identical to n separate extractions. */
int  mc33cu_classify_sweep(mc33cu_ctx *ctx, const double *isos, int n);
int  mc33cu_count_set_async(mc33cu_ctx *ctx, int set, uint32_t *dev_counts4);
int  mc33cu_extract_set_device(mc33cu_ctx *ctx, int set, const mc33cu_out *out);
/* every set keeps its own count state, so all the sets can be counted first (one all-gather of the
 * counts per sweep across z-slabs) and emitted afterwards: emit pre-classified, counted set `set` */
int  mc33cu_emit_set_device(mc33cu_ctx *ctx, int set, const mc33cu_out *out);
int  mc33cu_sync(mc33cu_ctx *ctx);
int  mc33cu_get_counts(mc33cu_ctx *ctx, mc33cu_counts *counts);
/* emit the mesh of the last mc33cu_count into HOST arrays of at least
 * counts.nV / counts.nT entries (device staging is owned by the context). */
int  mc33cu_emit_host(mc33cu_ctx *ctx, void *V, float *N, int32_t *color, uint32_t *T,
                      int32_t color_value);
/* the same for one z-slab of several, without waiting: the slab's vertices / triangles go to the given host
 * positions (the caller offsets the arrays by the slab's vertex / triangle base), triangles carry global ids
 * (vbase, vbase_next as in mc33cu_out); the triangle download overlaps the vertex kernel on a second stream.
 * mc33cu_sync waits for everything and reports overflow. */
int  mc33cu_emit_host_async(mc33cu_ctx *ctx, void *V, float *N, int32_t *color, uint32_t *T, uint32_t vbase,
                            uint32_t vbase_next, int32_t color_value);

/* Page-locked host memory for result arrays, pooled across calls (device-to-host
 * copies into it run at PCIe speed).  mc33cu_host_free returns MC33CU_ERR_ARG for a
 * pointer that did not come from mc33cu_host_alloc, so a caller holding arrays of
 * either origin can fall back to free(). */
int  mc33cu_host_alloc(size_t bytes, void **out);
int  mc33cu_host_free(void *p);

/* per-kernel device times of the most recent extraction, in milliseconds:
 * [0] classify [1] count (with its look-back scan) [2] the gap the round-1 row-scan kernel used to fill (~0)
 * [3] emit cells (triangles + centre vertices; both cell kernels when the device picks) [4] emit vertices.
 * Only measured while timing is enabled (event records between the kernels, which always run one after
 * the other on the context's stream). */
int  mc33cu_enable_timing(mc33cu_ctx *ctx, int on);
int  mc33cu_kernel_times(mc33cu_ctx *ctx, float ms[5]);
/* number of kernel launches issued by this context so far */
uint64_t mc33cu_launch_count(const mc33cu_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
