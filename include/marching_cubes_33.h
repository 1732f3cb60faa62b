/*
 * marching_cubes_33.h -- public C API of the B200-native Marching Cubes 33 library.
 *
 * Source- and ABI-compatible with the header of the same name in the MC33 C
 * library v5.5 by D. Vega and J. Abache (reference include/marching_cubes_33.h:
 * typedefs :66-88, _GRD :111-124, surface :133-152, MC33 :154-179, prototypes
 * :189-329): same type names, same struct layouts (field order, types, sizes),
 * same function signatures and the same three compile-time switches
 *
 *     INTEGER_GRD      grid samples are unsigned integers, MC33_real is float
 *     GRD_TYPE_SIZE    1, 2, 4 (integer) or 8 (double samples, MC33_real double)
 *     GRD_ORTHOGONAL   drop the inclined-grid members from _GRD and MC33
 *
 * so a program written against the reference recompiles and relinks unchanged
 * (-lMC33_b200_<variant>).  The extraction itself runs on an NVIDIA B200 through
 * the C-ABI in mc33cu.h; there is no CPU implementation behind these functions.
 *
 * Typical use (identical to the reference, README.md:84-155 there):
 *
 *     _GRD    *G = grid_from_data_pointer(Nx, Ny, Nz, samples);
 *     MC33    *M = create_MC33(G);
 *     surface *S = calculate_isosurface(M, isovalue);
 *     ... S->nV, S->V, S->N, S->color, S->nT, S->T ...
 *     free_surface_memory(S); free_MC33(M); free_memory_grd(G);
 */
#ifndef marching_cubes_33_h
#define marching_cubes_33_h

#define MC33C_VERSION_MAJOR 5
#define MC33C_VERSION_MINOR 5

/* ---- element type selection (compile time, as in the reference) ---------- */
#if defined(INTEGER_GRD)
typedef float MC33_real;
#  if GRD_TYPE_SIZE == 4
typedef unsigned int GRD_data_type;
#  elif GRD_TYPE_SIZE == 2
typedef unsigned short int GRD_data_type;
#  elif GRD_TYPE_SIZE == 1
typedef unsigned char GRD_data_type;
#  else
#    error "GRD_TYPE_SIZE must be 1, 2 or 4 when INTEGER_GRD is defined"
#  endif
#elif GRD_TYPE_SIZE == 8
typedef double GRD_data_type;
typedef double MC33_real;
#else
typedef float GRD_data_type;
typedef float MC33_real;
#  undef GRD_TYPE_SIZE
#  define GRD_TYPE_SIZE 4
#endif

#if !defined(marching_cubes_33_c) && defined(__cplusplus) && !defined(mc33_no_lib)
extern "C" {
#endif

/*
 * Regular scalar grid.  F[k][j][i] is the sample at grid point (i, j, k); x is the
 * fastest index.  N[] holds the number of INTERVALS per axis, so there are
 * (N[0]+1)*(N[1]+1)*(N[2]+1) samples.  r0 is the position of sample (0,0,0), d
 * the spacing per axis, L the edge lengths.  For an inclined grid nonortho is 1
 * and _A / A_ map inclined to orthogonal coordinates and back.  periodic, L,
 * Ang and title are carried for compatibility and not used by the extractor.
 */
typedef struct {
	GRD_data_type ***F;
	unsigned int N[3];
	double r0[3], d[3];
	float L[3];
#ifndef GRD_ORTHOGONAL
	float Ang[3];
	int nonortho;
	double _A[3][3], A_[3][3];
#endif
	int periodic;
	int internal_data;
	char title[160];
} _GRD;

/*
 * Indexed triangle mesh: nV vertices with position V, unit normal N and RGBA
 * colour; nT triangles as vertex index triples T.  The arrays are ordinary
 * malloc'ed host memory owned by the surface (release with free_surface_memory);
 * callers may modify them in place.  capt / capv are the allocated capacities.
 */
typedef struct {
	unsigned int (*T)[3];
	MC33_real (*V)[3];
	float (*N)[3];
	int *color;
	unsigned int nV, nT;
	unsigned int capt, capv;
	MC33_real iso;
	union {
		void *p;
		long long ul;
		int i[2];
		short si[4];
		char c[8];
		float f[2];
		double df;
	} user;
} surface;

/*
 * Extractor bound to one grid.  The leading members mirror `surface`; the rest
 * is the grid geometry snapshot taken by create_MC33.  Dx..Lz (the reference's
 * slice-to-slice vertex reuse tables) are kept for layout compatibility and
 * are NULL here: vertex sharing is resolved on the GPU by edge ownership.
 */
typedef struct {
	unsigned int (*T)[3];
	MC33_real (*V)[3];
	float (*N)[3];
	int *color;
	unsigned int nV, nT;
	unsigned int capt, capv;
	MC33_real iso;

	int memoryfault;            /* non-zero after an out-of-memory condition */

	const GRD_data_type ***F;   /* borrowed from the _GRD: it must outlive the MC33 */
	MC33_real O[3], D[3], ca, cb;
	unsigned int nx, ny, nz;
	unsigned int (*store)(void *, MC33_real *);
#ifndef GRD_ORTHOGONAL
	double _A[3][3], A_[3][3];
#endif
	unsigned int **Dx, **Dy, **Ux, **Uy, **Lz;
} MC33;

/* colour given to every vertex, 0xAABBGGRR */
extern int DefaultColorMC;

#ifndef GRD_ORTHOGONAL
/* c = A b (t == 0) or c = transpose(A) b (t != 0); the TSA form assumes A upper
 * triangular.  mult_Abf selects which one the inclined-grid store uses. */
void _multTSA_bf(const double (*A)[3], MC33_real *b, MC33_real *c, int t);
void _multA_bf(const double (*A)[3], MC33_real *b, MC33_real *c, int t);
extern void (*mult_Abf)(const double (*)[3], MC33_real *, MC33_real *, int);
#endif

/* ---- surface files: 0 on success, -1 on failure; read returns NULL on failure */
int write_bin_s(surface *S, const char *filename);
surface *read_bin_s(const char *filename);
int write_txt_s(surface *S, const char *filename);
int write_obj_s(surface *S, const char *filename);
int write_ply_s(surface *S, const char *filename, const char *author, const char *object);

/* ---- extractor life cycle -------------------------------------------------- */
/* NULL on error (NULL grid, out of memory, no usable CUDA device). */
MC33 *create_MC33(_GRD *G);
/* Isosurface for isovalue iso.  NULL when memory (host or device) runs out; an
 * empty isosurface is not an error and yields a zero-filled surface. */
surface *calculate_isosurface(MC33 *M, MC33_real iso);
/* Byte size the surface would occupy, and optionally its vertex / triangle count,
 * without building it. */
unsigned long long size_of_isosurface(MC33 *M, MC33_real iso, unsigned int *nV, unsigned int *nT);
void free_MC33(MC33 *M);
void free_surface_memory(surface *S);
/* shrink the arrays of S to exactly nV / nT entries */
void adjustvectorlenght_s(surface *S);

/* ---- grids ----------------------------------------------------------------- */
void free_memory_grd(_GRD *Z);
/* allocate Z->F for the interval counts already stored in Z->N; -1 on failure */
int alloc_F(_GRD *Z);
_GRD *read_grd(const char *filename);
_GRD *read_grd_binary(const char *filename);
_GRD *read_scanfiles(const char *filename, unsigned int res, int order);
_GRD *read_raw_file(const char *filename, unsigned int *N, int byte, int isfloat);
_GRD *read_dat_file(const char *filename);
/* wrap Nx*Ny*Nz caller-owned samples (x fastest); the samples are not copied and
 * not freed by free_memory_grd */
_GRD *grid_from_data_pointer(unsigned int Nx, unsigned int Ny, unsigned int Nz, GRD_data_type *data);
/* sample fn on a box */
_GRD *generate_grid_from_fn(
	double x_initial, double y_initial, double z_initial,
	double x_final, double y_final, double z_final,
	double x_step, double y_step, double z_step,
	double (*fn)(double x, double y, double z));

#if !defined(marching_cubes_33_c) && defined(__cplusplus) && !defined(mc33_no_lib)
}
#endif

#endif /* marching_cubes_33_h */
