/*
 * mc33_io.c -- the file side of the marching_cubes_33.h API:
 * surface writers / reader and the grid readers.  Plain host C; nothing here touches
 * the GPU except through the page-locked block the dense readers load into.
 *
 * Formats follow the reference byte for byte (a file written by either library is
 * read by the other):
 *   write_bin_s / read_bin_s     reference source/marching_cubes_33.c:139-209
 *   write_txt_s                  :211-248      write_obj_s  :250-286
 *   write_ply_s                  :288-327
 *   read_grd                     reference source/MC33_util_grd.c:181-262
 *   read_grd_binary              :267-321      read_scanfiles :329-413
 *   read_raw_file                :420-519      read_dat_file  :524-576
 *
 * Grid ingest (f2): read_raw_file, read_dat_file and read_grd_binary load the samples
 * into ONE contiguous x-fastest block of page-locked memory with row pointers over it
 * (mc33_alloc_F_block), reading whole rows / slices with single fread calls, so that
 * calculate_isosurface uploads the grid with one DMA at link speed.  read_grd and
 * read_scanfiles keep the reference's one-malloc-per-row layout (text parsing and
 * per-slice files dominate there, and read_scanfiles reorders slice tables).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mc33_internal.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#if defined(__BYTE_ORDER__) && __BYTE_ORDER__ == __ORDER_BIG_ENDIAN__
#define MC33_HOST_BE 1
#else
#define MC33_HOST_BE 0
#endif

/* ".sup" (float positions) / ".sud" (double positions), little endian ints */
#define MAGIC_F32 0x7075732e
#define MAGIC_F64 0x6575732e
#if GRD_TYPE_SIZE == 8
#define MAGIC_OWN MAGIC_F64
#define MAGIC_OTHER MAGIC_F32
typedef float other_real;
#else
#define MAGIC_OWN MAGIC_F32
#define MAGIC_OTHER MAGIC_F64
typedef double other_real;
#endif

/* ------------------------------------------------------------------------- */
/* surfaces                                                                   */
/* ------------------------------------------------------------------------- */
static int put(FILE *f, const void *p, size_t bytes)
{
	return bytes == 0 || fwrite(p, bytes, 1, f) == 1;
}

int write_bin_s(surface *S, const char *filename)
{
	if (!S || !filename) return -1;
	adjustvectorlenght_s(S);
	FILE *f = fopen(filename, "wb");
	if (!f) return -1;
	const int magic = MAGIC_OWN;
	int ok = put(f, &magic, sizeof magic) && put(f, &S->iso, sizeof(MC33_real)) &&
	         put(f, &S->nV, sizeof(int)) && put(f, &S->nT, sizeof(int)) &&
	         put(f, S->T, (size_t)S->nT * 3 * sizeof(int)) &&
	         put(f, S->V, (size_t)S->nV * 3 * sizeof(MC33_real)) &&
	         put(f, S->N, (size_t)S->nV * 3 * sizeof(float)) &&
	         put(f, S->color, (size_t)S->nV * sizeof(int));
	/* the reference reports only whether the last block (colours) went out; an empty
	 * surface therefore fails there (fwrite of 0 bytes returns 0) and here alike */
	if (S->nV == 0) ok = 0;
	if (fclose(f)) ok = 0;
	return ok ? 0 : -1;
}

static int get(FILE *f, void *p, size_t bytes)
{
	return bytes == 0 || fread(p, bytes, 1, f) == 1;
}

surface *read_bin_s(const char *filename)
{
	if (!filename) return 0;
	FILE *f = fopen(filename, "rb");
	if (!f) return 0;
	int magic = 0;
	surface *S = 0;
	if (!get(f, &magic, sizeof magic) || (magic != MAGIC_OWN && magic != MAGIC_OTHER)) goto bad;
	S = (surface *)calloc(1, sizeof(surface));
	if (!S) goto bad;
	if (magic == MAGIC_OWN) {
		if (!get(f, &S->iso, sizeof(MC33_real))) goto bad;
	} else {
		other_real iso;
		if (!get(f, &iso, sizeof iso)) goto bad;
		S->iso = (MC33_real)iso;
	}
	if (!get(f, &S->nV, sizeof(int)) || !get(f, &S->nT, sizeof(int))) goto bad;
	S->capv = S->nV; S->capt = S->nT;
	S->T = (unsigned int (*)[3])malloc((size_t)S->nT * 3 * sizeof(int) + 1);
	S->V = (MC33_real (*)[3])malloc((size_t)S->nV * 3 * sizeof(MC33_real) + 1);
	S->N = (float (*)[3])malloc((size_t)S->nV * 3 * sizeof(float) + 1);
	S->color = (int *)malloc((size_t)S->nV * sizeof(int) + 1);
	if (!S->T || !S->V || !S->N || !S->color) goto bad;
	if (!get(f, S->T, (size_t)S->nT * 3 * sizeof(int))) goto bad;
	if (magic == MAGIC_OWN) {
		if (!get(f, S->V, (size_t)S->nV * 3 * sizeof(MC33_real))) goto bad;
	} else {
		/* the file holds positions of the other precision: convert in bounded chunks */
		enum { CH = 4096 };
		other_real *tmp = (other_real *)malloc(CH * 3 * sizeof(other_real));
		if (!tmp) goto bad;
		for (size_t v0 = 0; v0 < S->nV; v0 += CH) {
			const size_t n = S->nV - v0 < CH ? S->nV - v0 : CH;
			if (!get(f, tmp, n * 3 * sizeof(other_real))) { free(tmp); goto bad; }
			for (size_t i = 0; i < 3 * n; i++) (&S->V[v0][0])[i] = (MC33_real)tmp[i];
		}
		free(tmp);
	}
	if (!get(f, S->N, (size_t)S->nV * 3 * sizeof(float))) goto bad;
	/* like the reference, a file without a colour block (e.g. nV == 0) is rejected */
	if (S->nV == 0 || !get(f, S->color, (size_t)S->nV * sizeof(int))) goto bad;
	fclose(f);
	return S;
bad:
	fclose(f);
	free_surface_memory(S);
	return 0;
}

int write_txt_s(surface *S, const char *filename)
{
	if (!S || !filename) return -1;
	adjustvectorlenght_s(S);
	FILE *f = fopen(filename, "w");
	if (!f) return -1;
	fprintf(f, "isovalue: %10.5E\n\nVERTICES:\n%d\n\n", (double)S->iso, S->nV);
	for (unsigned int i = 0; i < S->nV; i++)
		fprintf(f, "%9.6f %9.6f %9.6f\n", (double)S->V[i][0], (double)S->V[i][1], (double)S->V[i][2]);
	fprintf(f, "\n\nTRIANGLES:\n%d\n\n", S->nT);
	for (unsigned int i = 0; i < S->nT; i++) fprintf(f, "%8d %8d %8d\n", S->T[i][0], S->T[i][1], S->T[i][2]);
	fprintf(f, "\n\nNORMALS:\n");
	for (unsigned int i = 0; i < S->nV; i++)
		fprintf(f, "%9.6f %9.6f %9.6f\n", (double)S->N[i][0], (double)S->N[i][1], (double)S->N[i][2]);
	fprintf(f, "\n\nCOLORS:\n");
	for (unsigned int i = 0; i < S->nV; i++) fprintf(f, "%d\n", S->color[i]);
	const int tail = fprintf(f, "\nEND\n");
	const int bad = ferror(f);
	if (fclose(f) || bad || tail < 5) return -1;
	return 0;
}

int write_obj_s(surface *S, const char *filename)
{
	if (!S || !filename) return -1;
	adjustvectorlenght_s(S);
	FILE *f = fopen(filename, "w");
	if (!f) return -1;
	fprintf(f, "# isovalue: %10.5E\n# VERTICES %d:\n", (double)S->iso, S->nV);
	for (unsigned int i = 0; i < S->nV; i++)
		fprintf(f, "v %f %f %f\n", (double)S->V[i][0], (double)S->V[i][1], (double)S->V[i][2]);
	fprintf(f, "# NORMALS:\n");
	for (unsigned int i = 0; i < S->nV; i++)
		fprintf(f, "vn %f %f %f\n", (double)S->N[i][0], (double)S->N[i][1], (double)S->N[i][2]);
	fprintf(f, "# TRIANGLES %d:\n", S->nT);
	for (unsigned int i = 0; i < S->nT; i++) {
		/* OBJ indices are 1-based; position and normal share the index */
		const int a = (int)(S->T[i][0] + 1), b = (int)(S->T[i][1] + 1), c = (int)(S->T[i][2] + 1);
		fprintf(f, "f %d//%d %d//%d %d//%d\n", a, a, b, b, c, c);
	}
	const int tail = fprintf(f, "# END");
	const int bad = ferror(f);
	if (fclose(f) || bad || tail < 5) return -1;
	return 0;
}

int write_ply_s(surface *S, const char *filename, const char *author, const char *object)
{
	if (!S || !filename) return -1;
	adjustvectorlenght_s(S);
	FILE *f = fopen(filename, "w");
	if (!f) return -1;
	fprintf(f, "ply\nformat ascii 1.0\ncomment author: %s\ncomment object: %s\n", author ? author : "", object ? object : "");
	fprintf(f, "element vertex %d\nproperty float x\nproperty float y\nproperty float z", S->nV);
	fprintf(f, "\nproperty float nx\nproperty float ny\nproperty float nz");
	fprintf(f, "\nproperty uchar red\nproperty uchar green\nproperty uchar blue\nelement face");
	fprintf(f, " %d\nproperty list uchar int vertex_index\nend_header", S->nT);
	for (unsigned int i = 0; i < S->nV; i++) {
		/* colour bytes in memory order, as the reference takes them (0xAABBGGRR on little endian) */
		const unsigned char *c = (const unsigned char *)&S->color[i];
		fprintf(f, "\n%f %f %f %f %f %f %d %d %d", (double)S->V[i][0], (double)S->V[i][1], (double)S->V[i][2],
		        (double)S->N[i][0], (double)S->N[i][1], (double)S->N[i][2], c[0], c[1], c[2]);
	}
	for (unsigned int i = 0; i < S->nT; i++) fprintf(f, "\n3 %d %d %d", S->T[i][0], S->T[i][1], S->T[i][2]);
	const int tail = fprintf(f, "\n");
	const int bad = ferror(f);
	if (fclose(f) || bad || tail < 1) return -1;
	return 0;
}

/* ------------------------------------------------------------------------- */
/* grids                                                                      */
/* ------------------------------------------------------------------------- */
static _GRD *new_grd(void)
{
	_GRD *Z = (_GRD *)calloc(1, sizeof(_GRD));
	if (!Z) return 0;
	for (int i = 0; i < 3; i++) Z->d[i] = 1.0;
#ifndef GRD_ORTHOGONAL
	for (int i = 0; i < 3; i++) { Z->Ang[i] = 90.0f; Z->_A[i][i] = 1.0; Z->A_[i][i] = 1.0; }
#endif
	return Z;
}

static void set_counts(_GRD *Z, unsigned int npx, unsigned int npy, unsigned int npz)
{
	const unsigned int np[3] = {npx, npy, npz};
	for (int i = 0; i < 3; i++) { Z->N[i] = np[i] - 1; Z->L[i] = (float)(np[i] - 1); }
}

#ifndef GRD_ORTHOGONAL
/* cell matrix of a triclinic lattice (a along x, b in the xy plane) and the matrix the
 * reference pairs with it for the normals: same expressions, same order of operations
 * as MC33_util_grd.c:225-246 (its A_[0][2] is kept as written there) */
static void lattice_matrices(_GRD *Z)
{
	const double rad = M_PI / 180.0;
	const double cal = cos(Z->Ang[0] * rad), cbe = cos(Z->Ang[1] * rad);
	const double ga = Z->Ang[2] * rad, sg = sin(ga), cg = cos(ga);
	const double p = cal - cbe * cg;
	const double h = sqrt(sg * sg + 2 * cal * cbe * cg - cal * cal - cbe * cbe);
	const double isg = 1.0 / sg, ih = 1.0 / h;
	memset(Z->_A, 0, sizeof Z->_A);
	memset(Z->A_, 0, sizeof Z->A_);
	Z->_A[0][0] = 1.0; Z->_A[0][1] = cg;  Z->_A[0][2] = cbe;
	Z->_A[1][1] = sg;  Z->_A[1][2] = p * isg;
	Z->_A[2][2] = h * isg;
	Z->A_[0][0] = 1.0; Z->A_[0][1] = -cg * isg; Z->A_[0][2] = (cg * p - cal * sg * sg) * isg * ih;
	Z->A_[1][1] = isg; Z->A_[1][2] = -p * isg * ih;
	Z->A_[2][2] = sg * ih;
}
#endif

#if GRD_TYPE_SIZE == 8
#define SCAN_SAMPLE "%lf"
typedef double scan_t;
#else
#define SCAN_SAMPLE "%f"
typedef float scan_t;
#endif

/* DMol .grd text files */
_GRD *read_grd(const char *filename)
{
	if (!filename) return 0;
	FILE *f = fopen(filename, "r");
	if (!f) return 0;
	_GRD *Z = new_grd();
	if (!Z) { fclose(f); return 0; }
	char line[128];
	float ang[3] = {90, 90, 90};
	int order = 0, x0[3] = {0, 0, 0}, np[3] = {0, 0, 0};
	int ok = fgets(Z->title, 159, f) != 0 && fgets(line, 60, f) != 0 && fgets(line, 60, f) != 0 &&
	         sscanf(line, "%f %f %f %f %f %f", &Z->L[0], &Z->L[1], &Z->L[2], &ang[0], &ang[1], &ang[2]) >= 3 &&
	         fgets(line, 60, f) != 0 && sscanf(line, "%d %d %d", &np[0], &np[1], &np[2]) == 3 &&
	         fgets(line, 60, f) != 0 && sscanf(line, "%d %d %*d %d %*d %d %*d", &order, &x0[0], &x0[1], &x0[2]) == 4;
	if (!ok || np[0] < 2 || np[1] < 2 || np[2] < 2 || (order != 1 && order != 3)) { fclose(f); free(Z); return 0; }
	for (int i = 0; i < 3; i++) {
		Z->N[i] = (unsigned int)np[i];          /* the header gives interval counts */
		Z->d[i] = Z->L[i] / Z->N[i];
		Z->r0[i] = x0[i] * Z->d[i];
		if (x0[i] == 0) Z->periodic |= 1 << i;
	}
#ifndef GRD_ORTHOGONAL
	for (int i = 0; i < 3; i++) Z->Ang[i] = ang[i];
	Z->nonortho = ang[0] != 90 || ang[1] != 90 || ang[2] != 90;
	if (Z->nonortho) lattice_matrices(Z);
#endif
	if (alloc_F(Z)) { fclose(f); free_memory_grd(Z); return 0; }
	/* order 1: x fastest; order 3: y fastest within a slice */
	const unsigned int n_in = order == 1 ? Z->N[0] : Z->N[1], n_out = order == 1 ? Z->N[1] : Z->N[0];
	for (unsigned int k = 0; k <= Z->N[2]; k++)
		for (unsigned int o = 0; o <= n_out; o++)
			for (unsigned int q = 0; q <= n_in; q++) {
				scan_t v = 0;
				if (fscanf(f, SCAN_SAMPLE, &v) != 1) v = 0;
				if (order == 1) Z->F[k][o][q] = (GRD_data_type)v;
				else Z->F[k][q][o] = (GRD_data_type)v;
			}
	fclose(f);
	return Z;
}

/* the library's own binary grid format ("_GRD") */
_GRD *read_grd_binary(const char *filename)
{
	if (!filename) return 0;
	FILE *f = fopen(filename, "rb");
	if (!f) return 0;
	unsigned int w = 0;
	if (!get(f, &w, 4) || w != 0x4452475fu || !get(f, &w, 4) || w > 159) { fclose(f); return 0; }
	_GRD *Z = new_grd();
	if (!Z) { fclose(f); return 0; }
	int nono = 0;
	int ok = get(f, Z->title, w) && get(f, Z->N, sizeof Z->N) && get(f, Z->L, sizeof Z->L) &&
	         get(f, Z->r0, sizeof Z->r0) && get(f, Z->d, sizeof Z->d) && get(f, &nono, sizeof nono);
#ifndef GRD_ORTHOGONAL
	if (ok && nono) {
		ok = get(f, Z->Ang, sizeof Z->Ang) && get(f, Z->_A, sizeof Z->_A) && get(f, Z->A_, sizeof Z->A_);
		mult_Abf = _multA_bf;       /* as the reference does: files carry full matrices */
	}
	Z->nonortho = nono;
#else
	if (ok && nono) ok = fseek(f, 3 * sizeof(float) + 18 * sizeof(double), SEEK_CUR) == 0;
#endif
	if (Z->r0[0] == 0 && Z->r0[1] == 0 && Z->r0[2] == 0) Z->periodic = 1;
	if (!ok || !Z->N[0] || !Z->N[1] || !Z->N[2] || mc33_alloc_F_block(Z)) { fclose(f); free_memory_grd(Z); return 0; }
	/* rows are contiguous in the block: one read for the whole volume */
	const size_t total = ((size_t)Z->N[0] + 1) * ((size_t)Z->N[1] + 1) * ((size_t)Z->N[2] + 1);
	const size_t got = fread(Z->F[0][0], sizeof(GRD_data_type), total, f);
	(void)got;                      /* a short file leaves the tail unset, as in the reference */
	fclose(f);
	return Z;
}

/* One file per slice of res x res 16-bit samples; the name ends in the slice number
 * and every existing file from that number on is read; slices end up in reverse file order. */
_GRD *read_scanfiles(const char *filename, unsigned int res, int order)
{
	if (!filename || res < 2) return 0;
	_GRD *Z = new_grd();
	if (!Z) return 0;
	Z->internal_data = MC33_GRD_ROWS;
	const size_t len = strlen(filename);
	size_t stem = len;
	while (stem > 0 && filename[stem - 1] >= '0' && filename[stem - 1] <= '9') stem--;
	unsigned int number = (unsigned int)atoi(filename + stem);
	char *name = (char *)malloc(stem + 16);
	uint16_t *buf = (uint16_t *)malloc((size_t)res * sizeof(uint16_t));
	if (!name || !buf) { free(name); free(buf); free(Z); return 0; }
	memcpy(name, filename, stem);
	const int swap = MC33_HOST_BE ? !order : order != 0;
	unsigned int nslices = 0, cap = 0;
	GRD_data_type ***F = 0;
	for (;; number++) {
		sprintf(name + stem, "%-d", (int)number);
		FILE *f = fopen(name, "rb");
		if (!f) break;
		if (nslices == cap) {
			GRD_data_type ***g = (GRD_data_type ***)realloc(F, ((size_t)cap + 64) * sizeof(void *));
			if (!g) { fclose(f); break; }
			F = g; cap += 64;
		}
		GRD_data_type **rows = (GRD_data_type **)calloc(res, sizeof(void *));
		int good = rows != 0;
		for (unsigned int j = 0; good && j < res; j++) {
			rows[j] = (GRD_data_type *)malloc((size_t)res * sizeof(GRD_data_type));
			if (!rows[j]) { good = 0; break; }
			size_t n = fread(buf, sizeof(uint16_t), res, f);
			for (size_t i = n; i < res; i++) buf[i] = 0;
			for (unsigned int i = 0; i < res; i++)
				rows[j][i] = (GRD_data_type)(swap ? (uint16_t)((buf[i] >> 8) | (buf[i] << 8)) : buf[i]);
		}
		fclose(f);
		if (!good) {
			if (rows) { for (unsigned int j = 0; j < res; j++) free(rows[j]); free(rows); }
			break;
		}
		F[nslices++] = rows;
	}
	free(name); free(buf);
	if (nslices < 2) {
		/* not a volume: release what was read */
		for (unsigned int k = 0; k < nslices; k++) { for (unsigned int j = 0; j < res; j++) free(F[k][j]); free(F[k]); }
		free(F); free(Z);
		return 0;
	}
	/* slice tables swapped end for end over the first (last >> 1) positions, exactly as the
	 * reference does (MC33_util_grd.c:402-407): with an even number of slices its middle pair
	 * stays in file order, and so it does here */
	for (unsigned int a = 0, last = nslices - 1; a < (last >> 1); a++) { GRD_data_type **t = F[a]; F[a] = F[last - a]; F[last - a] = t; }
	Z->F = F;
	set_counts(Z, res, res, nslices);
	return Z;
}

static uint16_t swap16(uint16_t v) { return (uint16_t)((v >> 8) | (v << 8)); }
static uint32_t swap32(uint32_t v) { return (v >> 24) | ((v >> 8) & 0xFF00u) | ((v << 8) & 0xFF0000u) | (v << 24); }
static uint64_t swap64(uint64_t v) { return ((uint64_t)swap32((uint32_t)v) << 32) | swap32((uint32_t)(v >> 32)); }

/* one raw element (1, 2, 4 or 8 bytes, maybe byte swapped, integer or IEEE) -> sample */
static GRD_data_type convert_raw(const unsigned char *p, int size, int swap, int isfloat)
{
	if (size == 1) return (GRD_data_type)p[0];
	if (size == 2) { uint16_t v; memcpy(&v, p, 2); if (swap) v = swap16(v); return (GRD_data_type)v; }
	if (size == 4) {
		uint32_t v; memcpy(&v, p, 4);
		if (swap) v = swap32(v);
		if (!isfloat) return (GRD_data_type)v;
		float x; memcpy(&x, &v, 4);
		return (GRD_data_type)x;
	}
	uint64_t v; memcpy(&v, p, 8);
	if (swap) v = swap64(v);
	double x; memcpy(&x, &v, 8);
	return (GRD_data_type)x;
}

/* Headerless volume of N[0] x N[1] x N[2] samples, x fastest.  byte: element size
 * (1, 2, 4 integer; 4, 8 when isfloat), negative for big-endian files. */
_GRD *read_raw_file(const char *filename, unsigned int *N, int byte, int isfloat)
{
	if (!filename || !N) return 0;
	int size = abs(byte);
	if (isfloat ? (size != 4 && size != 8) : (size < 1 || size > 4 || size == 3)) return 0;
	if (!N[0] || !N[1] || !N[2]) return 0;
	int swap = byte < 0;
	if (MC33_HOST_BE) swap = !swap;
	FILE *f = fopen(filename, "rb");
	if (!f) return 0;
	_GRD *Z = new_grd();
	if (!Z) { fclose(f); return 0; }
	set_counts(Z, N[0], N[1], N[2]);
	if (mc33_alloc_F_block(Z)) { fclose(f); free_memory_grd(Z); return 0; }
	GRD_data_type *dst = Z->F[0][0];
	const size_t total = (size_t)N[0] * N[1] * N[2];
#ifdef GRD_INTEGER
	const int same = !isfloat && !swap && size == (int)sizeof(GRD_data_type);
#else
	const int same = isfloat && !swap && size == (int)sizeof(GRD_data_type);
#endif
	if (same) {
		size_t got = fread(dst, sizeof(GRD_data_type), total, f);
		(void)got;
	} else {
		/* any other element type / byte order is converted sample by sample (the reference has
		 * no branch for a byte-swapped file of the grid's own float type and leaves the samples
		 * unset there, MC33_util_grd.c:467-493; here it is swapped like every other case) */
		enum { CH = 1 << 16 };
		unsigned char *buf = (unsigned char *)malloc((size_t)CH * 8);
		if (!buf) { fclose(f); free_memory_grd(Z); return 0; }
		for (size_t i0 = 0; i0 < total; i0 += CH) {
			const size_t n = total - i0 < CH ? total - i0 : CH;
			const size_t got = fread(buf, (size_t)size, n, f);
			for (size_t i = 0; i < got; i++) dst[i0 + i] = convert_raw(buf + i * size, size, swap, isfloat);
			if (got < n) break;
		}
		free(buf);
	}
	fclose(f);
	return Z;
}

/* .dat volumes (TU Wien): three little-endian uint16 sizes, then uint16 samples; the
 * first slice of the file is the TOP slice of the grid */
_GRD *read_dat_file(const char *filename)
{
	if (!filename) return 0;
	FILE *f = fopen(filename, "rb");
	if (!f) return 0;
	uint16_t hdr[3];
	if (!get(f, hdr, sizeof hdr)) { fclose(f); return 0; }
	if (MC33_HOST_BE) for (int i = 0; i < 3; i++) hdr[i] = swap16(hdr[i]);
	if (!hdr[0] || !hdr[1] || !hdr[2]) { fclose(f); return 0; }
	_GRD *Z = new_grd();
	if (!Z) { fclose(f); return 0; }
	set_counts(Z, hdr[0], hdr[1], hdr[2]);
	if (mc33_alloc_F_block(Z)) { fclose(f); free_memory_grd(Z); return 0; }
	const size_t slice = (size_t)hdr[0] * hdr[1];
#if defined(GRD_INTEGER) && GRD_TYPE_SIZE == 2
	/* samples already have the grid's element type: read slices in place */
	for (unsigned int k = hdr[2]; k-- > 0;) {
		uint16_t *dst = (uint16_t *)Z->F[k][0];
		const size_t got = fread(dst, 2, slice, f);
		if (MC33_HOST_BE) for (size_t i = 0; i < got; i++) dst[i] = swap16(dst[i]);
		if (got < slice) break;
	}
#else
	uint16_t *buf = (uint16_t *)malloc(slice * sizeof(uint16_t));
	if (!buf) { fclose(f); free_memory_grd(Z); return 0; }
	for (unsigned int k = hdr[2]; k-- > 0;) {
		GRD_data_type *dst = Z->F[k][0];
		const size_t got = fread(buf, 2, slice, f);
		for (size_t i = 0; i < got; i++) dst[i] = (GRD_data_type)(MC33_HOST_BE ? swap16(buf[i]) : buf[i]);
		if (got < slice) break;
	}
	free(buf);
#endif
	fclose(f);
	return Z;
}
