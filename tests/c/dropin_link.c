/* dropin_link.c -- TEST PROGRAM: a C translation unit that includes the public header
 * and links -lMC33_b200_<variant> exactly as a user of the reference links -lMC33
 * (reference README.md:84-88, :135-155).  Built by tests/test_c_link.py with the
 * variant's -D flags.  It checks at compile time the struct layouts SURVEY.md 8(a)
 * lists, then runs the README flow:
 *   generate_grid_from_fn / grid_from_data_pointer -> create_MC33 -> calculate_isosurface
 *   -> write_bin_s / read_bin_s / write_obj_s / write_ply_s / write_txt_s -> free_*
 * and prints one line per step.  Without a CUDA device create_MC33 returns NULL
 * (there is no CPU fallback); the program then prints NO_DEVICE and still exercises
 * the host-only entry points, so it is useful in the CPU test suite too. */
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <marching_cubes_33.h>

/* ---- layouts (probe of the reference header, x86-64) ---- */
_Static_assert(sizeof(surface) == 64, "surface size");
_Static_assert(offsetof(surface, T) == 0 && offsetof(surface, V) == 8 && offsetof(surface, N) == 16 &&
               offsetof(surface, color) == 24 && offsetof(surface, nV) == 32 && offsetof(surface, nT) == 36 &&
               offsetof(surface, capt) == 40 && offsetof(surface, capv) == 44 && offsetof(surface, iso) == 48 &&
               offsetof(surface, user) == 56, "surface offsets");
_Static_assert(offsetof(_GRD, F) == 0 && offsetof(_GRD, N) == 8 && offsetof(_GRD, r0) == 24 && offsetof(_GRD, d) == 48 &&
               offsetof(_GRD, L) == 72, "_GRD head offsets");
#ifdef GRD_ORTHOGONAL
_Static_assert(sizeof(_GRD) == 256, "_GRD size (GRD_ORTHOGONAL)");
_Static_assert(sizeof(MC33) == 160, "MC33 size (GRD_ORTHOGONAL)");
#else
_Static_assert(sizeof(_GRD) == 416, "_GRD size");
_Static_assert(offsetof(_GRD, Ang) == 84 && offsetof(_GRD, nonortho) == 96 && offsetof(_GRD, _A) == 104 &&
               offsetof(_GRD, A_) == 176 && offsetof(_GRD, periodic) == 248 && offsetof(_GRD, internal_data) == 252 &&
               offsetof(_GRD, title) == 256, "_GRD offsets");
#if GRD_TYPE_SIZE == 8
_Static_assert(sizeof(MC33) == 344, "MC33 size (double)");
#else
_Static_assert(sizeof(MC33) == 304, "MC33 size");
#endif
#endif
_Static_assert(offsetof(MC33, T) == 0 && offsetof(MC33, nV) == 32 && offsetof(MC33, iso) == 48, "MC33 mirrors the surface head");

#if !defined(GRD_INTEGER)
static double fn(double x, double y, double z) { return cos(x) + cos(y) + cos(z); }
#endif

static int same_surface(const surface *a, const surface *b)
{
	return a->nV == b->nV && a->nT == b->nT && a->iso == b->iso &&
	       !memcmp(a->T, b->T, (size_t)a->nT * 12) && !memcmp(a->V, b->V, (size_t)a->nV * 3 * sizeof(MC33_real)) &&
	       !memcmp(a->N, b->N, (size_t)a->nV * 12) && !memcmp(a->color, b->color, (size_t)a->nV * 4);
}

int main(int argc, char **argv)
{
	const char *dir = argc > 1 ? argv[1] : ".";
	char path[1024];
	int fails = 0;

	/* ---- host-only entry points: NULL safety (reference: every free_* accepts NULL) ---- */
	free_surface_memory(0); free_MC33(0); free_memory_grd(0); adjustvectorlenght_s(0);
	if (create_MC33(0) || grid_from_data_pointer(0, 2, 2, 0)) { puts("FAIL null args"); fails++; }

	/* ---- grids ---- */
#if defined(GRD_INTEGER)
	const unsigned int n = 24;
	GRD_data_type *data = (GRD_data_type *)malloc((size_t)n * n * n * sizeof(GRD_data_type));
	for (unsigned int k = 0; k < n; k++)
		for (unsigned int j = 0; j < n; j++)
			for (unsigned int i = 0; i < n; i++) {
				const double dx = i - 11.5, dy = j - 11.5, dz = k - 11.5;
				const double v = 100.0 - 8.0 * sqrt(dx * dx + dy * dy + dz * dz) + 0.5;
				data[((size_t)k * n + j) * n + i] = (GRD_data_type)(v < 0 ? 0 : v);
			}
	_GRD *G = grid_from_data_pointer(n, n, n, data);
	const MC33_real iso = 40.0f;
#else
	GRD_data_type *data = 0;
	/* README.md:135-155 -> 201^3 samples, r0 = -4, d = 0.04 (BASELINE config 1) */
	_GRD *G = generate_grid_from_fn(-4, -4, -4, 4, 4, 4, .04, .04, .04, fn);
	const MC33_real iso = 0;
#endif
	if (!G) { puts("FAIL grid"); return 1; }
	printf("GRID %u %u %u internal_data=%d\n", G->N[0], G->N[1], G->N[2], G->internal_data);

	MC33 *M = create_MC33(G);
	if (!M) {
		puts("NO_DEVICE");
	} else {
		unsigned int nV = 0, nT = 0;
		unsigned long long bytes = size_of_isosurface(M, iso, &nV, &nT);
		printf("SIZE %llu %u %u\n", bytes, nV, nT);
		surface *S = calculate_isosurface(M, iso);
		if (!S) { puts("FAIL calculate_isosurface"); return 1; }
		printf("SURFACE %u %u capv=%u capt=%u iso=%g color0=%08x\n", S->nV, S->nT, S->capv, S->capt, (double)S->iso,
		       S->nV ? (unsigned)S->color[0] : 0u);
		if (S->nV != nV || S->nT != nT) { puts("FAIL size_of_isosurface disagrees"); fails++; }
		/* callers mutate the arrays in place (reference demos recolour and flip normals) */
		for (unsigned int i = 0; i < S->nV; i += 7) S->color[i] = (int)0xff0000ffu;
		snprintf(path, sizeof path, "%s/s.bin", dir);
		if (write_bin_s(S, path)) { puts("FAIL write_bin_s"); fails++; }
		surface *R = read_bin_s(path);
		if (!R || !same_surface(S, R)) { puts("FAIL read_bin_s round trip"); fails++; } else puts("BIN_ROUNDTRIP ok");
		free_surface_memory(R);
		snprintf(path, sizeof path, "%s/s.obj", dir); if (write_obj_s(S, path)) { puts("FAIL write_obj_s"); fails++; }
		snprintf(path, sizeof path, "%s/s.ply", dir); if (write_ply_s(S, path, "mc33-b200", "dropin_link")) { puts("FAIL write_ply_s"); fails++; }
		snprintf(path, sizeof path, "%s/s.txt", dir); if (write_txt_s(S, path)) { puts("FAIL write_txt_s"); fails++; }
		/* a second isovalue on the same MC33, then an empty one (zero-filled struct, not NULL) */
		surface *S2 = calculate_isosurface(M, iso + (MC33_real)0.5);
		if (!S2) { puts("FAIL second isovalue"); fails++; } else printf("SURFACE2 %u %u\n", S2->nV, S2->nT);
		free_surface_memory(S2);
		surface *E = calculate_isosurface(M, (MC33_real)1e6);
		if (!E || E->nV || E->nT || E->T || E->V || E->iso != 0) { puts("FAIL empty surface"); fails++; } else puts("EMPTY ok");
		free_surface_memory(E);
		free_surface_memory(S);
		free_MC33(M);
	}
	free_memory_grd(G);
	free(data);

	/* ---- grid readers on files this program writes itself ---- */
	{
		unsigned int N[3] = {5, 4, 3};
		unsigned short raw[60];
		for (int i = 0; i < 60; i++) raw[i] = (unsigned short)(i * 37 % 1000);
		snprintf(path, sizeof path, "%s/v.raw", dir);
		FILE *f = fopen(path, "wb"); fwrite(raw, 2, 60, f); fclose(f);
		_GRD *Z = read_raw_file(path, N, 2, 0);
		if (!Z || Z->N[0] != 4 || Z->N[1] != 3 || Z->N[2] != 2 || Z->F[2][3][4] != (GRD_data_type)raw[59] ||
		    Z->F[1][2][3] != (GRD_data_type)raw[(1 * 4 + 2) * 5 + 3]) { puts("FAIL read_raw_file"); fails++; } else puts("RAW ok");
		free_memory_grd(Z);
		snprintf(path, sizeof path, "%s/v.dat", dir);
		unsigned short hdr[3] = {5, 4, 3};
		f = fopen(path, "wb"); fwrite(hdr, 2, 3, f); fwrite(raw, 2, 60, f); fclose(f);
		Z = read_dat_file(path);
		/* first slice of the file is the top slice */
		if (!Z || Z->N[2] != 2 || Z->F[2][0][0] != (GRD_data_type)raw[0] || Z->F[0][3][4] != (GRD_data_type)raw[59]) { puts("FAIL read_dat_file"); fails++; } else puts("DAT ok");
		free_memory_grd(Z);
		if (read_raw_file(path, N, 3, 0) || read_dat_file("/nonexistent/x") || read_grd("/nonexistent/x") ||
		    read_grd_binary("/nonexistent/x") || read_bin_s("/nonexistent/x")) { puts("FAIL error returns"); fails++; }
	}
	printf("DONE fails=%d\n", fails);
	return fails != 0;
}
