/* mc33_oracle_impl.h -- TEST INFRASTRUCTURE ONLY (see mc33_oracle.h).
 *
 * Type-generic body, included once per element type by mc33_oracle.c with
 *   REAL    the reference's MC33_real   (include/marching_cubes_33.h:66-88)
 *   SAMPLE  the reference's GRD_data_type
 *   DIFF_T  the C type in which `F[a] - F[b]` on raw samples is evaluated
 *           (float, double, int for u8/u16 after integer promotion, unsigned
 *           for u32 -- the reference's wrap-around quirk)
 *   FN(x)   name mangling
 * All floating point here is compiled with -O2 -ffp-contract=off so that every
 * product, sum and quotient is individually rounded, as in the reference's
 * strict build.
 */

/* IEEE sign bit, the reference's signbf(): marching_cubes_33.c:392-409 */
static inline unsigned FN(sgn)(REAL x)
{
	if (sizeof(REAL) == 8) {
		uint64_t u; double d = (double)x; memcpy(&u, &d, 8); return (unsigned)(u >> 63);
	} else {
		uint32_t u; float f = (float)x; memcpy(&u, &f, 4); return u >> 31;
	}
}

/* one face test: is v[a]*v[b] < v[c]*v[d], both products rounded to REAL
 * (marching_cubes_33.c:349-364, 373-384) */
static inline int FN(face_lt)(const REAL *v, int f)
{
	static const unsigned char q[6][4] = {
		{0, 5, 1, 4}, {1, 6, 2, 5}, {3, 6, 2, 7}, {0, 7, 3, 4}, {0, 2, 1, 3}, {4, 6, 5, 7}};
	REAL l = v[q[f][0]] * v[q[f][1]];
	REAL r = v[q[f][2]] * v[q[f][3]];
	return l < r;
}

/* per face: index mask, the corner pair pattern when the "gate" corner of the
 * face is set and when it is clear.  Gate corner is 0 for faces 0,3,4 and 6
 * for faces 1,2,5 (marching_cubes_33.c:348-365). */
static const unsigned char FN(fmask)[6] = {0xCC, 0x66, 0x33, 0x99, 0xF0, 0x0F};
static const unsigned char FN(fset)[6]  = {0x84, 0x42, 0x12, 0x81, 0xA0, 0x0A};
static const unsigned char FN(fclr)[6]  = {0x48, 0x24, 0x21, 0x18, 0x50, 0x05};
static const unsigned char FN(fgate)[6] = {0x80, 0x02, 0x02, 0x80, 0x80, 0x02};

/* marching_cubes_33.c:347-367 */
static int FN(face_tests)(int *f, unsigned ind, const REAL *v)
{
	int s = 0;
	for (int j = 0; j < 6; j++) {
		unsigned m = ind & FN(fmask)[j];
		int r = 0;
		if (ind & FN(fgate)[j]) {
			if (m == FN(fset)[j]) r = FN(face_lt)(v, j) ? -1 : 1;
		} else {
			if (m == FN(fclr)[j]) r = FN(face_lt)(v, j) ? 1 : -1;
		}
		f[j] = r;
		s += r;
	}
	return s;
}

/* marching_cubes_33.c:371-386 */
static unsigned FN(face_test1)(int face, const REAL *v)
{
	return FN(face_lt)(v, face) ? FN(fclr)[face] : FN(fset)[face];
}

/* marching_cubes_33.c:431-462 ; operation order kept */
static int FN(interior)(int i, int flag13, const REAL *v)
{
	REAL At = v[4] - v[0], Bt = v[5] - v[1], Ct = v[6] - v[2], Dt = v[7] - v[3];
	REAL p1 = At * Ct, p2 = Bt * Dt;
	REAL t = p1 - p2;
	if (FN(sgn)(t)) {
		if (i & 1) return 0;
	} else {
		if (!(i & 1) || t == 0) return 0;
	}
	{
		REAL a0 = v[3] * Bt, a1 = v[2] * At, a2 = v[1] * Dt, a3 = v[0] * Ct;
		REAL s = a0 - a1;
		s = s + a2;
		s = s - a3;
		s = (REAL)0.5f * s;
		t = s / t;
	}
	if (t > 0 && t < 1) {
		REAL m;
		m = At * t; At = v[0] + m;
		m = Bt * t; Bt = v[1] + m;
		m = Ct * t; Ct = v[2] + m;
		m = Dt * t; Dt = v[3] + m;
		Ct = Ct * At;
		Dt = Dt * Bt;
		if (i & 1) {
			if (Ct < Dt && FN(sgn)(Dt) == 0)
				return (FN(sgn)(Bt) == FN(sgn)(v[i])) + flag13;
		} else {
			if (Ct > Dt && FN(sgn)(Ct) == 0)
				return (FN(sgn)(At) == FN(sgn)(v[i])) + flag13;
		}
	}
	return 0;
}

/* Case selection, marching_cubes_33.c:683-779.  Returns the
 * pattern start in MC33_TRI (reference offset - 127); *mflag receives the
 * reference's "m". */
static unsigned FN(select)(unsigned i, const REAL *v, unsigned *mflag)
{
	unsigned c = MC33_CASE256[i];
	unsigned k = c & 0x7FF, m = (c >> 11) & 1;
	unsigned idx = m ? i : (i ^ 0xFF);
	int f[6];
	int off;
	*mflag = m;
	switch (c >> 12) {
	case 0:
		off = (int)k;
		break;
	case 1: /* MC33 case 3 */
		off = (idx & FN(face_test1)((int)(k >> 2), v)) ? 183 + 2 * (int)k : 159 + (int)k;
		break;
	case 2: /* case 4 */
		off = FN(interior)((int)k, 0, v) ? 239 + 6 * (int)k : 231 + 2 * (int)k;
		break;
	case 3: /* case 6 */
		if (idx & FN(face_test1)((int)(k % 6), v))
			off = 575 + 5 * (int)k;
		else
			off = FN(interior)((int)(k / 6), 0, v) ? 407 + 7 * (int)k : 335 + 3 * (int)k;
		break;
	case 4: /* case 7 */
		switch (FN(face_tests)(f, idx, v)) {
		case -3: off = 695 + 3 * (int)k; break;
		case -1: off = (f[4] + f[5] < 0 ? (f[0] + f[2] < 0 ? 759 : 799) : 719) + 5 * (int)k; break;
		case 1:  off = (f[4] + f[5] < 0 ? 983 : (f[0] + f[2] < 0 ? 839 : 911)) + 9 * (int)k; break;
		default: off = FN(interior)((int)(k >> 1), 0, v) ? 1095 + 9 * (int)k : 1055 + 5 * (int)k;
		}
		break;
	case 5: /* case 10 */
		switch (FN(face_tests)(f, idx, v)) {
		case -2:
			if (k == 2 ? FN(interior)(0, 0, v)
			           : (FN(interior)(0, 0, v) || FN(interior)(k ? 1 : 3, 0, v)))
				off = 1213 + 8 * (int)k;
			else
				off = 1189 + 4 * (int)k;
			break;
		case 0:
			off = (f[2 + k] < 0 ? 1261 : 1285) + 8 * (int)k;
			break;
		default:
			if (k == 2 ? FN(interior)(1, 0, v)
			           : (FN(interior)(2, 0, v) || FN(interior)(k ? 3 : 1, 0, v)))
				off = 1237 + 8 * (int)k;
			else
				off = 1201 + 4 * (int)k;
		}
		break;
	case 6: /* case 12 */
		switch (FN(face_tests)(f, idx, v)) {
		case -2:
			off = FN(interior)((int)((0xDA010Cu >> (2 * k)) & 3), 0, v) ? 1453 + 8 * (int)k : 1357 + 4 * (int)k;
			break;
		case 0:
			off = (f[k >> 1] < 0 ? 1645 : 1741) + 8 * (int)k;
			break;
		default:
			off = FN(interior)((int)((0xA7B7E5u >> (2 * k)) & 3), 0, v) ? 1549 + 8 * (int)k : 1405 + 4 * (int)k;
		}
		break;
	default: /* case 13 */
		{
			int s = FN(face_tests)(f, 165, v);
			if (s < 0) s = -s;
			switch (s) {
			case 0: {
				int kk = ((f[1] < 0) << 1) | (f[5] < 0);
				if (f[0] * f[1] == f[5])
					off = 2157 + 12 * kk;
				else {
					int cc = FN(interior)(kk, 1, v);
					off = 2285 + (cc ? 10 * kk - 40 * cc : 6 * kk);
				}
				break;
			}
			case 2:
				off = 1917 + 10 * ((f[0] < 0 ? (f[2] > 0) : 12 + (f[2] < 0)) +
				                   (f[1] < 0 ? (f[3] < 0) : 6 + (f[3] > 0)));
				if (f[4] > 0) off += 30;
				break;
			case 4: {
				int kk = 21 + 11 * f[0] + 4 * f[1] + 3 * f[2] + 2 * f[3] + f[4];
				if (kk >> 4) kk -= (kk & 32 ? 20 : 10);
				off = 1845 + 3 * kk;
				break;
			}
			default:
				off = 1839 + 2 * f[0];
			}
		}
	}
	return (unsigned)(off - 127);
}

/* ------------------------------------------------------------------------- */
typedef struct {
	const SAMPLE *F;
	uint32_t nx, ny, nz;
	uint64_t NX, NXY;
	REAL iso;
} FN(grid);

static inline SAMPLE FN(raw)(const FN(grid) *g, uint32_t x, uint32_t y, uint32_t z)
{
	return g->F[(uint64_t)z * g->NXY + (uint64_t)y * g->NX + x];
}
/* v = iso - F in MC33_real: marching_cubes_33.c:1840-1855 */
static inline REAL FN(val)(const FN(grid) *g, uint32_t x, uint32_t y, uint32_t z)
{
	return g->iso - FN(raw)(g, x, y, z);
}

/* transverse component of an edge vertex normal, SURVEY.md A.7.
 * (x,y,z) = lower end point P0 of the edge, a = edge axis, c = transverse axis */
static REAL FN(transverse)(const FN(grid) *g, uint32_t x, uint32_t y, uint32_t z, int a, int c, REAL t)
{
	uint32_t p0[3] = {x, y, z}, p1[3] = {x, y, z};
	uint32_t n[3] = {g->nx, g->ny, g->nz};
	REAL one_t = 1 - t;
	p1[a] += 1;
	if (p0[c] == 0) {
		uint32_t q0[3] = {p0[0], p0[1], p0[2]}, q1[3] = {p1[0], p1[1], p1[2]};
		q0[c] += 1; q1[c] += 1;
		REAL d0 = FN(val)(g, q0[0], q0[1], q0[2]) - FN(val)(g, p0[0], p0[1], p0[2]);
		REAL d1 = FN(val)(g, q1[0], q1[1], q1[2]) - FN(val)(g, p1[0], p1[1], p1[2]);
		REAL m0 = d0 * one_t, m1 = d1 * t;
		return m0 + m1;
	} else if (p0[c] == n[c]) {
		uint32_t q0[3] = {p0[0], p0[1], p0[2]}, q1[3] = {p1[0], p1[1], p1[2]};
		q0[c] -= 1; q1[c] -= 1;
		REAL d0 = FN(val)(g, p0[0], p0[1], p0[2]) - FN(val)(g, q0[0], q0[1], q0[2]);
		REAL d1 = FN(val)(g, p1[0], p1[1], p1[2]) - FN(val)(g, q1[0], q1[1], q1[2]);
		REAL m0 = d0 * one_t, m1 = d1 * t;
		return m0 + m1;
	} else {
		uint32_t l0[3] = {p0[0], p0[1], p0[2]}, h0[3] = {p0[0], p0[1], p0[2]};
		uint32_t l1[3] = {p1[0], p1[1], p1[2]}, h1[3] = {p1[0], p1[1], p1[2]};
		l0[c] -= 1; h0[c] += 1; l1[c] -= 1; h1[c] += 1;
		DIFF_T e0 = FN(raw)(g, l0[0], l0[1], l0[2]) - FN(raw)(g, h0[0], h0[1], h0[2]);
		DIFF_T e1 = FN(raw)(g, l1[0], l1[1], l1[2]) - FN(raw)(g, h1[0], h1[1], h1[2]);
		REAL m0 = e0 * one_t, m1 = e1 * t;
		REAL s = m0 + m1;
		return (REAL)0.5f * s;
	}
}

/* MC33_surfint: marching_cubes_33.c:628-649 */
static REAL FN(point_grad)(const FN(grid) *g, uint32_t x, uint32_t y, uint32_t z, int c)
{
	uint32_t p[3] = {x, y, z}, lo[3] = {x, y, z}, hi[3] = {x, y, z};
	uint32_t n[3] = {g->nx, g->ny, g->nz};
	if (p[c] == 0) {
		hi[c] += 1;
		DIFF_T e = FN(raw)(g, x, y, z) - FN(raw)(g, hi[0], hi[1], hi[2]);
		return (REAL)e;
	} else if (p[c] == n[c]) {
		lo[c] -= 1;
		DIFF_T e = FN(raw)(g, lo[0], lo[1], lo[2]) - FN(raw)(g, x, y, z);
		return (REAL)e;
	} else {
		lo[c] -= 1; hi[c] += 1;
		DIFF_T e = FN(raw)(g, lo[0], lo[1], lo[2]) - FN(raw)(g, hi[0], hi[1], hi[2]);
		return (REAL)(0.5f * e);
	}
}

/* MC33_spn0/A/B/C: marching_cubes_33.c:485-621, MC33_util_grd.c:87-112.
 * r[0..2] index-space position, r[3..5] un-normalised normal. */
static void FN(store)(const mc33o_geom *gm, REAL *r, REAL *V, float *N)
{
	REAL D[3] = {(REAL)gm->D[0], (REAL)gm->D[1], (REAL)gm->D[2]};
	REAL O[3] = {(REAL)gm->O[0], (REAL)gm->O[1], (REAL)gm->O[2]};
	switch (gm->store) {
	case MC33O_SPN0:
		for (int i = 0; i < 3; i++) V[i] = r[i];
		break;
	case MC33O_SPNB:
		r[3] = r[3] * (REAL)gm->ca;
		r[4] = r[4] * (REAL)gm->cb;
		/* fall through */
	case MC33O_SPNA:
		for (int i = 0; i < 3; i++) { REAL m = r[i] * D[i]; V[i] = m + O[i]; }
		break;
	default: {
		const double *A = gm->A, *B = gm->Ai;
		REAL c0, c1, c2;
		if (gm->tsa) {
			c0 = (REAL)(A[0] * r[0] + A[1] * r[1] + A[2] * r[2]);
			c1 = (REAL)(A[4] * r[1] + A[5] * r[2]);
			c2 = (REAL)(A[8] * r[2]);
		} else {
			double u = A[0] * r[0] + A[1] * r[1] + A[2] * r[2];
			double w = A[3] * r[0] + A[4] * r[1] + A[5] * r[2];
			c2 = (REAL)(A[6] * r[0] + A[7] * r[1] + A[8] * r[2]);
			c0 = (REAL)u; c1 = (REAL)w;
		}
		V[0] = c0 + O[0]; V[1] = c1 + O[1]; V[2] = c2 + O[2];
		if (gm->tsa) {
			/* transposed upper triangular; evaluation order of the reference
			 * overwrites in place from index 2 downwards */
			c2 = (REAL)(B[2] * r[3] + B[5] * r[4] + B[8] * r[5]);
			c1 = (REAL)(B[1] * r[3] + B[4] * r[4]);
			c0 = (REAL)(B[0] * r[3]);
		} else {
			double u = B[0] * r[3] + B[3] * r[4] + B[6] * r[5];
			double w = B[1] * r[3] + B[4] * r[4] + B[7] * r[5];
			c2 = (REAL)(B[2] * r[3] + B[5] * r[4] + B[8] * r[5]);
			c0 = (REAL)u; c1 = (REAL)w;
		}
		r[3] = c0; r[4] = c1; r[5] = c2;
	}
	}
	{
		REAL a = r[3] * r[3], b = r[4] * r[4], c = r[5] * r[5];
		REAL s = a + b;
		s = s + c;
		/* exact reciprocal square root; the reference uses rsqrtss (<=3.1e-4
		 * length error) so lengths are compared loosely */
		float t = 1.0f / sqrtf((float)s);
		if (gm->normal_neg) t = -t;
		N[0] = t * (float)r[3]; N[1] = t * (float)r[4]; N[2] = t * (float)r[5];
	}
}

static const unsigned char FN(EA)[12] = {0, 1, 3, 0, 4, 5, 7, 4, 0, 1, 2, 3};
static const unsigned char FN(EB)[12] = {1, 2, 2, 3, 5, 6, 6, 7, 4, 5, 6, 7};
static const unsigned char FN(EAX)[12] = {1, 2, 1, 2, 1, 2, 1, 2, 0, 0, 0, 0};
static const unsigned char FN(CX)[8] = {0, 0, 0, 0, 1, 1, 1, 1};
static const unsigned char FN(CY)[8] = {0, 1, 1, 0, 0, 1, 1, 0};
static const unsigned char FN(CZ)[8] = {0, 0, 1, 1, 0, 0, 1, 1};

static void FN(corners)(const FN(grid) *g, uint32_t x, uint32_t y, uint32_t z, REAL *v)
{
	for (int c = 0; c < 8; c++)
		v[c] = FN(val)(g, x + FN(CX)[c], y + FN(CY)[c], z + FN(CZ)[c]);
}

static int FN(extract)(const SAMPLE *data, uint32_t nx, uint32_t ny, uint32_t nz, double iso_d,
                       const mc33o_geom *gm, int count_only, mc33o_mesh *out, uint16_t *pat_out)
{
	FN(grid) G = {data, nx, ny, nz, (uint64_t)nx + 1, ((uint64_t)nx + 1) * ((uint64_t)ny + 1), (REAL)iso_d};
	const FN(grid) *g = &G;
	const uint64_t NX = G.NX, NXY = G.NXY, NP = NXY * ((uint64_t)nz + 1);
	const uint64_t NC = (uint64_t)nx * ny * nz;
	int rc = -1;
	unsigned char *S = (unsigned char *)malloc(NP);   /* bit0: F>iso, bit1: on-iso */
	uint32_t *vid = 0;                                 /* 3 ids per point        */
	uint64_t nShared = 0, nPoint = 0, nCentre = 0, nT = 0, nActive = 0, capT = 0;
	uint32_t *T = 0; uint64_t *tcell = 0; uint16_t *tpat = 0;
	uint64_t *ckeys = 0; uint64_t capC = 0;            /* centre cells           */
	if (!S) goto done;
	memset(out, 0, sizeof(*out));

	for (uint32_t z = 0; z <= nz; z++)
		for (uint32_t y = 0; y <= ny; y++)
			for (uint32_t x = 0; x <= nx; x++) {
				REAL v = FN(val)(g, x, y, z);
				S[(uint64_t)z * NXY + (uint64_t)y * NX + x] = (unsigned char)(FN(sgn)(v) | ((v == 0) << 1));
			}
	if (!pat_out) {
		vid = (uint32_t *)malloc(NP * 3 * sizeof(uint32_t));
		if (!vid) goto done;
		/* shared vertices: SURVEY.md A.6.  Canonical numbering: per point row all
		 * X-plane vertices (X edges and POINT vertices) by x, then Y, then Z. */
		for (uint64_t k = 0; k < NP * 3; k++) vid[k] = 0xFFFFFFFFu;
		for (uint32_t z = 0; z <= nz; z++)
			for (uint32_t y = 0; y <= ny; y++)
				for (int pl = 0; pl < 3; pl++)
					for (uint32_t x = 0; x <= nx; x++) {
						uint64_t p = (uint64_t)z * NXY + (uint64_t)y * NX + x;
						unsigned s = S[p];
						uint32_t *id = vid + 3 * p;
						if (s & 2) {
							unsigned any = 0;
							if (pl != 0) continue;
							if (x > 0) any |= S[p - 1];
							if (x < nx) any |= S[p + 1];
							if (y > 0) any |= S[p - NX];
							if (y < ny) any |= S[p + NX];
							if (z > 0) any |= S[p - NXY];
							if (z < nz) any |= S[p + NXY];
							if (any & 1) { id[0] = (uint32_t)nShared++; nPoint++; }
						} else if (pl == 0) {
							if (x < nx && !(S[p + 1] & 2) && ((S[p + 1] ^ s) & 1)) id[0] = (uint32_t)nShared++;
						} else if (pl == 1) {
							if (y < ny && !(S[p + NX] & 2) && ((S[p + NX] ^ s) & 1)) id[1] = (uint32_t)nShared++;
						} else {
							if (z < nz && !(S[p + NXY] & 2) && ((S[p + NXY] ^ s) & 1)) id[2] = (uint32_t)nShared++;
						}
					}
	}
	/* cells, in the reference's sweep order */
	for (uint32_t z = 0; z < nz; z++)
		for (uint32_t y = 0; y < ny; y++)
			for (uint32_t x = 0; x < nx; x++) {
				uint64_t cell = ((uint64_t)z * ny + y) * nx + x;
				uint64_t pc[8];
				unsigned i = 0, zmask = 0;
				for (int c = 0; c < 8; c++) {
					pc[c] = (uint64_t)(z + FN(CZ)[c]) * NXY + (uint64_t)(y + FN(CY)[c]) * NX + (x + FN(CX)[c]);
					i |= (unsigned)(S[pc[c]] & 1) << (7 - c);
					zmask |= (unsigned)((S[pc[c]] >> 1) & 1) << c;
				}
				if (pat_out) pat_out[cell] = 0xFFFF;
				if (i == 0 || i == 0xFF) continue;
				nActive++;
				REAL v[8];
				unsigned m;
				FN(corners)(g, x, y, z, v);
				unsigned ps = FN(select)(i, v, &m);
				if (pat_out) { pat_out[cell] = (uint16_t)ps; continue; }
				uint32_t centre_id = 0xFFFFFFFFu;
				if (MC33_PAT_CENTRE[ps]) {
					if (nCentre == capC) {
						capC = capC ? capC * 2 : 1024;
						ckeys = (uint64_t *)realloc(ckeys, capC * sizeof(uint64_t));
						if (!ckeys) goto done;
					}
					ckeys[nCentre] = cell;
					centre_id = (uint32_t)nCentre++; /* rebased by nShared below */
				}
				for (unsigned w = ps;; w++) {
					unsigned tw = MC33_TRI[w];
					uint32_t ti[3];
					unsigned key[3];
					unsigned e3[3] = {(tw >> 8) & 15, (tw >> 4) & 15, tw & 15};
					for (int j = 0; j < 3; j++) {
						unsigned e = e3[j];
						if (e == 12) { ti[j] = 0x80000000u | centre_id; key[j] = 12; continue; }
						unsigned a = FN(EA)[e], b = FN(EB)[e];
						if (zmask & (1u << a)) { key[j] = 16 + a; ti[j] = count_only ? 0 : vid[3 * pc[a]]; }
						else if (zmask & (1u << b)) { key[j] = 16 + b; ti[j] = count_only ? 0 : vid[3 * pc[b]]; }
						else { key[j] = e; ti[j] = count_only ? 0 : vid[3 * pc[a] + FN(EAX)[e]]; }
					}
					/* zero-area drop: marching_cubes_33.c:1235 */
					if (key[0] != key[1] && key[0] != key[2] && key[1] != key[2]) {
						if (!count_only) {
							if (nT == capT) {
								capT = capT ? capT * 2 : 4096;
								T = (uint32_t *)realloc(T, capT * 3 * sizeof(uint32_t));
								tcell = (uint64_t *)realloc(tcell, capT * sizeof(uint64_t));
								tpat = (uint16_t *)realloc(tpat, capT * sizeof(uint16_t));
								if (!T || !tcell || !tpat) goto done;
							}
							/* winding: marching_cubes_33.c:1246-1250 with ti[0]=nibble2,
							 * ti[1]=nibble1, ti[2]=nibble0; n = !m */
							uint32_t a0 = m ? ti[0] : ti[1], a1 = m ? ti[1] : ti[0];
							if (gm->normal_neg) { uint32_t s = a0; a0 = a1; a1 = s; }
							T[3 * nT] = a0; T[3 * nT + 1] = a1; T[3 * nT + 2] = ti[2];
							tcell[nT] = cell;
							tpat[nT] = (uint16_t)ps;
						}
						nT++;
					}
					if (!(tw >> 12)) break;
				}
			}
	if (pat_out) { rc = 0; goto done; }
	out->nShared = nShared; out->nCentre = nCentre; out->nPoint = nPoint;
	out->nV = nShared + nCentre; out->nT = nT; out->nActive = nActive;
	if (count_only) { rc = 0; goto done; }
	if (out->nV >= 0x80000000ull) goto done;
	for (uint64_t k = 0; k < 3 * nT; k++)
		if (T[k] & 0x80000000u) T[k] = (uint32_t)(nShared + (T[k] & 0x7FFFFFFFu));

	{
		uint64_t nV = out->nV ? out->nV : 1;
		REAL *V = (REAL *)malloc(nV * 3 * sizeof(REAL));
		float *N = (float *)malloc(nV * 3 * sizeof(float));
		uint64_t *vkey = (uint64_t *)malloc(nV * sizeof(uint64_t));
		out->V = V; out->N = N; out->vkey = vkey;
		if (!V || !N || !vkey) goto done;
		for (uint32_t z = 0; z <= nz; z++)
			for (uint32_t y = 0; y <= ny; y++)
				for (uint32_t x = 0; x <= nx; x++) {
					uint64_t p = (uint64_t)z * NXY + (uint64_t)y * NX + x;
					const uint32_t *id = vid + 3 * p;
					uint32_t P[3] = {x, y, z};
					REAL r[6];
					if (S[p] & 2) {
						if (id[0] == 0xFFFFFFFFu) continue;
						r[0] = (REAL)x; r[1] = (REAL)y; r[2] = (REAL)z;
						for (int c = 0; c < 3; c++) r[3 + c] = FN(point_grad)(g, x, y, z, c);
						FN(store)(gm, r, V + 3 * (uint64_t)id[0], N + 3 * (uint64_t)id[0]);
						vkey[id[0]] = p * 4;
						continue;
					}
					for (int a = 0; a < 3; a++) {
						if (id[a] == 0xFFFFFFFFu) continue;
						uint32_t Q[3] = {x, y, z};
						Q[a] += 1;
						REAL va = FN(val)(g, x, y, z), vb = FN(val)(g, Q[0], Q[1], Q[2]);
						REAL den = va - vb;
						REAL t = va / den;
						for (int c = 0; c < 3; c++) {
							if (c == a) { r[c] = (REAL)P[c] + t; r[3 + c] = vb - va; }
							else { r[c] = (REAL)P[c]; r[3 + c] = FN(transverse)(g, x, y, z, a, c, t); }
						}
						FN(store)(gm, r, V + 3 * (uint64_t)id[a], N + 3 * (uint64_t)id[a]);
						vkey[id[a]] = p * 4 + (unsigned)a;
					}
				}
		for (uint64_t k = 0; k < nCentre; k++) {
			uint64_t cell = ckeys[k];
			uint32_t x = (uint32_t)(cell % nx), y = (uint32_t)((cell / nx) % ny), z = (uint32_t)(cell / ((uint64_t)nx * ny));
			REAL v[8], r[6];
			FN(corners)(g, x, y, z, v);
			/* marching_cubes_33.c:1226-1229 */
			r[0] = (REAL)x + 0.5f; r[1] = (REAL)y + 0.5f; r[2] = (REAL)z + 0.5f;
			r[3] = v[4] + v[5] + v[6] + v[7] - v[0] - v[1] - v[2] - v[3];
			r[4] = v[1] + v[2] + v[5] + v[6] - v[0] - v[3] - v[4] - v[7];
			r[5] = v[2] + v[3] + v[6] + v[7] - v[0] - v[1] - v[4] - v[5];
			FN(store)(gm, r, V + 3 * (nShared + k), N + 3 * (nShared + k));
			vkey[nShared + k] = cell * 4 + 3;
		}
	}
	out->T = T; out->tcell = tcell; out->tpat = tpat;
	T = 0; tcell = 0; tpat = 0;
	rc = 0;
done:
	free(S); free(vid); free(ckeys); free(T); free(tcell); free(tpat);
	(void)NC;
	if (rc) mc33o_free(out);
	return rc;
}
