// mc33_core.cuh -- per-word building blocks of the B200 Marching Cubes 33 path.
//
// Everything here is __host__ __device__ so that the kernels in mc33_kernels.cu
// stay thin and the same logic can be stepped on a CPU by the test-only harness
// tests/hostemu (never linked into the product).
//
// Design: the grid is streamed ONCE by the classify kernel, which
// leaves two bitmaps per point row: S (sample > iso, i.e. IEEE sign bit of
// iso - F, reference marching_cubes_33.c:1840-1859 and :392-409) and Z (sample
// exactly on the isovalue).  All topology -- which grid edges carry a vertex,
// which cells are active, the 8-bit case index of a cell -- is then derived
// with 32-cells-per-thread word operations from those bitmaps; sample values
// are only touched again for ambiguous cells (face / interior tests) and for
// the vertices themselves.
//
// Vertex ownership replaces the reference's slice-to-slice reuse tables
// (Dx,Dy,Ux,Uy,Lz; marching_cubes_33.c:780-1253):
//   grid point p owns  X: edge p->p+ex (or the POINT vertex when the sample at
//                         p is exactly on the isovalue),
//                      Y: edge p->p+ey,   Z: edge p->p+ez;
//   a cell owns its CENTRE vertex (edge code 12).
// Canonical numbering: per point row (z,y): all X-plane vertices by x, then the
// Y-plane, then the Z-plane; rows in (z,y) order; centre vertices after all
// shared ones, by cell in (z,y,x) order.  Triangles by cell in (z,y,x) order
// (the reference's sweep order) then table order.
#pragma once
#include <stdint.h>
#include <math.h>
#include "mc33_tables.h"

#if defined(__CUDACC__)
#define MC_HD __host__ __device__ __forceinline__
#define MC_HDN __host__ __device__
#else
#define MC_HD inline
#define MC_HDN
#endif

namespace mc33 {

enum { DT_F32 = 0, DT_F64 = 1, DT_U8 = 2, DT_U16 = 3, DT_U32 = 4 };
enum { STORE_SPN0 = 0, STORE_SPNA = 1, STORE_SPNB = 2, STORE_SPNC = 3 };

// ---------------------------------------------------------------------------
// individually rounded arithmetic: no FMA contraction, no reassociation, so
// that case selection and interpolation match the reference bit for bit
// (SURVEY.md 7.3-1).  The .cu files are also compiled with -fmad=false.
// ---------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
MC_HD float  rmul(float a, float b)   { return __fmul_rn(a, b); }
MC_HD float  radd(float a, float b)   { return __fadd_rn(a, b); }
MC_HD float  rsub(float a, float b)   { return __fsub_rn(a, b); }
MC_HD float  rdiv(float a, float b)   { return __fdiv_rn(a, b); }
MC_HD double rmul(double a, double b) { return __dmul_rn(a, b); }
MC_HD double radd(double a, double b) { return __dadd_rn(a, b); }
MC_HD double rsub(double a, double b) { return __dsub_rn(a, b); }
MC_HD double rdiv(double a, double b) { return __ddiv_rn(a, b); }
MC_HD int popc32(uint32_t v) { return __popc(v); }
MC_HD int popc64(uint64_t v) { return __popcll(v); }
MC_HD int ffs32(uint32_t v) { return __ffs((int)v) - 1; }
#else
MC_HD float  rmul(float a, float b)   { volatile float r = a * b; return r; }
MC_HD float  radd(float a, float b)   { volatile float r = a + b; return r; }
MC_HD float  rsub(float a, float b)   { volatile float r = a - b; return r; }
MC_HD float  rdiv(float a, float b)   { volatile float r = a / b; return r; }
MC_HD double rmul(double a, double b) { volatile double r = a * b; return r; }
MC_HD double radd(double a, double b) { volatile double r = a + b; return r; }
MC_HD double rsub(double a, double b) { volatile double r = a - b; return r; }
MC_HD double rdiv(double a, double b) { volatile double r = a / b; return r; }
MC_HD int popc32(uint32_t v) { return __builtin_popcount(v); }
MC_HD int popc64(uint64_t v) { return __builtin_popcountll(v); }
MC_HD int ffs32(uint32_t v) { return __builtin_ffs((int)v) - 1; }
#endif

MC_HD unsigned sgn(float x)
{
#if defined(__CUDA_ARCH__)
	return __float_as_uint(x) >> 31;
#else
	uint32_t u; __builtin_memcpy(&u, &x, 4); return u >> 31;
#endif
}
MC_HD unsigned sgn(double x)
{
#if defined(__CUDA_ARCH__)
	return (unsigned)(__double2hiint(x)) >> 31;
#else
	uint64_t u; __builtin_memcpy(&u, &x, 8); return (unsigned)(u >> 63);
#endif
}

// element type traits: Real = the reference's MC33_real for that build
// (include/marching_cubes_33.h:66-88)
template <typename S> struct Traits;
template <> struct Traits<float>    { typedef float Real;  enum { code = DT_F32 }; };
template <> struct Traits<double>   { typedef double Real; enum { code = DT_F64 }; };
template <> struct Traits<uint8_t>  { typedef float Real;  enum { code = DT_U8 }; };
template <> struct Traits<uint16_t> { typedef float Real;  enum { code = DT_U16 }; };
template <> struct Traits<uint32_t> { typedef float Real;  enum { code = DT_U32 }; };

// raw sample difference as the reference's C expression `F[a] - F[b]` yields it
// (float/double: that type; u8/u16: int after promotion; u32: wraps)
MC_HD float  rawdiff(float a, float b)       { return rsub(a, b); }
MC_HD double rawdiff(double a, double b)     { return rsub(a, b); }
MC_HD float  rawdiff(uint8_t a, uint8_t b)   { return (float)((int)a - (int)b); }
MC_HD float  rawdiff(uint16_t a, uint16_t b) { return (float)((int)a - (int)b); }
MC_HD float  rawdiff(uint32_t a, uint32_t b) { return (float)(uint32_t)(a - b); }

// ---------------------------------------------------------------------------
struct Geom {
	int store, normal_neg, tsa, pad_;
	double O[3], D[3], ca, cb;   // values already narrowed to Real by the caller
	double A[9], Ai[9];          // scaled matrices of create_MC33 (c:1763-1769)
};

struct Totals {                   // written by the scan kernel
	uint32_t nShared;             // shared vertices owned by this slab
	uint32_t nCentre;
	uint32_t nT;
	uint32_t nSharedHalo;         // shared vertices of the halo slice (next slab's first)
	uint32_t overflow;            // set by emit kernels if a capacity was exceeded
	uint32_t pad_[3];
};

struct Tables {
	const uint16_t *case256, *simple256, *tri;
	const uint8_t *pat;           // ntri | centre << 7, at pattern starts
};

struct Params {
	const void *data;             // samples of slices [zlo,zhi), x fastest
	uint32_t nx, ny, nz;          // GLOBAL interval counts (_GRD.N)
	uint32_t NX, NY;              // nx+1, ny+1
	uint32_t zlo, zhi;            // global sample slices held in `data`
	uint32_t pz0, pz1;            // point slices whose shared vertices this slab owns
	uint32_t cz0, cz1;            // cell layers this slab owns
	uint32_t hz;                  // halo point slice numbered in the next slab (== pz1) or 0xFFFFFFFF
	uint32_t W, WC, WP;           // words per point row, per cell row, row stride (>= W+1)
	uint32_t Lrows;               // (zhi-zlo)*NY
	uint32_t *S, *Z;              // bitmaps [Lrows][WP]
	uint8_t *rowZ;                // [Lrows] any Z bit in the row
	uint64_t *wpreV;              // [Lrows][W]  row-local exclusive prefix: X | Y<<21 | Z<<42
	uint64_t *wpreC;              // [Lrows][W]  row-local exclusive prefix: T | C<<32
	uint32_t *rowNX, *rowNY, *rowNZ, *rowNC, *rowNT;   // [Lrows] counts
	uint32_t *rowBX, *rowBY, *rowBZ, *rowBC, *rowBT;   // [Lrows] exclusive bases (scan)
	Totals *totals;
	double iso;
	Geom geom;
	// outputs (device)
	void *V; float *N; int32_t *color; uint32_t *T;
	uint64_t *vkey, *tcell;       // optional canonical keys (tests)
	uint32_t capV, capT;
	uint32_t vbase, vbase_next;   // global vertex id of this / the next slab's first vertex (0 on one GPU)
	const uint32_t *dbases;       // optional device copy {vbase, vbase_next}: overrides the two above
	int32_t color_value;
};

// corner c of a cell -> offsets; edge e -> end points, axis
#define MC_CX(c) (((c) >> 2) & 1)
#define MC_CY(c) ((0x66 >> (c)) & 1)
#define MC_CZ(c) ((0xCC >> (c)) & 1)

template <typename Sample>
MC_HD Sample ld_sample(const Params &P, uint32_t x, uint32_t y, uint32_t z)
{
	const Sample *F = (const Sample *)P.data;
	return F[((uint64_t)(z - P.zlo) * P.NY + y) * P.NX + x];
}
template <typename Sample>
MC_HD typename Traits<Sample>::Real ld_val(const Params &P, typename Traits<Sample>::Real iso, uint32_t x, uint32_t y, uint32_t z)
{
	typedef typename Traits<Sample>::Real Real;
	return rsub(iso, (Real)ld_sample<Sample>(P, x, y, z));
}

// ---------------------------------------------------------------------------
// MC33 disambiguation (reference marching_cubes_33.c:347-386, :431-462, :683-779)
// ---------------------------------------------------------------------------
template <typename Real>
MC_HD bool face_lt(const Real *v, int f)
{
	// corner quadruples {a,b,c,d}: test v[a]*v[b] < v[c]*v[d]
	const uint32_t qa = 0x400310u, qb = 0x627665u, qc = 0x513221u, qd = 0x734754u; // nibble f
	int a = (qa >> (4 * f)) & 15, b = (qb >> (4 * f)) & 15, c = (qc >> (4 * f)) & 15, d = (qd >> (4 * f)) & 15;
	return rmul(v[a], v[b]) < rmul(v[c], v[d]);
}

MC_HD unsigned face_mask(int f) { return (0x0FF0993366CCull >> (8 * f)) & 0xFF; }
MC_HD unsigned face_set(int f)  { return (0x0AA081124284ull >> (8 * f)) & 0xFF; }
MC_HD unsigned face_clr(int f)  { return (0x055018212448ull >> (8 * f)) & 0xFF; }
MC_HD unsigned face_gate(int f) { return (0x028080020280ull >> (8 * f)) & 0xFF; }

template <typename Real>
MC_HD int face_tests(int *fr, unsigned ind, const Real *v)
{
	int s = 0;
#pragma unroll
	for (int j = 0; j < 6; j++) {
		unsigned m = ind & face_mask(j);
		int r = 0;
		if (ind & face_gate(j)) {
			if (m == face_set(j)) r = face_lt(v, j) ? -1 : 1;
		} else {
			if (m == face_clr(j)) r = face_lt(v, j) ? 1 : -1;
		}
		fr[j] = r;
		s += r;
	}
	return s;
}

template <typename Real>
MC_HD unsigned face_test1(int f, const Real *v)
{
	return face_lt(v, f) ? face_clr(f) : face_set(f);
}

template <typename Real>
MC_HD int interior_test(int i, int flag13, const Real *v)
{
	Real At = rsub(v[4], v[0]), Bt = rsub(v[5], v[1]), Ct = rsub(v[6], v[2]), Dt = rsub(v[7], v[3]);
	Real t = rsub(rmul(At, Ct), rmul(Bt, Dt));
	if (sgn(t)) {
		if (i & 1) return 0;
	} else {
		if (!(i & 1) || t == (Real)0) return 0;
	}
	Real s = rsub(rmul(v[3], Bt), rmul(v[2], At));
	s = radd(s, rmul(v[1], Dt));
	s = rsub(s, rmul(v[0], Ct));
	s = rmul((Real)0.5f, s);
	t = rdiv(s, t);
	if (t > (Real)0 && t < (Real)1) {
		At = radd(v[0], rmul(At, t));
		Bt = radd(v[1], rmul(Bt, t));
		Ct = radd(v[2], rmul(Ct, t));
		Dt = radd(v[3], rmul(Dt, t));
		Ct = rmul(Ct, At);
		Dt = rmul(Dt, Bt);
		if (i & 1) {
			if (Ct < Dt && sgn(Dt) == 0) return (int)(sgn(Bt) == sgn(v[i])) + flag13;
		} else {
			if (Ct > Dt && sgn(Ct) == 0) return (int)(sgn(At) == sgn(v[i])) + flag13;
		}
	}
	return 0;
}

// -> pattern start in MC33_TRI; m = the reference's winding flag
template <typename Real>
MC_HDN unsigned select_pattern(const Tables &tb, unsigned i, const Real *v, unsigned *mflag)
{
	unsigned c = tb.case256[i];
	int k = (int)(c & 0x7FF);
	unsigned m = (c >> 11) & 1;
	unsigned idx = m ? i : (i ^ 0xFF);
	int f[6];
	int off;
	*mflag = m;
	switch (c >> 12) {
	case 0:
		off = k;
		break;
	case 1:
		off = (idx & face_test1(k >> 2, v)) ? 183 + 2 * k : 159 + k;
		break;
	case 2:
		off = interior_test(k, 0, v) ? 239 + 6 * k : 231 + 2 * k;
		break;
	case 3:
		if (idx & face_test1(k % 6, v))
			off = 575 + 5 * k;
		else
			off = interior_test(k / 6, 0, v) ? 407 + 7 * k : 335 + 3 * k;
		break;
	case 4:
		switch (face_tests(f, idx, v)) {
		case -3: off = 695 + 3 * k; break;
		case -1: off = (f[4] + f[5] < 0 ? (f[0] + f[2] < 0 ? 759 : 799) : 719) + 5 * k; break;
		case 1:  off = (f[4] + f[5] < 0 ? 983 : (f[0] + f[2] < 0 ? 839 : 911)) + 9 * k; break;
		default: off = interior_test(k >> 1, 0, v) ? 1095 + 9 * k : 1055 + 5 * k;
		}
		break;
	case 5:
		switch (face_tests(f, idx, v)) {
		case -2:
			if (k == 2 ? interior_test(0, 0, v) != 0
			           : (interior_test(0, 0, v) != 0 || interior_test(k ? 1 : 3, 0, v) != 0))
				off = 1213 + 8 * k;
			else
				off = 1189 + 4 * k;
			break;
		case 0:
			off = (f[2 + k] < 0 ? 1261 : 1285) + 8 * k;
			break;
		default:
			if (k == 2 ? interior_test(1, 0, v) != 0
			           : (interior_test(2, 0, v) != 0 || interior_test(k ? 3 : 1, 0, v) != 0))
				off = 1237 + 8 * k;
			else
				off = 1201 + 4 * k;
		}
		break;
	case 6:
		switch (face_tests(f, idx, v)) {
		case -2:
			off = interior_test((int)((0xDA010Cu >> (2 * k)) & 3), 0, v) ? 1453 + 8 * k : 1357 + 4 * k;
			break;
		case 0:
			off = (f[k >> 1] < 0 ? 1645 : 1741) + 8 * k;
			break;
		default:
			off = interior_test((int)((0xA7B7E5u >> (2 * k)) & 3), 0, v) ? 1549 + 8 * k : 1405 + 4 * k;
		}
		break;
	default: {
		int s = face_tests(f, 165u, v);
		if (s < 0) s = -s;
		if (s == 0) {
			int kk = ((f[1] < 0) << 1) | (f[5] < 0);
			if (f[0] * f[1] == f[5])
				off = 2157 + 12 * kk;
			else {
				int cc = interior_test(kk, 1, v);
				off = 2285 + (cc ? 10 * kk - 40 * cc : 6 * kk);
			}
		} else if (s == 2) {
			off = 1917 + 10 * ((f[0] < 0 ? (int)(f[2] > 0) : 12 + (int)(f[2] < 0)) +
			                   (f[1] < 0 ? (int)(f[3] < 0) : 6 + (int)(f[3] > 0)));
			if (f[4] > 0) off += 30;
		} else if (s == 4) {
			int kk = 21 + 11 * f[0] + 4 * f[1] + 3 * f[2] + 2 * f[3] + f[4];
			if (kk >> 4) kk -= (kk & 32 ? 20 : 10);
			off = 1845 + 3 * kk;
		} else {
			off = 1839 + 2 * f[0];
		}
	}
	}
	return (unsigned)(off - 127);
}

// ---------------------------------------------------------------------------
// bitmap helpers.  Rows are addressed by LOCAL row index lr = (z - zlo)*NY + y.
// Bits beyond x = nx are zero; word W (one past the last) exists and is zero.
// ---------------------------------------------------------------------------
MC_HD uint32_t ldw(const uint32_t *B, const Params &P, uint32_t lr, uint32_t w)
{
	return B[(uint64_t)lr * P.WP + w];
}
MC_HD uint32_t shr1(uint32_t lo, uint32_t hi) { return (lo >> 1) | (hi << 31); }

// mask of the bits of word w that denote x <= lim
MC_HD uint32_t mask_le(uint32_t w, uint32_t lim)
{
	uint32_t x0 = w << 5;
	if (x0 > lim) return 0u;
	uint32_t n = lim - x0;          // highest valid bit
	return n >= 31 ? 0xFFFFFFFFu : ((2u << n) - 1u);
}

struct Planes { uint32_t X, Y, Z; };

// which points of word w in point row (z,y) own an X/POINT, Y, Z vertex
MC_HDN Planes planes(const Params &P, uint32_t z, uint32_t y, uint32_t w)
{
	Planes r;
	const uint32_t lr = (z - P.zlo) * P.NY + y;
	const bool hasY = y < P.ny, hasZ = z < P.nz;
	const uint32_t vp = mask_le(w, P.nx);
	const uint32_t vx = P.nx ? mask_le(w, P.nx - 1) : 0u;
	const uint32_t s0 = ldw(P.S, P, lr, w), s0n = ldw(P.S, P, lr, w + 1);
	const uint32_t sx = shr1(s0, s0n);
	const uint32_t sy = hasY ? ldw(P.S, P, lr + 1, w) : s0;
	const uint32_t sz = hasZ ? ldw(P.S, P, lr + P.NY, w) : s0;
	r.X = (s0 ^ sx) & vx;
	r.Y = (s0 ^ sy) & vp;
	r.Z = (s0 ^ sz) & vp;
	unsigned zf = P.rowZ[lr];
	if (hasY) zf |= P.rowZ[lr + 1];
	if (hasZ) zf |= P.rowZ[lr + P.NY];
	if (zf) {
		const uint32_t z0 = ldw(P.Z, P, lr, w), z0n = ldw(P.Z, P, lr, w + 1);
		const uint32_t zx = shr1(z0, z0n);
		const uint32_t zy = hasY ? ldw(P.Z, P, lr + 1, w) : 0u;
		const uint32_t zz = hasZ ? ldw(P.Z, P, lr + P.NY, w) : 0u;
		r.X &= ~(z0 | zx);
		r.Y &= ~(z0 | zy);
		r.Z &= ~(z0 | zz);
		if (z0) {
			// POINT vertex: on-iso sample with at least one of its <=6 axis
			// neighbours above the isovalue
			uint32_t nb = sx | (s0 << 1) | (w ? ldw(P.S, P, lr, w - 1) >> 31 : 0u);
			if (hasY) nb |= sy;
			if (y > 0) nb |= ldw(P.S, P, lr - 1, w);
			if (hasZ) nb |= sz;
			if (z > 0) nb |= ldw(P.S, P, lr - P.NY, w);
			r.X |= z0 & nb;
		}
	}
	return r;
}

// corner sign words of the 32 cells of word w in cell row (z,y): c[k] bit b =
// index bit of corner k of cell x = 32w+b; zc[k] likewise for "on-iso".
struct CellWords { uint32_t c[8]; uint32_t zc[8]; uint32_t active; uint32_t zany; };

MC_HDN void cell_words(const Params &P, uint32_t z, uint32_t y, uint32_t w, CellWords &cw)
{
	const uint32_t lr = (z - P.zlo) * P.NY + y;
	const uint32_t r00 = lr, r10 = lr + 1, r01 = lr + P.NY, r11 = lr + P.NY + 1;
	uint32_t a, b;
	a = ldw(P.S, P, r00, w); b = ldw(P.S, P, r00, w + 1); cw.c[0] = a; cw.c[4] = shr1(a, b);
	a = ldw(P.S, P, r10, w); b = ldw(P.S, P, r10, w + 1); cw.c[1] = a; cw.c[5] = shr1(a, b);
	a = ldw(P.S, P, r11, w); b = ldw(P.S, P, r11, w + 1); cw.c[2] = a; cw.c[6] = shr1(a, b);
	a = ldw(P.S, P, r01, w); b = ldw(P.S, P, r01, w + 1); cw.c[3] = a; cw.c[7] = shr1(a, b);
	uint32_t any = 0, all = 0xFFFFFFFFu;
#pragma unroll
	for (int k = 0; k < 8; k++) { any |= cw.c[k]; all &= cw.c[k]; }
	cw.active = any & ~all & (P.nx ? mask_le(w, P.nx - 1) : 0u);
	cw.zany = 0;
	if (P.rowZ[r00] | P.rowZ[r10] | P.rowZ[r01] | P.rowZ[r11]) {
		a = ldw(P.Z, P, r00, w); b = ldw(P.Z, P, r00, w + 1); cw.zc[0] = a; cw.zc[4] = shr1(a, b);
		a = ldw(P.Z, P, r10, w); b = ldw(P.Z, P, r10, w + 1); cw.zc[1] = a; cw.zc[5] = shr1(a, b);
		a = ldw(P.Z, P, r11, w); b = ldw(P.Z, P, r11, w + 1); cw.zc[2] = a; cw.zc[6] = shr1(a, b);
		a = ldw(P.Z, P, r01, w); b = ldw(P.Z, P, r01, w + 1); cw.zc[3] = a; cw.zc[7] = shr1(a, b);
#pragma unroll
		for (int k = 0; k < 8; k++) cw.zany |= cw.zc[k];
	} else {
#pragma unroll
		for (int k = 0; k < 8; k++) cw.zc[k] = 0;
	}
}

MC_HD unsigned cell_index(const CellWords &cw, int b)
{
	unsigned i = 0;
#pragma unroll
	for (int k = 0; k < 8; k++) i |= ((cw.c[k] >> b) & 1u) << (7 - k);
	return i;
}
MC_HD unsigned cell_zmask(const CellWords &cw, int b)
{
	unsigned i = 0;
#pragma unroll
	for (int k = 0; k < 8; k++) i |= ((cw.zc[k] >> b) & 1u) << k;
	return i;
}

template <typename Sample>
MC_HD void cell_values(const Params &P, typename Traits<Sample>::Real iso, uint32_t x, uint32_t y, uint32_t z,
                       typename Traits<Sample>::Real *v)
{
#pragma unroll
	for (int c = 0; c < 8; c++) v[c] = ld_val<Sample>(P, iso, x + MC_CX(c), y + MC_CY(c), z + MC_CZ(c));
}

// end points of edge e and the key used for the zero-area test
MC_HD unsigned edge_a(unsigned e) { return (0x321047540310ull >> (4 * e)) & 15; }
MC_HD unsigned edge_b(unsigned e) { return (0x765476653221ull >> (4 * e)) & 15; }

MC_HD unsigned vertex_key(unsigned e, unsigned zmask)
{
	if (e == 12) return 12;
	unsigned a = edge_a(e), b = edge_b(e);
	if ((zmask >> a) & 1) return 16 + a;
	if ((zmask >> b) & 1) return 16 + b;
	return e;
}

// pattern of one active cell: start, winding flag, triangle count after the
// zero-area drop (marching_cubes_33.c:1235), centre vertex flag
struct CellPattern { unsigned start, m, ntri, centre; };

template <typename Sample>
MC_HDN CellPattern cell_pattern(const Params &P, const Tables &tb, uint32_t x, uint32_t y, uint32_t z,
                                unsigned idx, unsigned zmask)
{
	typedef typename Traits<Sample>::Real Real;
	CellPattern cp;
	unsigned e = tb.simple256[idx];
	if (e != 0xFFFFu) {
		cp.start = e & 0xFFF;
		cp.ntri = e >> 12;
		cp.m = (tb.case256[idx] >> 11) & 1;
		cp.centre = 0;
	} else {
		Real v[8];
		cell_values<Sample>(P, (Real)P.iso, x, y, z, v);
		cp.start = select_pattern<Real>(tb, idx, v, &cp.m);
		unsigned pi = tb.pat[cp.start];
		cp.ntri = pi & 0x7F;
		cp.centre = pi >> 7;
	}
	if (zmask) {
		unsigned n = 0;
		for (unsigned w = cp.start;; w++) {
			unsigned tw = tb.tri[w];
			unsigned k0 = vertex_key((tw >> 8) & 15, zmask), k1 = vertex_key((tw >> 4) & 15, zmask),
			         k2 = vertex_key(tw & 15, zmask);
			n += (k0 != k1 && k0 != k2 && k1 != k2);
			if (!(tw >> 12)) break;
		}
		cp.ntri = n;
	}
	return cp;
}

// ---------------------------------------------------------------------------
// count step for one (row, word): vertices owned by the 32 points, triangles and
// centre vertices of the 32 cells.  Returns packed counts:
//   cv = nX | nY<<21 | nZ<<42        cc = nT | nC<<32
// ---------------------------------------------------------------------------
template <typename Sample>
MC_HDN void count_word(const Params &P, const Tables &tb, uint32_t z, uint32_t y, uint32_t w,
                       bool own_points, bool own_cells, uint64_t &cv, uint64_t &cc)
{
	cv = 0; cc = 0;
	if (own_points) {
		Planes pl = planes(P, z, y, w);
		cv = (uint64_t)popc32(pl.X) | ((uint64_t)popc32(pl.Y) << 21) | ((uint64_t)popc32(pl.Z) << 42);
	}
	if (own_cells && w < P.WC) {
		CellWords cw;
		cell_words(P, z, y, w, cw);
		uint32_t act = cw.active;
		uint32_t nt = 0, nc = 0;
		while (act) {
			int b = ffs32(act);
			act &= act - 1;
			unsigned idx = cell_index(cw, b);
			unsigned zm = cw.zany ? cell_zmask(cw, b) : 0u;
			unsigned e = tb.simple256[idx];
			if (e != 0xFFFFu && !zm) {
				nt += e >> 12;
			} else {
				CellPattern cp = cell_pattern<Sample>(P, tb, (w << 5) + b, y, z, idx, zm);
				nt += cp.ntri;
				nc += cp.centre;
			}
		}
		cc = (uint64_t)nt | ((uint64_t)nc << 32);
	}
}

// ---------------------------------------------------------------------------
// vertex store: MC33_spn0/A/B/C (marching_cubes_33.c:485-621,
// MC33_util_grd.c:87-112); r[0..2] index-space position, r[3..5] = -grad F
// ---------------------------------------------------------------------------
template <typename Real>
MC_HDN void store_vertex(const Params &P, Real *r, uint32_t id)
{
	const Geom &g = P.geom;
	Real p[3];
	if (g.store == STORE_SPN0) {
		p[0] = r[0]; p[1] = r[1]; p[2] = r[2];
	} else if (g.store == STORE_SPNC) {
		const double *A = g.A, *B = g.Ai;
		Real c0, c1, c2;
		double r0 = r[0], r1 = r[1], r2 = r[2];
		if (g.tsa) {
			c0 = (Real)radd(radd(rmul(A[0], r0), rmul(A[1], r1)), rmul(A[2], r2));
			c1 = (Real)radd(rmul(A[4], r1), rmul(A[5], r2));
			c2 = (Real)rmul(A[8], r2);
		} else {
			c0 = (Real)radd(radd(rmul(A[0], r0), rmul(A[1], r1)), rmul(A[2], r2));
			c1 = (Real)radd(radd(rmul(A[3], r0), rmul(A[4], r1)), rmul(A[5], r2));
			c2 = (Real)radd(radd(rmul(A[6], r0), rmul(A[7], r1)), rmul(A[8], r2));
		}
		p[0] = radd(c0, (Real)g.O[0]); p[1] = radd(c1, (Real)g.O[1]); p[2] = radd(c2, (Real)g.O[2]);
		double n0 = r[3], n1 = r[4], n2 = r[5];
		if (g.tsa) {
			c2 = (Real)radd(radd(rmul(B[2], n0), rmul(B[5], n1)), rmul(B[8], n2));
			c1 = (Real)radd(rmul(B[1], n0), rmul(B[4], n1));
			c0 = (Real)rmul(B[0], n0);
		} else {
			c0 = (Real)radd(radd(rmul(B[0], n0), rmul(B[3], n1)), rmul(B[6], n2));
			c1 = (Real)radd(radd(rmul(B[1], n0), rmul(B[4], n1)), rmul(B[7], n2));
			c2 = (Real)radd(radd(rmul(B[2], n0), rmul(B[5], n1)), rmul(B[8], n2));
		}
		r[3] = c0; r[4] = c1; r[5] = c2;
	} else {
		if (g.store == STORE_SPNB) {
			r[3] = rmul(r[3], (Real)g.ca);
			r[4] = rmul(r[4], (Real)g.cb);
		}
#pragma unroll
		for (int i = 0; i < 3; i++) p[i] = radd(rmul(r[i], (Real)g.D[i]), (Real)g.O[i]);
	}
	Real s = radd(radd(rmul(r[3], r[3]), rmul(r[4], r[4])), rmul(r[5], r[5]));
	// exact 1/sqrt: the reference's rsqrtss is only good to 3e-4
	float t = rdiv(1.0f, sqrtf((float)s));
	if (g.normal_neg) t = -t;
	Real *V = (Real *)P.V + 3 * (uint64_t)id;
	float *N = P.N + 3 * (uint64_t)id;
	V[0] = p[0]; V[1] = p[1]; V[2] = p[2];
	N[0] = rmul(t, (float)r[3]); N[1] = rmul(t, (float)r[4]); N[2] = rmul(t, (float)r[5]);
	P.color[id] = P.color_value;
}

// transverse component of the edge normal (e.g. c:993-998)
template <typename Sample>
MC_HDN typename Traits<Sample>::Real transverse(const Params &P, typename Traits<Sample>::Real iso,
                                               uint32_t x, uint32_t y, uint32_t z, int a, int c,
                                               typename Traits<Sample>::Real t)
{
	typedef typename Traits<Sample>::Real Real;
	const uint32_t n[3] = {P.nx, P.ny, P.nz};
	uint32_t p0[3] = {x, y, z}, p1[3] = {x, y, z};
	p1[a] += 1;
	const Real one_t = rsub((Real)1, t);
	const uint32_t q = p0[c];
	if (q == 0 || q == n[c]) {
		uint32_t q0[3] = {p0[0], p0[1], p0[2]}, q1[3] = {p1[0], p1[1], p1[2]};
		Real d0, d1;
		if (q == 0) {
			q0[c] += 1; q1[c] += 1;
			d0 = rsub(ld_val<Sample>(P, iso, q0[0], q0[1], q0[2]), ld_val<Sample>(P, iso, p0[0], p0[1], p0[2]));
			d1 = rsub(ld_val<Sample>(P, iso, q1[0], q1[1], q1[2]), ld_val<Sample>(P, iso, p1[0], p1[1], p1[2]));
		} else {
			q0[c] -= 1; q1[c] -= 1;
			d0 = rsub(ld_val<Sample>(P, iso, p0[0], p0[1], p0[2]), ld_val<Sample>(P, iso, q0[0], q0[1], q0[2]));
			d1 = rsub(ld_val<Sample>(P, iso, p1[0], p1[1], p1[2]), ld_val<Sample>(P, iso, q1[0], q1[1], q1[2]));
		}
		return radd(rmul(d0, one_t), rmul(d1, t));
	}
	uint32_t l0[3] = {p0[0], p0[1], p0[2]}, h0[3] = {p0[0], p0[1], p0[2]};
	uint32_t l1[3] = {p1[0], p1[1], p1[2]}, h1[3] = {p1[0], p1[1], p1[2]};
	l0[c] -= 1; h0[c] += 1; l1[c] -= 1; h1[c] += 1;
	Real e0 = rawdiff(ld_sample<Sample>(P, l0[0], l0[1], l0[2]), ld_sample<Sample>(P, h0[0], h0[1], h0[2]));
	Real e1 = rawdiff(ld_sample<Sample>(P, l1[0], l1[1], l1[2]), ld_sample<Sample>(P, h1[0], h1[1], h1[2]));
	return rmul((Real)0.5f, radd(rmul(e0, one_t), rmul(e1, t)));
}

// MC33_surfint gradient (marching_cubes_33.c:628-647)
template <typename Sample>
MC_HDN typename Traits<Sample>::Real point_grad(const Params &P, uint32_t x, uint32_t y, uint32_t z, int c)
{
	typedef typename Traits<Sample>::Real Real;
	const uint32_t n[3] = {P.nx, P.ny, P.nz};
	uint32_t p[3] = {x, y, z}, lo[3] = {x, y, z}, hi[3] = {x, y, z};
	if (p[c] == 0) {
		hi[c] += 1;
		return rawdiff(ld_sample<Sample>(P, x, y, z), ld_sample<Sample>(P, hi[0], hi[1], hi[2]));
	} else if (p[c] == n[c]) {
		lo[c] -= 1;
		return rawdiff(ld_sample<Sample>(P, lo[0], lo[1], lo[2]), ld_sample<Sample>(P, x, y, z));
	}
	lo[c] -= 1; hi[c] += 1;
	// 0.5f*(F - F): float product for float and integer grids, double for double
	return (Real)rmul((Real)0.5f, rawdiff(ld_sample<Sample>(P, lo[0], lo[1], lo[2]), ld_sample<Sample>(P, hi[0], hi[1], hi[2])));
}

template <typename Sample>
MC_HDN void emit_edge_vertex(const Params &P, uint32_t x, uint32_t y, uint32_t z, int a, uint32_t id)
{
	typedef typename Traits<Sample>::Real Real;
	const Real iso = (Real)P.iso;
	uint32_t q[3] = {x, y, z};
	const uint32_t p[3] = {x, y, z};
	q[a] += 1;
	const Real va = ld_val<Sample>(P, iso, x, y, z), vb = ld_val<Sample>(P, iso, q[0], q[1], q[2]);
	const Real t = rdiv(va, rsub(va, vb));
	Real r[6];
#pragma unroll
	for (int c = 0; c < 3; c++) {
		if (c == a) { r[c] = radd((Real)p[c], t); r[3 + c] = rsub(vb, va); }
		else { r[c] = (Real)p[c]; r[3 + c] = transverse<Sample>(P, iso, x, y, z, a, c, t); }
	}
	store_vertex<Real>(P, r, id);
}

template <typename Sample>
MC_HDN void emit_point_vertex(const Params &P, uint32_t x, uint32_t y, uint32_t z, uint32_t id)
{
	typedef typename Traits<Sample>::Real Real;
	Real r[6] = {(Real)x, (Real)y, (Real)z, 0, 0, 0};
#pragma unroll
	for (int c = 0; c < 3; c++) r[3 + c] = point_grad<Sample>(P, x, y, z, c);
	store_vertex<Real>(P, r, id);
}

template <typename Sample>
MC_HDN void emit_centre_vertex(const Params &P, uint32_t x, uint32_t y, uint32_t z, uint32_t id)
{
	typedef typename Traits<Sample>::Real Real;
	Real v[8], r[6];
	cell_values<Sample>(P, (Real)P.iso, x, y, z, v);
	r[0] = radd((Real)x, (Real)0.5f); r[1] = radd((Real)y, (Real)0.5f); r[2] = radd((Real)z, (Real)0.5f);
	// marching_cubes_33.c:1227-1229, left to right
	r[3] = rsub(rsub(rsub(rsub(radd(radd(radd(v[4], v[5]), v[6]), v[7]), v[0]), v[1]), v[2]), v[3]);
	r[4] = rsub(rsub(rsub(rsub(radd(radd(radd(v[1], v[2]), v[5]), v[6]), v[0]), v[3]), v[4]), v[7]);
	r[5] = rsub(rsub(rsub(rsub(radd(radd(radd(v[2], v[3]), v[6]), v[7]), v[0]), v[1]), v[4]), v[5]);
	store_vertex<Real>(P, r, id);
}

// is the row one whose shared vertices this slab numbers (own or halo)?
MC_HD bool row_points_owned(const Params &P, uint32_t z) { return z >= P.pz0 && z < P.pz1; }
MC_HD bool row_points_halo(const Params &P, uint32_t z) { return z == P.hz; }
MC_HD bool row_cells_owned(const Params &P, uint32_t z, uint32_t y) { return z >= P.cz0 && z < P.cz1 && y < P.ny; }

// slab-local index of the first vertex of plane `pl` (0 X,1 Y,2 Z) in word w of
// point row (z,y).  Vertex ARRAYS are indexed locally; triangle CONTENTS are
// global ids (local + vbase).
MC_HD uint32_t plane_base_local(const Params &P, uint32_t z, uint32_t y, uint32_t w, int pl)
{
	const uint32_t lr = (z - P.zlo) * P.NY + y;
	const uint64_t pre = P.wpreV[(uint64_t)lr * P.W + w];
	const uint32_t rb = pl == 0 ? P.rowBX[lr] : (pl == 1 ? P.rowBY[lr] : P.rowBZ[lr]);
	return rb + (uint32_t)((pre >> (21 * pl)) & 0x1FFFFF);
}
MC_HD uint32_t plane_base_global(const Params &P, uint32_t z, uint32_t y, uint32_t w, int pl)
{
	const uint32_t local = plane_base_local(P, z, y, w, pl);
	// rows of the halo slice are numbered in the next slab's index space
	if (z == P.hz) return (P.dbases ? P.dbases[1] : P.vbase_next) + (local - P.totals->nShared);
	return (P.dbases ? P.dbases[0] : P.vbase) + local;
}

// ---------------------------------------------------------------------------
// vertex emit for one (row, word)
// ---------------------------------------------------------------------------
template <typename Sample>
MC_HDN void emit_vertices_word(const Params &P, uint32_t z, uint32_t y, uint32_t w)
{
	const Planes pl = planes(P, z, y, w);
	if (!(pl.X | pl.Y | pl.Z)) return;
	const uint32_t lr = (z - P.zlo) * P.NY + y;
	const uint32_t zw = P.rowZ[lr] ? ldw(P.Z, P, lr, w) : 0u;
	const uint64_t prow = ((uint64_t)z * P.NY + y) * P.NX;
	uint32_t m, id;
	for (int a = 0; a < 3; a++) {
		m = a == 0 ? pl.X : (a == 1 ? pl.Y : pl.Z);
		if (!m) continue;
		id = plane_base_local(P, z, y, w, a);
		while (m) {
			int b = ffs32(m);
			m &= m - 1;
			uint32_t x = (w << 5) + b;
			if (id < P.capV) {
				if (a == 0 && ((zw >> b) & 1)) emit_point_vertex<Sample>(P, x, y, z, id);
				else emit_edge_vertex<Sample>(P, x, y, z, a, id);
				if (P.vkey) P.vkey[id] = (prow + x) * 4 + (unsigned)a;
			} else {
				P.totals->overflow = 1;
			}
			id++;
		}
	}
}

// ---------------------------------------------------------------------------
// triangle (+ centre vertex) emit for one (cell row, word).
// scr: per-thread scratch of 8 (plane mask, base id) pairs, element k of this thread at
// scr_mask[k*stride], scr_base[k*stride]  (shared memory in the kernel).
// plane combos: 0 X00  1 Y00  2 Z00  3 X10  4 Z10  5 X01  6 Y01  7 X11
// (row suffix = dy dz of the point row relative to the cell row)
// ---------------------------------------------------------------------------
MC_HD unsigned combo_of_edge(unsigned e)  { return (0x573026412641ull >> (4 * e)) & 15; }
MC_HD unsigned cx_of_edge(unsigned e)     { return (0x0F0u >> e) & 1; }
MC_HD unsigned combo_of_corner(unsigned c) { return (0x57305730u >> (4 * c)) & 15; }

template <typename Sample>
MC_HDN void emit_triangles_word(const Params &P, const Tables &tb, uint32_t z, uint32_t y, uint32_t w,
                                uint32_t *scr_mask, uint32_t *scr_base, uint32_t stride)
{
	CellWords cw;
	cell_words(P, z, y, w, cw);
	uint32_t act = cw.active;
	if (!act) return;
	const uint32_t lr = (z - P.zlo) * P.NY + y;
	{
		// plane masks of word w and the id of their first vertex
		const int dy[8] = {0, 0, 0, 1, 1, 0, 0, 1}, dz[8] = {0, 0, 0, 0, 0, 1, 1, 1}, pln[8] = {0, 1, 2, 0, 2, 0, 1, 0};
		Planes p0[4];          // per point row (dy + 2*dz) at word w
		for (int r = 0; r < 4; r++) p0[r] = planes(P, z + (r >> 1), y + (r & 1), w);
		for (int k = 0; k < 8; k++) {
			int r = dy[k] + 2 * dz[k];
			// ranks only count bits below the queried one (offset <= 32), so the
			// 32 bits of word w are all that is needed even for x = 32w+32
			scr_mask[k * stride] = pln[k] == 0 ? p0[r].X : (pln[k] == 1 ? p0[r].Y : p0[r].Z);
			scr_base[k * stride] = plane_base_global(P, z + dz[k], y + dy[k], w, pln[k]);
		}
	}
	const uint64_t pre = P.wpreC[(uint64_t)lr * P.W + w];
	uint32_t tid = P.rowBT[lr] + (uint32_t)(pre & 0xFFFFFFFFu);                         // slab-local
	uint32_t cid = P.totals->nShared + P.rowBC[lr] + (uint32_t)(pre >> 32);             // slab-local
	const uint64_t crow = ((uint64_t)z * P.ny + y) * P.nx;
	while (act) {
		int b = ffs32(act);
		act &= act - 1;
		const uint32_t x = (w << 5) + b;
		const unsigned idx = cell_index(cw, b);
		const unsigned zm = cw.zany ? cell_zmask(cw, b) : 0u;
		const CellPattern cp = cell_pattern<Sample>(P, tb, x, y, z, idx, zm);
		uint32_t centre_id = 0;
		if (cp.centre) {
			const uint32_t cl = cid++;
			if (cl < P.capV) {
				emit_centre_vertex<Sample>(P, x, y, z, cl);
				if (P.vkey) P.vkey[cl] = (crow + x) * 4 + 3;
			} else {
				P.totals->overflow = 1;
			}
			centre_id = (P.dbases ? P.dbases[0] : P.vbase) + cl;
		}
		for (unsigned tw_i = cp.start;; tw_i++) {
			const unsigned tw = tb.tri[tw_i];
			uint32_t ti[3];
			unsigned key[3];
#pragma unroll
			for (int j = 0; j < 3; j++) {
				const unsigned e = (tw >> (8 - 4 * j)) & 15;
				if (e == 12) { ti[j] = centre_id; key[j] = 12; continue; }
				const unsigned a = edge_a(e), bb = edge_b(e);
				unsigned combo, off;
				if (zm && ((zm >> a) & 1)) { key[j] = 16 + a; combo = combo_of_corner(a); off = b + MC_CX(a); }
				else if (zm && ((zm >> bb) & 1)) { key[j] = 16 + bb; combo = combo_of_corner(bb); off = b + MC_CX(bb); }
				else { key[j] = e; combo = combo_of_edge(e); off = b + cx_of_edge(e); }
				const uint32_t mk = scr_mask[combo * stride];
				ti[j] = scr_base[combo * stride] + (uint32_t)popc32(off >= 32 ? mk : (mk & ((1u << off) - 1u)));
			}
			if (key[0] != key[1] && key[0] != key[2] && key[1] != key[2]) {
				if (tid < P.capT) {
					uint32_t a0 = cp.m ? ti[0] : ti[1], a1 = cp.m ? ti[1] : ti[0];
					if (P.geom.normal_neg) { uint32_t s = a0; a0 = a1; a1 = s; }
					uint32_t *T = P.T + 3 * (uint64_t)tid;
					T[0] = a0; T[1] = a1; T[2] = ti[2];
					if (P.tcell) P.tcell[tid] = crow + x;
				} else {
					P.totals->overflow = 1;
				}
				tid++;
			}
			if (!(tw >> 12)) break;
		}
	}
}

}  // namespace mc33
