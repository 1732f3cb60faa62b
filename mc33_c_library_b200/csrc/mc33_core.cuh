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
// cold paths (ambiguous cells, on-iso samples, inclined grids) are kept out of line:
// inlined at every use they made the emit kernel 730 KB of SASS and the warps
// spent 70 % of their time waiting for instruction fetches
#define MC_COLD __host__ __device__ __noinline__
#else
#define MC_HD inline
#define MC_HDN
#define MC_COLD __attribute__((noinline))
#endif

// The mesh (V, N, color, T) is written once and never read again on the device: streaming stores (st.global.cs, evict
// first), and the vertex tasks are read with a streaming load, so that gigabytes of output do not push the sample lines
// and bitmap words that the next slice's work needs out of the L2.  Measured (round 2): vertex kernel on a 1026-slice
// slab of cfg4 6.23 -> 3.99 ms (DRAM reads were 32 GB for a 17 GB slab), 258 slices 1.05 -> 0.95, cfg2 0.087 -> 0.086;
// cell kernel cfg2 0.167 -> 0.163.  (Vertex tasks written with streaming stores: no difference, left as plain stores.)
#ifndef MC33_STREAM_STORES
#define MC33_STREAM_STORES 1
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
	float Of[3], Df[3], caf, cbf;   // the same O, D, ca, cb as float (Real == float builds: no conversion per vertex)
};
MC_HD float  geomO(const Geom &g, int i, float)  { return g.Of[i]; }
MC_HD double geomO(const Geom &g, int i, double) { return g.O[i]; }
MC_HD float  geomD(const Geom &g, int i, float)  { return g.Df[i]; }
MC_HD double geomD(const Geom &g, int i, double) { return g.D[i]; }
MC_HD float  geomCa(const Geom &g, float)  { return g.caf; }
MC_HD double geomCa(const Geom &g, double) { return g.ca; }
MC_HD float  geomCb(const Geom &g, float)  { return g.cbf; }
MC_HD double geomCb(const Geom &g, double) { return g.cb; }

struct Totals {                   // written by the scan kernel
	uint32_t nShared;             // shared vertices owned by this slab
	uint32_t nCentre;
	uint32_t nT;
	uint32_t nSharedHalo;         // shared vertices of the halo slice (next slab's first)
	uint32_t overflow;            // set by emit kernels if a capacity was exceeded
	uint32_t nSharedAll;          // shared vertices counted (own + halo slice)
	uint32_t range;               // 1: more than 2^32-1 vertices or triangles
	uint32_t anyZ;                // classify: some sample is exactly on the isovalue
	uint32_t ticket;              // next row group of the cell kernel (dynamic distribution; the vertex kernel re-arms it)
	uint32_t ticket2;             // ... of the vertex kernel (the cell kernel re-arms it)
	uint32_t nslow_acc, nslow;    // quads with an on-iso sample in reach: accumulated by the count kernel, published by its last block

};

struct Tables {
	const uint16_t *case256, *simple256, *tri;
	const uint8_t *pat;           // ntri | centre << 7, at pattern starts
	const uint32_t *cinfo;        // per case index: simple256 entry (start | ntri << 12, or 0xFFFF) | winding flag m << 16
	const uint16_t *pord;         // [pattern start] ordinal of the pattern (0 .. MC33_NPATTERNS-1)
	const uint16_t *keep;         // [ordinal][on-iso corner mask]: which triangles of the pattern survive the zero-area drop
};
#define MC33_NPATTERNS 383

struct Params {
	const void *data;             // samples of slices [zlo,zhi), x fastest
	uint32_t nx, ny, nz;          // GLOBAL interval counts (_GRD.N)
	uint32_t NX, NY;              // nx+1, ny+1
	uint32_t zlo, zhi;            // global sample slices held in `data`
	uint32_t pz0, pz1;            // point slices whose shared vertices this slab owns
	uint32_t cz0, cz1;            // cell layers this slab owns
	uint32_t hz;                  // halo point slice numbered in the next slab (== pz1) or 0xFFFFFFFF
	uint32_t W, WC;               // words per point row, per cell row
	uint32_t Q, WP;               // quads (4 words, one 16-byte load) per row; row stride in words = 4Q+4
	uint32_t G;                   // rows per warp: a warp takes G whole rows, one lane per quad (Q <= 32),
	                              // or one row in passes of 32 quads (G == 1)
	uint32_t Lrows;               // (zhi-zlo)*NY
	uint32_t mQ, mNY;             // floor(2^32/Q), floor(2^32/NY) for fastdiv
	uint32_t *S, *Z;              // bitmaps [Lrows][WP]; words >= W of a row are zero
	uint32_t *A;                  // [Lrows][WP] visit bitmap left by the count kernel for the cell kernel: active
	                              // cells of the row, and grid points that own a vertex this slab emits
	uint32_t *D;                  // dirty bits of Z (k_classify_sweep: one per 128-sample group, set while the group holds a
	                              // non-zero Z word, so that clean groups need not be rewritten with zeros)
	uint32_t *anyZp;              // one word: set by the classify kernel when some sample is exactly on the isovalue
	                              // (totals->anyZ, or the slot of a pre-classified sweep set)
	uint32_t *rowZ;               // [Lrows] hint: == zepoch when the row has an on-iso sample in this extraction
	uint32_t zepoch;
	uint64_t *wpreV;              // [Lrows][WP] X | Y<<21 | Z<<42: row-local index of each plane's first vertex
	                              // in word k (entries W..4Q: one past the plane's last vertex)
	uint32_t *rowBV, *rowBT, *rowBC;   // [Lrows+1] slab-local exclusive bases: vertices, triangles, centres
	Totals *totals;
	double iso;
	uint32_t dbg_noz;             // measurement hook (MC33_B200_DEBUG_NOZ=1, tools/ only): classify ignores on-iso samples
	uint32_t ithr, ieq_lo, ieq_hi, inone;   // integer grids: sample >= ithr <=> above iso; [ieq_lo,ieq_hi] on iso
	Geom geom;
	// outputs (device)
	void *V; float *N; int32_t *color; uint32_t *T;
	uint64_t *vkey, *tcell;       // optional canonical keys (tests)
	uint64_t *vtask;              // [capV] vertex tasks left by the cell kernel for the vertex kernel
	uint16_t *pcache;             // [Lrows][32 WP] pattern start of every COMPLEX cell (face / interior tests decided it),
	                              // left by the count kernel so that the cell kernel does not run the tests again
	uint32_t capV, capT;
	uint32_t vbase, vbase_next;   // global vertex id of this / the next slab's first vertex (0 on one GPU)
	const uint32_t *dbases;       // optional device copy {vbase, vbase_next}: overrides the two above
	int32_t color_value;
};

// corner c of a cell -> offsets; edge e -> end points, axis
#define MC_CX(c) (((c) >> 2) & 1)
#define MC_CY(c) ((0x66 >> (c)) & 1)
#define MC_CZ(c) ((0xCC >> (c)) & 1)

// samples are never written by the kernels: read-only (non-coherent) loads, which
// the compiler may hoist above the mesh stores
template <typename Sample>
MC_HD Sample ldro(const Sample *p)
{
#if defined(__CUDA_ARCH__)
	return __ldg(p);
#else
	return *p;
#endif
}
template <typename Sample>
MC_HD Sample ld_sample(const Params &P, uint32_t x, uint32_t y, uint32_t z)
{
	const Sample *F = (const Sample *)P.data;
	return ldro(F + ((uint64_t)(z - P.zlo) * P.NY + y) * P.NX + x);
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

// -> pattern start in MC33_TRI; m = the reference's winding flag.
//
// Same decisions as MC33_findCase (marching_cubes_33.c:693-779), arranged for a warp whose
// lanes hold cells of DIFFERENT base cases: the six face comparisons are made once,
// branch free; each case then only names the interior tests it needs (at most two), the
// tests themselves run at two common call sites where all the lanes that need one meet
// again, and a last per-case step turns the answers into the table offset.  (With a call
// inside every case branch the lanes ran the tests one after the other: 2 active threads
// per instruction on white noise.)  The tests are pure functions of the corner values, so
// running the second one where the reference short-circuits changes nothing.
template <typename Real>
MC_COLD unsigned select_pattern(const Tables &tb, unsigned i, const Real *v, unsigned *mflag)
{
	const unsigned c = tb.case256[i];
	const int k = (int)(c & 0x7FF);
	const unsigned m = (c >> 11) & 1;
	const unsigned idx = m ? i : (i ^ 0xFF);
	const int cs = (int)(c >> 12);
	*mflag = m;
	if (cs == 0) return (unsigned)(k - 127);
	// face comparisons and MC33_faceTests results as bit masks (bit j: face j)
	unsigned lt = 0, fneg = 0, fpos = 0;
	const unsigned ind = cs == 7 ? 165u : idx;
	int s = 0;
#pragma unroll
	for (int j = 0; j < 6; j++) {
		const unsigned l = face_lt(v, j) ? 1u : 0u;
		lt |= l << j;
		const unsigned mm = ind & face_mask(j);
		int r = 0;
		if (ind & face_gate(j)) {
			if (mm == face_set(j)) r = l ? -1 : 1;
		} else {
			if (mm == face_clr(j)) r = l ? 1 : -1;
		}
		fneg |= (r < 0 ? 1u : 0u) << j;
		fpos |= (r > 0 ? 1u : 0u) << j;
		s += r;
	}
#define MC_F(j) ((int)((fpos >> (j)) & 1u) - (int)((fneg >> (j)) & 1u))
#define MC_FT1(f) (((lt >> (f)) & 1u) ? face_clr(f) : face_set(f))
	int ia = -1, ib = -1, f13 = 0, off = 0, kk = 0;
	switch (cs) {
	case 1:
		off = (idx & MC_FT1(k >> 2)) ? 183 + 2 * k : 159 + k;
		break;
	case 2:
		ia = k;
		break;
	case 3:
		if (idx & MC_FT1(k % 6)) off = 575 + 5 * k; else ia = k / 6;
		break;
	case 4:
		if (s == -3) off = 695 + 3 * k;
		else if (s == -1) off = (MC_F(4) + MC_F(5) < 0 ? (MC_F(0) + MC_F(2) < 0 ? 759 : 799) : 719) + 5 * k;
		else if (s == 1) off = (MC_F(4) + MC_F(5) < 0 ? 983 : (MC_F(0) + MC_F(2) < 0 ? 839 : 911)) + 9 * k;
		else ia = k >> 1;
		break;
	case 5:
		if (s == -2) { ia = 0; ib = k == 2 ? -1 : (k ? 1 : 3); }
		else if (s == 0) off = (MC_F(2 + k) < 0 ? 1261 : 1285) + 8 * k;
		else if (k == 2) ia = 1;
		else { ia = 2; ib = k ? 3 : 1; }
		break;
	case 6:
		if (s == -2) ia = (int)((0xDA010Cu >> (2 * k)) & 3);
		else if (s == 0) off = (MC_F(k >> 1) < 0 ? 1645 : 1741) + 8 * k;
		else ia = (int)((0xA7B7E5u >> (2 * k)) & 3);
		break;
	default: {
		const int sa = s < 0 ? -s : s;
		if (sa == 0) {
			kk = ((MC_F(1) < 0) << 1) | (MC_F(5) < 0);
			if (MC_F(0) * MC_F(1) == MC_F(5)) off = 2157 + 12 * kk;
			else { ia = kk; f13 = 1; }
		} else if (sa == 2) {
			off = 1917 + 10 * ((MC_F(0) < 0 ? (int)(MC_F(2) > 0) : 12 + (int)(MC_F(2) < 0)) +
			                   (MC_F(1) < 0 ? (int)(MC_F(3) < 0) : 6 + (int)(MC_F(3) > 0)));
			if (MC_F(4) > 0) off += 30;
		} else if (sa == 4) {
			int q = 21 + 11 * MC_F(0) + 4 * MC_F(1) + 3 * MC_F(2) + 2 * MC_F(3) + MC_F(4);
			if (q >> 4) q -= (q & 32 ? 20 : 10);
			off = 1845 + 3 * q;
		} else {
			off = 1839 + 2 * MC_F(0);
		}
	}
	}
#undef MC_F
#undef MC_FT1
	// the interior tests, at common call sites
	int ra = 0, rb = 0;
	if (ia >= 0) ra = interior_test(ia, f13, v);
	if (ib >= 0) rb = interior_test(ib, 0, v);
	if (ia >= 0) {
		switch (cs) {
		case 2: off = ra ? 239 + 6 * k : 231 + 2 * k; break;
		case 3: off = ra ? 407 + 7 * k : 335 + 3 * k; break;
		case 4: off = ra ? 1095 + 9 * k : 1055 + 5 * k; break;
		case 5:
			if (s == -2) off = (ra || rb) ? 1213 + 8 * k : 1189 + 4 * k;
			else off = (ra || rb) ? 1237 + 8 * k : 1201 + 4 * k;
			break;
		case 6:
			if (s == -2) off = ra ? 1453 + 8 * k : 1357 + 4 * k;
			else off = ra ? 1549 + 8 * k : 1405 + 4 * k;
			break;
		default:
			off = 2285 + (ra ? 10 * kk - 40 * ra : 6 * kk);
		}
	}
	return (unsigned)(off - 127);
}

// ---------------------------------------------------------------------------
// bitmap helpers.  Rows are addressed by LOCAL row index lr = (z - zlo)*NY + y.
// Bits beyond x = nx are zero; word W (one past the last) exists and is zero.
// Word indices fit 32 bits (checked at context creation).
// ---------------------------------------------------------------------------
MC_HD uint32_t shr1(uint32_t lo, uint32_t hi) { return (lo >> 1) | (hi << 31); }

// mask of the bits of word w that denote x <= lim
MC_HD uint32_t mask_le(uint32_t w, uint32_t lim)
{
	uint32_t x0 = w << 5;
	if (x0 > lim) return 0u;
	uint32_t n = lim - x0;          // highest valid bit
	return n >= 31 ? 0xFFFFFFFFu : ((2u << n) - 1u);
}

// exact n / d for 32-bit n given m = floor(2^32 / d) -- one mulhi + fix-up
MC_HD uint32_t fastdiv(uint32_t n, uint32_t d, uint32_t m)
{
	if (d == 1) return n;
#if defined(__CUDA_ARCH__)
	uint32_t q = __umulhi(n, m);
#else
	uint32_t q = (uint32_t)(((uint64_t)n * m) >> 32);
#endif
	uint32_t r = n - q * d;
	if (r >= d) { q++; r -= d; }
	if (r >= d) { q++; }
	return q;
}

// What the count kernel leaves per (point row, word): which of the 32 points own
// an X/POINT, Y, Z vertex, and which of the 32 cells (same x range, cell row of
// the same (z,y)) are active.
struct WordRec { uint32_t X, Y, Z, act; };

// corner sign words of the 32 cells of word w in cell row (z,y): c[k] bit b =
// index bit of corner k of cell x = 32w+b; zc[k] likewise for "on-iso".
struct CellWords { uint32_t c[8]; uint32_t zc[8]; uint32_t zany; };

// One pass over the bitmaps for word w of row (z,y).  gz: the grid has at least
// one on-iso sample (uniform flag from the classify kernel); when it is false
// the Z bitmap is not touched at all.
// the on-iso part of word_masks(): rare, kept out of line
MC_COLD void word_masks_z(const Params &P, uint32_t z, uint32_t y, uint32_t w, WordRec &rec, CellWords &cw)
{
	const uint32_t lr = (z - P.zlo) * P.NY + y;
	const bool hasY = y < P.ny, hasZ = z < P.nz;
	const uint32_t dY = hasY ? P.WP : 0u, dZ = hasZ ? P.NY * P.WP : 0u;
	const uint32_t i00 = lr * P.WP + w, i10 = i00 + dY, i01 = i00 + dZ, i11 = i01 + dY;
	const bool f00 = P.rowZ[lr] == P.zepoch, f10 = hasY && P.rowZ[lr + 1] == P.zepoch;
	const bool f01 = hasZ && P.rowZ[lr + P.NY] == P.zepoch, f11 = hasY && hasZ && P.rowZ[lr + P.NY + 1] == P.zepoch;
	if (!(f00 | f10 | f01 | f11)) return;
	const uint32_t s00 = cw.c[0], x00 = cw.c[4], s10 = cw.c[1], s01 = cw.c[3];
	const uint32_t z00 = P.Z[i00], z10 = hasY ? P.Z[i10] : 0u, z01 = hasZ ? P.Z[i01] : 0u;
	const uint32_t z11 = (hasY && hasZ) ? P.Z[i11] : 0u;
	const uint32_t zx00 = shr1(z00, P.Z[i00 + 1]);
	rec.X &= ~(z00 | zx00);
	rec.Y &= ~(z00 | z10);
	rec.Z &= ~(z00 | z01);
	if (z00) {
		// POINT vertex: on-iso sample with at least one of its <= 6 axis neighbours
		// above the isovalue; it lives in the X plane
		uint32_t nb = x00 | (s00 << 1) | (w ? P.S[i00 - 1] >> 31 : 0u);
		if (hasY) nb |= s10;
		if (y > 0) nb |= P.S[i00 - P.WP];
		if (hasZ) nb |= s01;
		if (z > 0) nb |= P.S[i00 - P.NY * P.WP];
		rec.X |= z00 & nb;
	}
	if (rec.act) {
		cw.zc[0] = z00; cw.zc[4] = zx00;
		cw.zc[1] = z10; cw.zc[5] = shr1(z10, P.Z[i10 + 1]);
		cw.zc[2] = z11; cw.zc[6] = shr1(z11, P.Z[i11 + 1]);
		cw.zc[3] = z01; cw.zc[7] = shr1(z01, P.Z[i01 + 1]);
#pragma unroll
		for (int k = 0; k < 8; k++) cw.zany |= cw.zc[k];
	}
}

MC_HDN void word_masks(const Params &P, uint32_t z, uint32_t y, uint32_t w, bool gz, WordRec &rec, CellWords &cw)
{
	const uint32_t lr = (z - P.zlo) * P.NY + y;
	const bool hasY = y < P.ny, hasZ = z < P.nz;
	const uint32_t dY = hasY ? P.WP : 0u, dZ = hasZ ? P.NY * P.WP : 0u;
	const uint32_t i00 = lr * P.WP + w, i10 = i00 + dY, i01 = i00 + dZ, i11 = i01 + dY;
	const uint32_t vp = mask_le(w, P.nx), vx = mask_le(w, P.nx - 1);
	const uint32_t s00 = P.S[i00], s10 = P.S[i10], s01 = P.S[i01], s11 = P.S[i11];
	const uint32_t x00 = shr1(s00, P.S[i00 + 1]), x10 = shr1(s10, P.S[i10 + 1]);
	const uint32_t x01 = shr1(s01, P.S[i01 + 1]), x11 = shr1(s11, P.S[i11 + 1]);
	rec.X = (s00 ^ x00) & vx;
	rec.Y = (s00 ^ s10) & vp;       // zero when !hasY (s10 is s00 then)
	rec.Z = (s00 ^ s01) & vp;
	cw.c[0] = s00; cw.c[4] = x00; cw.c[1] = s10; cw.c[5] = x10;
	cw.c[2] = s11; cw.c[6] = x11; cw.c[3] = s01; cw.c[7] = x01;
	const uint32_t any = s00 | s10 | s01 | s11 | x00 | x10 | x01 | x11;
	const uint32_t all = s00 & s10 & s01 & s11 & x00 & x10 & x01 & x11;
	rec.act = (hasY && hasZ) ? (any & ~all & vx) : 0u;
	cw.zany = 0;
#pragma unroll
	for (int k = 0; k < 8; k++) cw.zc[k] = 0;
	if (gz) word_masks_z(P, z, y, w, rec, cw);
}

// out-of-line form for the (rare) groups with on-iso samples
MC_COLD void word_masks_generic(const Params &P, uint32_t z, uint32_t y, uint32_t w, WordRec &rec, CellWords &cw)
{
	word_masks(P, z, y, w, true, rec, cw);
}

// ---------------------------------------------------------------------------
// quads: four consecutive bitmap words of a row (one 16-byte load) plus the
// word after them (for the x+1 neighbour of bit 31 of the fourth word)
// ---------------------------------------------------------------------------
struct Quad { uint32_t s[5]; };

MC_HD Quad load_quad(const uint32_t *B, uint64_t i)   // i = word index, multiple of 4
{
	Quad q;
#if defined(__CUDA_ARCH__)
	const uint4 v = *reinterpret_cast<const uint4 *>(B + i);
	q.s[0] = v.x; q.s[1] = v.y; q.s[2] = v.z; q.s[3] = v.w;
#else
	q.s[0] = B[i]; q.s[1] = B[i + 1]; q.s[2] = B[i + 2]; q.s[3] = B[i + 3];
#endif
	q.s[4] = B[i + 4];
	return q;
}

// Is there an on-iso sample among the points whose signs decide word k of a cell row:
// the four corner rows at this word, or their first point of the next word?  When there
// is none, the no-on-iso formulas (quad_word, cell_fast) yield exactly what the generic
// walk yields for this word -- every mask they build only involves these points -- so the
// slow path is taken per WORD, not per row group.
MC_HD uint32_t quad_oniso(const Quad &z00, const Quad &z10, const Quad &z01, const Quad &z11, int k)
{
	return (z00.s[k] | z10.s[k] | z01.s[k] | z11.s[k]) | ((z00.s[k + 1] | z10.s[k + 1] | z01.s[k + 1] | z11.s[k + 1]) & 1u);
}

// ... for the four words of a quad at once (bit k = word k); out of line: only row groups
// near an on-iso sample come here, and the count kernel has no registers to spare
MC_COLD uint32_t quad_oniso_mask(const uint32_t *Zb, uint64_t i00, uint64_t dY, uint64_t dZ)
{
	const Quad z00 = load_quad(Zb, i00), z10 = load_quad(Zb, i00 + dY);
	const Quad z01 = load_quad(Zb, i00 + dZ), z11 = load_quad(Zb, i00 + dY + dZ);
	uint32_t m = 0;
#pragma unroll
	for (int k = 0; k < 4; k++) m |= quad_oniso(z00, z10, z01, z11, k) ? (1u << k) : 0u;
	return m;
}

// the same predicate for word w of cell / point row (z,y), straight from the Z bitmap
MC_HD bool word_oniso(const Params &P, uint32_t z, uint32_t y, uint32_t w)
{
	const uint32_t uy = y < P.ny ? 1u : 0u, uz = z < P.nz ? P.NY : 0u;
	const uint32_t l00 = (z - P.zlo) * P.NY + y;
	const uint32_t i00 = l00 * P.WP + w, i10 = (l00 + uy) * P.WP + w, i01 = (l00 + uz) * P.WP + w, i11 = (l00 + uz + uy) * P.WP + w;
	return ((P.Z[i00] | P.Z[i10] | P.Z[i01] | P.Z[i11]) | ((P.Z[i00 + 1] | P.Z[i10 + 1] | P.Z[i01 + 1] | P.Z[i11 + 1]) & 1u)) != 0u;
}

// word k (0..3, compile time) of quad q of row (z,y) when the grid has NO on-iso
// sample: the same record and corner words word_masks() yields.  q10 / q01 / q11
// are the quads of rows y+1, z+1, (y+1,z+1); where such a row does not exist the
// caller passes the row itself, which zeroes the corresponding plane.
MC_HD void quad_word(const Params &P, const Quad &q00, const Quad &q10, const Quad &q01, const Quad &q11, int k,
                     uint32_t w, bool cells, WordRec &rec, uint32_t *c)
{
	const uint32_t vp = mask_le(w, P.nx), vx = mask_le(w, P.nx - 1);
	const uint32_t s00 = q00.s[k], s10 = q10.s[k], s01 = q01.s[k], s11 = q11.s[k];
	const uint32_t x00 = shr1(s00, q00.s[k + 1]), x10 = shr1(s10, q10.s[k + 1]);
	const uint32_t x01 = shr1(s01, q01.s[k + 1]), x11 = shr1(s11, q11.s[k + 1]);
	rec.X = (s00 ^ x00) & vx;
	rec.Y = (s00 ^ s10) & vp;
	rec.Z = (s00 ^ s01) & vp;
	c[0] = s00; c[4] = x00; c[1] = s10; c[5] = x10;
	c[2] = s11; c[6] = x11; c[3] = s01; c[7] = x01;
	const uint32_t any = s00 | s10 | s01 | s11 | x00 | x10 | x01 | x11;
	const uint32_t all = s00 & s10 & s01 & s11 & x00 & x10 & x01 & x11;
	rec.act = cells ? (any & ~all & vx) : 0u;
}

MC_HD unsigned cell_index(const uint32_t *c, uint32_t stride, int b)
{
	unsigned i = 0;
#pragma unroll
	for (int k = 0; k < 8; k++) i = (i << 1) | ((c[k * stride] >> b) & 1u);
	return i;
}
MC_HD unsigned cell_zmask(const uint32_t *zc, uint32_t stride, int b)
{
	unsigned i = 0;
#pragma unroll
	for (int k = 0; k < 8; k++) i |= ((zc[k * stride] >> b) & 1u) << k;
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

// ambiguous index (face / interior tests on the corner values) or on-iso corners
template <typename Sample>
MC_COLD CellPattern cell_pattern_slow(const Params &P, const Tables &tb, uint32_t x, uint32_t y, uint32_t z,
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

template <typename Sample>
MC_HD CellPattern cell_pattern(const Params &P, const Tables &tb, uint32_t x, uint32_t y, uint32_t z,
                               unsigned idx, unsigned zmask)
{
	const unsigned e = tb.simple256[idx];
	if (e != 0xFFFFu && !zmask) {
		CellPattern cp;
		cp.start = e & 0xFFF;
		cp.ntri = e >> 12;
		cp.m = (tb.case256[idx] >> 11) & 1;
		cp.centre = 0;
		return cp;
	}
	return cell_pattern_slow<Sample>(P, tb, x, y, z, idx, zmask);
}

// ---------------------------------------------------------------------------
// count step.  Packed counts of one (row, word):
//   cv = nX | nY<<21 | nZ<<42   vertices owned by the 32 points
//   cc = nT | nC<<32            triangles / centre vertices of the 32 cells
// ---------------------------------------------------------------------------
MC_HD uint64_t pack_planes(const WordRec &rec)
{
	return (uint64_t)popc32(rec.X) | ((uint64_t)popc32(rec.Y) << 21) | ((uint64_t)popc32(rec.Z) << 42);
}

// walk the active cells of word w of cell row (z,y): c[k] / zc[k] = corner sign /
// on-iso words (zc only read when zany)
MC_HD uint32_t count_simple_cells(const uint32_t *c, uint32_t act, uint32_t &cx);

template <typename Sample>
MC_HDN uint64_t count_cells(const Params &P, const Tables &tb, uint32_t z, uint32_t y, uint32_t w, uint32_t act,
                            const uint32_t *c, const uint32_t *zc, uint32_t zany)
{
	uint32_t nc = 0;
	// cells without an on-iso corner follow the index alone: the simple ones are counted 32 at a
	// time, and only complex cells and cells with an on-iso corner are walked (this is the word
	// that makes one lane of the count kernel late, so it should be short)
	uint32_t zcells = 0;
	if (zany) zcells = (zc[0] | zc[1] | zc[2] | zc[3] | zc[4] | zc[5] | zc[6] | zc[7]) & act;
	uint32_t cx = 0;
	uint32_t nt = count_simple_cells(c, act & ~zcells, cx);
	act = cx | zcells;
	while (act) {
		int b = ffs32(act);
		act &= act - 1;
		unsigned idx = cell_index(c, 1, b);
		unsigned zm = zany ? cell_zmask(zc, 1, b) : 0u;
		unsigned e = tb.simple256[idx];
		if (e != 0xFFFFu && !zm) {
			nt += e >> 12;
		} else {
			CellPattern cp = cell_pattern_slow<Sample>(P, tb, (w << 5) + b, y, z, idx, zm);
			nt += cp.ntri;
			nc += cp.centre;
		}
	}
	return (uint64_t)nt | ((uint64_t)nc << 32);
}

// the active cells of the four words of a quad in ONE loop (a separate loop per
// word would make a warp pay the longest word four times); no on-iso samples.
// Each iteration takes the lowest active cell of the lowest non-empty word.
// (only the COMPLEX cells come here -- the simple ones are counted 32 at a time -- so it is
// out of line and loads the four sign quads again instead of taking them by reference, which
// kept them in local memory in the caller's hot path)
template <typename Sample>
MC_COLD uint64_t count_cells_quad(const Params &P, const Tables &tb, uint32_t z, uint32_t y, uint32_t q,
                                  uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint64_t i00, uint64_t dY, uint64_t dZ)
{
	uint32_t nt = 0, nc = 0;
	const Quad q00 = load_quad(P.S, i00), q10 = load_quad(P.S, i00 + dY);
	const Quad q01 = load_quad(P.S, i00 + dZ), q11 = load_quad(P.S, i00 + dY + dZ);
	while (a0 | a1 | a2 | a3) {
		const int k = a0 ? 0 : (a1 ? 1 : (a2 ? 2 : 3));
		const uint32_t m = k == 0 ? a0 : (k == 1 ? a1 : (k == 2 ? a2 : a3));
		const int b = ffs32(m);
		const uint32_t cl = m & (m - 1);
		if (k == 0) a0 = cl; else if (k == 1) a1 = cl; else if (k == 2) a2 = cl; else a3 = cl;
		// the cell's 2 x 4 corner bits: bit b of word k and bit b+1 (bit 0 of word k+1 when b == 31)
		const uint64_t t00 = ((uint64_t)(k == 0 ? q00.s[1] : (k == 1 ? q00.s[2] : (k == 2 ? q00.s[3] : q00.s[4]))) << 32) |
		                     (k == 0 ? q00.s[0] : (k == 1 ? q00.s[1] : (k == 2 ? q00.s[2] : q00.s[3])));
		const uint64_t t10 = ((uint64_t)(k == 0 ? q10.s[1] : (k == 1 ? q10.s[2] : (k == 2 ? q10.s[3] : q10.s[4]))) << 32) |
		                     (k == 0 ? q10.s[0] : (k == 1 ? q10.s[1] : (k == 2 ? q10.s[2] : q10.s[3])));
		const uint64_t t01 = ((uint64_t)(k == 0 ? q01.s[1] : (k == 1 ? q01.s[2] : (k == 2 ? q01.s[3] : q01.s[4]))) << 32) |
		                     (k == 0 ? q01.s[0] : (k == 1 ? q01.s[1] : (k == 2 ? q01.s[2] : q01.s[3])));
		const uint64_t t11 = ((uint64_t)(k == 0 ? q11.s[1] : (k == 1 ? q11.s[2] : (k == 2 ? q11.s[3] : q11.s[4]))) << 32) |
		                     (k == 0 ? q11.s[0] : (k == 1 ? q11.s[1] : (k == 2 ? q11.s[2] : q11.s[3])));
		const uint32_t p00 = (uint32_t)(t00 >> b) & 3u, p10 = (uint32_t)(t10 >> b) & 3u;
		const uint32_t p01 = (uint32_t)(t01 >> b) & 3u, p11 = (uint32_t)(t11 >> b) & 3u;
		// corners 0..3 at x: rows 00 10 11 01 -> index bits 7..4; corners 4..7 at x+1 -> bits 3..0
		const unsigned idx = ((p00 & 1u) << 7) | ((p10 & 1u) << 6) | ((p11 & 1u) << 5) | ((p01 & 1u) << 4) |
		                     ((p00 >> 1) << 3) | ((p10 >> 1) << 2) | ((p11 >> 1) << 1) | (p01 >> 1);
		const unsigned e = tb.simple256[idx];
		if (e != 0xFFFFu) {
			nt += e >> 12;
		} else {
			CellPattern cp = cell_pattern_slow<Sample>(P, tb, (((q << 2) + (uint32_t)k) << 5) + (uint32_t)b, y, z, idx, 0u);
			nt += cp.ntri;
			nc += cp.centre;
		}
	}
	return (uint64_t)nt | ((uint64_t)nc << 32);
}

// ---------------------------------------------------------------------------
// Triangles of the SIMPLE cells of a word, 32 cells at a time, without looking at any
// cell.  The cases the reference triangulates without a test (base cases 1, 2, 5, 8, 9,
// 11, 14 of MC33_all_tables, marching_cubes_33.c:693-696) are exactly the cells whose
// surface is one disc, and a disc through k crossed edges has k - 2 triangles; every
// other active cell has an ambiguous face (corner signs alternating round the face) or
// is case 4 (two corners on a body diagonal against the other six) -- checked against
// the table for all 256 indices in tests/test_hostemu_vs_oracle.py.  So
//   triangles(simple cells) = sum over the 12 edges of popc(crossed & simple) - 2 popc(simple)
// and only the complex cells (returned in cx) are walked one by one.
// c[k] = sign word of corner k (bit b = cell b), act = active cells.
// ---------------------------------------------------------------------------
MC_HD uint32_t count_simple_cells(const uint32_t *c, uint32_t act, uint32_t &cx)
{
	const uint32_t e01 = c[0] ^ c[1], e12 = c[1] ^ c[2], e32 = c[3] ^ c[2], e03 = c[0] ^ c[3];
	const uint32_t e45 = c[4] ^ c[5], e56 = c[5] ^ c[6], e76 = c[7] ^ c[6], e47 = c[4] ^ c[7];
	const uint32_t e04 = c[0] ^ c[4], e15 = c[1] ^ c[5], e26 = c[2] ^ c[6], e37 = c[3] ^ c[7];
	// ambiguous faces {0,1,5,4} {1,2,6,5} {3,2,6,7} {0,3,7,4} {0,1,2,3} {4,5,6,7}
	uint32_t m = (e01 & e15 & e45) | (e12 & e26 & e56) | (e32 & e26 & e76) |
	             (e03 & e37 & e47) | (e01 & e12 & e32) | (e45 & e56 & e76);
	// case 4, diagonal {0,6} {1,7} {2,4} {3,5}: both differ from a neighbour, the other six are equal
	m |= e01 & e26 & ~(e12 | e32 | e37 | e47 | e45);
	m |= e01 & e37 & ~(e03 | e32 | e26 | e56 | e45);
	m |= e12 & e04 & ~(e01 | e03 | e37 | e76 | e56);
	m |= e03 & e15 & ~(e01 | e12 | e26 | e76 | e47);
	cx = act & m;
	const uint32_t sm = act & ~m;
	const int n = popc32(e01 & sm) + popc32(e12 & sm) + popc32(e32 & sm) + popc32(e03 & sm) +
	              popc32(e45 & sm) + popc32(e56 & sm) + popc32(e76 & sm) + popc32(e47 & sm) +
	              popc32(e04 & sm) + popc32(e15 & sm) + popc32(e26 & sm) + popc32(e37 & sm);
	return (uint32_t)(n - 2 * popc32(sm));
}

MC_HD bool row_points_owned(const Params &P, uint32_t z);

// generic form (any grid, on-iso samples included) straight from the bitmaps
template <typename Sample>
MC_COLD void count_word(const Params &P, const Tables &tb, uint32_t z, uint32_t y, uint32_t w, bool gz,
                       bool own_points, bool own_cells, uint64_t &cv, uint64_t &cc, uint32_t &visit)
{
	WordRec rec;
	CellWords cw;
	if (gz) word_masks_generic(P, z, y, w, rec, cw); else word_masks(P, z, y, w, false, rec, cw);
	if (!own_points) { rec.X = rec.Y = rec.Z = 0; }
	if (!own_cells) rec.act = 0;
	visit = rec.act | (row_points_owned(P, z) ? (rec.X | rec.Y | rec.Z) : 0u);
	cv = pack_planes(rec);
	cc = count_cells<Sample>(P, tb, z, y, w, rec.act, cw.c, cw.zc, cw.zany);
}

// ---------------------------------------------------------------------------
// vertex store: MC33_spn0/A/B/C (marching_cubes_33.c:485-621,
// MC33_util_grd.c:87-112); r[0..2] index-space position, r[3..5] = -grad F
// ---------------------------------------------------------------------------
// inclined grids, MC33_spnC (marching_cubes_33.c:595-621): p = _A r + O, n = A_^T n in double.
// Six values in, six out, all by value: the hot caller keeps them in registers.
template <typename Real> struct Vtx { Real p0, p1, p2, n0, n1, n2; };

template <typename Real>
MC_COLD Vtx<Real> store_spnc(const Geom &g, Vtx<Real> v)
{
	const double *A = g.A, *B = g.Ai;
	Real c0, c1, c2;
	const double r0 = v.p0, r1 = v.p1, r2 = v.p2;
	if (g.tsa) {
		c0 = (Real)radd(radd(rmul(A[0], r0), rmul(A[1], r1)), rmul(A[2], r2));
		c1 = (Real)radd(rmul(A[4], r1), rmul(A[5], r2));
		c2 = (Real)rmul(A[8], r2);
	} else {
		c0 = (Real)radd(radd(rmul(A[0], r0), rmul(A[1], r1)), rmul(A[2], r2));
		c1 = (Real)radd(radd(rmul(A[3], r0), rmul(A[4], r1)), rmul(A[5], r2));
		c2 = (Real)radd(radd(rmul(A[6], r0), rmul(A[7], r1)), rmul(A[8], r2));
	}
	Vtx<Real> o;
	o.p0 = radd(c0, (Real)g.O[0]); o.p1 = radd(c1, (Real)g.O[1]); o.p2 = radd(c2, (Real)g.O[2]);
	const double n0 = v.n0, n1 = v.n1, n2 = v.n2;
	if (g.tsa) {
		c2 = (Real)radd(radd(rmul(B[2], n0), rmul(B[5], n1)), rmul(B[8], n2));
		c1 = (Real)radd(rmul(B[1], n0), rmul(B[4], n1));
		c0 = (Real)rmul(B[0], n0);
	} else {
		c0 = (Real)radd(radd(rmul(B[0], n0), rmul(B[3], n1)), rmul(B[6], n2));
		c1 = (Real)radd(radd(rmul(B[1], n0), rmul(B[4], n1)), rmul(B[7], n2));
		c2 = (Real)radd(radd(rmul(B[2], n0), rmul(B[5], n1)), rmul(B[8], n2));
	}
	o.n0 = c0; o.n1 = c1; o.n2 = c2;
	return o;
}

// r[0..2] index-space position, r[3..5] = -grad F  ->  stored position Vo[3] and unit normal No[3]
template <typename Real>
MC_HD void transform_vertex(const Params &P, const Real *r, Real *Vo, float *No)
{
	const Geom &g = P.geom;
	Vtx<Real> v;
	v.p0 = r[0]; v.p1 = r[1]; v.p2 = r[2]; v.n0 = r[3]; v.n1 = r[4]; v.n2 = r[5];
	if (g.store == STORE_SPNC) {
		v = store_spnc<Real>(g, v);
	} else if (g.store != STORE_SPN0) {
		if (g.store == STORE_SPNB) {
			v.n0 = rmul(v.n0, geomCa(g, Real()));
			v.n1 = rmul(v.n1, geomCb(g, Real()));
		}
		v.p0 = radd(rmul(v.p0, geomD(g, 0, Real())), geomO(g, 0, Real()));
		v.p1 = radd(rmul(v.p1, geomD(g, 1, Real())), geomO(g, 1, Real()));
		v.p2 = radd(rmul(v.p2, geomD(g, 2, Real())), geomO(g, 2, Real()));
	}
	const Real s = radd(radd(rmul(v.n0, v.n0), rmul(v.n1, v.n1)), rmul(v.n2, v.n2));
	// the reference's rsqrtss is only good to 3e-4; the device uses the
	// hardware reciprocal square root (2 ulp), the host emulation the exact quotient
#if defined(__CUDA_ARCH__)
	float t = rsqrtf((float)s);
#else
	float t = rdiv(1.0f, sqrtf((float)s));
#endif
	if (g.normal_neg) t = -t;
	Vo[0] = v.p0; Vo[1] = v.p1; Vo[2] = v.p2;
	No[0] = rmul(t, (float)v.n0); No[1] = rmul(t, (float)v.n1); No[2] = rmul(t, (float)v.n2);
}

template <typename Real>
MC_HD void store_vertex(const Params &P, const Real *r, uint32_t id)
{
	Real Vo[3];
	float No[3];
	transform_vertex<Real>(P, r, Vo, No);
	Real *V = (Real *)P.V + 3 * (uint64_t)id;
	float *N = P.N + 3 * (uint64_t)id;
#if defined(__CUDA_ARCH__) && MC33_STREAM_STORES
	// the mesh is written once and not read again on the device: streaming stores (evict first), so that it does not push
	// the sample lines the next slice's vertices will need out of the L2
	__stcs(V, Vo[0]); __stcs(V + 1, Vo[1]); __stcs(V + 2, Vo[2]);
	__stcs(N, No[0]); __stcs(N + 1, No[1]); __stcs(N + 2, No[2]);
	__stcs(P.color + id, P.color_value);
#else
	V[0] = Vo[0]; V[1] = Vo[1]; V[2] = Vo[2];
	N[0] = No[0]; N[1] = No[1]; N[2] = No[2];
	P.color[id] = P.color_value;
#endif
}

// ---------------------------------------------------------------------------
// EDGE vertex of plane a (0 X, 1 Y, 2 Z) at grid point (x,y,z): reference
// marching_cubes_33.c:780-1224 (e.g. :990-1000), normal stencil SURVEY.md A.7.
// Written without a per-plane code path (a is a run-time value selected by
// predication) so that a warp whose lanes hold vertices of different planes
// does not diverge; only grid-boundary points take a separate branch.
// ---------------------------------------------------------------------------
template <typename Sample>
MC_HD void edge_vertex_r(const Params &P, uint32_t x, uint32_t y, uint32_t z, int a, typename Traits<Sample>::Real *r)
{
	typedef typename Traits<Sample>::Real Real;
	const Real iso = (Real)P.iso;
	// offsets within +-1 slice of the point fit 32 bits (a slice holds < 2^31 samples, checked
	// at context creation): one 64-bit row address, then 32-bit element offsets
	const int32_t sy = (int32_t)P.NX, sz = (int32_t)(P.NX * P.NY);
	// rotated axes: a along the edge, b = a+1, c = a+2 (mod 3) across it
	const int32_t sa = a == 0 ? 1 : (a == 1 ? sy : sz);
	const int32_t sb = a == 0 ? sy : (a == 1 ? sz : 1);
	const int32_t sc = a == 0 ? sz : (a == 1 ? 1 : sy);
	const uint32_t qa = a == 0 ? x : (a == 1 ? y : z), qb = a == 0 ? y : (a == 1 ? z : x), qc = a == 0 ? z : (a == 1 ? x : y);
	const uint32_t nb = a == 0 ? P.ny : (a == 1 ? P.nz : P.nx), nc = a == 0 ? P.nz : (a == 1 ? P.nx : P.ny);
	const Sample *p0 = (const Sample *)P.data + ((uint64_t)(z - P.zlo) * P.NY + y) * P.NX + x;
	const Sample *p1 = p0 + sa;
	const Real va = rsub(iso, (Real)ldro(p0)), vb = rsub(iso, (Real)ldro(p1));
	const Real t = rdiv(va, rsub(va, vb));
	const Real one_t = rsub((Real)1, t);
	// Across-axis differences.  The four neighbours of each end point are loaded
	// unconditionally with the step clamped at the grid faces, so that all ten loads of a
	// vertex are in flight together (a branch around them serialised two round trips):
	// at a face one of the two samples is the end point itself and
	// (iso - F(+)) - (iso - F(-)) is exactly the reference's one-sided difference of the
	// iso-subtracted values (forward e.g. c:813, backward e.g. c:995 else-branch).
	Real g[2];
#pragma unroll
	for (int k = 0; k < 2; k++) {
		const int32_t s = k ? sc : sb;
		const uint32_t q = k ? qc : qb, n = k ? nc : nb;
		const bool face = q == 0 || q == n;
		const int32_t sm = q == 0 ? 0 : -s, sp = q == n ? 0 : s;
		const Sample f0m = ldro(p0 + sm), f0p = ldro(p0 + sp), f1m = ldro(p0 + (sa + sm)), f1p = ldro(p0 + (sa + sp));
		if (!face) {
			// central difference on raw samples (e.g. c:993-994)
			const Real e0 = rawdiff(f0m, f0p), e1 = rawdiff(f1m, f1p);
			g[k] = rmul((Real)0.5f, radd(rmul(e0, one_t), rmul(e1, t)));
		} else {
			const Real d0 = rsub(rsub(iso, (Real)f0p), rsub(iso, (Real)f0m));
			const Real d1 = rsub(rsub(iso, (Real)f1p), rsub(iso, (Real)f1m));
			g[k] = radd(rmul(d0, one_t), rmul(d1, t));
		}
	}
	const Real ga = rsub(vb, va), pa = radd((Real)qa, t);
	r[0] = a == 0 ? pa : (Real)x;
	r[1] = a == 1 ? pa : (Real)y;
	r[2] = a == 2 ? pa : (Real)z;
	r[3] = a == 0 ? ga : (a == 1 ? g[1] : g[0]);
	r[4] = a == 0 ? g[0] : (a == 1 ? ga : g[1]);
	r[5] = a == 0 ? g[1] : (a == 1 ? g[0] : ga);
}

template <typename Sample>
MC_HDN void emit_edge_vertex(const Params &P, uint32_t x, uint32_t y, uint32_t z, int a, uint32_t id)
{
	typedef typename Traits<Sample>::Real Real;
	Real r[6];
	edge_vertex_r<Sample>(P, x, y, z, a, r);
	store_vertex<Real>(P, r, id);
}

// POINT vertex: MC33_surfint (marching_cubes_33.c:628-649)
template <typename Sample>
MC_COLD void point_vertex_r(const Params &P, uint32_t x, uint32_t y, uint32_t z, typename Traits<Sample>::Real *r)
{
	typedef typename Traits<Sample>::Real Real;
	const int64_t sy = (int64_t)P.NX, sz = (int64_t)P.NX * P.NY;
	const Sample *p = (const Sample *)P.data + ((uint64_t)(z - P.zlo) * P.NY + y) * P.NX + x;
	r[0] = (Real)x; r[1] = (Real)y; r[2] = (Real)z; r[3] = 0; r[4] = 0; r[5] = 0;
#pragma unroll
	for (int c = 0; c < 3; c++) {
		const int64_t sc = c == 0 ? (int64_t)1 : (c == 1 ? sy : sz);
		const uint32_t q = c == 0 ? x : (c == 1 ? y : z);
		const uint32_t nq = c == 0 ? P.nx : (c == 1 ? P.ny : P.nz);
		if (q == 0) r[3 + c] = rawdiff(p[0], p[sc]);
		else if (q == nq) r[3 + c] = rawdiff(p[-sc], p[0]);
		// 0.5f*(F - F): float product for float and integer grids, double for double
		else r[3 + c] = (Real)rmul((Real)0.5f, rawdiff(p[-sc], p[sc]));
	}
}

template <typename Sample>
MC_COLD void emit_point_vertex(const Params &P, uint32_t x, uint32_t y, uint32_t z, uint32_t id)
{
	typedef typename Traits<Sample>::Real Real;
	Real r[6];
	point_vertex_r<Sample>(P, x, y, z, r);
	store_vertex<Real>(P, r, id);
}

// CENTRE vertex (edge code 12): marching_cubes_33.c:1225-1230
template <typename Sample>
MC_COLD void emit_centre_vertex(const Params &P, uint32_t x, uint32_t y, uint32_t z, uint32_t id)
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

// ---------------------------------------------------------------------------
// vertex numbering.  Within a row: X plane by x, then Y, then Z.  wpreV[lr][w]
// packs, per plane (21 bits each), the row-local index of the plane's first
// vertex in word w -- i.e. the vertices of that plane in words < w, plus the
// totals of the planes in front of it; rowBV[lr] is the slab-local id of the
// row's first vertex.  Vertex ARRAYS are indexed locally; triangle CONTENTS are
// global ids (local + vbase).
// ---------------------------------------------------------------------------
MC_HD uint32_t fldV(uint64_t p, int pl) { return (uint32_t)(p >> (21 * pl)) & 0x1FFFFFu; }

MC_HD uint32_t plane_base_local(const Params &P, uint32_t lr, uint32_t w, int pl)
{
	return P.rowBV[lr] + fldV(P.wpreV[(uint64_t)lr * P.WP + w], pl);
}

// what the count kernel adds to every prefix of a row whose plane totals are tot:
// the Y plane starts after the X plane, the Z plane after both
MC_HD uint64_t plane_offsets(uint64_t tot)
{
	return ((uint64_t)fldV(tot, 0) << 21) | ((uint64_t)(fldV(tot, 0) + fldV(tot, 1)) << 42);
}
MC_HD uint32_t local_to_global(const Params &P, uint32_t z, uint32_t local)
{
	// rows of the halo slice are numbered in the next slab's index space
	if (z == P.hz) return (P.dbases ? P.dbases[1] : P.vbase_next) + (local - P.totals->nShared);
	return (P.dbases ? P.dbases[0] : P.vbase) + local;
}

// one vertex: plane a of grid point (x,y,z) -> slab-local vertex index id
template <typename Sample, bool KEYS = true>
MC_HDN void emit_vertex_task(const Params &P, uint32_t x, uint32_t y, uint32_t z, int a, bool is_point, uint32_t id)
{
	if (id >= P.capV) { P.totals->overflow = 1; return; }
	if (is_point) emit_point_vertex<Sample>(P, x, y, z, id);
	else emit_edge_vertex<Sample>(P, x, y, z, a, id);
	if (KEYS && P.vkey) P.vkey[id] = (((uint64_t)z * P.NY + y) * P.NX + x) * 4 + (unsigned)a;
}

// ---------------------------------------------------------------------------
// Vertex tasks.  The cell kernel already knows, for the grid point at the low corner
// of each cell it visits, which of the point's three planes carry a vertex and what
// their ids are, so it leaves one 8-byte task per vertex in vtask[id]:
// local row | (x | plane << 16 | is_point << 18) << 32.  The vertex kernel is then
// dense: thread id reads task id and computes vertex id.
// ---------------------------------------------------------------------------
MC_HD void put_vertex_task(const Params &P, uint32_t id, uint32_t lr, uint32_t x, unsigned a, bool is_point)
{
	// (ids beyond the capacity: the vertex kernel raises the overflow flag)
	if (id < P.capV) P.vtask[id] = (uint64_t)lr | ((uint64_t)(x | (a << 16) | ((uint32_t)is_point << 18)) << 32);
}

template <typename Sample, bool KEYS = true>
MC_HD void run_vertex_task(const Params &P, uint32_t id)
{
#if defined(__CUDA_ARCH__) && MC33_STREAM_STORES
	const uint64_t t = __ldcs(reinterpret_cast<const unsigned long long *>(P.vtask) + id);      // (read once)
#else
	const uint64_t t = P.vtask[id];
#endif
	const uint32_t lr = (uint32_t)t, e = (uint32_t)(t >> 32);
	const uint32_t zl = fastdiv(lr, P.NY, P.mNY), y = lr - zl * P.NY, z = zl + P.zlo;
	emit_vertex_task<Sample, KEYS>(P, e & 0xFFFFu, y, z, (int)((e >> 16) & 3u), ((e >> 18) & 1u) != 0, id);
}

// tasks of the vertices owned by grid point (x,y,z) on a grid WITH on-iso samples
// (generic path): rec = plane masks of the point's own row and word, POINT vertices flagged
MC_HDN void put_vertex_tasks_rec(const Params &P, uint32_t x, uint32_t y, uint32_t z, const WordRec &rec)
{
	const uint32_t lr = (z - P.zlo) * P.NY + y, w = x >> 5, b = x & 31u;
	const uint32_t zw = P.rowZ[lr] == P.zepoch ? P.Z[(uint64_t)lr * P.WP + w] : 0u;
	const uint32_t lo = (1u << b) - 1u;
#pragma unroll
	for (int a = 0; a < 3; a++) {
		const uint32_t m = a == 0 ? rec.X : (a == 1 ? rec.Y : rec.Z);
		if ((m >> b) & 1u)
			put_vertex_task(P, plane_base_local(P, lr, w, a) + (uint32_t)popc32(m & lo), lr, x, (unsigned)a,
			                a == 0 && ((zw >> b) & 1u));
	}
}

// ---------------------------------------------------------------------------
// Vertex ids referenced by the cells of word w of cell row (z,y): eight
// (plane mask, global id of the plane's first vertex in this word) pairs
// plane combos: 0 X00  1 Y00  2 Z00  3 X10  4 Z10  5 X01  6 Y01  7 X11
// (suffix = dy dz of the point row relative to the cell row).  Ranks only count
// bits below the queried one (offset <= 32), so the 32 bits of word w are all
// that is needed even for x = 32w+32.
// ---------------------------------------------------------------------------
MC_HD unsigned combo_of_edge(unsigned e)  { return (0x573026412641ull >> (4 * e)) & 15; }
MC_HD unsigned cx_of_edge(unsigned e)     { return (0x0F0u >> e) & 1; }
MC_HD unsigned combo_of_corner(unsigned c) { return (0x57305730u >> (4 * c)) & 15; }

struct CellPairs { uint32_t mask[8], base[8]; };

MC_COLD void cell_pairs(const Params &P, uint32_t z, uint32_t y, uint32_t w, bool gz, const WordRec &rec00,
                       const CellWords &cw, CellPairs &cp)
{
	const uint32_t lr = (z - P.zlo) * P.NY + y;
	WordRec r10, r01, r11;
	if (!gz) {
		// no on-iso sample anywhere: plane masks straight from the corner sign words
		// (c[1],c[5] = row y+1 at x, x+1; c[3],c[7] = row z+1; c[2],c[6] = row (y+1,z+1))
		const uint32_t vp = mask_le(w, P.nx), vx = mask_le(w, P.nx - 1);
		r10.X = (cw.c[1] ^ cw.c[5]) & vx;
		r01.X = (cw.c[3] ^ cw.c[7]) & vx;
		r11.X = (cw.c[2] ^ cw.c[6]) & vx;
		r10.Z = (cw.c[1] ^ cw.c[2]) & vp;
		r01.Y = (cw.c[3] ^ cw.c[2]) & vp;
	} else {
		CellWords dummy;
		word_masks_generic(P, z, y + 1, w, r10, dummy);
		word_masks_generic(P, z + 1, y, w, r01, dummy);
		word_masks_generic(P, z + 1, y + 1, w, r11, dummy);
	}
	const uint32_t l10 = lr + 1, l01 = lr + P.NY, l11 = l01 + 1;
	cp.mask[0] = rec00.X; cp.base[0] = local_to_global(P, z, plane_base_local(P, lr, w, 0));
	cp.mask[1] = rec00.Y; cp.base[1] = local_to_global(P, z, plane_base_local(P, lr, w, 1));
	cp.mask[2] = rec00.Z; cp.base[2] = local_to_global(P, z, plane_base_local(P, lr, w, 2));
	cp.mask[3] = r10.X;   cp.base[3] = local_to_global(P, z, plane_base_local(P, l10, w, 0));
	cp.mask[4] = r10.Z;   cp.base[4] = local_to_global(P, z, plane_base_local(P, l10, w, 2));
	cp.mask[5] = r01.X;   cp.base[5] = local_to_global(P, z + 1, plane_base_local(P, l01, w, 0));
	cp.mask[6] = r01.Y;   cp.base[6] = local_to_global(P, z + 1, plane_base_local(P, l01, w, 1));
	cp.mask[7] = r11.X;   cp.base[7] = local_to_global(P, z + 1, plane_base_local(P, l11, w, 0));
}

// KEYS: the optional per-triangle / per-vertex canonical keys (tests); the kernels are also
// built without them so that the hot loops carry no code for them
template <bool KEYS = true>
MC_HD void write_triangle(const Params &P, uint32_t tid, const uint32_t *ti, unsigned m, uint64_t cell)
{
	if (tid >= P.capT) { P.totals->overflow = 1; return; }
	// winding: marching_cubes_33.c:1246-1250 (ti[0] = nibble 2, ti[1] = nibble 1, ti[2] = nibble 0)
	uint32_t a0 = m ? ti[0] : ti[1], a1 = m ? ti[1] : ti[0];
	if (P.geom.normal_neg) { uint32_t t = a0; a0 = a1; a1 = t; }
	uint32_t *T = P.T + 3 * (uint64_t)tid;
#if defined(__CUDA_ARCH__) && MC33_STREAM_STORES
	__stcs(T, a0); __stcs(T + 1, a1); __stcs(T + 2, ti[2]);
#else
	T[0] = a0; T[1] = a1; T[2] = ti[2];
#endif
	if (KEYS && P.tcell) P.tcell[tid] = cell;
}

// ---------------------------------------------------------------------------
// Fast path for a cell of a grid WITHOUT on-iso samples: the 8-bit case index,
// which vertices its low corner point owns, and
// the global ids of the vertices on its 12 edges, straight from the sign bitmap,
// the word prefixes and the row bases.  Every edge is evaluated (no dependence on
// the pattern), ids of edges that carry no vertex are meaningless and never read.
// g0 / g1: what turns a slab-local id of slice z / z+1 into a global one.
// ---------------------------------------------------------------------------
MC_HDN unsigned cell_fast(const Params &P, uint32_t x, uint32_t y, uint32_t z, uint32_t g0, uint32_t g1, uint32_t *id, unsigned &own)
{
	const uint32_t w = x >> 5, b = x & 31u;
	// (the point rows of the grid's high faces are visited for the vertices they own: a row
	// that does not exist is replaced by the row itself, which empties the plane towards it)
	const uint32_t uy = y < P.ny ? 1u : 0u, uz = z < P.nz ? P.NY : 0u;
	const uint32_t l00 = (z - P.zlo) * P.NY + y, l10 = l00 + uy, l01 = l00 + uz, l11 = l01 + uy;
	// word indices fit 32 bits (checked at context creation)
	const uint32_t i00 = l00 * P.WP + w, i10 = l10 * P.WP + w, i01 = l01 * P.WP + w, i11 = l11 * P.WP + w;
	const uint32_t s00 = P.S[i00], s10 = P.S[i10], s01 = P.S[i01], s11 = P.S[i11];
	const uint32_t x00 = shr1(s00, P.S[i00 + 1]), x10 = shr1(s10, P.S[i10 + 1]);
	const uint32_t x01 = shr1(s01, P.S[i01 + 1]), x11 = shr1(s11, P.S[i11 + 1]);
	const uint64_t p00 = P.wpreV[i00], p10 = P.wpreV[i10], p01 = P.wpreV[i01], p11 = P.wpreV[i11];
	const uint32_t r00 = P.rowBV[l00] + g0, r10 = P.rowBV[l10] + g0, r01 = P.rowBV[l01] + g1, r11 = P.rowBV[l11] + g1;
	// bits beyond x = nx are zero in S, so only the X plane needs a mask: the last point has no X edge
	const uint32_t vx = mask_le(w, P.nx - 1);
	const uint32_t lo0 = (1u << b) - 1u;                            // bits below x
	// plane masks (suffix = dy dz of the point row) and the id of their first vertex in word w
	const uint32_t mX00 = (s00 ^ x00) & vx, mY00 = s00 ^ s10, mZ00 = s00 ^ s01;
	const uint32_t mX10 = (s10 ^ x10) & vx, mZ10 = s10 ^ s11;
	const uint32_t mX01 = (s01 ^ x01) & vx, mY01 = s01 ^ s11;
	const uint32_t mX11 = (s11 ^ x11) & vx;
	const uint32_t bX00 = r00 + fldV(p00, 0), bY00 = r00 + fldV(p00, 1), bZ00 = r00 + fldV(p00, 2);
	const uint32_t bX10 = r10 + fldV(p10, 0), bZ10 = r10 + fldV(p10, 2);
	const uint32_t bX01 = r01 + fldV(p01, 0), bY01 = r01 + fldV(p01, 1);
	const uint32_t bX11 = r11 + fldV(p11, 0);
	// edges: 0:(0,1)y 1:(1,2)z 2:(3,2)y 3:(0,3)z 4:(4,5)y 5:(5,6)z 6:(7,6)y 7:(4,7)z 8:(0,4)x 9:(1,5)x 10:(2,6)x 11:(3,7)x
	// (the edge at x+1 follows the one at x in the same plane: one more if the point x carries a vertex)
	id[0] = bY00 + (uint32_t)popc32(mY00 & lo0);  id[4] = id[0] + ((mY00 >> b) & 1u);
	id[1] = bZ10 + (uint32_t)popc32(mZ10 & lo0);  id[5] = id[1] + ((mZ10 >> b) & 1u);
	id[2] = bY01 + (uint32_t)popc32(mY01 & lo0);  id[6] = id[2] + ((mY01 >> b) & 1u);
	id[3] = bZ00 + (uint32_t)popc32(mZ00 & lo0);  id[7] = id[3] + ((mZ00 >> b) & 1u);
	id[8] = bX00 + (uint32_t)popc32(mX00 & lo0);
	id[9] = bX10 + (uint32_t)popc32(mX10 & lo0);
	id[10] = bX11 + (uint32_t)popc32(mX11 & lo0);
	id[11] = bX01 + (uint32_t)popc32(mX01 & lo0);
	// planes of the point (x,y,z) itself that carry a vertex: bit 0 X (id[8]), 1 Y (id[0]), 2 Z (id[3])
	own = ((mX00 >> b) & 1u) | (((mY00 >> b) & 1u) << 1) | (((mZ00 >> b) & 1u) << 2);
	// corner k -> index bit 7-k (corners 0..3 at x: rows 00 10 11 01; 4..7 at x+1)
	return (((s00 >> b) & 1u) << 7) | (((s10 >> b) & 1u) << 6) | (((s11 >> b) & 1u) << 5) | (((s01 >> b) & 1u) << 4) |
	       (((x00 >> b) & 1u) << 3) | (((x10 >> b) & 1u) << 2) | (((x11 >> b) & 1u) << 1) | ((x01 >> b) & 1u);
}

// What a visited cell of row lr needs about its four point rows (suffix = dy dz), built once per row group: word
// offsets of the rows in the bitmaps / prefixes, id of each row's first vertex as a GLOBAL id (the halo slice is
// numbered by the next slab), and which of the row's points / cells this slab owns.  (A row that does not exist --
// beyond the grid's high faces -- is replaced by the row itself, which empties the plane towards it.)
struct RowT { uint32_t i00, i10, i01, i11, r00, r10, r01, r11, y, z, flags, g0; };

MC_HD RowT make_rowt(const Params &P, uint32_t lr, uint32_t vb, uint32_t vbn)
{
	RowT t;
	const uint32_t zl = fastdiv(lr, P.NY, P.mNY), y = lr - zl * P.NY, z = zl + P.zlo;
	const uint32_t uy = y < P.ny ? 1u : 0u, uz = z < P.nz ? P.NY : 0u;
	const uint32_t l00 = lr, l10 = l00 + uy, l01 = l00 + uz, l11 = l01 + uy;
	const uint32_t g0 = z == P.hz ? vbn : vb, g1 = z + 1 == P.hz ? vbn : vb;
	t.i00 = l00 * P.WP; t.i10 = l10 * P.WP; t.i01 = l01 * P.WP; t.i11 = l11 * P.WP;
	t.r00 = P.rowBV[l00] + g0; t.r10 = P.rowBV[l10] + g0; t.r01 = P.rowBV[l01] + g1; t.r11 = P.rowBV[l11] + g1;
	t.y = y; t.z = z;
	t.flags = (row_points_owned(P, z) ? 1u : 0u) | (row_cells_owned(P, z, y) ? 2u : 0u);
	t.g0 = g0;
	return t;
}

// cell_fast (mc33_core.cuh) with the per-row work taken from the row table: 12 loads, 8 popcounts.  (No mask for the
// X plane's last point: a rank only counts bits below the cell's own, and the spurious bit of the row's last point
// sits above every cell of its word; the caller clears it from `own`.)
MC_HD unsigned cell_fast_rt(const Params &P, const RowT &t, uint32_t x, uint32_t *id, unsigned &own)
{
	const uint32_t w = x >> 5, b = x & 31u;
	const uint32_t i00 = t.i00 + w, i10 = t.i10 + w, i01 = t.i01 + w, i11 = t.i11 + w;
	const uint32_t s00 = P.S[i00], s10 = P.S[i10], s01 = P.S[i01], s11 = P.S[i11];
	const uint32_t x00 = shr1(s00, P.S[i00 + 1]), x10 = shr1(s10, P.S[i10 + 1]);
	const uint32_t x01 = shr1(s01, P.S[i01 + 1]), x11 = shr1(s11, P.S[i11 + 1]);
	const uint64_t p00 = P.wpreV[i00], p10 = P.wpreV[i10], p01 = P.wpreV[i01], p11 = P.wpreV[i11];
	const uint32_t lo0 = (1u << b) - 1u;
	const uint32_t mX00 = s00 ^ x00, mY00 = s00 ^ s10, mZ00 = s00 ^ s01;
	const uint32_t mX10 = s10 ^ x10, mZ10 = s10 ^ s11;
	const uint32_t mX01 = s01 ^ x01, mY01 = s01 ^ s11;
	const uint32_t mX11 = s11 ^ x11;
	const uint32_t bY = (mY00 >> b) & 1u, bZ = (mZ00 >> b) & 1u;
	id[0] = t.r00 + fldV(p00, 1) + (uint32_t)popc32(mY00 & lo0);  id[4] = id[0] + bY;
	id[1] = t.r10 + fldV(p10, 2) + (uint32_t)popc32(mZ10 & lo0);  id[5] = id[1] + ((mZ10 >> b) & 1u);
	id[2] = t.r01 + fldV(p01, 1) + (uint32_t)popc32(mY01 & lo0);  id[6] = id[2] + ((mY01 >> b) & 1u);
	id[3] = t.r00 + fldV(p00, 2) + (uint32_t)popc32(mZ00 & lo0);  id[7] = id[3] + bZ;
	id[8] = t.r00 + fldV(p00, 0) + (uint32_t)popc32(mX00 & lo0);
	id[9] = t.r10 + fldV(p10, 0) + (uint32_t)popc32(mX10 & lo0);
	id[10] = t.r11 + fldV(p11, 0) + (uint32_t)popc32(mX11 & lo0);
	id[11] = t.r01 + fldV(p01, 0) + (uint32_t)popc32(mX01 & lo0);
	own = ((mX00 >> b) & 1u) | (bY << 1) | (bZ << 2);
	return (((s00 >> b) & 1u) << 7) | (((s10 >> b) & 1u) << 6) | (((s11 >> b) & 1u) << 5) | (((s01 >> b) & 1u) << 4) |
	       (((x00 >> b) & 1u) << 3) | (((x10 >> b) & 1u) << 2) | (((x11 >> b) & 1u) << 1) | ((x01 >> b) & 1u);
}

// triangle j of a cell whose 13 vertex ids (12 edges + centre) sit at ids[e * stride]
template <bool KEYS = true>
MC_HD void emit_triangle_fast(const Params &P, unsigned tw, unsigned m, const uint32_t *ids, uint32_t stride, uint32_t tid,
                              uint64_t cell)
{
	uint32_t ti[3];
	ti[0] = ids[((tw >> 8) & 15u) * stride];
	ti[1] = ids[((tw >> 4) & 15u) * stride];
	ti[2] = ids[(tw & 15u) * stride];
	write_triangle<KEYS>(P, tid, ti, m, cell);
}

// global id of the vertex a triangle corner refers to: edge code e (0..11) of the
// cell at bit b; zm = on-iso corner mask of the cell; key = identity used by the
// zero-area test (marching_cubes_33.c:1235)
MC_HD uint32_t corner_vertex(const uint32_t *pmask, const uint32_t *pbase, uint32_t stride, unsigned e, unsigned b,
                             unsigned zm, unsigned &key)
{
	unsigned combo = combo_of_edge(e), off = b + cx_of_edge(e);
	key = e;
	if (zm) {
		const unsigned a = edge_a(e), bb = edge_b(e);
		if ((zm >> a) & 1) { key = 16 + a; combo = combo_of_corner(a); off = b + MC_CX(a); }
		else if ((zm >> bb) & 1) { key = 16 + bb; combo = combo_of_corner(bb); off = b + MC_CX(bb); }
	}
	const uint32_t mk = pmask[combo * stride];
	return pbase[combo * stride] + (uint32_t)popc32(off >= 32 ? mk : (mk & ((1u << off) - 1u)));
}

// all triangles of a cell WITH on-iso corners (zero-area triangles are dropped);
// only ids in [lo, hi) are written.  Returns the number of triangles kept.
MC_COLD uint32_t emit_cell_triangles_z(const Params &P, const Tables &tb, unsigned b, const CellPattern &cp, unsigned zm,
                                      uint32_t centre_id, const uint32_t *pmask, const uint32_t *pbase, uint32_t stride,
                                      uint32_t tid, uint32_t lo, uint32_t hi, uint64_t cell)
{
	uint32_t n = 0;
	for (unsigned tw_i = cp.start;; tw_i++) {
		const unsigned tw = tb.tri[tw_i];
		uint32_t ti[3];
		unsigned key[3];
#pragma unroll
		for (int j = 0; j < 3; j++) {
			const unsigned e = (tw >> (8 - 4 * j)) & 15;
			if (e == 12) { ti[j] = centre_id; key[j] = 12; }
			else ti[j] = corner_vertex(pmask, pbase, stride, e, b, zm, key[j]);
		}
		if (key[0] != key[1] && key[0] != key[2] && key[1] != key[2]) {
			const uint32_t id = tid + n;
			if (id >= lo && id < hi) write_triangle(P, id, ti, cp.m, cell);
			n++;
		}
		if (!(tw >> 12)) break;
	}
	return n;
}

// The complex cells of one quad (words without on-iso samples: not bit k of slow), from
// scratch: the count kernel defers them to the end of its pass, where almost nothing is
// live, and only remembers that the quad has some.  -> nT | nC << 32
template <typename Sample>
MC_COLD uint64_t count_quad_complex(const Params &P, const Tables &tb, uint32_t lr, uint32_t q, uint32_t slow)
{
	const uint32_t zl = fastdiv(lr, P.NY, P.mNY), y = lr - zl * P.NY, z = zl + P.zlo;
	const bool hasY = y < P.ny, hasZ = z < P.nz;
	const uint64_t dY = hasY ? P.WP : 0u, dZ = hasZ ? (uint64_t)P.NY * P.WP : 0u;
	const uint64_t i00 = (uint64_t)lr * P.WP + 4 * q;
	const Quad q00 = load_quad(P.S, i00), q10 = load_quad(P.S, i00 + dY);
	const Quad q01 = load_quad(P.S, i00 + dZ), q11 = load_quad(P.S, i00 + dY + dZ);
	const bool cells = row_cells_owned(P, z, y) && hasZ;
	uint32_t cx[4] = {0, 0, 0, 0};
#pragma unroll
	for (int k = 0; k < 4; k++) {
		if ((slow >> k) & 1u) continue;
		WordRec rec;
		uint32_t c[8];
		quad_word(P, q00, q10, q01, q11, k, 4 * q + k, cells, rec, c);
		if (rec.act) count_simple_cells(c, rec.act, cx[k]);
	}
	return count_cells_quad<Sample>(P, tb, z, y, q, cx[0], cx[1], cx[2], cx[3], i00, dY, dZ);
}

// the words of a quad that hold an on-iso sample (bit k of slow), generic rules; one call per
// quad so that the count kernel's hot loop stays free of calls
struct QuadSlow { uint64_t pv[4]; uint64_t cc; uint32_t vis[4]; };

template <typename Sample>
MC_COLD QuadSlow count_quad_slow(const Params &P, const Tables &tb, uint32_t z, uint32_t y, uint32_t q, uint32_t slow,
                                 bool own_points, bool own_cells)
{
	QuadSlow r;
	r.cc = 0;
	for (int k = 0; k < 4; k++) {
		r.pv[k] = 0; r.vis[k] = 0;
		const uint32_t w = 4 * q + (uint32_t)k;
		if (((slow >> k) & 1u) && w < P.W) {
			uint64_t cc;
			count_word<Sample>(P, tb, z, y, w, true, own_points, own_cells, r.pv[k], cc, r.vis[k]);
			r.cc += cc;
		}
	}
	return r;
}

// ---------------------------------------------------------------------------
// One visited cell of a word that holds an on-iso sample (generic rules): the vertex
// tasks of its low corner point and, if the cell is active, its pattern and on-iso
// corner mask.  Without an on-iso corner its 12 edge vertex ids go to ids[e * stride]
// exactly like cell_fast's, so it joins the dense triangle loop; with one, its own lane
// writes the triangles afterwards (cell_slow_triangles), dropping the zero-area ones.
// ---------------------------------------------------------------------------
template <typename Sample>
MC_COLD CellPattern cell_slow(const Params &P, const Tables &tb, uint32_t x, uint32_t y, uint32_t z, bool ownp, bool cellok,
                              uint32_t *ids, uint32_t stride, unsigned &zm)
{
	CellPattern pat;
	pat.start = 0; pat.m = 0; pat.ntri = 0; pat.centre = 0;
	zm = 0;
	const uint32_t w = x >> 5, b = x & 31u;
	WordRec rec; CellWords cw;
	word_masks_generic(P, z, y, w, rec, cw);         // once: the point's tasks and the cell share it
	if (ownp) put_vertex_tasks_rec(P, x, y, z, rec);
	if (!cellok) return pat;
	// (with on-iso samples a point can own a vertex while its cell is inactive)
	if (!((rec.act >> b) & 1u)) return pat;
	const unsigned idx = cell_index(cw.c, 1, (int)b);
	zm = cw.zany ? cell_zmask(cw.zc, 1, (int)b) : 0u;
	pat = cell_pattern<Sample>(P, tb, x, y, z, idx, zm);
	if (!zm) {
		CellPairs cp;
		cell_pairs(P, z, y, w, true, rec, cw, cp);
		for (unsigned e = 0; e < 12; e++) {
			unsigned key;
			ids[e * stride] = corner_vertex(cp.mask, cp.base, 1, e, b, 0u, key);
		}
	}
	return pat;
}

template <typename Sample>
MC_COLD void cell_slow_triangles(const Params &P, const Tables &tb, uint32_t x, uint32_t y, uint32_t z, const CellPattern &pat,
                                 unsigned zm, uint32_t centre_id, uint32_t tid, uint64_t cell)
{
	WordRec rec; CellWords cw; CellPairs cp;
	word_masks_generic(P, z, y, x >> 5, rec, cw);
	cell_pairs(P, z, y, x >> 5, true, rec, cw, cp);
	emit_cell_triangles_z(P, tb, x & 31u, pat, zm, centre_id, cp.mask, cp.base, 1, tid, 0u, 0xFFFFFFFFu, cell);
}

}  // namespace mc33
