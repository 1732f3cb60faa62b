// mc33_kernels.cu -- sm_100a kernels and the C-ABI (include/mc33cu.h) of the
// B200 Marching Cubes 33 extractor.  See DESIGN.md for the pipeline:
//
//   K1 classify   stream the samples ONCE through a TMA (cp.async.bulk) ring in
//                 shared memory -> S / Z bitmaps (1 bit per sample); for an iso
//                 sweep, once for up to eight isovalues (k_classify_sweep)
//   K2 count      per (row, 32-point word): owned vertices per plane (popcounts),
//                 triangles (simple cells 32 at a time, complex ones walked) ->
//                 row-local word prefixes, visit bitmap, CTA-relative row bases
//   K2b rowscan   CTA-relative -> absolute row bases (vertices, triangles, centres)
//   K4 emit T     visited cells compacted per row group, one lane per cell (vertex
//                 ids, pattern, vertex tasks), then one lane per triangle
//   K3 emit V     dense: one thread per vertex id runs the task K4 left for it
//
// Replaces: reference source/marching_cubes_33.c:1816-1889 (calculate_isosurface),
// :673-1253 (MC33_findCase), :485-649 (store / surfint).  No CPU fallback.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mc33_core.cuh"
#include "mc33_pipeline.cuh"
#include "../../include/mc33cu.h"

using namespace mc33;

// ---------------------------------------------------------------------------
// case tables in device memory (copied to shared memory by the kernels that
// index them divergently)
// ---------------------------------------------------------------------------
#define TBL_TRI_BYTES ((MC33_NTRI_WORDS * 2 + 15) / 16 * 16)
#define TBL_PAT_BYTES ((MC33_NTRI_WORDS + 15) / 16 * 16)
#define TBL_CINFO_BYTES 1024
#define TBL_BYTES (512 + 512 + TBL_TRI_BYTES + TBL_PAT_BYTES + TBL_CINFO_BYTES)
#define TBL2_BYTES (TBL_TRI_BYTES + TBL_PAT_BYTES + TBL_CINFO_BYTES)     // what the round-2 cell kernel keeps in shared memory

// one image: case256 | simple256 | tri | pat | cinfo, copied to shared memory with 16-byte loads
__device__ __align__(16) unsigned char d_tables[TBL_BYTES];
// pattern ordinals and the keep masks of every (pattern, on-iso corner mask): read in place (only cells with an on-iso corner)
__device__ uint16_t d_pord[MC33_NTRI_WORDS];
__device__ uint16_t d_keep[MC33_NPATTERNS * 256];

// the same tables read in place (kernels that only index them on cold paths)
__device__ __forceinline__ Tables global_tables()
{
	Tables tb;
	tb.case256 = (const uint16_t *)d_tables;
	tb.simple256 = tb.case256 + 256;
	tb.tri = tb.simple256 + 256;
	tb.pat = (const uint8_t *)tb.tri + TBL_TRI_BYTES;
	tb.cinfo = (const uint32_t *)(tb.pat + TBL_PAT_BYTES);
	tb.pord = d_pord; tb.keep = d_keep;
	return tb;
}

// tri | pat | cinfo in shared memory (the cell kernel of round 2 never runs the MC33 tests: no case256 / simple256)
__device__ __forceinline__ Tables load_tables2(unsigned char *smem)
{
	const uint4 *src = reinterpret_cast<const uint4 *>(d_tables + 1024);
	uint4 *dst = reinterpret_cast<uint4 *>(smem);
	for (int i = threadIdx.x; i < TBL2_BYTES / 16; i += blockDim.x) dst[i] = src[i];
	__syncthreads();
	Tables tb;
	tb.case256 = (const uint16_t *)d_tables;
	tb.simple256 = tb.case256 + 256;
	tb.tri = (const uint16_t *)smem;
	tb.pat = (const uint8_t *)smem + TBL_TRI_BYTES;
	tb.cinfo = (const uint32_t *)(smem + TBL_TRI_BYTES + TBL_PAT_BYTES);
	tb.pord = d_pord; tb.keep = d_keep;
	return tb;
}

__device__ __forceinline__ Tables load_tables(unsigned char *smem)
{
	const uint4 *src = reinterpret_cast<const uint4 *>(d_tables);
	uint4 *dst = reinterpret_cast<uint4 *>(smem);
	for (int i = threadIdx.x; i < TBL_BYTES / 16; i += blockDim.x) dst[i] = src[i];
	__syncthreads();
	Tables tb;
	tb.case256 = (const uint16_t *)smem;
	tb.simple256 = tb.case256 + 256;
	tb.tri = tb.simple256 + 256;
	tb.pat = (const uint8_t *)tb.tri + TBL_TRI_BYTES;
	tb.cinfo = (const uint32_t *)(tb.pat + TBL_PAT_BYTES);
	tb.pord = d_pord; tb.keep = d_keep;
	return tb;
}

// ---------------------------------------------------------------------------
// TMA (bulk async copy) + mbarrier primitives, sm_90+/sm_100a PTX
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
	const uint32_t a = smem_u32(bar);
	uint32_t ok;
	do {
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
		             : "=r"(ok) : "r"(a), "r"(parity) : "memory");
	} while (!ok);
}
// global -> shared bulk copy (the 1-D TMA path: SASS UBLKCP); size and both
// addresses are multiples of 16 bytes; completion is signalled on the mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// ---------------------------------------------------------------------------
// K1: classify.
//
// The grid is cut into chunks of whole rows (or, for very long rows, pieces of
// one row): each chunk is one contiguous byte range.  A CTA walks its chunks
// through a CLS_STAGES-deep ring of shared-memory buffers filled by bulk async
// copies (one elected thread issues them; the <16-byte unaligned head and tail of
// a range are copied by a few threads with plain loads), so many KB per SM are in
// flight without holding registers.  A warp then turns a row into bitmap words:
// lane l reads sample 32*w + l from shared memory (conflict free) and
// __ballot_sync yields word w directly.
//
// The reference's index bit is the IEEE sign bit of iso - F and its on-iso test
// is iso - F == 0 (marching_cubes_33.c:1840-1859, :392-409).  With the isovalue
// folded onto +0.0 and no flush-to-zero, sign(iso - F) is set exactly when
// F > iso and iso - F == 0 exactly when F == iso, so the two bits are taken by
// comparison (integer grids: against integer thresholds precomputed on the
// host from the same float conversion), saving the subtraction.
//
// The Z bitmap is all zero for almost every row of real data: rows are only
// written when they have, or had in the previous extraction, an on-iso sample.
// ---------------------------------------------------------------------------
template <typename Sample> struct Cls {
	typename Traits<Sample>::Real iso;
	bool noz;
	__device__ __forceinline__ Cls(const Params &P) : iso((typename Traits<Sample>::Real)P.iso), noz(P.dbg_noz != 0) {}
	__device__ __forceinline__ bool gt(Sample f) const { return f > iso; }
	__device__ __forceinline__ bool eq(Sample f) const { return f == iso && !noz; }
};
template <typename Sample> struct ClsInt {
	uint32_t thr, eq_lo, eq_span;
	bool none, eq_none;
	__device__ __forceinline__ ClsInt(const Params &P)
		: thr(P.ithr), eq_lo(P.ieq_lo), eq_span(P.ieq_hi - P.ieq_lo), none(P.inone != 0), eq_none(P.ieq_hi < P.ieq_lo) {}
	__device__ __forceinline__ bool gt(Sample f) const { return (uint32_t)f >= thr && !none; }
	__device__ __forceinline__ bool eq(Sample f) const { return (uint32_t)f - eq_lo <= eq_span && !eq_none; }
};
template <> struct Cls<uint8_t> : ClsInt<uint8_t> { __device__ __forceinline__ Cls(const Params &P) : ClsInt<uint8_t>(P) {} };
template <> struct Cls<uint16_t> : ClsInt<uint16_t> { __device__ __forceinline__ Cls(const Params &P) : ClsInt<uint16_t>(P) {} };
template <> struct Cls<uint32_t> : ClsInt<uint32_t> { __device__ __forceinline__ Cls(const Params &P) : ClsInt<uint32_t>(P) {} };

#define CLS_THREADS 256
#define CLS_STAGES 4

struct ClsPlan {
	uint32_t rows, words;        // a chunk = `rows` whole rows (words == W) or `words` words of one row (rows == 1)
	uint32_t nwchunk;            // chunks per row (1 when rows are whole)
	uint32_t nchunks;
	uint32_t stage_bytes;        // multiple of 128
};

// byte range of chunk c within the sample array, and what it covers
struct ClsChunk { uint32_t lr0, nrows, w0, nw; uint64_t b0, b1; };

template <typename Sample>
__device__ __forceinline__ ClsChunk cls_chunk(const Params &P, const ClsPlan &pl, uint32_t c)
{
	ClsChunk k;
	if (pl.nwchunk == 1) {
		k.lr0 = c * pl.rows;
		k.nrows = min(pl.rows, P.Lrows - k.lr0);
		k.w0 = 0; k.nw = P.W;
		k.b0 = (uint64_t)k.lr0 * P.NX * sizeof(Sample);
		k.b1 = (uint64_t)(k.lr0 + k.nrows) * P.NX * sizeof(Sample);
	} else {
		k.lr0 = c / pl.nwchunk;
		k.nrows = 1;
		k.w0 = (c - k.lr0 * pl.nwchunk) * pl.words;
		k.nw = min(pl.words, P.W - k.w0);
		const uint64_t e0 = (uint64_t)k.lr0 * P.NX + ((uint64_t)k.w0 << 5);
		const uint64_t e1 = (uint64_t)k.lr0 * P.NX + min((uint64_t)P.NX, ((uint64_t)(k.w0 + k.nw) << 5));
		k.b0 = e0 * sizeof(Sample); k.b1 = e1 * sizeof(Sample);
	}
	return k;
}

template <typename Sample>
__global__ void __launch_bounds__(CLS_THREADS) k_classify(const __grid_constant__ Params P, ClsPlan pl)
{
	extern __shared__ __align__(128) unsigned char smem[];
	__shared__ uint64_t full[CLS_STAGES];
	const Cls<Sample> cls(P);
	const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const char *gbase = (const char *)P.data;
	const uint32_t nfull = P.NX >> 5, tail = P.NX & 31;

	if (threadIdx.x == 0) {
		for (int s = 0; s < CLS_STAGES; s++) mbar_init(&full[s], 1);
		fence_mbar_init();
	}
	__syncthreads();

	// fill stage `s` with chunk c: thread 0 issues the 16-byte aligned middle as a
	// bulk copy, threads 32.. copy the unaligned head / tail bytes
	auto issue = [&](uint32_t c, int s) {
		const ClsChunk k = cls_chunk<Sample>(P, pl, c);
		unsigned char *st = smem + (size_t)s * pl.stage_bytes;
		const uint64_t a0 = (uint64_t)(uintptr_t)gbase + k.b0, a1 = (uint64_t)(uintptr_t)gbase + k.b1;
		const uint64_t o = a0 & ~15ull;                    // address that maps to stage offset 0
		uint64_t m0 = (a0 + 15) & ~15ull, m1 = a1 & ~15ull;
		if (m1 < m0) { m0 = a1; m1 = a1; }                 // range inside one 16-byte block
		if (threadIdx.x == 0) {
			if (m1 > m0) {
				fence_proxy_async();
				mbar_arrive_expect_tx(&full[s], (uint32_t)(m1 - m0));
				bulk_g2s(st + (m0 - o), (const void *)(uintptr_t)m0, (uint32_t)(m1 - m0), &full[s]);
			} else {
				mbar_arrive(&full[s]);
			}
		} else if (threadIdx.x >= 32 && threadIdx.x < 48) {
			const uint64_t a = a0 + (threadIdx.x - 32);
			if (a < m0) st[a - o] = *(const unsigned char *)(uintptr_t)a;
		} else if (threadIdx.x >= 48 && threadIdx.x < 64) {
			const uint64_t a = m1 + (threadIdx.x - 48);
			if (a >= m0 && a < a1) st[a - o] = *(const unsigned char *)(uintptr_t)a;
		}
	};

	uint32_t nmine = blockIdx.x < pl.nchunks ? (pl.nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
	for (uint32_t k = 0; k < CLS_STAGES && k < nmine; k++) issue(blockIdx.x + k * gridDim.x, (int)k);
	__syncthreads();

	for (uint32_t k = 0; k < nmine; k++) {
		const int s = (int)(k % CLS_STAGES);
		const uint32_t c = blockIdx.x + k * gridDim.x;
		const ClsChunk ck = cls_chunk<Sample>(P, pl, c);
		// one warp polls the barrier, the others sleep at the CTA barrier (256 polling
		// threads were taking issue slots from the warps still classifying)
		if (wid == 0) mbar_wait(&full[s], (k / CLS_STAGES) & 1);
		__syncthreads();
		const unsigned char *st = smem + (size_t)s * pl.stage_bytes + (((uint64_t)(uintptr_t)gbase + ck.b0) & 15);
		{
			// general shape: one warp per row (or piece of a long row), partial last word
			const uint32_t nf = nfull > ck.w0 ? min(ck.nw, nfull - ck.w0) : 0u;   // words with 32 valid samples
			for (uint32_t rr = wid; rr < ck.nrows; rr += CLS_THREADS / 32) {
				const uint32_t lr = ck.lr0 + rr;
				// sample x of this row sits at src[x - 32*w0]
				const Sample *q = (const Sample *)st + (size_t)rr * P.NX + lane;
				uint32_t *Sr = P.S + (uint64_t)lr * P.WP + ck.w0;    // 16-byte aligned (WP, w0 multiples of 4)
				uint32_t *Zr = P.Z + (uint64_t)lr * P.WP + ck.w0;
				uint32_t zacc = 0;
				uint32_t g = 0;
				for (; g + 4 <= nf; g += 4, q += 128) {
					const Sample f0 = q[0], f1 = q[32], f2 = q[64], f3 = q[96];
					uint4 b, e;
					b.x = __ballot_sync(0xFFFFFFFFu, cls.gt(f0)); b.y = __ballot_sync(0xFFFFFFFFu, cls.gt(f1));
					b.z = __ballot_sync(0xFFFFFFFFu, cls.gt(f2)); b.w = __ballot_sync(0xFFFFFFFFu, cls.gt(f3));
					e.x = __ballot_sync(0xFFFFFFFFu, cls.eq(f0)); e.y = __ballot_sync(0xFFFFFFFFu, cls.eq(f1));
					e.z = __ballot_sync(0xFFFFFFFFu, cls.eq(f2)); e.w = __ballot_sync(0xFFFFFFFFu, cls.eq(f3));
					zacc |= e.x | e.y | e.z | e.w;
					if (lane == 0) { *reinterpret_cast<uint4 *>(Sr + g) = b; *reinterpret_cast<uint4 *>(Zr + g) = e; }
				}
				for (; g < ck.nw; g++, q += 32) {             // remaining words, the last one maybe partial
					const bool ok = g < nf || lane < tail;     // (g >= nf happens only for the row's partial last word)
					const Sample f = ok ? q[0] : (Sample)0;
					const uint32_t b = __ballot_sync(0xFFFFFFFFu, ok && cls.gt(f));
					const uint32_t e = __ballot_sync(0xFFFFFFFFu, ok && cls.eq(f));
					zacc |= e;
					if (lane == 0) { Sr[g] = b; Zr[g] = e; }
				}
				if (lane == 0 && zacc) { P.rowZ[lr] = P.zepoch; *P.anyZp = P.zepoch; }
			}
		}
		__syncthreads();                                     // every warp is done with stage s
		if (k + CLS_STAGES < nmine) issue(blockIdx.x + (k + CLS_STAGES) * gridDim.x, s);
	}
}

// ---------------------------------------------------------------------------
// K1, vector form: rows of whole NW-word groups and 16-byte aligned samples (the
// common power-of-two grids).  Same TMA ring; a lane reads 16 bytes = NS
// consecutive samples per load and turns them into NS bits, the LW = 32/NS lanes
// whose bits make up one bitmap word OR them together with log2(LW) butterfly
// shuffles, and a warp iteration covers NW words (f32: 128 samples with a single
// LDS.128).  Every warp takes a contiguous run of its chunk's iterations.
// ---------------------------------------------------------------------------
template <typename Sample>
__global__ void __launch_bounds__(CLS_THREADS) k_classify_vec(const __grid_constant__ Params P, uint32_t rows, uint32_t nchunks, uint32_t stage_bytes)
{
	extern __shared__ __align__(128) unsigned char smem[];
	__shared__ uint64_t full[CLS_STAGES];
	constexpr int NS = 16 / (int)sizeof(Sample);       // samples per 16-byte load
	constexpr int NL = NS >= 4 ? 1 : 4 / NS;            // loads per iteration (f64: 2)
	constexpr int NW = NS * NL;                         // words per iteration
	constexpr int LW = 32 / NS;                         // lanes per word
	const Cls<Sample> cls(P);
	const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const char *gbase = (const char *)P.data;
	const uint32_t rowb = P.NX * (uint32_t)sizeof(Sample), chunkb = rows * rowb;
	const uint32_t gpr = P.W / NW;                      // iterations per row
	const uint32_t ipw = (rows * gpr + CLS_THREADS / 32 - 1) / (CLS_THREADS / 32);   // iterations per warp and chunk
	const uint32_t i0 = wid * ipw, r0 = i0 / gpr, g0 = i0 - r0 * gpr;
	const unsigned sh = NS * (lane % LW);
	uint32_t *const Sb = P.S, *const Zb = P.Z;
	const uint32_t WP = P.WP, Lrows = P.Lrows;

	if (threadIdx.x == 0) {
		for (int s = 0; s < CLS_STAGES; s++) mbar_init(&full[s], 1);
		fence_mbar_init();
	}
	__syncthreads();
	auto issue = [&](uint32_t c, int s) {
		if (threadIdx.x == 0) {
			const uint32_t lr0 = c * rows;
			const uint32_t bytes = (lr0 + rows <= Lrows ? rows : Lrows - lr0) * rowb;
			mbar_arrive_expect_tx(&full[s], bytes);
			bulk_g2s(smem + (size_t)s * stage_bytes, gbase + (uint64_t)c * chunkb, bytes, &full[s]);
		}
	};
	const uint32_t nmine = blockIdx.x < nchunks ? (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
	for (uint32_t k = 0; k < CLS_STAGES && k < nmine; k++) issue(blockIdx.x + k * gridDim.x, (int)k);

	for (uint32_t k = 0; k < nmine; k++) {
		const int s = (int)(k % CLS_STAGES);
		const uint32_t lr0 = (blockIdx.x + k * gridDim.x) * rows;
		const uint32_t nit = (lr0 + rows <= Lrows ? rows : Lrows - lr0) * gpr;
		mbar_wait(&full[s], (k / CLS_STAGES) & 1);
		const uint4 *src = reinterpret_cast<const uint4 *>(smem + (size_t)s * stage_bytes) + (size_t)i0 * (32 * NL) + lane;
		uint32_t gq = g0, o = (lr0 + r0) * WP + g0 * NW;
		const uint32_t i1 = min(i0 + ipw, nit);
		for (uint32_t it = i0; it < i1; it++, src += 32 * NL) {
			uint32_t wS[NL];
			bool eqa = false;
			uint4 raw[NL];
#pragma unroll
			for (int j = 0; j < NL; j++) raw[j] = src[32 * j];
#pragma unroll
			for (int j = 0; j < NL; j++) {
				const Sample *f = reinterpret_cast<const Sample *>(&raw[j]);
				uint32_t nb = 0;
#pragma unroll
				for (int i = 0; i < NS; i++) { nb |= (uint32_t)cls.gt(f[i]) << i; eqa = eqa || cls.eq(f[i]); }
				uint32_t v = nb << sh;
#pragma unroll
				for (int d = LW / 2; d; d >>= 1) v |= __shfl_xor_sync(0xFFFFFFFFu, v, d);
				wS[j] = v;
			}
			// lane l holds word (l / LW) of load j: the first lane of each segment stores it
			if (lane % LW == 0) {
#pragma unroll
				for (int j = 0; j < NL; j++) Sb[o + j * NS + lane / LW] = wS[j];
			}
			if (__any_sync(0xFFFFFFFFu, eqa)) {
#pragma unroll
				for (int j = 0; j < NL; j++) {
					const Sample *f = reinterpret_cast<const Sample *>(&raw[j]);
					uint32_t nb = 0;
#pragma unroll
					for (int i = 0; i < NS; i++) nb |= (uint32_t)cls.eq(f[i]) << i;
					uint32_t v = nb << sh;
#pragma unroll
					for (int d = LW / 2; d; d >>= 1) v |= __shfl_xor_sync(0xFFFFFFFFu, v, d);
					if (lane % LW == 0) Zb[o + j * NS + lane / LW] = v;
				}
				if (lane == 0) { P.rowZ[o / WP] = P.zepoch; *P.anyZp = P.zepoch; }
			} else if (lane < NW) {
				// (every Z word is rewritten on every launch; tracking dirty units as k_classify_sweep
				// does measured slower here: one isovalue leaves this kernel enough slack for the zeros)
				Zb[o + lane] = 0u;
			}
			o += NW;
			if (++gq == gpr) { gq = 0; o += WP - gpr * NW; }
		}
		__syncthreads();                                     // every warp is done with stage s
		if (k + CLS_STAGES < nmine) issue(blockIdx.x + (k + CLS_STAGES) * gridDim.x, s);
	}
}

// ---------------------------------------------------------------------------
// K1, iso sweep: the samples are streamed ONCE for up to eight isovalues (same TMA ring as
// k_classify_vec; float grids with rows of whole 128-sample groups).  A lane loads samples
// lane, lane+32, lane+64, lane+96 of a group and packs its 4 x 8 comparison results into one
// word (bit 4j+i: sample 32i+lane above isovalue j); a 32 x 32 bit-matrix transpose across
// the warp (five shuffle stages) then leaves in lane 4j+i exactly bitmap word i of the group
// for isovalue j, stored to set j.  Equality (on-iso) is accumulated in one predicate; only a
// group that has a hit builds the Z words the same way.  Isovalues beyond n are +inf.
// ---------------------------------------------------------------------------
#define SWEEP_MAX 8
struct SweepSets {
	float iso[SWEEP_MAX];
	uint32_t *S, *Z, *rowZ, *any, *D; // set j at S + j * set_words, rowZ + j * Lrows, any + j, D + j * dwords
	uint64_t set_words, dwords;
};

// 32 x 32 bit-matrix transpose across the warp: out[l] bit s = in[s] bit l.  Five butterfly
// stages; a stage is one shuffle, one rotate and one bit-select (the rotate amount and the
// select mask depend on the lane only and are passed in).
struct TransposeLane { uint32_t sh[5], m[5]; };
__device__ __forceinline__ TransposeLane transpose_lane(unsigned lane)
{
	TransposeLane t;
#pragma unroll
	for (int i = 0; i < 5; i++) {
		const int k = 16 >> i;
		const uint32_t M = k == 16 ? 0xFFFF0000u : k == 8 ? 0xFF00FF00u : k == 4 ? 0xF0F0F0F0u : k == 2 ? 0xCCCCCCCCu : 0xAAAAAAAAu;
		t.sh[i] = (lane & k) ? 32u - k : (uint32_t)k;     // rotate left by k, or right by k
		t.m[i] = (lane & k) ? ~M : M;                     // bit positions taken from the partner
	}
	return t;
}
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t w, const TransposeLane &t)
{
#pragma unroll
	for (int i = 0; i < 5; i++) {
		const uint32_t x = __shfl_xor_sync(0xFFFFFFFFu, w, 16 >> i);
		const uint32_t r = __funnelshift_l(x, x, t.sh[i]);
		w = (w & ~t.m[i]) | (r & t.m[i]);
	}
	return w;
}
// all ones if a == b, for building bit fields without predicates
__device__ __forceinline__ uint32_t feq_mask(float a, float b)
{
	uint32_t d;
	asm("set.eq.u32.f32 %0, %1, %2;" : "=r"(d) : "f"(a), "f"(b));
	return d;
}

#ifndef SWEEP_THREADS
#define SWEEP_THREADS 256    // (A/B: 256 x unroll 2 measured best, 128 threads within 3 %; 512 threads spill)
#endif
#ifndef SWEEP_UNROLL
#define SWEEP_UNROLL 2
#endif
constexpr int kSweepUnroll = SWEEP_UNROLL;
__global__ void __launch_bounds__(SWEEP_THREADS, 3) k_classify_sweep(const __grid_constant__ Params P, const __grid_constant__ SweepSets ss,
                                                                uint32_t rows, uint32_t nchunks, uint32_t stage_bytes)
{
	extern __shared__ __align__(128) unsigned char smem[];
	__shared__ uint64_t full[CLS_STAGES];
	const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const char *gbase = (const char *)P.data;
	const uint32_t rowb = P.NX * 4u, chunkb = rows * rowb;
	const uint32_t gpr = P.W / 4;                       // 128-sample groups per row
	const uint32_t ipw = (rows * gpr + SWEEP_THREADS / 32 - 1) / (SWEEP_THREADS / 32);
	const uint32_t i0 = wid * ipw, r0 = i0 / gpr, g0 = i0 - r0 * gpr;
	const uint32_t WP = P.WP, Lrows = P.Lrows;
	const uint64_t lane_off = (uint64_t)(lane >> 2) * ss.set_words + (lane & 3u);     // set j = lane / 4, word i = lane % 4
	uint32_t *const Sl = ss.S + lane_off, *const Zl = ss.Z + lane_off;
	uint32_t *const Dl = ss.D + (uint64_t)(lane >> 2) * ss.dwords;
	float iso[SWEEP_MAX];
#pragma unroll
	for (int j = 0; j < SWEEP_MAX; j++) iso[j] = ss.iso[j];
	const TransposeLane tl = transpose_lane(lane);

	if (threadIdx.x == 0) {
		for (int s = 0; s < CLS_STAGES; s++) mbar_init(&full[s], 1);
		fence_mbar_init();
	}
	__syncthreads();
	auto issue = [&](uint32_t c, int s) {
		if (threadIdx.x == 0) {
			const uint32_t lr0 = c * rows;
			const uint32_t bytes = (lr0 + rows <= Lrows ? rows : Lrows - lr0) * rowb;
			mbar_arrive_expect_tx(&full[s], bytes);
			bulk_g2s(smem + (size_t)s * stage_bytes, gbase + (uint64_t)c * chunkb, bytes, &full[s]);
		}
	};
	const uint32_t nmine = blockIdx.x < nchunks ? (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
	for (uint32_t k = 0; k < CLS_STAGES && k < nmine; k++) issue(blockIdx.x + k * gridDim.x, (int)k);

	for (uint32_t k = 0; k < nmine; k++) {
		const int s = (int)(k % CLS_STAGES);
		const uint32_t lr0 = (blockIdx.x + k * gridDim.x) * rows;
		const uint32_t nit = (lr0 + rows <= Lrows ? rows : Lrows - lr0) * gpr;
		// Z is only rewritten where it has to change: a 128-sample group with an on-iso sample is
		// written and marked dirty, a clean group is zeroed only if it was dirty (writing eight
		// sets of zeros every time cost a third of the kernel).  A warp's groups of a chunk (at most
		// 32) span two dirty words, fetched before waiting for the samples; every lane tracks its own set.
		const uint32_t i1 = min(i0 + ipw, nit);
		uint32_t gi = lr0 * gpr + i0;
		const uint32_t dw0 = gi >> 5;
		uint64_t dd = 0;
		if (i0 < i1) dd = (uint64_t)Dl[dw0] | ((uint64_t)Dl[dw0 + 1] << 32);
		mbar_wait(&full[s], (k / CLS_STAGES) & 1);
		const float *src = reinterpret_cast<const float *>(smem + (size_t)s * stage_bytes) + (size_t)i0 * 128 + lane;
		uint32_t gq = g0, o = (lr0 + r0) * WP + g0 * 4;
#pragma unroll kSweepUnroll
		for (uint32_t it = i0; it < i1; it++, src += 128, gi++) {
			const bool dirty = (dd >> (gi - (dw0 << 5))) & 1u;
			const uint32_t dbit = 1u << (gi & 31u);
			const float f0 = src[0], f1 = src[32], f2 = src[64], f3 = src[96];
			// bit 4j+i = IEEE sign bit of iso_j - f_i, exactly the reference's index bit
			// (marching_cubes_33.c:392-409, :1856-1859; no flush-to-zero, so the difference is
			// zero only for equal operands): one subtraction and one funnel shift per bit,
			// highest bit first
			uint32_t w = 0;
			bool eq = false;
#pragma unroll
			for (int j = SWEEP_MAX - 1; j >= 0; j--) {
				const float d3 = __fsub_rn(iso[j], f3), d2 = __fsub_rn(iso[j], f2), d1 = __fsub_rn(iso[j], f1), d0 = __fsub_rn(iso[j], f0);
				w = __funnelshift_l(__float_as_uint(d3), w, 1);
				w = __funnelshift_l(__float_as_uint(d2), w, 1);
				w = __funnelshift_l(__float_as_uint(d1), w, 1);
				w = __funnelshift_l(__float_as_uint(d0), w, 1);
				// any exact zero among the four?  One product (on the otherwise idle FMA pipe) and one
				// test instead of four tests: a zero factor gives 0, or NaN against an infinite one,
				// and both fail |p| > 0.  An underflowing product only sends the group through the
				// exact on-iso path below for nothing.
				eq = eq || !(fabsf(__fmul_rn(__fmul_rn(d0, d1), __fmul_rn(d2, d3))) > 0.0f);
			}
			Sl[o] = warp_transpose32(w, tl);
			if (__any_sync(0xFFFFFFFFu, eq)) {
				uint32_t e = 0;
#pragma unroll
				for (int j = 0; j < SWEEP_MAX; j++) {
					e |= feq_mask(f0, iso[j]) & (1u << (4 * j));
					e |= feq_mask(f1, iso[j]) & (2u << (4 * j));
					e |= feq_mask(f2, iso[j]) & (4u << (4 * j));
					e |= feq_mask(f3, iso[j]) & (8u << (4 * j));
				}
				e = warp_transpose32(e, tl);
				// the four lanes of a set agree on whether the set has a hit in this group
				const bool hit = ((__ballot_sync(0xFFFFFFFFu, e != 0u) >> (lane & ~3u)) & 0xFu) != 0u;
				if (hit) {
					Zl[o] = e;
					if (e) { ss.rowZ[(uint64_t)(lane >> 2) * Lrows + o / WP] = P.zepoch; ss.any[lane >> 2] = P.zepoch; }
					if (!dirty && (lane & 3u) == 0) atomicOr(&Dl[gi >> 5], dbit);
				} else if (dirty) {
					Zl[o] = 0u;
					if ((lane & 3u) == 0) atomicAnd(&Dl[gi >> 5], ~dbit);
				}
			} else if (dirty) {
				Zl[o] = 0u;
				if ((lane & 3u) == 0) atomicAnd(&Dl[gi >> 5], ~dbit);
			}
			o += 4;
			if (++gq == gpr) { gq = 0; o += WP - gpr * 4; }
		}
		__syncthreads();                                     // every warp is done with stage s
		if (k + CLS_STAGES < nmine) issue(blockIdx.x + (k + CLS_STAGES) * gridDim.x, s);
	}
}

// ---------------------------------------------------------------------------
// block-wide exclusive scan of two 64-bit counters (256 threads)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void block_exscan2(uint64_t &a, uint64_t &b, uint64_t &ta, uint64_t &tb,
                                              uint64_t (*sw)[8])
{
	const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	uint64_t ia = a, ib = b;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		uint64_t xa = __shfl_up_sync(0xFFFFFFFFu, ia, d), xb = __shfl_up_sync(0xFFFFFFFFu, ib, d);
		if (lane >= (unsigned)d) { ia += xa; ib += xb; }
	}
	if (lane == 31) { sw[0][wid] = ia; sw[1][wid] = ib; }
	__syncthreads();
	uint64_t oa = 0, ob = 0, sa = 0, sb = 0;
#pragma unroll
	for (int k = 0; k < 8; k++) {
		uint64_t va = sw[0][k], vb = sw[1][k];
		if ((unsigned)k < wid) { oa += va; ob += vb; }
		sa += va; sb += vb;
	}
	__syncthreads();
	a = oa + ia - a; b = ob + ib - b;   // exclusive
	ta = sa; tb = sb;
}

__device__ __forceinline__ uint32_t sumV(uint64_t p) { return fldV(p, 0) + fldV(p, 1) + fldV(p, 2); }

// Does any row that the G rows starting at local row lr0 depend on hold an on-iso
// sample?  (rows y..y+G+1 of slices z..z+2: the cells' corner rows and the rows
// their vertex masks are built from.)  Warp-uniform; decides between the quad
// fast paths and the generic bitmap walk for the whole group.
__device__ __forceinline__ bool group_has_oniso(const Params &P, bool any, uint32_t lr0, unsigned lane)
{
	if (!any) return false;
	bool f = false;
	const uint32_t n = P.G + 2;
	for (uint32_t i = lane; i < 3 * n; i += 32) {
		const uint32_t dz = i / n, dy = i - dz * n;
		const uint64_t row = (uint64_t)lr0 + dy + (uint64_t)dz * P.NY;
		if (row < P.Lrows) f = f || P.rowZ[row] == P.zepoch;
	}
	return __any_sync(0xFFFFFFFFu, f);
}

// lane -> (row within the warp's group, quad) of pass `pass`
__device__ __forceinline__ void lane_item(const Params &P, unsigned lane, uint32_t pass, uint32_t &r, uint32_t &q)
{
	if (P.Q <= 32) { r = fastdiv(lane, P.Q, P.mQ); q = lane - r * P.Q; }
	else { r = 0; q = pass * 32 + lane; }
}

// ---------------------------------------------------------------------------
// K2: count.  Warp-synchronous: a warp takes G whole rows, one lane per quad
// (four bitmap words = one 16-byte load per neighbouring row), so the row-local
// word prefixes come out of one shuffle scan with no block barrier; rows longer
// than 128 words are walked in passes of 32 quads with a carry.  The per-row
// totals of the CTA's 8G rows are then scanned once per CTA, leaving CTA-relative
// row bases and one (V, T, C) sum per CTA for k_rowscan.
// ---------------------------------------------------------------------------
#define CNT_WARPS 8

#ifndef CNT_MINB
#define CNT_MINB 3     // resident CTAs per SM k_count is compiled for: 80 registers (A/B on the gyroid: 0.061 ms at 2, 0.064 at 3, 0.073 at 4, where the hot loop spills; white noise wants the warps: 5.5 ms at 2, 3.9 at 4)
#endif
// (integer grids are CT-like in practice: many complex / on-iso cells to walk, which wants resident warps more
// than registers -- 1024^3 u16: 1.31 ms at 4 CTAs per SM against 1.37 at 3)
template <typename Sample> struct CountMinBlocks { enum { value = CNT_MINB }; };
template <> struct CountMinBlocks<uint8_t> { enum { value = 4 }; };
template <> struct CountMinBlocks<uint16_t> { enum { value = 4 }; };
template <> struct CountMinBlocks<uint32_t> { enum { value = 4 }; };

template <typename Sample>
__global__ void __launch_bounds__(256, CountMinBlocks<Sample>::value) k_count(const __grid_constant__ Params P, uint32_t nblk, uint32_t GW, uint32_t *blkSum)
{
	__shared__ uint32_t s_row[3][CNT_WARPS * 32];
	__shared__ uint32_t s_cx[2][CNT_WARPS * 32];     // triangles / centres of the complex cells, per row (added late)
	__shared__ uint64_t s_w[2][8];
	// (the case tables are only touched by the complex-cell walk: read in place, no copy per CTA)
	const Tables tb = global_tables();
	const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const bool anyz = *P.anyZp == P.zepoch;     // (tagged with the classify epoch: never has to be cleared)
	const uint32_t RB = CNT_WARPS * GW * P.G;       // rows per block: GW groups of G rows per warp
	const uint32_t npass = (P.Q + 31) / 32;

	// (one block per CTA when the grid allows: the hardware hands CTAs out as slots free up, which
	// measured better than persistent CTAs pulling blocks by ticket, and than smaller blocks)
	for (uint32_t blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
		s_row[0][threadIdx.x] = 0; s_row[1][threadIdx.x] = 0; s_row[2][threadIdx.x] = 0;
		s_cx[0][threadIdx.x] = 0; s_cx[1][threadIdx.x] = 0;
		__syncthreads();
	  for (uint32_t sub = 0; sub < GW; sub++) {
		const uint32_t srow = (wid * GW + sub) * P.G;      // first row of the group within the block
		const uint32_t row0 = blk * RB + srow;
		if (row0 >= P.Lrows) break;
		const bool gz = group_has_oniso(P, anyz, row0, lane);
		uint64_t carryV = 0, carryT = 0;
		for (uint32_t pass = 0; pass < npass; pass++) {
			uint32_t r, q;
			lane_item(P, lane, pass, r, q);
			const uint32_t lr = row0 + r;
			const bool valid = r < P.G && q < P.Q && lr < P.Lrows;
			uint64_t pv0 = 0, pv1 = 0, pv2 = 0, pv3 = 0, tt = 0;
			bool has_cx = false;
			uint32_t slow_q = 0;
			if (valid) {
				const uint32_t zl = fastdiv(lr, P.NY, P.mNY), y = lr - zl * P.NY, z = zl + P.zlo;
				const bool own_p = row_points_owned(P, z) || row_points_halo(P, z);
				const bool own_c = row_cells_owned(P, z, y);
				if (own_p || own_c) {
					const bool hasY = y < P.ny, hasZ = z < P.nz;
					const uint64_t dY = hasY ? P.WP : 0u, dZ = hasZ ? (uint64_t)P.NY * P.WP : 0u;
					const uint64_t i00 = (uint64_t)lr * P.WP + 4 * q;
					const Quad q00 = load_quad(P.S, i00), q10 = load_quad(P.S, i00 + dY);
					const Quad q01 = load_quad(P.S, i00 + dZ), q11 = load_quad(P.S, i00 + dY + dZ);
					// groups near an on-iso sample: the words it touches take the generic walk
					const uint32_t slow = gz ? quad_oniso_mask(P.Z, i00, dY, dZ) : 0u;
					uint64_t pv[4];
					uint32_t act[4], vis[4];
					uint32_t nts = 0;
					const uint32_t pm = row_points_owned(P, z) ? 0xFFFFFFFFu : 0u;
#pragma unroll
					for (int k = 0; k < 4; k++) {
						act[k] = 0;
						if (!((slow >> k) & 1u)) {
							WordRec rec;
							uint32_t c[8];
							quad_word(P, q00, q10, q01, q11, k, 4 * q + k, own_c && hasZ, rec, c);
							vis[k] = rec.act | ((rec.X | rec.Y | rec.Z) & pm);
							if (!own_p) { rec.X = rec.Y = rec.Z = 0; }
							pv[k] = pack_planes(rec);
							// simple cells are counted 32 at a time; only the complex ones are walked
							if (rec.act) nts += count_simple_cells(c, rec.act, act[k]);
						} else {
							pv[k] = 0; vis[k] = 0;
						}
					}
					if (slow) {
						const QuadSlow r = count_quad_slow<Sample>(P, tb, z, y, q, slow, own_p, own_c);
#pragma unroll
						for (int k = 0; k < 4; k++)
							if ((slow >> k) & 1u) { pv[k] = r.pv[k]; vis[k] = r.vis[k]; }
						tt += r.cc;
					}
					pv0 = pv[0]; pv1 = pv[1]; pv2 = pv[2]; pv3 = pv[3];
					*reinterpret_cast<uint4 *>(P.A + i00) = make_uint4(vis[0], vis[1], vis[2], vis[3]);
					tt += nts;
					// complex cells: walked at the end of the pass (their triangles only enter the row
					// totals), so that the call does not sit in the middle of everything that is live here
					has_cx = (act[0] | act[1] | act[2] | act[3]) != 0;
					slow_q = slow;
				}
			}
			// lane-local exclusive prefix over the four words, then the warp scan
			const uint64_t e1 = pv0, e2 = e1 + pv1, e3 = e2 + pv2, tv = e3 + pv3;
			uint64_t lv, it, rst = 0, iv = 0;
			if (P.Q <= 32) {
				// whole rows in one pass: a warp's 32 quads hold at most 4096 vertices per plane, 4096
				// centres and 49152 triangles, so the five counters scan as three 32-bit words
				// (X | Y << 16, Z | C << 16, T) instead of two 64-bit ones
				const uint32_t a0 = fldV(tv, 0) | (fldV(tv, 1) << 16), b0 = fldV(tv, 2) | ((uint32_t)(tt >> 32) << 16), c0 = (uint32_t)tt;
				uint32_t ia = a0, ib = b0, ic = c0;
#pragma unroll
				for (int d = 1; d < 32; d <<= 1) {
					const uint32_t xa = __shfl_up_sync(0xFFFFFFFFu, ia, d), xb = __shfl_up_sync(0xFFFFFFFFu, ib, d),
					               xc = __shfl_up_sync(0xFFFFFFFFu, ic, d);
					if (lane >= (unsigned)d) { ia += xa; ib += xb; ic += xc; }
				}
				// exclusive, relative to the first quad of the lane's row
				const int srcl = (int)min(r * P.Q, 31u);
				const uint32_t ea = ia - a0, eb = ib - b0, ec = ic - c0;
				const uint32_t ra = ea - __shfl_sync(0xFFFFFFFFu, ea, srcl), rb = eb - __shfl_sync(0xFFFFFFFFu, eb, srcl);
				const uint32_t rc0 = __shfl_sync(0xFFFFFFFFu, ec, srcl);
				lv = (uint64_t)(ra & 0xFFFFu) | ((uint64_t)(ra >> 16) << 21) | ((uint64_t)(rb & 0xFFFFu) << 42);
				// fold the plane offsets in: Y ids follow the row's X ids, Z ids follow both
				const int lastl = (int)min(r * P.Q + P.Q - 1, 31u);
				lv += plane_offsets(__shfl_sync(0xFFFFFFFFu, lv + tv, lastl));
				// row totals so far (inclusive of this quad): triangles, centres
				it = (uint64_t)(ic - rc0) | ((uint64_t)((eb >> 16) + (b0 >> 16) - (__shfl_sync(0xFFFFFFFFu, eb, srcl) >> 16)) << 32);
			} else {
				uint64_t jt = tt;
				iv = tv;
#pragma unroll
				for (int d = 1; d < 32; d <<= 1) {
					const uint64_t xv = __shfl_up_sync(0xFFFFFFFFu, iv, d), xt = __shfl_up_sync(0xFFFFFFFFu, jt, d);
					if (lane >= (unsigned)d) { iv += xv; jt += xt; }
				}
				lv = carryV + iv - tv;                           // row-local prefix in front of this quad
				it = jt;
			}
			if (valid) {
				uint64_t *pw = P.wpreV + (uint64_t)lr * P.WP + 4 * q;
				*reinterpret_cast<ulonglong2 *>(pw) = make_ulonglong2(lv, lv + e1);
				*reinterpret_cast<ulonglong2 *>(pw + 2) = make_ulonglong2(lv + e2, lv + e3);
				if (q == P.Q - 1) {
					const uint64_t rowT = P.Q <= 32 ? it : carryT + it - rst;
					pw[4] = lv + tv;
					s_row[0][srow + r] = P.Q <= 32 ? fldV(lv + tv, 2) : sumV(lv + tv);
					s_row[1][srow + r] = (uint32_t)rowT;
					s_row[2][srow + r] = (uint32_t)(rowT >> 32);
				}
			}
			if (P.Q > 32) { carryV += __shfl_sync(0xFFFFFFFFu, iv, 31); carryT += __shfl_sync(0xFFFFFFFFu, it, 31); }
			if (has_cx) {
				const uint64_t cc = count_quad_complex<Sample>(P, tb, lr, q, slow_q);
				atomicAdd(&s_cx[0][srow + r], (uint32_t)cc);
				atomicAdd(&s_cx[1][srow + r], (uint32_t)(cc >> 32));
			}
		}
		if (P.Q > 32 && row0 < P.Lrows) {
			// long rows: the plane totals are only known now; add the offsets in a second sweep
			const uint64_t add = plane_offsets(carryV);
			for (uint32_t i = lane; i <= 4 * P.Q; i += 32) P.wpreV[(uint64_t)row0 * P.WP + i] += add;
		}
	  }
		__syncthreads();
		// CTA-relative row bases
		{
			const uint32_t t = threadIdx.x;
			uint64_t a = (uint64_t)s_row[0][t] | ((uint64_t)(s_row[2][t] + s_cx[1][t]) << 32), b = s_row[1][t] + s_cx[0][t], ta, tb2;
			block_exscan2(a, b, ta, tb2, s_w);
			const uint32_t lr = blk * RB + t;
			if (t < RB && lr < P.Lrows) {
				P.rowBV[lr] = (uint32_t)a; P.rowBC[lr] = (uint32_t)(a >> 32); P.rowBT[lr] = (uint32_t)b;
			}
			if (t == 0) { blkSum[blk] = (uint32_t)ta; blkSum[nblk + blk] = (uint32_t)tb2; blkSum[2 * nblk + blk] = (uint32_t)(ta >> 32); }
		}
	}
}

// ---------------------------------------------------------------------------
// K2b: row scan.  Thread t of CTA j owns k_count block 256j + t: the CTA first
// sums every earlier block (at most a few thousand values), scans its own 256,
// and turns the CTA-relative row bases into slab-local ones: the implicit running
// M->nV++ / nT++ of the reference (marching_cubes_33.c:487, :1245).
// ---------------------------------------------------------------------------
#ifndef RS_BLOCKS
#define RS_BLOCKS 4      // k_count blocks per k_rowscan CTA
#endif

__global__ void __launch_bounds__(256) k_rowscan(const __grid_constant__ Params P, uint32_t nblk, uint32_t RB, const uint32_t *blkSum, uint32_t owned_end_row)
{
	__shared__ uint64_t s_w[2][8];
	__shared__ uint64_t s_base[3][RS_BLOCKS];
	const uint32_t b0 = blockIdx.x * RS_BLOCKS, me = b0 + threadIdx.x;
	const bool mine = threadIdx.x < RS_BLOCKS && me < nblk;
	uint64_t aV = 0, aT = 0, aC = 0;
	for (uint32_t i = threadIdx.x; i < b0; i += 256) { aV += blkSum[i]; aT += blkSum[nblk + i]; aC += blkSum[2 * nblk + i]; }
	uint64_t mV = mine ? blkSum[me] : 0, mT = mine ? blkSum[nblk + me] : 0, mC = mine ? blkSum[2 * nblk + me] : 0;
	uint64_t preV, preT, preC, totV, totT, totC;
	block_exscan2(aV, aT, preV, preT, s_w);              // totals = sums over all earlier blocks
	block_exscan2(aC, mV, preC, totV, s_w);              // mV becomes this block's exclusive prefix within the CTA
	block_exscan2(mT, mC, totT, totC, s_w);
	if (threadIdx.x < RS_BLOCKS) {
		s_base[0][threadIdx.x] = preV + mV; s_base[1][threadIdx.x] = preT + mT; s_base[2][threadIdx.x] = preC + mC;
	}
	__syncthreads();
	const uint64_t rows0 = (uint64_t)b0 * RB;
	for (uint32_t i = threadIdx.x; i < RS_BLOCKS * RB; i += 256) {
		const uint64_t lr = rows0 + i;
		if (lr >= P.Lrows) break;
		const uint32_t bl = i / RB;
		const uint32_t v = P.rowBV[lr] + (uint32_t)s_base[0][bl];
		P.rowBV[lr] = v;
		P.rowBT[lr] += (uint32_t)s_base[1][bl];
		P.rowBC[lr] += (uint32_t)s_base[2][bl];
		if (lr == owned_end_row) P.totals->nShared = v;
	}
	if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
		const uint64_t tv = preV + totV, tt = preT + totT, tc = preC + totC;
		P.rowBV[P.Lrows] = (uint32_t)tv; P.rowBT[P.Lrows] = (uint32_t)tt; P.rowBC[P.Lrows] = (uint32_t)tc;
		if (owned_end_row >= P.Lrows) P.totals->nShared = (uint32_t)tv;
		P.totals->nCentre = (uint32_t)tc;
		P.totals->nT = (uint32_t)tt;
		P.totals->nSharedAll = (uint32_t)tv;
		// 32-bit index range check (include/marching_cubes_33.h:140 uses unsigned int)
		P.totals->range = (tv + tc >= 0xFFFFFFFFull || tt >= 0xFFFFFFFFull) ? 1u : 0u;
	}
}

// ---------------------------------------------------------------------------
// K3 / K4: emit.
//
// cells     (K4, k_emit_cells) warp-synchronous, one warp per group of G rows, no block
//           barrier.  Fill: the active cells -- and the grid points that own a vertex --
//           are compacted in sweep order (shuffle scan of the popcounts).  Drain: one lane
//           per visited CELL gets the ids of its 12 edge vertices from bitmaps + prefixes,
//           leaves a task for each vertex its low corner point owns, picks the MC33
//           pattern, and after a shuffle scan of the triangle counts the triangles (and
//           centre vertices) are written at the ids the reference's sweep order gives them.
// vertices  (K3, k_emit_vertices) dense over the vertex ids: thread id runs the task left
//           in slot id, so consecutive threads write consecutive V / N / color.
// Groups with more cells than a queue holds take several windows.
// ---------------------------------------------------------------------------
#define EM_WARPS 8
#define CQ 256
#ifndef EMC_TRI_STAGED
#define EMC_TRI_STAGED 0     // 1: every cell lane writes its own triangles into a shared-memory window that goes out as whole words.
                            // Measured slower (cfg2 0.184 against 0.175 ms, cfg5 25.0 against 20.6): the per-lane pattern walk runs
                            // as long as the busiest lane of the warp.  Kept for A/B.
#endif
#define EM_TW 96         // triangles per staging window
#if EMC_TRI_STAGED
#define EM_SCR (13 * 32 + 3 * EM_TW)      // per-warp scratch words: 13 ids x 32 lanes + the triangle window
#else
#define EM_SCR (13 * 32 + 96)             // per-warp scratch words: 13 ids x 32 lanes + one owner byte per triangle of a round (<= 32 x 12)
#endif
#ifndef EMV_MINB
#define EMV_MINB 5      // resident CTAs per SM the emit kernels are compiled for (register cap)
#endif
#ifndef EMC_MINB
#define EMC_MINB 4
#endif
#ifndef EMC_REVERSE
#define EMC_REVERSE 1
#endif
#ifndef EMC_V2
#define EMC_V2 1        // 1: per-row values of the cell kernel (row offsets, id bases, ownership) come from a table built once per
                        // group, the pattern from the packed case table, the capacity check is hoisted out of the triangle loop
#endif
// per-warp row table words for groups of G rows (G <= 32): RowT / (y, z) of the group's rows.  Sized by the grid's G at
// launch: shared memory the kernel does not take stays L1 cache -- for this kernel and for whatever runs next to it.
#if EMC_V2
#define EM_ROWT(G) (12u * (G))
#else
#define EM_ROWT(G) 64u
#endif
#define EMC_SMEM(G) (TBL_BYTES + EM_WARPS * (CQ + EM_SCR + EM_ROWT(G)) * 4)

// K3, dense: thread id computes vertex id from the task the cell kernel left in the
// vertex's own slot (mc33_core.cuh "Vertex tasks"): no scan, no queue, consecutive
// threads write consecutive V / N / color.
// What bounds it (round 2, ncu): DRAM traffic in scattered 32-byte sectors -- the ten samples of a vertex arrive as
// 8.6 sectors, the surface touches about half of the grid's sectors (275 MB read for 126 MB written at cfg2: 4.6 TB/s,
// 70 % of the copy bandwidth).  Measured and dropped: 16-byte stores of four lanes' components after a shuffle (0.098
// against 0.087 ms), V / N through a shared-memory transpose (0.093), aligned 4-sample loads instead of scalar gathers
// (0.109: a 128-bit load of scattered lanes costs as many L1 wavefronts as the two loads it replaces, and the planes'
// code paths diverge), a task-free form that rebuilds the plane masks per row group (0.18), a tile order of the ids for
// large slices (slower).  Kept: streaming stores / loads for what is touched once (mc33_core.cuh, MC33_STREAM_STORES),
// which keeps the sample lines of the next slice in L2 -- 36 % off on 2048 x 2048 slices.
template <typename Sample, bool KEYS>
__global__ void __launch_bounds__(256, EMV_MINB) k_emit_vertices(const __grid_constant__ Params P)
{
	const uint32_t nS = P.totals->nShared, n = min(nS, P.capV);
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		if (nS > P.capV) P.totals->overflow = 1;
		P.totals->ticket = 0;                  // re-arm the cell kernel's group counter (it has finished: stream order)
	}
	for (uint32_t id = blockIdx.x * 256u + threadIdx.x; id < n; id += gridDim.x * 256u) run_vertex_task<Sample, KEYS>(P, id);
}

template <typename Sample, bool KEYS>
__global__ void __launch_bounds__(256, EMC_MINB) k_emit_cells(const __grid_constant__ Params P, uint32_t row_begin, uint32_t row_end, uint32_t ngroups,
                                                                uint32_t ncoarse, uint32_t gfine, uint32_t pick_nquads)
{
	// (round 2: launched next to k_emit_cells2; the count kernel's on-iso statistic decides which of the two runs)
	if (pick_nquads && cells_pick_records(P, pick_nquads)) return;
	extern __shared__ __align__(128) unsigned char smem[];
	const Tables tb = load_tables(smem);
	const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	uint32_t *cq = (uint32_t *)(smem + TBL_BYTES) + wid * CQ;
	uint32_t *scr = (uint32_t *)(smem + TBL_BYTES) + EM_WARPS * CQ + wid * EM_SCR;
#if EMC_V2
	RowT *rowt = reinterpret_cast<RowT *>((uint32_t *)(smem + TBL_BYTES) + EM_WARPS * (CQ + EM_SCR) + wid * EM_ROWT(P.G));
	const bool swap_all = P.geom.normal_neg != 0;
#else
	uint2 *rowt = reinterpret_cast<uint2 *>((uint32_t *)(smem + TBL_BYTES) + EM_WARPS * (CQ + EM_SCR) + wid * EM_ROWT(P.G));
#endif
	const bool anyz = *P.anyZp == P.zepoch;     // (tagged with the classify epoch: never has to be cleared)
	const uint32_t nShared = P.totals->nShared;
	const uint32_t vb = P.dbases ? P.dbases[0] : P.vbase;
	const uint32_t vbn = (P.dbases ? P.dbases[1] : P.vbase_next) - nShared;   // halo slice: ids of the next slab
	const uint32_t npass = (P.Q + 31) / 32;

	// Row groups are handed out by a ticket counter: the work of a group follows the surface,
	// and with a fixed assignment the slowest warp set the kernel time (a quarter of the warp
	// slots sat empty on the gyroid).  The next ticket is fetched while the current group runs.
	// (every warp's first unit is its own index: no burst of same-address atomics at the start)
	const uint32_t nwarps = gridDim.x * EM_WARPS;
	uint32_t g = blockIdx.x * EM_WARPS + wid, gnext = 0;
	// The first ncoarse tickets are whole groups of G rows; the rows after them are handed
	// out gfine at a time, so that the warps which finish early at the end of the kernel still
	// find work and the last units are short (a warp takes ~20 us over a whole group).
	for (; g < ngroups; g = __shfl_sync(0xFFFFFFFFu, gnext, 0)) {
		if (lane == 0) gnext = nwarps + atomicAdd(&P.totals->ticket, 1u);
#if EMC_REVERSE
		// whole groups are taken from the high rows down: the count kernel wrote A / wpreV / row bases from the low rows
		// up, so its last rows are the ones still in L2 when this kernel starts
		const uint32_t lr0 = row_begin + (g < ncoarse ? (ncoarse - 1u - g) * P.G : ncoarse * P.G + (g - ncoarse) * gfine);
#else
		const uint32_t lr0 = row_begin + (g < ncoarse ? g * P.G : ncoarse * P.G + (g - ncoarse) * gfine);
#endif
		const uint32_t lrE = min(lr0 + (g < ncoarse ? P.G : gfine), row_end);
		const bool gz = group_has_oniso(P, anyz, lr0, lane);
#if EMC_V2
		__syncwarp();                               // (the previous group's readers are done)
		if (lane < P.G && lr0 + lane < lrE) rowt[lane] = make_rowt(P, lr0 + lane, vb, vbn);
#else
		if (lane < P.G) {
			const uint32_t lr = lr0 + lane, zl = fastdiv(lr, P.NY, P.mNY);
			rowt[lane] = make_uint2(lr - zl * P.NY, zl + P.zlo);
		}
#endif
		__syncwarp();
		// ======================= cells =======================
		const uint32_t tbase = P.rowBT[lr0], cloc0 = nShared + P.rowBC[lr0];
		uint32_t runT = 0, runC = 0;                             // triangles / centres of the group so far
		for (uint32_t pass = 0; pass < npass; pass++) {
			uint32_t r, q;
			lane_item(P, lane, pass, r, q);
			uint32_t act0 = 0, act1 = 0, act2 = 0, act3 = 0;
			{
				const uint32_t lr = lr0 + r;
				if (r < P.G && q < P.Q && lr < lrE) {
					// visited: active cells, and grid points that own a vertex (on the high faces of
					// the grid those are points without a cell); the count kernel left the mask,
					// rows this slab neither has cells nor emits vertices for stay zero
					const uint4 a = *reinterpret_cast<const uint4 *>(P.A + (uint64_t)lr * P.WP + 4 * q);
					act0 = a.x; act1 = a.y; act2 = a.z; act3 = a.w;
				}
			}
			const uint32_t na = (uint32_t)(__popc(act0) + __popc(act1) + __popc(act2) + __popc(act3));
			uint32_t ia = na;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const uint32_t x = __shfl_up_sync(0xFFFFFFFFu, ia, d);
				if (lane >= (unsigned)d) ia += x;
			}
			const uint32_t pos0 = ia - na, ncp = __shfl_sync(0xFFFFFFFFu, ia, 31);
			for (uint32_t cw0 = 0; cw0 < ncp; cw0 += CQ) {
				// fill: cells cw0 .. cw0+CQ-1 of this pass, in sweep order
				if (na && pos0 < cw0 + CQ && pos0 + na > cw0) {
					uint32_t slot = pos0 - cw0;
#pragma unroll
					for (int k = 0; k < 4; k++) {
						uint32_t m = k == 0 ? act0 : (k == 1 ? act1 : (k == 2 ? act2 : act3));
						while (m) {
							const int b = __ffs((int)m) - 1;
							m &= m - 1;
							if (slot < CQ) cq[slot] = (((4 * q + k) << 5) + b) | (r << 16);
							slot++;
						}
					}
				}
				__syncwarp();
				const uint32_t ncw = min((uint32_t)CQ, ncp - cw0);
				for (uint32_t j0 = 0; j0 < ncw; j0 += 32) {
					const bool on = j0 + lane < ncw;
#if EMC_V2
					// (scalars, not a CellPattern: the struct went through local memory)
					unsigned zm = 0;
					uint32_t x = 0, y = 0, z = 0, pstart = 0, pm = 0, pntri = 0, pcentre = 0;
					if (on) {
						const uint32_t e = cq[j0 + lane];
						x = e & 0xFFFFu;
						const RowT rt = rowt[e >> 16];
						y = rt.y; z = rt.z;
						const bool ownp = (rt.flags & 1u) != 0u, cellok = (rt.flags & 2u) != 0u && x < P.nx;
						// groups near an on-iso sample: only the words it touches take the generic rules
						if (!(gz && word_oniso(P, z, y, x >> 5))) {
							uint32_t id[12];
							unsigned own;
							const unsigned idx = cell_fast_rt(P, rt, x, id, own);
#pragma unroll
							for (int k = 0; k < 12; k++) scr[k * 32 + lane] = id[k];
							if (cellok && idx != 0u && idx != 255u) {
								const uint32_t ci = tb.cinfo[idx];      // simple256 entry | winding flag << 16
								pm = (ci >> 16) & 1u;
								if ((ci & 0xFFFFu) != 0xFFFFu) {
									pstart = ci & 0xFFFu; pntri = (ci >> 12) & 15u;
								} else if (P.pcache) {
									// complex cell: the count kernel of round 2 ran the MC33 tests and kept the pattern
									pstart = P.pcache[(uint64_t)(lr0 + (e >> 16)) * (P.WP * 32u) + x];
									const unsigned pi = tb.pat[pstart];
									pntri = pi & 0x7Fu; pcentre = pi >> 7;
								} else {
									const CellPattern cp = cell_pattern_slow<Sample>(P, tb, x, y, z, idx, 0u);
									pstart = cp.start; pm = cp.m; pntri = cp.ntri; pcentre = cp.centre;
								}
							}
							if (ownp) {
								const uint32_t lr = lr0 + (e >> 16);
								if ((own & 1u) && x < P.nx) put_vertex_task(P, id[8] - rt.g0, lr, x, 0u, false);
								if (own & 2u) put_vertex_task(P, id[0] - rt.g0, lr, x, 1u, false);
								if (own & 4u) put_vertex_task(P, id[3] - rt.g0, lr, x, 2u, false);
							}
						} else {
							const CellPattern cp = cell_slow<Sample>(P, tb, x, y, z, ownp, cellok, scr + lane, 32, zm);
							pstart = cp.start; pm = cp.m; pntri = cp.ntri; pcentre = cp.centre;
						}
					}
					CellPattern pat;
					pat.start = pstart; pat.m = pm; pat.ntri = pntri; pat.centre = pcentre;
#else
					CellPattern pat;
					pat.start = 0; pat.m = 0; pat.ntri = 0; pat.centre = 0;
					unsigned zm = 0, b = 0;
					uint32_t x = 0, y = 0, z = 0;
					bool cellok = false;
					if (on) {
						const uint32_t e = cq[j0 + lane];
						x = e & 0xFFFFu; b = x & 31u;
						const uint32_t lr = lr0 + (e >> 16);
						const uint2 yz = rowt[e >> 16];
						y = yz.x; z = yz.y;
						const bool ownp = row_points_owned(P, z);
						cellok = row_cells_owned(P, z, y) && x < P.nx;
						// groups near an on-iso sample: only the words it touches take the generic rules
						if (!(gz && word_oniso(P, z, y, x >> 5))) {
							uint32_t id[12];
							unsigned own;
							const uint32_t g0 = z == P.hz ? vbn : vb;
							const unsigned idx = cell_fast(P, x, y, z, g0, z + 1 == P.hz ? vbn : vb, id, own);
#pragma unroll
							for (int k = 0; k < 12; k++) scr[k * 32 + lane] = id[k];
							if (cellok) {
								if (P.pcache && tb.simple256[idx] == 0xFFFFu && idx != 0u && idx != 255u) {
									// complex cell: the count kernel of round 2 ran the MC33 tests and kept the pattern
									pat.start = P.pcache[(uint64_t)lr * (P.WP * 32u) + x];
									pat.m = (tb.case256[idx] >> 11) & 1u;
									const unsigned pi = tb.pat[pat.start];
									pat.ntri = pi & 0x7Fu; pat.centre = pi >> 7;
								} else {
									pat = cell_pattern<Sample>(P, tb, x, y, z, idx, 0u);
								}
							}
							if (ownp) {
								if (own & 1u) put_vertex_task(P, id[8] - g0, lr, x, 0u, false);
								if (own & 2u) put_vertex_task(P, id[0] - g0, lr, x, 1u, false);
								if (own & 4u) put_vertex_task(P, id[3] - g0, lr, x, 2u, false);
							}
						} else {
							pat = cell_slow<Sample>(P, tb, x, y, z, ownp, cellok, scr + lane, 32, zm);
						}
					}
#endif
					// triangle / centre offsets: shuffle scan in sweep order
					const uint32_t v = pat.ntri | (pat.centre << 16);
					uint32_t iv = v;
#pragma unroll
					for (int d = 1; d < 32; d <<= 1) {
						const uint32_t xs = __shfl_up_sync(0xFFFFFFFFu, iv, d);
						if (lane >= (unsigned)d) iv += xs;
					}
					const uint32_t ex = iv - v, tot = __shfl_sync(0xFFFFFFFFu, iv, 31);
					const uint32_t tid = tbase + runT + (ex & 0xFFFFu), cl = cloc0 + runC + (ex >> 16);
					const uint64_t cell = ((uint64_t)z * P.ny + y) * P.nx + x;
					if (on && pat.centre) {
						if (cl < P.capV) {
							emit_centre_vertex<Sample>(P, x, y, z, cl);
							if (KEYS && P.vkey) P.vkey[cl] = cell * 4 + 3;
						} else {
							P.totals->overflow = 1;
						}
					}
#if EMC_TRI_STAGED
					{
						// Triangles of this round.  Every cell lane writes its own triangles (pattern walk, ids from its own
						// column of the scratch, winding) into a shared-memory window at their positions of the round, then the
						// warp copies the window out as consecutive words: whole-sector stores, and no search for the owner cell
						// of a triangle (round 1: one lane per triangle with a byte map of the owners, 16 % of the kernel's
						// instructions).  Cells with an on-iso corner write their own triangles afterwards, over their slots.
						const uint32_t e0 = ex & 0xFFFFu, ntot = tot & 0xFFFFu;
						uint32_t *tst = scr + 13 * 32;
						if (on) scr[12 * 32 + lane] = vb + cl;
						const bool swp = (pat.m != 0u) != (P.geom.normal_neg != 0);
						for (uint32_t t0 = 0; t0 < ntot; t0 += EM_TW) {
							if (on && !zm && e0 < t0 + EM_TW && e0 + pat.ntri > t0) {
								const uint32_t jlo = t0 > e0 ? t0 - e0 : 0u, jhi = min(pat.ntri, t0 + EM_TW - e0);
								for (uint32_t j = jlo; j < jhi; j++) {
									const unsigned tw = tb.tri[pat.start + j];
									// winding: marching_cubes_33.c:1246-1250 (nibble 2, nibble 1, nibble 0; m swaps the first two)
									const uint32_t i0 = scr[((tw >> 8) & 15u) * 32 + lane], i1 = scr[((tw >> 4) & 15u) * 32 + lane];
									uint32_t *d = tst + 3 * (e0 + j - t0);
									d[0] = swp ? i0 : i1; d[1] = swp ? i1 : i0; d[2] = scr[(tw & 15u) * 32 + lane];
									if (KEYS && P.tcell) { const uint32_t tj = tbase + runT + e0 + j; if (tj < P.capT) P.tcell[tj] = cell; }
								}
							}
							__syncwarp();
							const uint32_t nw = 3u * min((uint32_t)EM_TW, ntot - t0);
							const uint64_t wbase = (uint64_t)(tbase + runT + t0) * 3u, wcap = (uint64_t)P.capT * 3u;
							for (uint32_t w = lane; w < nw; w += 32) {
								if (wbase + w < wcap) P.T[wbase + w] = tst[w];
								else P.totals->overflow = 1;
							}
							__syncwarp();
						}
						if (zm) cell_slow_triangles<Sample>(P, tb, x, y, z, pat, zm, vb + cl, tid, cell);
					}
#else
					{
						// Triangles of this round, one lane per TRIANGLE: each owner lane marks its
						// slots in the round's triangle range, then lane t fetches the owner's pattern
						// by shuffle and its three vertex ids from the owner's column of the scratch;
						// consecutive lanes write consecutive triangles.  (Cells with an on-iso corner
						// write their own triangles afterwards: bit 31 of the pattern word.)
						uint8_t *own = reinterpret_cast<uint8_t *>(scr + 13 * 32);
						const uint32_t e0 = ex & 0xFFFFu, ntot = tot & 0xFFFFu;
						if (on) {
							scr[12 * 32 + lane] = vb + cl;
							for (uint32_t j = 0; j < pat.ntri; j++) own[e0 + j] = (uint8_t)lane;
						}
						__syncwarp();
#if EMC_V2
						// bit 12: the triangle's first two corners change places (the pattern's winding flag against the grid's
						// orientation, marching_cubes_33.c:1246-1250); the capacity is checked once for the round's triangles
						const uint32_t sm = pat.start | (((pat.m != 0u) != swap_all ? 1u : 0u) << 12) | (zm ? 0x80000000u : 0u);
						const uint32_t tfirst = tbase + runT;
						const bool fits = (uint64_t)tfirst + ntot <= (uint64_t)P.capT;
						uint32_t *const Tp = P.T + 3 * (uint64_t)tfirst;
						for (uint32_t t0 = 0; t0 < ntot; t0 += 32) {
							const uint32_t t = t0 + lane;
							const bool act = t < ntot;
							const int c = act ? (int)own[t] : 0;
							const uint32_t csm = __shfl_sync(0xFFFFFFFFu, sm, c), ce0 = __shfl_sync(0xFFFFFFFFu, e0, c);
							uint64_t ccell = 0;
							if (KEYS && P.tcell) ccell = __shfl_sync(0xFFFFFFFFu, cell, c);
							if (act && !(csm >> 31)) {
								const unsigned tw = tb.tri[(csm & 0xFFFu) + (t - ce0)];
								if (fits) {
									const uint32_t *ids = scr + c;
									const uint32_t i0 = ids[((tw >> 8) & 15u) * 32], i1 = ids[((tw >> 4) & 15u) * 32], i2 = ids[(tw & 15u) * 32];
									const bool sw = ((csm >> 12) & 1u) != 0u;
									uint32_t *T = Tp + 3 * t;
#if MC33_STREAM_STORES
									__stcs(T, sw ? i0 : i1); __stcs(T + 1, sw ? i1 : i0); __stcs(T + 2, i2);     // (written once, read by nobody here)
#else
									T[0] = sw ? i0 : i1; T[1] = sw ? i1 : i0; T[2] = i2;
#endif
									if (KEYS && P.tcell) P.tcell[tfirst + t] = ccell;
								} else {
									emit_triangle_fast<KEYS>(P, tw, (((csm >> 12) & 1u) != 0u) != swap_all ? 1u : 0u, scr + c, 32, tfirst + t, ccell);
								}
							}
						}
#else
						const uint32_t sm = pat.start | (pat.m << 12) | (zm ? 0x80000000u : 0u);
						for (uint32_t t0 = 0; t0 < ntot; t0 += 32) {
							const uint32_t t = t0 + lane;
							const bool act = t < ntot;
							const int c = act ? (int)own[t] : 0;
							const uint32_t csm = __shfl_sync(0xFFFFFFFFu, sm, c), ce0 = __shfl_sync(0xFFFFFFFFu, e0, c);
							uint64_t ccell = 0;
							if (KEYS && P.tcell) ccell = __shfl_sync(0xFFFFFFFFu, cell, c);
							if (act && !(csm >> 31))
								emit_triangle_fast<KEYS>(P, tb.tri[(csm & 0xFFFu) + (t - ce0)], (csm >> 12) & 1u, scr + c, 32, tbase + runT + t, ccell);
						}
#endif
						__syncwarp();
						if (zm) cell_slow_triangles<Sample>(P, tb, x, y, z, pat, zm, vb + cl, tid, cell);
					}
#endif
					runT += tot & 0xFFFFu; runC += tot >> 16;
				}
				__syncwarp();
			}
		}
	}
}

// ---------------------------------------------------------------------------
// round-2 kernels: the bodies live in mc33_pipeline.cuh (shared with the CPU test harness)
// ---------------------------------------------------------------------------
#ifndef CNT2_MINB
#define CNT2_MINB 4     // 64 registers, 4 CTAs per SM (A/B at cfg2 / cfg3 / cfg5: 0.076 / 1.51 / 9.99 ms against 0.081 / 1.66 / 10.6 at 3 and 0.092 / 2.02 / 13.3 at 2)
#endif
template <typename Sample>
__global__ void __launch_bounds__(256, CNT2_MINB) k_count2(const __grid_constant__ Params P, const __grid_constant__ CountArgs A)
{
	extern __shared__ __align__(128) unsigned char smem[];
	const DevCtx cx(smem);
	count_body<Sample>(cx, P, global_tables(), A);
}

#define EMV2_SMEM (P2_VX_WARPS * P2_VX_WARP_BYTES)
template <typename Sample, bool KEYS>
__global__ void __launch_bounds__(256, 4) k_emit_vertices2(const __grid_constant__ Params P, const __grid_constant__ VertexArgs A)
{
	extern __shared__ __align__(128) unsigned char smem[];
	const DevCtx cx(smem);
	emit_vertices_body<Sample, KEYS>(cx, P, A, smem + (threadIdx.x >> 5) * P2_VX_WARP_BYTES);
}

#define EMC2_SMEM (TBL2_BYTES + P2_EM_WARPS * P2_EM_WARP_BYTES)
template <typename Sample, bool KEYS>
__global__ void __launch_bounds__(256, 3) k_emit_cells2(const __grid_constant__ Params P, const __grid_constant__ EmitArgs A)
{
	extern __shared__ __align__(128) unsigned char smem[];
	const Tables tb = load_tables2(smem);
	const DevCtx cx(smem);
	emit_cells_body<Sample, KEYS>(cx, P, tb, A, smem + TBL2_BYTES + (threadIdx.x >> 5) * P2_EM_WARP_BYTES);
}

// ===========================================================================
// host side
// ===========================================================================
static thread_local char g_err[512] = "";
static int fail(int code, const char *fmt, const char *a = "", const char *b = "")
{
	snprintf(g_err, sizeof g_err, fmt, a, b);
	return code;
}
#define CU(call)                                                                                \
	do {                                                                                        \
		cudaError_t e_ = (call);                                                                \
		if (e_ != cudaSuccess) return fail(e_ == cudaErrorMemoryAllocation ? MC33CU_ERR_NOMEM : MC33CU_ERR_CUDA, \
		                                   "%s: %s", #call, cudaGetErrorString(e_));             \
	} while (0)

#define MC33CU_MAX_STREAMS 4
struct mc33cu_ctx {
	mc33cu_desc d;
	int device;
	int n_sm;
	cudaStream_t own_stream, stream;
	Params P;
	ClsPlan cls;
	size_t sample_size, real_size;
	uint64_t n_samples;
	void *grid_owned;        // device copy made by the upload calls
	// second stream: the triangle download of mc33cu_emit_host_async runs on it next to the vertex kernel
	cudaStream_t copy_stream; cudaEvent_t ev_cells, ev_copy; bool copy_pending;
	cudaEvent_t ev_chain;    // mc33cu_stream_wait: "everything issued on this context's stream so far"
	uint32_t *dlT; uint64_t dlT_words;     // where the triangles go on the host (set by mc33cu_emit_host_async for one emit)
	void *pinned; size_t pinned_bytes;   // staging for row-wise uploads
	// k_count blocks (CNT_WARPS * G rows each) and their (V, T, C) sums
	uint32_t *blk_sum; uint32_t nblk, cnt_gw, cnt_rb;
	// host mirror of totals
	Totals *h_all;           // pinned: [0] single-isovalue state, [1 + j] sweep set j
	Totals *h_totals;        // = &h_all[counted_set + 1]
	bool counted;
	uint32_t counted_mask;   // bit (state + 1): the state has a valid count (cleared for the sets by a new sweep classify)
	uint32_t pending_mask;   // ... has count / emit launches whose flags mc33cu_sync has not looked at yet
	uint32_t hvalid_mask;    // ... its totals are in the host mirror
	// staging outputs for the host path
	void *oV; float *oN; int32_t *oC; uint32_t *oT; uint64_t ocapV, ocapT;
	// timing
	bool timing; cudaEvent_t ev[6]; bool ev_valid;
	uint64_t launches;
	// resident CTAs per SM of the two emit kernels (persistent grids)
	uint32_t emc_per_sm, emv_per_sm, emc2_per_sm, emv2_per_sm;
	// what the device picked for an isovalue of a state (from the totals of an earlier extraction that was synchronised):
	// the same isovalue on the same state launches only that kernel (unconditionally: a stale entry costs time, never correctness)
	double pick_iso[SWEEP_MAX + 1][4]; int8_t pick_val[SWEEP_MAX + 1][4]; uint8_t pick_pos[SWEEP_MAX + 1];
	double state_iso[SWEEP_MAX + 1];       // isovalue of the last count of each state
	int cells;                             // cell kernel: 0 both launched, the device picks by the on-iso statistic; 1 / 2 force one
	int vtx;                               // 2: vertices straight from the bitmaps (no vertex tasks); 1: the round-1 task form
	uint32_t fine_pct, fine_rows;          // k_emit_cells: share of the rows handed out in small units at the end, unit size
	bool fine_env;                         // ... set by the environment (A/B runs): no automatic choice for small grids
	// vertex tasks (written by the cell kernel, read by the vertex kernel of the same extraction), grown to the largest
	// output capacity seen; one buffer per stream the context has been used on, so that extractions of different sweep
	// sets may run side by side on different streams
	struct { cudaStream_t s; uint64_t *p; uint64_t cap; cudaEvent_t done; uint64_t stamp; } vt[MC33CU_MAX_STREAMS];
	int n_vt, vt_cur;
	uint64_t vt_clock;
	// bitmaps of the single-isovalue path (P.S / P.Z / P.rowZ point here or into a sweep set)
	uint32_t *S0, *Z0, *rowZ0, *D0;
	size_t dwords;                         // words of one Z dirty-bit array
	// who wrote a Z array last: 1 k_classify_sweep (dirty bits valid), 2 a single-isovalue kernel (every word
	// rewritten each time, dirty bits not kept): switching 2 -> 1 clears the array first
	uint8_t zmode0, zmode_sw[SWEEP_MAX], *zmode_cur;
	uint32_t epoch;                        // on-iso hint epoch, bumped by every classify launch
	uint32_t epoch0;                       // ... of the last single-isovalue classify
	int counted_set;                       // what the last count phase ran on: -1 single path, else the sweep set
	int cur_state;                         // ... and what the kernel parameters point at right now
	uint32_t emits_since_count[SWEEP_MAX + 1];   // emit launches since the count phase of each state (index set + 1)
	// iso sweep: up to SWEEP_MAX pre-classified bitmap sets (allocated by the first sweep)
	uint32_t *swS, *swZ, *swRowZ, *swAny, *swD;
	// ... and per-set count state (visit bitmap, word prefixes, row bases, totals, block sums), so that all
	// the sets can be counted before any is emitted (one all-gather of the counts per sweep across slabs)
	uint32_t *A0, *rowBV0, *blk0; uint64_t *wpreV0; Totals *totals0;
	uint32_t *swA, *swRowB, *swBlk; uint64_t *swWpre; Totals *swTotals;
	double sw_iso[SWEEP_MAX]; int sw_n; uint32_t sw_epoch;
	bool sweep_ready;                      // all the sweep-set arrays are allocated
	// round-2 pipeline (count with look-back, record-based cell kernel); MC33_B200_PIPE=1 selects the round-1 kernels
	int pipe;
	unsigned long long *lb0, *swLb, *lb_cur;   // look-back words [nblk][6] per state
	uint32_t *lbt0, *swLbt, *lbt_cur;          // {ticket, finished} per state
	uint16_t *pcache0, *swPcache;              // pattern of every complex cell, per state
	uint32_t lb_tag;                           // 16-bit launch tag of the look-back words
	uint32_t *export4;                         // where the next count leaves {nV, nT, nShared, nCentre} on the device (or null)
	EmitArgs emit_shape;
};

extern "C" const char *mc33cu_last_error(void) { return g_err; }

extern "C" int mc33cu_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

static int upload_tables()
{
	alignas(16) static unsigned char img[TBL_BYTES];
	memset(img, 0, sizeof img);
	memcpy(img, MC33_CASE256, 512);
	memcpy(img + 512, MC33_SIMPLE256, 512);
	memcpy(img + 1024, MC33_TRI, sizeof(MC33_TRI));
	unsigned char *pat = img + 1024 + TBL_TRI_BYTES;
	for (int i = 0; i < MC33_NTRI_WORDS; i++) pat[i] = (uint8_t)(MC33_PAT_NTRI[i] | (MC33_PAT_CENTRE[i] << 7));
	uint32_t *cinfo = (uint32_t *)(pat + TBL_PAT_BYTES);
	for (int i = 0; i < 256; i++) cinfo[i] = (uint32_t)MC33_SIMPLE256[i] | ((uint32_t)((MC33_CASE256[i] >> 11) & 1u) << 16);
	CU(cudaMemcpyToSymbol(d_tables, img, sizeof img));
	{
		// keep masks: the zero-area drop (marching_cubes_33.c:1235) of every pattern under every on-iso corner mask
		static uint16_t pord[MC33_NTRI_WORDS], keep[MC33_NPATTERNS * 256];
		Tables tb;
		tb.case256 = MC33_CASE256; tb.simple256 = MC33_SIMPLE256; tb.tri = MC33_TRI; tb.pat = pat; tb.cinfo = cinfo; tb.pord = pord; tb.keep = keep;
		unsigned n = 0;
		for (int i = 0; i < MC33_NTRI_WORDS; i++) {
			pord[i] = 0;
			if (MC33_PAT_NTRI[i] && n < MC33_NPATTERNS) {
				pord[i] = (uint16_t)n;
				for (unsigned zm = 0; zm < 256; zm++) keep[n * 256 + zm] = (uint16_t)keep_mask_walk(tb, (unsigned)i, zm);
				n++;
			}
		}
		if (n != MC33_NPATTERNS) return fail(MC33CU_ERR_ARG, "pattern table: unexpected number of patterns");
		CU(cudaMemcpyToSymbol(d_pord, pord, sizeof pord));
		CU(cudaMemcpyToSymbol(d_keep, keep, sizeof keep));
	}
	return MC33CU_OK;
}

template <typename T> static int dalloc(T **p, size_t n)
{
	*p = nullptr;
	CU(cudaMalloc((void **)p, n ? n * sizeof(T) : sizeof(T)));
	return MC33CU_OK;
}

extern "C" void mc33cu_destroy(mc33cu_ctx *c)
{
	if (!c) return;
	cudaSetDevice(c->device);
	if (c->own_stream) cudaStreamSynchronize(c->own_stream);
	Params &P = c->P;
	if (!c->S0) { c->S0 = P.S; c->Z0 = P.Z; c->rowZ0 = P.rowZ; c->A0 = P.A; c->wpreV0 = P.wpreV; c->rowBV0 = P.rowBV; c->totals0 = P.totals; c->blk0 = c->blk_sum; }
	cudaFree(c->S0); cudaFree(c->Z0); cudaFree(c->A0); cudaFree(c->rowZ0); cudaFree(c->wpreV0);
	cudaFree(c->swA); cudaFree(c->swWpre); cudaFree(c->swRowB); cudaFree(c->swTotals); cudaFree(c->swBlk);
	cudaFree(c->swS); cudaFree(c->swZ); cudaFree(c->swRowZ); cudaFree(c->swAny); cudaFree(c->swD); cudaFree(c->D0);
	cudaFree(c->rowBV0); cudaFree(c->totals0);
	cudaFree(c->blk0);
	for (int i = 0; i < c->n_vt; i++) { cudaFree(c->vt[i].p); cudaEventDestroy(c->vt[i].done); }
	cudaFree(c->lb0); cudaFree(c->lbt0); cudaFree(c->pcache0); cudaFree(c->swLb); cudaFree(c->swLbt); cudaFree(c->swPcache);
	cudaFree(c->grid_owned);
	if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
	if (c->ev_cells) cudaEventDestroy(c->ev_cells);
	if (c->ev_copy) cudaEventDestroy(c->ev_copy);
	if (c->ev_chain) cudaEventDestroy(c->ev_chain);
	cudaFree(c->oV); cudaFree(c->oN); cudaFree(c->oC); cudaFree(c->oT);
	if (c->pinned) cudaFreeHost(c->pinned);
	if (c->h_all) cudaFreeHost(c->h_all);
	for (int i = 0; i < 6; i++) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
	if (c->own_stream) cudaStreamDestroy(c->own_stream);
	free(c);
}

static int set_geom(mc33cu_ctx *c, const mc33cu_desc *d);

// (the attribute belongs to the function, not to a context: contexts of different row
// lengths coexist, so it is set to the largest ring any plan can ask for)
#define CLS_MAX_SMEM ((16384 + 512 + 32 + 127) / 128 * 128 * CLS_STAGES)

template <typename Sample> static int set_kernel_attrs(const ClsPlan &pl)
{
	if ((size_t)pl.stage_bytes * CLS_STAGES > CLS_MAX_SMEM) return fail(MC33CU_ERR_ARG, "classify plan exceeds the shared memory ring");
	CU(cudaFuncSetAttribute(k_classify<Sample>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CLS_MAX_SMEM));
	CU(cudaFuncSetAttribute(k_classify_vec<Sample>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CLS_MAX_SMEM));
	CU(cudaFuncSetAttribute(k_emit_cells<Sample, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EMC_SMEM(32)));
	CU(cudaFuncSetAttribute(k_emit_cells<Sample, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EMC_SMEM(32)));
	CU(cudaFuncSetAttribute(k_emit_cells2<Sample, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EMC2_SMEM));
	CU(cudaFuncSetAttribute(k_emit_cells2<Sample, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EMC2_SMEM));
	CU(cudaFuncSetAttribute(k_classify_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CLS_MAX_SMEM));
	return MC33CU_OK;
}

extern "C" int mc33cu_create(const mc33cu_desc *d, int device, mc33cu_ctx **out)
{
	if (!d || !out) return fail(MC33CU_ERR_ARG, "null argument");
	*out = nullptr;
	if (d->dtype < MC33CU_F32 || d->dtype > MC33CU_U32) return fail(MC33CU_ERR_ARG, "bad dtype");
	if (!d->nx || !d->ny || !d->nz) return fail(MC33CU_ERR_ARG, "empty grid");
	if (d->store < MC33CU_SPN0 || d->store > MC33CU_SPNC) return fail(MC33CU_ERR_ARG, "bad store variant");
	const uint32_t NZ = d->nz + 1;
	if (!(d->cell_z0 < d->cell_z1 && d->cell_z1 <= d->nz && d->z_lo < d->z_hi && d->z_hi <= NZ))
		return fail(MC33CU_ERR_ARG, "bad slab range");
	if (d->is_last && d->cell_z1 != d->nz) return fail(MC33CU_ERR_ARG, "is_last but cell_z1 != nz");
	{
		const uint32_t need_lo = (d->cell_z0 > 0 ? d->cell_z0 : 1) - 1;
		const uint32_t need_hi = d->cell_z1 + 2 < NZ ? d->cell_z1 + 2 : NZ;
		if (d->z_lo > need_lo || d->z_hi < need_hi) return fail(MC33CU_ERR_ARG, "slab halo too small");
	}
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
		return fail(MC33CU_ERR_CUDA, "no CUDA device: %s", cudaGetErrorString(cudaGetLastError()));
	if (device < 0 || device >= ndev) return fail(MC33CU_ERR_ARG, "bad device index");
	CU(cudaSetDevice(device));
	mc33cu_ctx *c = (mc33cu_ctx *)calloc(1, sizeof(mc33cu_ctx));
	if (!c) return fail(MC33CU_ERR_NOMEM, "calloc");
	c->d = *d;
	c->device = device;
	c->emc_per_sm = EMC_MINB; c->emv_per_sm = EMV_MINB; c->emc2_per_sm = 3;
	c->emv2_per_sm = 4; c->vtx = 1;
	if (const char *e = getenv("MC33_B200_EMV2_PER_SM")) { int v = atoi(e); if (v >= 1 && v <= 8) c->emv2_per_sm = (uint32_t)v; }
	if (const char *e = getenv("MC33_B200_VTX")) { if (atoi(e) == 2) c->vtx = 2; }
	c->cells = 0;
	if (const char *e = getenv("MC33_B200_CELLS")) { int v = atoi(e); if (v >= 0 && v <= 2) c->cells = v; }
	if (const char *e = getenv("MC33_B200_EMC2_PER_SM")) { int v = atoi(e); if (v >= 1 && v <= 8) c->emc2_per_sm = (uint32_t)v; }
	c->P.dbg_noz = getenv("MC33_B200_DEBUG_NOZ") ? 1u : 0u;
	c->fine_pct = 8; c->fine_rows = 4;
	if (const char *e = getenv("MC33_B200_FINE_PCT")) { int v = atoi(e); if (v >= 0 && v <= 100) { c->fine_pct = (uint32_t)v; c->fine_env = true; } }
	if (const char *e = getenv("MC33_B200_FINE_ROWS")) { int v = atoi(e); if (v >= 1 && v <= 32) { c->fine_rows = (uint32_t)v; c->fine_env = true; } }
	if (const char *e = getenv("MC33_B200_EMC_PER_SM")) { int v = atoi(e); if (v >= 1 && v <= 8) c->emc_per_sm = (uint32_t)v; }
	if (const char *e = getenv("MC33_B200_EMV_PER_SM")) { int v = atoi(e); if (v >= 1 && v <= 8) c->emv_per_sm = (uint32_t)v; }
	int rc = upload_tables();
	if (rc) { free(c); return rc; }
	if (cudaDeviceGetAttribute(&c->n_sm, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || c->n_sm <= 0) c->n_sm = 148;
	Params &P = c->P;
	P.nx = d->nx; P.ny = d->ny; P.nz = d->nz;
	P.NX = d->nx + 1; P.NY = d->ny + 1;
	P.zlo = d->z_lo; P.zhi = d->z_hi;
	P.cz0 = d->cell_z0; P.cz1 = d->cell_z1;
	P.pz0 = d->cell_z0; P.pz1 = d->is_last ? NZ : d->cell_z1;
	P.hz = d->is_last ? 0xFFFFFFFFu : d->cell_z1;
	P.W = (P.NX + 31) / 32; P.WC = (P.nx + 31) / 32;
	P.Q = (P.W + 3) / 4;
	P.WP = 4 * P.Q + 4;
	P.G = P.Q <= 32 ? 32 / P.Q : 1;
	P.Lrows = (P.zhi - P.zlo) * P.NY;
	if (P.NX > 65535) { free(c); return fail(MC33CU_ERR_ARG, "rows longer than 65535 samples are not supported"); }
	if ((uint64_t)P.NX * P.NY > 0x7FFFFFFFull) { free(c); return fail(MC33CU_ERR_ARG, "slices of more than 2^31-1 samples are not supported"); }
	if ((uint64_t)(P.zhi - P.zlo) * P.NY > 0x7FFFFFFFull / P.WP) { free(c); return fail(MC33CU_ERR_ARG, "too many rows"); }
	P.mQ = P.Q >= 2 ? (uint32_t)(0x100000000ull / P.Q) : 0u;
	P.mNY = (uint32_t)(0x100000000ull / P.NY);
	set_geom(c, d);
	static const size_t ssz[5] = {4, 8, 1, 2, 4};
	c->sample_size = ssz[d->dtype];
	c->real_size = d->dtype == MC33CU_F64 ? 8 : 4;
	c->n_samples = (uint64_t)P.Lrows * P.NX;
	{
		// Rows per count block (8 warps x GW groups of G rows, at most 256).  The blocks of a wave finish together, so the
		// kernel takes (number of waves) x (a block's rows + its fixed tail: ticket, barriers, look-back, about 300 quads'
		// worth); the block size is chosen for the fewest row-times, larger blocks on a tie.  Measured: cfg5 (768^3, 2458
		// blocks of 240 rows = 4.15 waves) 10.3 -> 8.9 ms with 200 rows (4.98 waves), a 258-slice slab of cfg4 0.454 ->
		// 0.418 ms with 224 rows; cfg2 / cfg3 keep 256.  MC33_B200_CNT_RPW=<rows per warp> overrides.
		const uint32_t gw_max = 32 / P.G ? 32 / P.G : 1;
		const uint64_t slots = (uint64_t)c->n_sm * CNT2_MINB;
		const uint32_t tail_rows = (300 + P.Q - 1) / P.Q;
		uint32_t best = gw_max;
		uint64_t best_cost = ~0ull;
		for (uint32_t gw = gw_max; gw >= 1; gw--) {
			const uint64_t rb = (uint64_t)CNT_WARPS * gw * P.G, nb = (P.Lrows + rb - 1) / rb;
			const uint64_t cost = ((nb + slots - 1) / slots) * (rb + tail_rows);
			if (cost < best_cost) { best_cost = cost; best = gw; }
			if (gw == gw_max && nb <= slots) break;          // one wave of the largest blocks: smaller ones only add tails (cfg1)
		}
		c->cnt_gw = best;
		if (const char *e = getenv("MC33_B200_CNT_RPW")) { int v = atoi(e); if (v >= 1 && v <= 32) c->cnt_gw = (uint32_t)v / P.G ? (uint32_t)v / P.G : 1; }
		c->cnt_rb = CNT_WARPS * c->cnt_gw * P.G;
		c->nblk = (P.Lrows + c->cnt_rb - 1) / c->cnt_rb;
	}
	{
		// classify chunks: ~16 KB of whole rows (a multiple of the CTA's warp count
		// when possible), or 16 KB pieces of one long row (a multiple of 4 words)
		ClsPlan &pl = c->cls;
		const size_t rowb = (size_t)P.NX * c->sample_size, target = 16384 + 512;
		if (rowb <= target) {
			pl.rows = (uint32_t)(target / rowb); pl.words = P.W; pl.nwchunk = 1;
			if (pl.rows > CLS_THREADS / 32) pl.rows -= pl.rows % (CLS_THREADS / 32);
			if (pl.rows > P.Lrows) pl.rows = P.Lrows;
			pl.nchunks = (P.Lrows + pl.rows - 1) / pl.rows;
			pl.stage_bytes = (uint32_t)(((size_t)pl.rows * rowb + 32 + 127) & ~(size_t)127);
		} else {
			pl.rows = 1; pl.words = (uint32_t)(16384 / (32 * c->sample_size));
			pl.nwchunk = (P.W + pl.words - 1) / pl.words;
			pl.nchunks = P.Lrows * pl.nwchunk;
			pl.stage_bytes = (uint32_t)(((size_t)pl.words * 32 * c->sample_size + 32 + 127) & ~(size_t)127);
		}
	}
	switch (d->dtype) {
	case MC33CU_F32: rc = set_kernel_attrs<float>(c->cls); break;
	case MC33CU_F64: rc = set_kernel_attrs<double>(c->cls); break;
	case MC33CU_U8:  rc = set_kernel_attrs<uint8_t>(c->cls); break;
	case MC33CU_U16: rc = set_kernel_attrs<uint16_t>(c->cls); break;
	default:         rc = set_kernel_attrs<uint32_t>(c->cls); break;
	}
	if (rc) { free(c); return rc; }

#define TRY(x) do { rc = (x); if (rc) { mc33cu_destroy(c); return rc; } } while (0)
#define TRYCU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { mc33cu_destroy(c); \
	return fail(e_ == cudaErrorMemoryAllocation ? MC33CU_ERR_NOMEM : MC33CU_ERR_CUDA, "%s: %s", #x, cudaGetErrorString(e_)); } } while (0)
	TRYCU(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
	c->stream = c->own_stream;
	const size_t bm = (size_t)P.Lrows * P.WP;
	TRY(dalloc(&P.S, bm)); TRY(dalloc(&P.Z, bm));
	TRYCU(cudaMemsetAsync(P.S, 0, bm * 4, c->stream)); TRYCU(cudaMemsetAsync(P.Z, 0, bm * 4, c->stream));
	TRY(dalloc(&P.A, bm));
	TRYCU(cudaMemsetAsync(P.A, 0, bm * 4, c->stream));
	TRY(dalloc(&P.rowZ, (size_t)P.Lrows));
	TRYCU(cudaMemsetAsync(P.rowZ, 0, (size_t)P.Lrows * 4, c->stream));
	P.zepoch = 0; c->epoch = 0;
	TRY(dalloc(&P.wpreV, (size_t)P.Lrows * P.WP));
	TRYCU(cudaMemsetAsync(P.wpreV, 0, (size_t)P.Lrows * P.WP * 8, c->stream));
	TRY(dalloc(&P.rowBV, ((size_t)P.Lrows + 1) * 3));
	P.rowBT = P.rowBV + (P.Lrows + 1); P.rowBC = P.rowBT + (P.Lrows + 1);
	TRY(dalloc(&P.totals, 1));
	TRYCU(cudaMemsetAsync(P.totals, 0, sizeof(Totals), c->stream));
	TRY(dalloc(&c->blk_sum, (size_t)c->nblk * 3));
	TRYCU(cudaMallocHost((void **)&c->h_all, sizeof(Totals) * (SWEEP_MAX + 1)));
	memset(c->h_all, 0, sizeof(Totals) * (SWEEP_MAX + 1));
	c->h_totals = c->h_all;
	for (int i = 0; i < 6; i++) TRYCU(cudaEventCreate(&c->ev[i]));
	TRYCU(cudaStreamSynchronize(c->stream));
#undef TRY
#undef TRYCU
	c->counted_set = -1; c->cur_state = -1;
	c->S0 = P.S; c->Z0 = P.Z; c->rowZ0 = P.rowZ;
	c->A0 = P.A; c->wpreV0 = P.wpreV; c->rowBV0 = P.rowBV; c->totals0 = P.totals; c->blk0 = c->blk_sum;
	P.anyZp = &P.totals->anyZ;
	c->dwords = ((size_t)P.Lrows * (P.W / 4 + 1) + 31) / 32 + 2;
	if (cudaMalloc((void **)&c->D0, c->dwords * 4) != cudaSuccess || cudaMemset(c->D0, 0, c->dwords * 4) != cudaSuccess) {
		mc33cu_destroy(c);
		return fail(MC33CU_ERR_NOMEM, "cudaMalloc (Z dirty bits)");
	}
	P.D = c->D0; c->zmode_cur = &c->zmode0;
	c->pipe = 2;
	if (const char *e = getenv("MC33_B200_PIPE")) { if (atoi(e) == 1) c->pipe = 1; }
	{
		const size_t bmw = (size_t)P.Lrows * P.WP;
		if (cudaMalloc((void **)&c->lb0, (size_t)c->nblk * P2_LB_WORDS * 8) != cudaSuccess ||
		    cudaMemset(c->lb0, 0, (size_t)c->nblk * P2_LB_WORDS * 8) != cudaSuccess ||
		    cudaMalloc((void **)&c->lbt0, 8) != cudaSuccess || cudaMemset(c->lbt0, 0, 8) != cudaSuccess ||
		    cudaMalloc((void **)&c->pcache0, bmw * 32 * sizeof(uint16_t)) != cudaSuccess) {
			mc33cu_destroy(c);
			return fail(MC33CU_ERR_NOMEM, "cudaMalloc (look-back words / pattern cache)");
		}
		P.pcache = c->pipe == 2 ? c->pcache0 : nullptr; c->lb_cur = c->lb0; c->lbt_cur = c->lbt0;
		emit_geometry(P.Q, c->emit_shape);
	}
	*out = c;
	return MC33CU_OK;
}

static int set_geom(mc33cu_ctx *c, const mc33cu_desc *d)
{
	if (d->store < MC33CU_SPN0 || d->store > MC33CU_SPNC) return fail(MC33CU_ERR_ARG, "bad store variant");
	Params &P = c->P;
	P.geom.store = d->store; P.geom.normal_neg = d->normal_neg; P.geom.tsa = d->tsa;
	for (int i = 0; i < 3; i++) { P.geom.O[i] = d->O[i]; P.geom.D[i] = d->D[i]; }
	P.geom.ca = d->ca; P.geom.cb = d->cb;
	for (int i = 0; i < 3; i++) { P.geom.Of[i] = (float)d->O[i]; P.geom.Df[i] = (float)d->D[i]; }
	P.geom.caf = (float)d->ca; P.geom.cbf = (float)d->cb;
	for (int i = 0; i < 9; i++) { P.geom.A[i] = d->A[i]; P.geom.Ai[i] = d->Ai[i]; }
	return MC33CU_OK;
}

extern "C" int mc33cu_set_geometry(mc33cu_ctx *c, const mc33cu_desc *d)
{
	if (!c || !d) return fail(MC33CU_ERR_ARG, "null argument");
	return set_geom(c, d);
}

extern "C" int mc33cu_set_stream(mc33cu_ctx *c, void *s)
{
	if (!c) return fail(MC33CU_ERR_ARG, "null context");
	c->stream = s ? (cudaStream_t)s : c->own_stream;
	return MC33CU_OK;
}

extern "C" int mc33cu_grid_device(mc33cu_ctx *c, const void *dev)
{
	if (!c || !dev) return fail(MC33CU_ERR_ARG, "null argument");
	if ((uintptr_t)dev % c->sample_size) return fail(MC33CU_ERR_ARG, "sample pointer is not aligned to the sample size");
	c->P.data = dev;
	return MC33CU_OK;
}

static int ensure_grid(mc33cu_ctx *c)
{
	if (!c->grid_owned) {
		CU(cudaSetDevice(c->device));
		CU(cudaMalloc(&c->grid_owned, c->n_samples * c->sample_size));
	}
	c->P.data = c->grid_owned;
	return MC33CU_OK;
}

// Page-locking of CALLER memory is explicit (mc33cu_host_register): the upload calls never
// register anything behind the caller's back.  A registered block is copied by DMA at link
// speed and asynchronously; pageable memory goes through the driver's staging buffers.
static int upload_block(mc33cu_ctx *c, const void *host, bool wait)
{
	if (!c || !host) return fail(MC33CU_ERR_ARG, "null argument");
	int rc = ensure_grid(c);
	if (rc) return rc;
	CU(cudaSetDevice(c->device));
	const size_t bytes = c->n_samples * c->sample_size;
	CU(cudaMemcpyAsync(c->grid_owned, host, bytes, cudaMemcpyHostToDevice, c->stream));
	if (wait) CU(cudaStreamSynchronize(c->stream));
	return MC33CU_OK;
}

extern "C" int mc33cu_grid_upload(mc33cu_ctx *c, const void *host) { return upload_block(c, host, true); }
extern "C" int mc33cu_grid_upload_async(mc33cu_ctx *c, const void *host) { return upload_block(c, host, false); }

// ---------------------------------------------------------------------------
// pooled page-locked host memory for result arrays (the drop-in hands them to the
// caller as surface.V/N/T/color and gets them back in free_surface_memory)
// ---------------------------------------------------------------------------
#include <mutex>
#include <vector>
namespace {
struct HostBlock { void *p; size_t cap; bool used; uint64_t stamp; };    // stamp: when the block was handed back
std::mutex g_pool_mx;
std::vector<HostBlock> g_pool;
uint64_t g_pool_clock = 0;
const size_t POOL_MAX_FREE = 16;
}

extern "C" int mc33cu_host_alloc(size_t bytes, void **out)
{
	if (!out) return fail(MC33CU_ERR_ARG, "null argument");
	*out = nullptr;
	if (!bytes) bytes = 1;
	std::lock_guard<std::mutex> lk(g_pool_mx);
	int best = -1;
	for (size_t i = 0; i < g_pool.size(); i++)
		if (!g_pool[i].used && g_pool[i].cap >= bytes && g_pool[i].cap <= bytes + bytes / 2 + (1u << 20) &&
		    (best < 0 || g_pool[i].cap < g_pool[(size_t)best].cap)) best = (int)i;
	if (best >= 0) { g_pool[(size_t)best].used = true; *out = g_pool[(size_t)best].p; return MC33CU_OK; }
	const size_t cap = (bytes + (1u << 20) - 1) & ~(size_t)((1u << 20) - 1);
	void *p = nullptr;
	if (cudaHostAlloc(&p, cap, cudaHostAllocPortable) != cudaSuccess) {
		cudaGetLastError();
		return fail(MC33CU_ERR_NOMEM, "cudaHostAlloc failed");
	}
	g_pool.push_back({p, cap, true, 0});
	*out = p;
	return MC33CU_OK;
}

extern "C" int mc33cu_host_free(void *p)
{
	if (!p) return MC33CU_OK;
	std::lock_guard<std::mutex> lk(g_pool_mx);
	size_t nfree = 0;
	int found = -1, oldest = -1;
	for (size_t i = 0; i < g_pool.size(); i++) {
		if (g_pool[i].p == p) found = (int)i;
		else if (!g_pool[i].used) {
			nfree++;
			if (oldest < 0 || g_pool[i].stamp < g_pool[(size_t)oldest].stamp) oldest = (int)i;
		}
	}
	if (found < 0) return MC33CU_ERR_ARG;               // not ours: the caller frees it with free()
	g_pool[(size_t)found].used = false;
	g_pool[(size_t)found].stamp = ++g_pool_clock;
	if (nfree >= POOL_MAX_FREE) {
		// too many idle blocks: release the one that has been idle longest (sizes no caller asks for any more must
		// not crowd out the sizes in use -- page-locking a fresh block costs ~0.3 ms per MB)
		cudaFreeHost(g_pool[(size_t)oldest].p);
		g_pool.erase(g_pool.begin() + oldest);
	}
	return MC33CU_OK;
}

// process-wide registry of caller blocks page-locked through mc33cu_host_register (reference counted: several
// extractors may share one grid); cudaHostRegisterPortable makes the block DMA-able from every device
namespace {
struct RegBlock { const void *p; size_t bytes; int refs; };
std::vector<RegBlock> g_reg;
}

extern "C" int mc33cu_host_register(const void *p, size_t bytes)
{
	if (!p || !bytes) return fail(MC33CU_ERR_ARG, "null argument");
	std::lock_guard<std::mutex> lk(g_pool_mx);
	for (auto &r : g_reg)
		if (r.p == p && r.bytes == bytes) { r.refs++; return MC33CU_OK; }
	for (auto &b : g_pool)                               // pool memory is page-locked already
		if ((const char *)p >= (const char *)b.p && (const char *)p + bytes <= (const char *)b.p + b.cap) return MC33CU_OK;
	cudaError_t e = cudaHostRegister(const_cast<void *>(p), bytes, cudaHostRegisterPortable);
	if (e != cudaSuccess) {
		cudaGetLastError();
		return fail(MC33CU_ERR_CUDA, "cudaHostRegister: %s", cudaGetErrorString(e));
	}
	g_reg.push_back({p, bytes, 1});
	return MC33CU_OK;
}

extern "C" int mc33cu_host_unregister(const void *p)
{
	if (!p) return MC33CU_OK;
	std::lock_guard<std::mutex> lk(g_pool_mx);
	for (size_t i = 0; i < g_reg.size(); i++)
		if (g_reg[i].p == p) {
			if (--g_reg[i].refs == 0) {
				if (cudaHostUnregister(const_cast<void *>(p)) != cudaSuccess) cudaGetLastError();
				g_reg.erase(g_reg.begin() + (long)i);
			}
			return MC33CU_OK;
		}
	return MC33CU_OK;                                    // pool memory or never registered: nothing to do
}

// is the slab's part of F one contiguous block (grid_from_data_pointer layout, MC33_util_grd.c:600-612)?
static const char *rows_contiguous(const mc33cu_ctx *c, const void *const *const *F)
{
	const Params &P = c->P;
	const size_t rowb = (size_t)P.NX * c->sample_size;
	const char *first = (const char *)F[P.zlo][0];
	for (uint32_t z = P.zlo; z < P.zhi; z++)
		for (uint32_t y = 0; y < P.NY; y++)
			if ((const char *)F[z][y] != first + ((size_t)(z - P.zlo) * P.NY + y) * rowb) return nullptr;
	return first;
}

extern "C" int mc33cu_grid_rows_block(mc33cu_ctx *c, const void *const *const *F, const void **block, size_t *bytes)
{
	if (!c || !F || !block || !bytes) return fail(MC33CU_ERR_ARG, "null argument");
	*block = rows_contiguous(c, F);
	*bytes = *block ? c->n_samples * c->sample_size : 0;
	return MC33CU_OK;
}

// the same as mc33cu_grid_upload_rows without the final synchronisation when the rows form one block (the copy
// is then one asynchronous DMA on the context's stream: several contexts / devices upload side by side);
// separately allocated rows are staged and copied synchronously as in mc33cu_grid_upload_rows
extern "C" int mc33cu_grid_upload_rows_async(mc33cu_ctx *c, const void *const *const *F)
{
	if (!c || !F) return fail(MC33CU_ERR_ARG, "null argument");
	if (const char *first = rows_contiguous(c, F)) return upload_block(c, first, false);
	return mc33cu_grid_upload_rows(c, F);
}

extern "C" int mc33cu_grid_upload_rows(mc33cu_ctx *c, const void *const *const *F)
{
	if (!c || !F) return fail(MC33CU_ERR_ARG, "null argument");
	int rc = ensure_grid(c);
	if (rc) return rc;
	const Params &P = c->P;
	const size_t rowb = (size_t)P.NX * c->sample_size;
	// fast path: grid_from_data_pointer layout (MC33_util_grd.c:600-612), one block
	if (const char *first = rows_contiguous(c, F)) return mc33cu_grid_upload(c, first);
	// general path: alloc_F layout, one malloc per row (MC33_util_grd.c:147-169):
	// gather rows into pinned chunks, copy chunk by chunk (double buffered)
	const size_t chunk_rows = (32u << 20) / rowb ? (32u << 20) / rowb : 1;
	const size_t need = 2 * chunk_rows * rowb;
	if (c->pinned_bytes < need) {
		if (c->pinned) cudaFreeHost(c->pinned);
		c->pinned = nullptr; c->pinned_bytes = 0;
		CU(cudaMallocHost(&c->pinned, need));
		c->pinned_bytes = need;
	}
	cudaEvent_t done[2];
	CU(cudaEventCreate(&done[0])); CU(cudaEventCreate(&done[1]));
	bool used[2] = {false, false};
	size_t r = 0, buf = 0;
	int err = MC33CU_OK;
	while (r < P.Lrows) {
		const size_t n = (P.Lrows - r) < chunk_rows ? (P.Lrows - r) : chunk_rows;
		char *stage = (char *)c->pinned + buf * chunk_rows * rowb;
		if (used[buf] && cudaEventSynchronize(done[buf]) != cudaSuccess) { err = MC33CU_ERR_CUDA; break; }
		for (size_t k = 0; k < n; k++) {
			const size_t lr = r + k;
			memcpy(stage + k * rowb, F[P.zlo + lr / P.NY][lr % P.NY], rowb);
		}
		if (cudaMemcpyAsync((char *)c->grid_owned + r * rowb, stage, n * rowb, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
		    cudaEventRecord(done[buf], c->stream) != cudaSuccess) { err = MC33CU_ERR_CUDA; break; }
		used[buf] = true;
		buf ^= 1;
		r += n;
	}
	cudaError_t e = cudaStreamSynchronize(c->stream);
	cudaEventDestroy(done[0]); cudaEventDestroy(done[1]);
	if (err || e != cudaSuccess) return fail(MC33CU_ERR_CUDA, "row upload: %s", cudaGetErrorString(e));
	return MC33CU_OK;
}

// ---------------------------------------------------------------------------
template <typename Sample> static int launch_classify(mc33cu_ctx *c)
{
	Params &P = c->P;
	cudaStream_t s = c->stream;
	const ClsPlan &pl = c->cls;
	const size_t smem = (size_t)pl.stage_bytes * CLS_STAGES;
	uint32_t per_sm = (uint32_t)((220u << 10) / (smem + 1024));
	if (per_sm < 1) per_sm = 1;
	if (per_sm > 8) per_sm = 8;
	uint32_t grid = (uint32_t)c->n_sm * per_sm;
	if (grid > pl.nchunks) grid = pl.nchunks;
	// vector path: rows of whole NW-word groups, 16-byte aligned samples
	const uint32_t nwv = sizeof(Sample) >= 4 ? 4u : 16u / (uint32_t)sizeof(Sample);
	*c->zmode_cur = 2;                             // both kernels rewrite every Z word and keep no dirty bits
	if (pl.nwchunk == 1 && P.NX % (32 * nwv) == 0 && ((uintptr_t)P.data & 15) == 0)
		k_classify_vec<Sample><<<grid, CLS_THREADS, smem, s>>>(P, pl.rows, pl.nchunks, pl.stage_bytes);
	else k_classify<Sample><<<grid, CLS_THREADS, smem, s>>>(P, pl);
	c->launches++;
	return MC33CU_OK;
}

// a new epoch invalidates the per-row on-iso hints of every bitmap set
static int next_epoch(mc33cu_ctx *c, uint32_t *e)
{
	if (++c->epoch == 0) {
		CU(cudaMemsetAsync(c->rowZ0, 0, (size_t)c->P.Lrows * 4, c->stream));
		if (c->swRowZ) CU(cudaMemsetAsync(c->swRowZ, 0, (size_t)c->P.Lrows * 4 * SWEEP_MAX, c->stream));
		c->epoch = 1;
	}
	*e = c->epoch;
	return MC33CU_OK;
}

// point the kernel parameters at the bitmaps and count state of the single-isovalue path (set < 0) or of
// pre-classified sweep set `set`
static void select_state(mc33cu_ctx *c, int set)
{
	Params &P = c->P;
	c->cur_state = set;
	if (set < 0) {
		P.S = c->S0; P.Z = c->Z0; P.rowZ = c->rowZ0;
		P.D = c->D0; c->zmode_cur = &c->zmode0;
		P.A = c->A0; P.wpreV = c->wpreV0; P.rowBV = c->rowBV0; P.totals = c->totals0; c->blk_sum = c->blk0;
		P.anyZp = &P.totals->anyZ;
		P.zepoch = c->epoch0;
		P.pcache = c->pipe == 2 ? c->pcache0 : nullptr; c->lb_cur = c->lb0; c->lbt_cur = c->lbt0;
	} else {
		const size_t bm = (size_t)P.Lrows * P.WP, nr = (size_t)P.Lrows + 1;
		P.S = c->swS + (size_t)set * bm; P.Z = c->swZ + (size_t)set * bm; P.rowZ = c->swRowZ + (size_t)set * P.Lrows;
		P.anyZp = c->swAny + set;
		P.D = c->swD + (size_t)set * c->dwords; c->zmode_cur = &c->zmode_sw[set];
		P.zepoch = c->sw_epoch;
		P.A = c->swA + (size_t)set * bm; P.wpreV = c->swWpre + (size_t)set * bm; P.rowBV = c->swRowB + (size_t)set * nr * 3;
		P.totals = c->swTotals + set; c->blk_sum = c->swBlk + (size_t)set * c->nblk * 3;
		P.pcache = c->pipe == 2 ? c->swPcache + (size_t)set * bm * 32 : nullptr; c->lb_cur = c->swLb + (size_t)set * c->nblk * P2_LB_WORDS; c->lbt_cur = c->swLbt + 2 * set;
	}
	P.rowBT = P.rowBV + (P.Lrows + 1); P.rowBC = P.rowBT + (P.Lrows + 1);
}

__global__ void k_export_counts(const Totals *t, uint32_t *out4);

// set < 0: classify the single-isovalue bitmaps now; set >= 0: use pre-classified sweep set
template <typename Sample> static int launch_count_phase(mc33cu_ctx *c, int set)
{
	Params &P = c->P;
	cudaStream_t s = c->stream;
	select_state(c, set);
	// round-1 kernels: re-arm the totals (overflow flag) with a memset; the round-2 count kernel does it itself
	if (c->pipe == 1) CU(cudaMemsetAsync(P.totals, 0, offsetof(Totals, anyZ), s));
	if (c->timing) CU(cudaEventRecord(c->ev[0], s));
	if (set < 0) {
		int rc = next_epoch(c, &c->epoch0);
		if (rc) return rc;
		P.zepoch = c->epoch0;
		launch_classify<Sample>(c);
	}
	c->counted_set = set; c->emits_since_count[set + 1] = 0;
	c->state_iso[set + 1] = P.iso;
	c->h_totals = c->h_all + (set + 1);
	c->counted_mask |= 1u << (set + 1); c->pending_mask |= 1u << (set + 1); c->hvalid_mask &= ~(1u << (set + 1));
	if (c->timing) CU(cudaEventRecord(c->ev[1], s));
	const uint32_t owned_end = (P.pz1 - P.zlo) * P.NY;
	if (c->pipe == 1) {
		uint32_t grid = (uint32_t)c->n_sm * 8;
		if (grid > c->nblk) grid = c->nblk;
		k_count<Sample><<<grid, 256, 0, s>>>(P, c->nblk, c->cnt_gw, c->blk_sum);
		c->launches++;
		if (c->timing) CU(cudaEventRecord(c->ev[2], s));
		k_rowscan<<<(c->nblk + RS_BLOCKS - 1) / RS_BLOCKS, 256, 0, s>>>(P, c->nblk, c->cnt_rb, c->blk_sum, owned_end);
		c->launches++;
		if (c->export4) { k_export_counts<<<1, 1, 0, s>>>(P.totals, c->export4); c->launches++; }
	} else {
		// one kernel: counts, row bases by decoupled look-back, totals, optional export of the counts
		if (++c->lb_tag > 0xFFFFu) {
			// the 16-bit tag wraps: words of 65535 launches ago must not match
			CU(cudaMemsetAsync(c->lb0, 0, (size_t)c->nblk * P2_LB_WORDS * 8, s));
			if (c->swLb) CU(cudaMemsetAsync(c->swLb, 0, (size_t)c->nblk * P2_LB_WORDS * 8 * SWEEP_MAX, s));
			c->lb_tag = 1;
		}
		CountArgs A;
		A.nblk = c->nblk; A.GW = c->cnt_gw; A.lb = c->lb_cur; A.lb_ticket = c->lbt_cur; A.tag = c->lb_tag;
		A.owned_end_row = owned_end; A.export4 = c->export4; A.dbg_noprefix = 0;
		k_count2<Sample><<<c->nblk, 256, P2_CNT_SMEM, s>>>(P, A);
		c->launches++;
		if (c->timing) CU(cudaEventRecord(c->ev[2], s));
	}
	c->export4 = nullptr;
	if (c->timing) CU(cudaEventRecord(c->ev[3], s));
	CU(cudaGetLastError());
	return MC33CU_OK;
}

template <typename Sample> static int launch_emit_phase(mc33cu_ctx *c)
{
	Params &P = c->P;
	cudaStream_t s = c->stream;
	// (the count phase re-arms the totals; a second emit of the same count must not see the first one's overflow flag)
	if (c->emits_since_count[c->cur_state + 1]++) CU(cudaMemsetAsync(&P.totals->overflow, 0, sizeof(uint32_t), s));
	c->pending_mask |= 1u << (c->cur_state + 1); c->hvalid_mask &= ~(1u << (c->cur_state + 1));
	{
		// cell rows of the slab, plus the point rows above them whose vertices it owns
		// (the grid's last slice on the last slab)
		const uint32_t zend = P.pz1 > P.cz1 ? P.pz1 : P.cz1;
		const uint32_t rb = (P.cz0 - P.zlo) * P.NY, re = (zend - P.zlo) * P.NY;
		// work units: whole groups of G rows first, the last rows (about one group per resident
		// warp's worth ... c->fine_pct per cent) in units of gfine rows
		const uint32_t nrows = re - rb;
		uint32_t gfine = P.G >= 4 ? (c->fine_rows < P.G ? c->fine_rows : P.G) : P.G;
		uint32_t ncoarse = (uint32_t)((uint64_t)nrows * (100u - c->fine_pct) / 100u) / P.G;
		// a small grid has fewer whole groups than the GPU has warps to run them (201^3: 2525 groups of 16 rows against
		// 4736 warp slots), and the kernel takes as long as its busiest warp: hand out all of it in small units
		// (cfg1: cells 0.037 -> 0.029 ms with units of 8 rows)
		if (!c->fine_env && nrows / P.G < 4u * (uint32_t)c->n_sm * c->emc_per_sm * EM_WARPS && P.G >= 4) {
			gfine = P.G < 8u ? P.G : 8u;
			ncoarse = 0;
		}
		if (gfine == P.G) ncoarse = nrows / P.G;
		const uint32_t ngroups = ncoarse + (nrows - ncoarse * P.G + gfine - 1) / gfine;
		uint32_t grid = (uint32_t)c->n_sm * c->emc_per_sm;
		if (grid > (ngroups + EM_WARPS - 1) / EM_WARPS) grid = (ngroups + EM_WARPS - 1) / EM_WARPS;
		const uint32_t nquads = P.Lrows * P.Q;
		int which = c->pipe == 1 ? 1 : c->cells;                  // 1 direct kernel, 2 record kernel, 0 both (the device picks)
		if (which == 0)
			for (int i = 0; i < 4; i++)
				if (c->pick_val[c->cur_state + 1][i] && c->pick_iso[c->cur_state + 1][i] == P.iso) which = c->pick_val[c->cur_state + 1][i];
		if (which != 2) {
			const uint32_t pk = which == 0 ? nquads : 0u;
			if (P.vkey || P.tcell) k_emit_cells<Sample, true><<<grid, 256, EMC_SMEM(P.G), s>>>(P, rb, re, ngroups, ncoarse, gfine, pk);
			else k_emit_cells<Sample, false><<<grid, 256, EMC_SMEM(P.G), s>>>(P, rb, re, ngroups, ncoarse, gfine, pk);
			if (which == 0) c->launches++;
		}
		if (which != 1) {
			EmitArgs A;
			A.f = c->emit_shape.f; A.z = c->emit_shape.z;
			A.pick = which == 0 ? 1u : 0u; A.nquads = nquads;
			A.row_begin = rb; A.row_end = re;
			A.ngroups = (nrows + A.f.Ge - 1) / A.f.Ge;
			A.nunits = (A.ngroups + P2_EM_UNIT - 1) / P2_EM_UNIT;
			uint32_t g2 = (uint32_t)c->n_sm * c->emc2_per_sm;
			if (g2 > (A.nunits + P2_EM_WARPS - 1) / P2_EM_WARPS) g2 = (A.nunits + P2_EM_WARPS - 1) / P2_EM_WARPS;
			if (g2 < 1) g2 = 1;
			if (P.vkey || P.tcell) k_emit_cells2<Sample, true><<<g2, 256, EMC2_SMEM, s>>>(P, A);
			else k_emit_cells2<Sample, false><<<g2, 256, EMC2_SMEM, s>>>(P, A);
		}
		c->launches++;
	}
	if (c->dlT) {
		// mc33cu_emit_host_async: the triangles are complete; their download runs on the second stream
		// while the vertex kernel runs on this one
		CU(cudaEventRecord(c->ev_cells, s));
		CU(cudaStreamWaitEvent(c->copy_stream, c->ev_cells, 0));
		CU(cudaMemcpyAsync(c->dlT, P.T, c->dlT_words * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->copy_stream));
		c->copy_pending = true;
		c->dlT = nullptr;
	}
	if (c->timing) CU(cudaEventRecord(c->ev[4], s));
	{
		if (c->pipe == 2 && c->vtx == 2) {
			// shared vertices straight from the bitmaps, row group by row group
			VertexArgs A;
			A.row_begin = (P.pz0 - P.zlo) * P.NY; A.row_end = (P.pz1 - P.zlo) * P.NY;
			A.Gv = vertex_group_rows(P.Q);
			A.ngroups = (A.row_end - A.row_begin + A.Gv - 1) / A.Gv;
			uint32_t g2 = (uint32_t)c->n_sm * c->emv2_per_sm;
			if (g2 > (A.ngroups + P2_VX_WARPS - 1) / P2_VX_WARPS) g2 = (A.ngroups + P2_VX_WARPS - 1) / P2_VX_WARPS;
			if (g2 < 1) g2 = 1;
			if (P.vkey) k_emit_vertices2<Sample, true><<<g2, 256, EMV2_SMEM, s>>>(P, A);
			else k_emit_vertices2<Sample, false><<<g2, 256, EMV2_SMEM, s>>>(P, A);
		} else {
			// dense over the vertex ids (the count is only known on the device: persistent grid)
			if (P.vkey) k_emit_vertices<Sample, true><<<(uint32_t)c->n_sm * c->emv_per_sm, 256, 0, s>>>(P);
			else k_emit_vertices<Sample, false><<<(uint32_t)c->n_sm * c->emv_per_sm, 256, 0, s>>>(P);
		}
		c->launches++;
	}
	if (c->timing) { CU(cudaEventRecord(c->ev[5], s)); c->ev_valid = true; }
	if (P.vtask) CU(cudaEventRecord(c->vt[c->vt_cur].done, s));     // (mc33cu_sync and a slot that changes hands wait for it)
	CU(cudaGetLastError());
	return MC33CU_OK;
}

static int dispatch_count(mc33cu_ctx *c, int set = -1)
{
	switch (c->d.dtype) {
	case MC33CU_F32: return launch_count_phase<float>(c, set);
	case MC33CU_F64: return launch_count_phase<double>(c, set);
	case MC33CU_U8:  return launch_count_phase<uint8_t>(c, set);
	case MC33CU_U16: return launch_count_phase<uint16_t>(c, set);
	default:         return launch_count_phase<uint32_t>(c, set);
	}
}
static int dispatch_emit(mc33cu_ctx *c)
{
	switch (c->d.dtype) {
	case MC33CU_F32: return launch_emit_phase<float>(c);
	case MC33CU_F64: return launch_emit_phase<double>(c);
	case MC33CU_U8:  return launch_emit_phase<uint8_t>(c);
	case MC33CU_U16: return launch_emit_phase<uint16_t>(c);
	default:         return launch_emit_phase<uint32_t>(c);
	}
}

static void set_iso(mc33cu_ctx *c, double iso)
{
	// iso in MC33_real; -0.0 is folded onto +0.0 (DESIGN.md "isovalue -0.0")
	Params &P = c->P;
	if (c->d.dtype == MC33CU_F64) { P.iso = iso + 0.0; return; }
	const float f = (float)iso + 0.0f;
	P.iso = (double)f;
	if (c->d.dtype == MC33CU_F32) return;
	// integer grids: the classify kernel compares samples against integer
	// thresholds that reproduce (float)F > iso and (float)F == iso exactly
	// (the conversion is monotone, so both sets are intervals)
	const uint64_t top = c->d.dtype == MC33CU_U8 ? 0x100ull : c->d.dtype == MC33CU_U16 ? 0x10000ull : 0x100000000ull;
	auto first = [&](bool strict) {          // smallest F in [0,top] with (float)F > iso / >= iso
		uint64_t lo = 0, hi = top;
		while (lo < hi) {
			uint64_t mid = (lo + hi) >> 1;
			float v = (float)(uint32_t)mid;
			bool ok = strict ? (v > f) : (v >= f);
			if (ok) hi = mid; else lo = mid + 1;
		}
		return lo;
	};
	const uint64_t gt0 = first(true), ge0 = first(false);
	P.inone = gt0 >= 0x100000000ull;
	P.ithr = P.inone ? 0xFFFFFFFFu : (uint32_t)gt0;
	if (ge0 < gt0) { P.ieq_lo = (uint32_t)ge0; P.ieq_hi = (uint32_t)(gt0 - 1); }
	else { P.ieq_lo = 1; P.ieq_hi = 0; }
}

static int set_out(mc33cu_ctx *c, const mc33cu_out *o)
{
	Params &P = c->P;
	const bool need_tasks = !(c->pipe == 2 && c->vtx == 2);
	P.vtask = nullptr;
	if (need_tasks) {
		int k = 0;
		while (k < c->n_vt && c->vt[k].s != c->stream) k++;
		if (k == c->n_vt) {
			if (k < MC33CU_MAX_STREAMS) {
				c->vt[k].p = nullptr; c->vt[k].cap = 0;
				CU(cudaEventCreateWithFlags(&c->vt[k].done, cudaEventDisableTiming));
				c->n_vt++;
			} else {
				// more streams than slots: the slot used longest ago changes hands once its last extraction is through
				k = 0;
				for (int i = 1; i < c->n_vt; i++) if (c->vt[i].stamp < c->vt[k].stamp) k = i;
				CU(cudaEventSynchronize(c->vt[k].done));
			}
			c->vt[k].s = c->stream;
		}
		c->vt[k].stamp = ++c->vt_clock;
		c->vt_cur = k;
		if (o->capV > c->vt[k].cap) {
			// (first extraction on this stream, or a larger output than any before: not on the steady-state path)
			CU(cudaStreamSynchronize(c->stream));
			cudaFree(c->vt[k].p);
			c->vt[k].p = nullptr; c->vt[k].cap = 0;
			const uint64_t cap = (uint64_t)o->capV + o->capV / 8 + 1024;
			CU(cudaMalloc((void **)&c->vt[k].p, cap * sizeof(uint64_t)));
			c->vt[k].cap = cap;
		}
		P.vtask = c->vt[k].p;
	}
	P.V = o->V; P.N = o->N; P.color = o->color; P.T = o->T; P.vkey = o->vkey; P.tcell = o->tcell;
	P.capV = o->capV; P.capT = o->capT;
	P.vbase = o->vbase; P.vbase_next = o->vbase_next; P.dbases = o->dev_bases;
	P.color_value = o->color_value;
	return MC33CU_OK;
}

static Totals *state_totals(mc33cu_ctx *c, int state) { return state < 0 ? c->totals0 : c->swTotals + state; }

// Bring the totals (counts, overflow / range flags) of every state with launches since the last
// fetch to the host, plus those of the last counted state; one synchronisation.  *flags = OR of
// {1: range, 2: overflow} over the fetched states.
static int fetch_totals(mc33cu_ctx *c, uint32_t *flags = nullptr)
{
	uint32_t want = c->pending_mask | (c->counted ? 1u << (c->counted_set + 1) : 0u);
	if (!c->counted && !want) want = 1u;
	want &= ~c->hvalid_mask | c->pending_mask;
	// (extractions may have been issued on other streams of this context: their totals are read after them)
	for (int i = 0; i < c->n_vt; i++)
		if (c->vt[i].s != c->stream) CU(cudaEventSynchronize(c->vt[i].done));
	for (int st = -1; st < SWEEP_MAX; st++)
		if ((want >> (st + 1)) & 1u) {
			if (st >= 0 && !c->swTotals) continue;
			CU(cudaMemcpyAsync(c->h_all + (st + 1), state_totals(c, st), sizeof(Totals), cudaMemcpyDeviceToHost, c->stream));
		}
	CU(cudaStreamSynchronize(c->stream));
	uint32_t f = 0;
	for (int st = -1; st < SWEEP_MAX; st++)
		if ((want >> (st + 1)) & 1u) {
			const Totals &t = c->h_all[st + 1];
			if (t.range) f |= 1u;
			if (t.overflow) f |= 2u;
			if (c->pipe == 2 && ((c->counted_mask >> (st + 1)) & 1u)) {
				// remember which cell kernel the device statistic picks for this state's isovalue
				const int8_t pv = (uint64_t)t.nslow * 64u > (uint64_t)c->P.Lrows * c->P.Q ? 2 : 1;
				int slot = -1;
				for (int i = 0; i < 4; i++) if (c->pick_val[st + 1][i] && c->pick_iso[st + 1][i] == c->state_iso[st + 1]) slot = i;
				if (slot < 0) { slot = c->pick_pos[st + 1]; c->pick_pos[st + 1] = (uint8_t)((slot + 1) & 3); }
				c->pick_iso[st + 1][slot] = c->state_iso[st + 1]; c->pick_val[st + 1][slot] = pv;
			}
		}
	c->hvalid_mask |= want;
	c->pending_mask = 0;
	if (flags) *flags = f;
	return MC33CU_OK;
}

static void fill_counts(const mc33cu_ctx *c, mc33cu_counts *k)
{
	const Totals &t = *c->h_totals;
	k->nShared = t.nShared; k->nCentre = t.nCentre; k->nT = t.nT;
	k->nSharedHalo = t.nSharedAll - t.nShared;
	k->nV = (uint64_t)t.nShared + t.nCentre;
}

extern "C" int mc33cu_count(mc33cu_ctx *c, double iso, mc33cu_counts *k)
{
	if (!c) return fail(MC33CU_ERR_ARG, "null context");
	if (!c->P.data) return fail(MC33CU_ERR_STATE, "no grid bound");
	CU(cudaSetDevice(c->device));
	set_iso(c, iso);
	c->ev_valid = false;
	int rc = dispatch_count(c);
	if (rc) return rc;
	rc = fetch_totals(c);
	if (rc) return rc;
	c->counted = true;
	if (c->h_totals->range) return fail(MC33CU_ERR_RANGE, "more than 2^32-1 vertices or triangles");
	if (k) fill_counts(c, k);
	return MC33CU_OK;
}

__global__ void k_export_counts(const Totals *t, uint32_t *out4)
{
	out4[0] = t->nShared + t->nCentre; out4[1] = t->nT; out4[2] = t->nShared; out4[3] = t->nCentre;
}

extern "C" int mc33cu_count_async(mc33cu_ctx *c, double iso, uint32_t *dev_counts4)
{
	if (!c) return fail(MC33CU_ERR_ARG, "null context");
	if (!c->P.data) return fail(MC33CU_ERR_STATE, "no grid bound");
	CU(cudaSetDevice(c->device));
	set_iso(c, iso);
	c->ev_valid = false;
	c->export4 = dev_counts4;
	int rc = dispatch_count(c);
	if (rc) return rc;
	c->counted = true;
	return MC33CU_OK;
}

// (64-bit sums: the per-slab range check of the scan does not see an overflow of the GLOBAL vertex count)
__global__ void k_slab_bases(const uint32_t *all4, uint32_t stride, int rank, int world, uint32_t *bases2, Totals *tot)
{
	uint64_t v = 0, all = 0;
	for (int r = 0; r < world; r++) {
		if (r == rank) { bases2[0] = (uint32_t)v; bases2[1] = (uint32_t)(v + all4[(size_t)stride * r]); }
		if (r <= rank) v += all4[(size_t)stride * r];
		all += all4[(size_t)stride * r];
	}
	if (all >= 0xFFFFFFFFull) tot->range = 1u;
}

extern "C" int mc33cu_slab_bases_strided(mc33cu_ctx *c, const uint32_t *dev_counts_all, uint32_t stride_words, int rank, int world,
                                         uint32_t *dev_bases2)
{
	if (!c || !dev_counts_all || !dev_bases2) return fail(MC33CU_ERR_ARG, "null argument");
	if (world < 1 || rank < 0 || rank >= world || stride_words < 4) return fail(MC33CU_ERR_ARG, "bad rank / world / stride");
	CU(cudaSetDevice(c->device));
	k_slab_bases<<<1, 1, 0, c->stream>>>(dev_counts_all, stride_words, rank, world, dev_bases2, c->P.totals);
	c->pending_mask |= 1u << (c->cur_state + 1);
	c->launches++;
	CU(cudaGetLastError());
	return MC33CU_OK;
}

extern "C" int mc33cu_slab_bases(mc33cu_ctx *c, const uint32_t *dev_counts_all, int rank, int world, uint32_t *dev_bases2)
{
	if (!c || !dev_counts_all || !dev_bases2) return fail(MC33CU_ERR_ARG, "null argument");
	if (world < 1 || rank < 0 || rank >= world) return fail(MC33CU_ERR_ARG, "bad rank / world");
	CU(cudaSetDevice(c->device));
	k_slab_bases<<<1, 1, 0, c->stream>>>(dev_counts_all, 4u, rank, world, dev_bases2, c->P.totals);
	c->pending_mask |= 1u << (c->cur_state + 1);
	c->launches++;
	CU(cudaGetLastError());
	return MC33CU_OK;
}

extern "C" int mc33cu_emit_device(mc33cu_ctx *c, const mc33cu_out *o)
{
	if (!c || !o) return fail(MC33CU_ERR_ARG, "null argument");
	if (!c->counted || !((c->counted_mask >> (c->counted_set + 1)) & 1u))
		return fail(MC33CU_ERR_STATE, "mc33cu_count has not run (or its sweep set was re-classified since)");
	CU(cudaSetDevice(c->device));
	select_state(c, c->counted_set);               // the mesh of the LAST count, whatever ran in between
	int rc = set_out(c, o);
	if (rc) return rc;
	return dispatch_emit(c);
}

extern "C" int mc33cu_extract_device(mc33cu_ctx *c, double iso, const mc33cu_out *o)
{
	if (!c || !o) return fail(MC33CU_ERR_ARG, "null argument");
	if (!c->P.data) return fail(MC33CU_ERR_STATE, "no grid bound");
	CU(cudaSetDevice(c->device));
	set_iso(c, iso);
	int rc = set_out(c, o);
	if (rc) return rc;
	rc = dispatch_count(c);
	if (rc) return rc;
	rc = dispatch_emit(c);
	if (rc) return rc;
	c->counted = true;
	return MC33CU_OK;
}

// ---------------------------------------------------------------------------
// iso sweep: classify once for up to SWEEP_MAX isovalues, then count / emit per set
// ---------------------------------------------------------------------------
template <typename Sample> static int classify_one_into_set(mc33cu_ctx *c, int j)
{
	// general shapes / element types: the single-isovalue kernel, once per set
	select_state(c, j);
	return launch_classify<Sample>(c);
}

extern "C" int mc33cu_classify_sweep(mc33cu_ctx *c, const double *isos, int n)
{
	if (!c || !isos) return fail(MC33CU_ERR_ARG, "null argument");
	if (n < 1 || n > SWEEP_MAX) return fail(MC33CU_ERR_ARG, "a sweep holds 1..8 isovalues");
	if (!c->P.data) return fail(MC33CU_ERR_STATE, "no grid bound");
	CU(cudaSetDevice(c->device));
	Params &P = c->P;
	cudaStream_t s = c->stream;
	const size_t bm = (size_t)P.Lrows * P.WP;
	if (!c->sweep_ready) {
		// (first sweep on this context: not on the steady-state path).  All or nothing: a failed
		// allocation releases what the earlier ones got, so that a later call starts over.
		CU(cudaStreamSynchronize(s));
		struct { void **p; size_t bytes; bool zero; } al[] = {
			{(void **)&c->swS, bm * 4 * SWEEP_MAX, true}, {(void **)&c->swZ, bm * 4 * SWEEP_MAX, true},
			{(void **)&c->swRowZ, (size_t)P.Lrows * 4 * SWEEP_MAX, true}, {(void **)&c->swAny, 4 * SWEEP_MAX, true},
			{(void **)&c->swD, c->dwords * 4 * SWEEP_MAX, true}, {(void **)&c->swA, bm * 4 * SWEEP_MAX, true},
			{(void **)&c->swWpre, bm * 8 * SWEEP_MAX, true}, {(void **)&c->swRowB, ((size_t)P.Lrows + 1) * 3 * 4 * SWEEP_MAX, false},
			{(void **)&c->swTotals, sizeof(Totals) * SWEEP_MAX, true}, {(void **)&c->swBlk, (size_t)c->nblk * 3 * 4 * SWEEP_MAX, false},
			{(void **)&c->swLb, (size_t)c->nblk * P2_LB_WORDS * 8 * SWEEP_MAX, true}, {(void **)&c->swLbt, 8 * SWEEP_MAX, true},
			{(void **)&c->swPcache, bm * 32 * sizeof(uint16_t) * SWEEP_MAX, false}};
		cudaError_t e = cudaSuccess;
		for (auto &a : al) {
			*a.p = nullptr;
			e = cudaMalloc(a.p, a.bytes);
			if (e == cudaSuccess && a.zero) e = cudaMemsetAsync(*a.p, 0, a.bytes, s);
			if (e != cudaSuccess) break;
		}
		if (e != cudaSuccess) {
			for (auto &a : al) { cudaFree(*a.p); *a.p = nullptr; }
			cudaGetLastError();
			return fail(e == cudaErrorMemoryAllocation ? MC33CU_ERR_NOMEM : MC33CU_ERR_CUDA, "sweep sets: %s", cudaGetErrorString(e));
		}
		c->sweep_ready = true;
	}
	int rc = next_epoch(c, &c->sw_epoch);
	if (rc) return rc;
	c->sw_n = n;
	c->counted_mask &= 1u;            // the sets' earlier counts belong to the previous sweep's bitmaps
	// (pick entries are keyed by isovalue, so a set that gets another isovalue simply misses)
	c->hvalid_mask &= 1u;
	for (int j = 0; j < n; j++) c->sw_iso[j] = isos[j];
	c->ev_valid = false;
	const ClsPlan &pl = c->cls;
	// (one-pass kernel: float rows of whole 128-sample groups; a warp's groups of one chunk must fit the
	// two dirty words it prefetches: a chunk holds at most 33 groups, far below the 32 per warp allowed)
	const uint32_t sweep_ipw = (pl.rows * (P.W / 4) + SWEEP_THREADS / 32 - 1) / (SWEEP_THREADS / 32);
	if (c->d.dtype == MC33CU_F32 && pl.nwchunk == 1 && P.NX % 128 == 0 && ((uintptr_t)P.data & 15) == 0 && sweep_ipw <= 32) {
		SweepSets ss;
		for (int j = 0; j < SWEEP_MAX; j++) ss.iso[j] = j < n ? (float)isos[j] + 0.0f : __builtin_inff();
		ss.S = c->swS; ss.Z = c->swZ; ss.rowZ = c->swRowZ; ss.any = c->swAny; ss.set_words = bm;
		ss.D = c->swD; ss.dwords = c->dwords;
		for (int j = 0; j < SWEEP_MAX; j++) {
			if (c->zmode_sw[j] == 2) {
				CU(cudaMemsetAsync(c->swZ + (size_t)j * bm, 0, bm * 4, s));
				CU(cudaMemsetAsync(c->swD + (size_t)j * c->dwords, 0, c->dwords * 4, s));
			}
			c->zmode_sw[j] = 1;
		}
		P.zepoch = c->sw_epoch;
		const size_t smem = (size_t)pl.stage_bytes * CLS_STAGES;
		uint32_t per_sm = (uint32_t)((220u << 10) / (smem + 1024));
		if (per_sm < 1) per_sm = 1;
		if (per_sm > 8) per_sm = 8;
		uint32_t grid = (uint32_t)c->n_sm * per_sm;
		if (grid > pl.nchunks) grid = pl.nchunks;
		k_classify_sweep<<<grid, SWEEP_THREADS, smem, s>>>(P, ss, pl.rows, pl.nchunks, pl.stage_bytes);
		c->launches++;
	} else {
		for (int j = 0; j < n; j++) {
			set_iso(c, isos[j]);
			switch (c->d.dtype) {
			case MC33CU_F32: classify_one_into_set<float>(c, j); break;
			case MC33CU_F64: classify_one_into_set<double>(c, j); break;
			case MC33CU_U8:  classify_one_into_set<uint8_t>(c, j); break;
			case MC33CU_U16: classify_one_into_set<uint16_t>(c, j); break;
			default:         classify_one_into_set<uint32_t>(c, j); break;
			}
		}
	}
	select_state(c, -1);
	CU(cudaGetLastError());
	return MC33CU_OK;
}

static int check_set(mc33cu_ctx *c, int set)
{
	if (!c) return fail(MC33CU_ERR_ARG, "null context");
	if (set < 0 || set >= c->sw_n) return fail(MC33CU_ERR_STATE, "no such pre-classified set: call mc33cu_classify_sweep first");
	return MC33CU_OK;
}

extern "C" int mc33cu_count_set_async(mc33cu_ctx *c, int set, uint32_t *dev_counts4)
{
	int rc = check_set(c, set);
	if (rc) return rc;
	CU(cudaSetDevice(c->device));
	set_iso(c, c->sw_iso[set]);
	c->ev_valid = false;
	c->export4 = dev_counts4;
	rc = dispatch_count(c, set);
	if (rc) return rc;
	c->counted = true;
	return MC33CU_OK;
}

extern "C" int mc33cu_emit_set_device(mc33cu_ctx *c, int set, const mc33cu_out *o)
{
	int rc = check_set(c, set);
	if (rc) return rc;
	if (!o) return fail(MC33CU_ERR_ARG, "null argument");
	if (!((c->counted_mask >> (set + 1)) & 1u)) return fail(MC33CU_ERR_STATE, "the set has not been counted since the last mc33cu_classify_sweep");
	CU(cudaSetDevice(c->device));
	set_iso(c, c->sw_iso[set]);
	select_state(c, set);
	rc = set_out(c, o);
	if (rc) return rc;
	return dispatch_emit(c);
}

extern "C" int mc33cu_extract_set_device(mc33cu_ctx *c, int set, const mc33cu_out *o)
{
	int rc = check_set(c, set);
	if (rc) return rc;
	if (!o) return fail(MC33CU_ERR_ARG, "null argument");
	CU(cudaSetDevice(c->device));
	set_iso(c, c->sw_iso[set]);
	rc = set_out(c, o);
	if (rc) return rc;
	rc = dispatch_count(c, set);
	if (rc) return rc;
	rc = dispatch_emit(c);
	if (rc) return rc;
	c->counted = true;
	return MC33CU_OK;
}

extern "C" int mc33cu_sync(mc33cu_ctx *c)
{
	if (!c) return fail(MC33CU_ERR_ARG, "null context");
	CU(cudaSetDevice(c->device));
	// every state (single-isovalue path, sweep sets) counted or emitted since the last sync is looked at:
	// an overflow of set 0 is not hidden by a later emit of set 1
	uint32_t flags = 0;
	int rc = fetch_totals(c, &flags);
	if (rc) return rc;
	if (c->copy_pending) { CU(cudaStreamSynchronize(c->copy_stream)); c->copy_pending = false; }
	if (flags & 1u) return fail(MC33CU_ERR_RANGE, "more than 2^32-1 vertices or triangles");
	if (flags & 2u) return fail(MC33CU_ERR_CAPACITY, "output capacity exceeded");
	return MC33CU_OK;
}

extern "C" int mc33cu_get_counts(mc33cu_ctx *c, mc33cu_counts *k)
{
	if (!c || !k) return fail(MC33CU_ERR_ARG, "null argument");
	if (!c->counted) return fail(MC33CU_ERR_STATE, "nothing counted yet");
	if (!((c->hvalid_mask >> (c->counted_set + 1)) & 1u)) {   // counted without a host fetch (*_async): fetch now
		int rc = fetch_totals(c);
		if (rc) return rc;
	}
	fill_counts(c, k);
	return MC33CU_OK;
}

static int emit_host_impl(mc33cu_ctx *c, void *V, float *N, int32_t *color, uint32_t *T, uint32_t vbase, uint32_t vbase_next,
                          int32_t color_value, bool wait)
{
	if (!c) return fail(MC33CU_ERR_ARG, "null context");
	if (!c->counted || !((c->counted_mask >> (c->counted_set + 1)) & 1u))
		return fail(MC33CU_ERR_STATE, "mc33cu_count has not run (or its sweep set was re-classified since)");
	CU(cudaSetDevice(c->device));
	select_state(c, c->counted_set);               // the mesh of the LAST count
	if (!((c->hvalid_mask >> (c->counted_set + 1)) & 1u)) {
		int rc0 = fetch_totals(c);
		if (rc0) return rc0;
	}
	mc33cu_counts k;
	fill_counts(c, &k);
	if (k.nV == 0 && k.nT == 0) return MC33CU_OK;
	if (!V || !N || !color || !T) return fail(MC33CU_ERR_ARG, "null output array");
	if (c->ocapV < k.nV) {
		CU(cudaStreamSynchronize(c->stream));
		cudaFree(c->oV); cudaFree(c->oN); cudaFree(c->oC);
		c->oV = nullptr; c->oN = nullptr; c->oC = nullptr; c->ocapV = 0;
		const uint64_t cap = k.nV + k.nV / 4 + 1024;      // head room: an iso sweep does not re-allocate at every step
		CU(cudaMalloc(&c->oV, cap * 3 * c->real_size));
		CU(cudaMalloc((void **)&c->oN, cap * 3 * sizeof(float)));
		CU(cudaMalloc((void **)&c->oC, cap * sizeof(int32_t)));
		c->ocapV = cap;
	}
	if (c->ocapT < k.nT) {
		CU(cudaStreamSynchronize(c->stream));
		cudaFree(c->oT);
		c->oT = nullptr; c->ocapT = 0;
		const uint64_t cap = k.nT + k.nT / 4 + 1024;
		CU(cudaMalloc((void **)&c->oT, cap * 3 * sizeof(uint32_t)));
		c->ocapT = cap;
	}
	if (!c->copy_stream) {
		CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
		CU(cudaEventCreateWithFlags(&c->ev_cells, cudaEventDisableTiming));
		CU(cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming));
	}
	mc33cu_out o;
	memset(&o, 0, sizeof o);
	o.V = c->oV; o.N = c->oN; o.color = c->oC; o.T = c->oT;
	o.capV = (uint32_t)k.nV; o.capT = (uint32_t)k.nT;
	o.vbase = vbase; o.vbase_next = vbase_next;
	o.color_value = color_value;
	int rc = set_out(c, &o);
	if (rc) return rc;
	c->dlT = T; c->dlT_words = k.nT * 3;
	rc = dispatch_emit(c);
	c->dlT = nullptr;
	if (rc) return rc;
	cudaStream_t s = c->stream;
	CU(cudaMemcpyAsync(V, c->oV, k.nV * 3 * c->real_size, cudaMemcpyDeviceToHost, s));
	CU(cudaMemcpyAsync(N, c->oN, k.nV * 3 * sizeof(float), cudaMemcpyDeviceToHost, s));
	CU(cudaMemcpyAsync(color, c->oC, k.nV * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
	if (wait) {
		CU(cudaStreamSynchronize(s));
		if (c->copy_pending) { CU(cudaStreamSynchronize(c->copy_stream)); c->copy_pending = false; }
	}
	return MC33CU_OK;
}

extern "C" int mc33cu_emit_host(mc33cu_ctx *c, void *V, float *N, int32_t *color, uint32_t *T, int32_t color_value)
{
	return emit_host_impl(c, V, N, color, T, 0u, 0u, color_value, true);
}

extern "C" int mc33cu_emit_host_async(mc33cu_ctx *c, void *V, float *N, int32_t *color, uint32_t *T, uint32_t vbase,
                                      uint32_t vbase_next, int32_t color_value)
{
	return emit_host_impl(c, V, N, color, T, vbase, vbase_next, color_value, false);
}

// Order what is issued on c's stream from now on behind everything issued so far on `after`'s stream (an event: the
// host does not wait).  The drop-in chains the uploads of the z-chunks that share a device, so that they cross the
// link one after the other and chunk k is counted, emitted and downloaded while chunk k+1 is still arriving.
extern "C" int mc33cu_stream_wait(mc33cu_ctx *c, mc33cu_ctx *after)
{
	if (!c || !after) return fail(MC33CU_ERR_ARG, "null context");
	if (c == after) return MC33CU_OK;
	CU(cudaSetDevice(after->device));
	if (!after->ev_chain) CU(cudaEventCreateWithFlags(&after->ev_chain, cudaEventDisableTiming));
	CU(cudaEventRecord(after->ev_chain, after->stream));
	CU(cudaSetDevice(c->device));
	CU(cudaStreamWaitEvent(c->stream, after->ev_chain, 0));
	return MC33CU_OK;
}

extern "C" int mc33cu_enable_timing(mc33cu_ctx *c, int on)
{
	if (!c) return fail(MC33CU_ERR_ARG, "null context");
	c->timing = on != 0;
	c->ev_valid = false;
	return MC33CU_OK;
}

extern "C" int mc33cu_kernel_times(mc33cu_ctx *c, float ms[5])
{
	if (!c || !ms) return fail(MC33CU_ERR_ARG, "null argument");
	if (!c->ev_valid) return fail(MC33CU_ERR_STATE, "no timed extraction");
	CU(cudaEventSynchronize(c->ev[5]));
	for (int i = 0; i < 5; i++) CU(cudaEventElapsedTime(&ms[i], c->ev[i], c->ev[i + 1]));
	return MC33CU_OK;
}

// measurement hook (tools/ only, not declared in include/mc33cu.h): the classify kernel alone
extern "C" int mc33cu_debug_classify(mc33cu_ctx *c, double iso)
{
	if (!c || !c->P.data) return MC33CU_ERR_ARG;
	set_iso(c, iso);
	switch (c->d.dtype) {
	case MC33CU_F32: return launch_classify<float>(c);
	case MC33CU_F64: return launch_classify<double>(c);
	case MC33CU_U8:  return launch_classify<uint8_t>(c);
	case MC33CU_U16: return launch_classify<uint16_t>(c);
	default:         return launch_classify<uint32_t>(c);
	}
}

extern "C" uint64_t mc33cu_launch_count(const mc33cu_ctx *c) { return c ? c->launches : 0; }
