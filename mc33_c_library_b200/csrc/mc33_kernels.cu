// mc33_kernels.cu -- sm_100a kernels and the C-ABI (include/mc33cu.h) of the
// B200 Marching Cubes 33 extractor.  See DESIGN.md for the pipeline:
//
//   K1 classify   stream the samples ONCE through a TMA (cp.async.bulk) ring in
//                 shared memory -> S / Z bitmaps (1 bit per sample)
//   K2 count      per (row, 32-point word): owned vertices per plane, triangles
//                 and centre vertices of the 32 cells -> row-local word prefixes;
//                 fused single-pass decoupled look-back scan over the batches
//                 -> per-row vertex / triangle / centre bases
//   K3 emit V     vertex tasks compacted per tile in shared memory, one thread
//                 per vertex, consecutive threads write consecutive vertices
//   K4 emit T     triangle tasks compacted the same way, one thread per triangle
//
// Replaces: reference source/marching_cubes_33.c:1816-1889 (calculate_isosurface),
// :673-1253 (MC33_findCase), :485-649 (store / surfint).  No CPU fallback.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mc33_core.cuh"
#include "../../include/mc33cu.h"

using namespace mc33;

// ---------------------------------------------------------------------------
// case tables in device memory (copied to shared memory by the kernels that
// index them divergently)
// ---------------------------------------------------------------------------
__device__ uint16_t d_case256[256];
__device__ uint16_t d_simple256[256];
__device__ uint16_t d_tri[MC33_NTRI_WORDS];
__device__ uint8_t d_pat[MC33_NTRI_WORDS];

#define TBL_TRI_BYTES ((MC33_NTRI_WORDS * 2 + 15) / 16 * 16)
#define TBL_PAT_BYTES ((MC33_NTRI_WORDS + 15) / 16 * 16)
#define TBL_BYTES (512 + 512 + TBL_TRI_BYTES + TBL_PAT_BYTES)

__device__ __forceinline__ Tables load_tables(unsigned char *smem)
{
	uint16_t *c = (uint16_t *)smem;
	uint16_t *s = c + 256;
	uint16_t *t = s + 256;
	uint8_t *p = (uint8_t *)t + TBL_TRI_BYTES;
	for (int i = threadIdx.x; i < 128; i += blockDim.x) {
		((uint32_t *)c)[i] = ((const uint32_t *)d_case256)[i];
		((uint32_t *)s)[i] = ((const uint32_t *)d_simple256)[i];
	}
	for (int i = threadIdx.x; i < MC33_NTRI_WORDS / 2; i += blockDim.x) ((uint32_t *)t)[i] = ((const uint32_t *)d_tri)[i];
	for (int i = threadIdx.x; i < MC33_NTRI_WORDS; i += blockDim.x) p[i] = d_pat[i];
	__syncthreads();
	Tables tb;
	tb.case256 = c; tb.simple256 = s; tb.tri = t; tb.pat = p;
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
	__device__ __forceinline__ Cls(const Params &P) : iso((typename Traits<Sample>::Real)P.iso) {}
	__device__ __forceinline__ bool gt(Sample f) const { return f > iso; }
	__device__ __forceinline__ bool eq(Sample f) const { return f == iso; }
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

#define CLS_THREADS 128
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
__global__ void __launch_bounds__(CLS_THREADS) k_classify(Params P, ClsPlan pl)
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
		mbar_wait(&full[s], (k / CLS_STAGES) & 1);
		const unsigned char *st = smem + (size_t)s * pl.stage_bytes + (((uint64_t)(uintptr_t)gbase + ck.b0) & 15);
		for (uint32_t rr = wid; rr < ck.nrows; rr += CLS_THREADS / 32) {
			const uint32_t lr = ck.lr0 + rr;
			// sample x of this row sits at src[x - 32*w0]
			const Sample *src = (const Sample *)st + (size_t)rr * P.NX + lane;
			uint32_t *Sr = P.S + (uint64_t)lr * P.WP + ck.w0;
			bool zl = false;                                  // this lane saw an on-iso sample
			for (uint32_t c0 = 0; c0 < ck.nw; c0 += 32) {     // 32 words per group: lane j keeps word c0+j
				const uint32_t wg = ck.w0 + c0;               // global index of the group's first word
				const uint32_t nwg = min(32u, ck.nw - c0);
				const uint32_t limf = nfull > wg ? min(nwg, nfull - wg) : 0u;   // words with 32 valid samples
				const Sample *q = src + ((size_t)c0 << 5);
				uint32_t sw = 0;
				uint32_t g = 0;
				for (; g + 4 <= limf; g += 4) {
					const Sample f0 = q[32 * g], f1 = q[32 * g + 32], f2 = q[32 * g + 64], f3 = q[32 * g + 96];
					const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, cls.gt(f0)), b1 = __ballot_sync(0xFFFFFFFFu, cls.gt(f1));
					const uint32_t b2 = __ballot_sync(0xFFFFFFFFu, cls.gt(f2)), b3 = __ballot_sync(0xFFFFFFFFu, cls.gt(f3));
					zl = zl || cls.eq(f0) || cls.eq(f1) || cls.eq(f2) || cls.eq(f3);
					sw = lane == g ? b0 : sw;
					sw = lane == g + 1 ? b1 : sw;
					sw = lane == g + 2 ? b2 : sw;
					sw = lane == g + 3 ? b3 : sw;
				}
				for (; g < limf; g++) {
					const Sample f = q[32 * g];
					const uint32_t b = __ballot_sync(0xFFFFFFFFu, cls.gt(f));
					zl = zl || cls.eq(f);
					sw = lane == g ? b : sw;
				}
				if (tail && g < nwg && wg + g == nfull) {      // the partial last word of the row
					const bool ok = lane < tail;
					const Sample f = ok ? q[32 * g] : (Sample)0;
					const uint32_t b = __ballot_sync(0xFFFFFFFFu, ok && cls.gt(f));
					zl = zl || (ok && cls.eq(f));
					sw = lane == g ? b : sw;
				}
				if (lane < nwg) Sr[c0 + lane] = sw;
			}
			const bool zany = __any_sync(0xFFFFFFFFu, zl);
			const bool whole = pl.nwchunk == 1;
			const bool zold = P.rowZ[lr] != 0;
			if (zany || zold || !whole) {                    // rare: (re)write this row's Z words
				uint32_t *Zr = P.Z + (uint64_t)lr * P.WP + ck.w0;
				for (uint32_t w = 0; w < ck.nw; w++) {
					const uint32_t x = ((ck.w0 + w) << 5) + lane;
					const bool ok = x < P.NX;
					const Sample f = ok ? src[(size_t)w << 5] : (Sample)0;
					const uint32_t b = __ballot_sync(0xFFFFFFFFu, ok && cls.eq(f));
					if (lane == 0) Zr[w] = b;
				}
				// pieces of one long row share the flag: it is only ever raised there
				if (lane == 0 && (whole || zany)) P.rowZ[lr] = zany;
				if (lane == 0 && zany) P.totals->anyZ = 1u;
			}
		}
		__syncthreads();                                     // every warp is done with stage s
		if (k + CLS_STAGES < nmine) issue(blockIdx.x + (k + CLS_STAGES) * gridDim.x, s);
	}
}

// ---------------------------------------------------------------------------
// block-wide exclusive scan of two packed 64-bit counters (256 threads)
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

// ---------------------------------------------------------------------------
// K2: count + scan.
//
// Batches of R whole rows (R*W <= 256, one thread per (row, word); a row longer
// than 256 words is one batch walked in passes) are taken in ticket order.  One
// block scan gives the row-local word prefixes; the batch totals then go through
// a single-pass decoupled look-back (status word = flag<<62 | epoch<<42 | value;
// flag 1 = batch aggregate, 2 = inclusive prefix; the epoch makes stale words of
// earlier extractions read as "not ready", so the array is never cleared), which
// yields the slab-local base of every row: the implicit running M->nV++ / nT++ of
// the reference (marching_cubes_33.c:487, :1245).
// ---------------------------------------------------------------------------
#define ST_FLAG(s) ((unsigned)((s) >> 62))
#define ST_EPOCH(s) ((uint32_t)((s) >> 42) & 0xFFFFFu)
#define ST_VAL(s) ((s) & ((1ull << 42) - 1))
#define ST_MAKE(flag, epoch, val) (((uint64_t)(flag) << 62) | ((uint64_t)((epoch) & 0xFFFFFu) << 42) | (val))

__device__ __forceinline__ uint32_t sumV(uint64_t p) { return fldV(p, 0) + fldV(p, 1) + fldV(p, 2); }

template <typename Sample>
__global__ void __launch_bounds__(256) k_count(Params P, uint64_t *status, uint32_t *ticket, uint32_t nbatch, uint32_t epoch,
                                               uint32_t owned_end_row)
{
	extern __shared__ __align__(128) unsigned char smem[];
	__shared__ uint64_t s_preV[257], s_preT[257];
	__shared__ uint64_t s_w[2][8];
	__shared__ uint64_t s_excl[3];
	__shared__ uint32_t s_batch;
	const Tables tb = load_tables(smem);
	const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const bool gz = P.totals->anyZ != 0;
	const uint32_t n = P.R * P.W;

	while (true) {
		if (threadIdx.x == 0) s_batch = atomicAdd(ticket, 1u);
		__syncthreads();
		const uint32_t batch = s_batch;
		if (batch >= nbatch) break;
		const uint32_t row0 = batch * P.R;
		uint64_t carryV = 0, carryT = 0;
		for (uint32_t base = 0; base < n; base += 256) {
			const uint32_t it = base + threadIdx.x;
			uint64_t cv = 0, ct = 0;
			uint32_t r = 0, w = 0, lr = 0xFFFFFFFFu;
			if (it < n) {
				r = P.R > 1 ? fastdiv(it, P.W, P.mCW) : 0u;
				w = it - r * P.W;
				if (row0 + r < P.Lrows) lr = row0 + r;
			}
			if (lr != 0xFFFFFFFFu) {
				const uint32_t zl = fastdiv(lr, P.NY, P.mNY), y = lr - zl * P.NY, z = zl + P.zlo;
				const bool own_p = row_points_owned(P, z) || row_points_halo(P, z);
				const bool own_c = row_cells_owned(P, z, y);
				if (own_p || own_c) {
					WordRec rec;
					count_word<Sample>(P, tb, z, y, w, gz, own_p, own_c, rec, cv, ct);
				}
			}
			const uint64_t mv = cv, mt = ct;
			uint64_t tv, tt;
			block_exscan2(cv, ct, tv, tt, s_w);
			cv += carryV; ct += carryT;
			s_preV[threadIdx.x] = cv; s_preT[threadIdx.x] = ct;
			__syncthreads();
			if (lr != 0xFFFFFFFFu) {
				const uint64_t rsV = P.R > 1 ? s_preV[r * P.W] : 0ull, rsT = P.R > 1 ? s_preT[r * P.W] : 0ull;
				uint64_t *pv = P.wpreV + (uint64_t)lr * P.W1, *pt = P.wpreT + (uint64_t)lr * P.W1;
				pv[w] = cv - rsV; pt[w] = ct - rsT;
				if (w == P.W - 1) { pv[P.W] = cv + mv - rsV; pt[P.W] = ct + mt - rsT; }
			}
			carryV += tv; carryT += tt;
			if (base + 256 < n) __syncthreads();
		}
		// batch aggregates -> look-back; warps 0..2 handle vertices / triangles / centres
		if (wid < 3) {
			const unsigned q = wid;
			const uint64_t agg = q == 0 ? (uint64_t)sumV(carryV) : (q == 1 ? (carryT & 0xFFFFFFFFull) : (carryT >> 32));
			volatile uint64_t *st = status;
			uint64_t excl = 0;
			if (batch == 0) {
				if (lane == 0) st[q] = ST_MAKE(2, epoch, agg);
			} else {
				if (lane == 0) st[3 * (uint64_t)batch + q] = ST_MAKE(1, epoch, agg);
				int64_t pos = (int64_t)batch - 1;
				while (true) {
					const int64_t idx = pos - (int64_t)lane;
					uint64_t s = ST_MAKE(2, epoch, 0);       // batches before the first one: prefix 0
					if (idx >= 0) {
						do { s = st[3 * (uint64_t)idx + q]; } while (ST_FLAG(s) == 0 || ST_EPOCH(s) != (epoch & 0xFFFFFu));
					}
					const unsigned pre_mask = __ballot_sync(0xFFFFFFFFu, ST_FLAG(s) == 2);
					// lanes up to and including the first inclusive prefix contribute
					const unsigned first = pre_mask ? (unsigned)__ffs((int)pre_mask) - 1u : 32u;
					uint64_t v = (lane <= first) ? ST_VAL(s) : 0;
#pragma unroll
					for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
					excl += v;
					if (pre_mask) break;
					pos -= 32;
				}
				if (lane == 0) st[3 * (uint64_t)batch + q] = ST_MAKE(2, epoch, excl + agg);
			}
			if (lane == 0) s_excl[q] = excl;
		}
		__syncthreads();
		const uint64_t eV = s_excl[0], eT = s_excl[1], eC = s_excl[2];
		if (threadIdx.x < P.R && row0 + threadIdx.x < P.Lrows) {
			const uint32_t lr = row0 + threadIdx.x;
			const uint64_t rsV = P.R > 1 ? s_preV[threadIdx.x * P.W] : 0ull, rsT = P.R > 1 ? s_preT[threadIdx.x * P.W] : 0ull;
			const uint64_t bv = eV + sumV(rsV);
			P.rowBV[lr] = (uint32_t)bv;
			P.rowBT[lr] = (uint32_t)(eT + (rsT & 0xFFFFFFFFull));
			P.rowBC[lr] = (uint32_t)(eC + (rsT >> 32));
			if (lr == owned_end_row) P.totals->nShared = (uint32_t)bv;
		}
		if (batch == nbatch - 1 && threadIdx.x == 0) {
			const uint64_t tv = eV + sumV(carryV), tt = eT + (carryT & 0xFFFFFFFFull), tc = eC + (carryT >> 32);
			P.rowBV[P.Lrows] = (uint32_t)tv; P.rowBT[P.Lrows] = (uint32_t)tt; P.rowBC[P.Lrows] = (uint32_t)tc;
			if (owned_end_row >= P.Lrows) P.totals->nShared = (uint32_t)tv;
			P.totals->nCentre = (uint32_t)tc;
			P.totals->nT = (uint32_t)tt;
			P.totals->nSharedAll = (uint32_t)tv;
			// 32-bit index range check (include/marching_cubes_33.h:140 uses unsigned int)
			P.totals->range = (tv + tc >= 0xFFFFFFFFull || tt >= 0xFFFFFFFFull) ? 1u : 0u;
		}
		__syncthreads();
	}
}

// ---------------------------------------------------------------------------
// K3: vertices.  A tile = R consecutive point rows; its vertices are the id range
// [rowBV[first row], rowBV[last row + 1]).  Pass 1: one thread per (row, word)
// that owns vertices (known from the word prefixes without touching the bitmaps)
// recomputes its plane masks and drops one task per vertex into the shared-memory
// list at slot id - window start.  Pass 2: thread t computes vertex
// window start + t, so consecutive threads write consecutive V / N / color
// entries.  Tiles with more vertices than the list holds take several windows.
// task word: x | plane << 16 | is_point << 18 | (row in tile) << 19
// ---------------------------------------------------------------------------
#define VCAP 2048

template <typename Sample>
__global__ void __launch_bounds__(256) k_emit_vertices(Params P, uint32_t row_begin, uint32_t row_end, uint32_t ntiles)
{
	__shared__ uint32_t vlist[VCAP];
	const bool gz = P.totals->anyZ != 0;
	const uint32_t n = P.R * P.W;
	for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
		const uint32_t lr0 = row_begin + tile * P.R, lrE = min(lr0 + P.R, row_end);
		const uint32_t vfirst = P.rowBV[lr0], vend = P.rowBV[lrE];
		for (uint32_t win0 = vfirst; win0 < vend; win0 += VCAP) {
			for (uint32_t it = threadIdx.x; it < n; it += 256) {
				const uint32_t r = P.R > 1 ? fastdiv(it, P.W, P.mCW) : 0u, w = it - r * P.W, lr = lr0 + r;
				if (lr >= lrE) continue;
				const uint64_t *pv = P.wpreV + (uint64_t)lr * P.W1;
				const uint64_t pre0 = pv[w], pre1 = pv[w + 1];
				if (pre0 == pre1) continue;
				const uint64_t tot = pv[P.W];
				uint32_t id[3], cnt[3];
				id[0] = P.rowBV[lr] + fldV(pre0, 0);
				id[1] = P.rowBV[lr] + fldV(tot, 0) + fldV(pre0, 1);
				id[2] = P.rowBV[lr] + fldV(tot, 0) + fldV(tot, 1) + fldV(pre0, 2);
				bool hit = false;
#pragma unroll
				for (int a = 0; a < 3; a++) {
					cnt[a] = fldV(pre1, a) - fldV(pre0, a);
					hit = hit || (cnt[a] && id[a] < win0 + VCAP && id[a] + cnt[a] > win0);
				}
				if (!hit) continue;
				const uint32_t zl = fastdiv(lr, P.NY, P.mNY), y = lr - zl * P.NY, z = zl + P.zlo;
				WordRec rec; CellWords cw;
				word_masks(P, z, y, w, gz, rec, cw);
				const uint32_t zw = (gz && P.rowZ[lr]) ? P.Z[(uint64_t)lr * P.WP + w] : 0u;
#pragma unroll
				for (int a = 0; a < 3; a++) {
					uint32_t m = a == 0 ? rec.X : (a == 1 ? rec.Y : rec.Z);
					uint32_t slot = id[a] - win0;
					while (m) {
						const int b = __ffs((int)m) - 1;
						m &= m - 1;
						if (slot < VCAP)
							vlist[slot] = ((w << 5) + b) | ((uint32_t)a << 16) | ((a == 0 ? (zw >> b) & 1u : 0u) << 18) | (r << 19);
						slot++;
					}
				}
			}
			__syncthreads();
			const uint32_t nt = min((uint32_t)VCAP, vend - win0);
			for (uint32_t t = threadIdx.x; t < nt; t += 256) {
				const uint32_t e = vlist[t];
				const uint32_t lr = lr0 + (e >> 19);
				const uint32_t zl = fastdiv(lr, P.NY, P.mNY), y = lr - zl * P.NY, z = zl + P.zlo;
				emit_vertex_task<Sample>(P, e & 0xFFFFu, y, z, (int)((e >> 16) & 3u), ((e >> 18) & 1u) != 0, win0 + t);
			}
			__syncthreads();
		}
	}
}

// ---------------------------------------------------------------------------
// K4: triangles (+ centre vertices).  A tile = R cell rows (or a 256-word piece
// of one long row); its triangles are a contiguous id range.  Pass 1: a thread
// whose word has active cells builds the eight (plane mask, first id) pairs its
// cells' vertex ids are read from, walks the cells, picks each cell's MC33
// pattern and drops one task per triangle into the list.  Pass 2: thread t writes
// triangle window start + t: 12 consecutive bytes per thread.
// task word: item | bit << 8 | (table word index) << 13 | m << 25 | (centre ordinal) << 26; bit 31 = skip
// (cells with on-iso corners emit in pass 1: the zero-area drop is order dependent)
// ---------------------------------------------------------------------------
#define TCAP 4096
#define TSKIP 0x80000000u
#define EMT_SMEM (TBL_BYTES + 17 * 256 * 4 + TCAP * 4)

template <typename Sample>
__global__ void __launch_bounds__(256) k_emit_triangles(Params P, uint32_t row_begin, uint32_t row_end, uint32_t ntiles, uint32_t nchunk)
{
	extern __shared__ __align__(128) unsigned char smem[];
	const Tables tb = load_tables(smem);
	uint32_t *pmask = (uint32_t *)(smem + TBL_BYTES);        // [8][256]
	uint32_t *pbase = pmask + 8 * 256;                       // [8][256]
	uint32_t *cbase = pbase + 8 * 256;                       // [256] global id of the word's first centre vertex
	uint32_t *tlist = cbase + 256;                           // [TCAP]
	const bool gz = P.totals->anyZ != 0;
	const uint32_t vb = P.dbases ? P.dbases[0] : P.vbase;
	const uint32_t nShared = P.totals->nShared;
	for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
		const uint32_t rb = tile / nchunk, ch = tile - rb * nchunk;
		const uint32_t lr0 = row_begin + rb * P.R, lrE = min(lr0 + P.R, row_end), w0 = ch * P.CW;
		uint32_t tfirst, tend;
		if (nchunk == 1) { tfirst = P.rowBT[lr0]; tend = P.rowBT[lrE]; }
		else {
			const uint64_t *pt = P.wpreT + (uint64_t)lr0 * P.W1;
			tfirst = P.rowBT[lr0] + (uint32_t)pt[w0];
			tend = P.rowBT[lr0] + (uint32_t)pt[min(w0 + P.CW, P.W)];
		}
		// item of this thread
		const uint32_t r = P.R > 1 ? fastdiv(threadIdx.x, P.CW, P.mCW) : 0u, wl = threadIdx.x - r * P.CW;
		const uint32_t w = w0 + wl, lr = lr0 + r;
		const bool valid = r < P.R && lr < lrE && w < P.WC;
		uint64_t pre0 = 0, pre1 = 0;
		uint32_t t0 = 0, y = 0, z = 0;
		if (valid) {
			const uint64_t *pt = P.wpreT + (uint64_t)lr * P.W1;
			pre0 = pt[w]; pre1 = pt[w + 1];
			t0 = P.rowBT[lr] + (uint32_t)pre0;
			const uint32_t zl = fastdiv(lr, P.NY, P.mNY);
			y = lr - zl * P.NY; z = zl + P.zlo;
		}
		const uint32_t t1 = t0 + (uint32_t)(pre1 - pre0);
		for (uint32_t win0 = tfirst; win0 == tfirst || win0 < tend; win0 += TCAP) {
			const uint32_t win1 = win0 + TCAP;
			const bool last = win1 >= tend;
			if (pre0 != pre1 && (t0 < win1 || last) && t1 >= win0) {
				WordRec rec; CellWords cw; CellPairs cp;
				word_masks(P, z, y, w, gz, rec, cw);
				cell_pairs(P, z, y, w, gz, rec, cp);
#pragma unroll
				for (int k = 0; k < 8; k++) { pmask[k * 256 + threadIdx.x] = cp.mask[k]; pbase[k * 256 + threadIdx.x] = cp.base[k]; }
				const uint32_t cloc = nShared + P.rowBC[lr] + (uint32_t)(pre0 >> 32);   // slab-local id of the word's first centre
				cbase[threadIdx.x] = vb + cloc;
				uint32_t act = rec.act, tid = t0, cord = 0;
				while (act) {
					const int b = __ffs((int)act) - 1;
					act &= act - 1;
					const unsigned idx = cell_index(cw.c, 1, b);
					const unsigned zm = cw.zany ? cell_zmask(cw.zc, 1, b) : 0u;
					const uint32_t x = (w << 5) + b;
					const CellPattern pat = cell_pattern<Sample>(P, tb, x, y, z, idx, zm);
					const uint64_t cell = ((uint64_t)z * P.ny + y) * P.nx + x;
					const bool mine = (tid >= win0 && tid < win1) || (last && tid >= win1);   // the window that owns this cell
					if (pat.centre && mine) {
						const uint32_t cl = cloc + cord;
						if (cl < P.capV) {
							emit_centre_vertex<Sample>(P, x, y, z, cl);
							if (P.vkey) P.vkey[cl] = cell * 4 + 3;
						} else {
							P.totals->overflow = 1;
						}
					}
					if (zm) {
						// (the pairs are read back from shared memory: indexing the register copy
						// dynamically would push it to local memory for every thread)
						const uint32_t kept = emit_cell_triangles_z(P, tb, (unsigned)b, pat, zm, vb + cloc + cord, pmask + threadIdx.x,
						                                            pbase + threadIdx.x, 256, tid, win0, win1, cell);
						for (uint32_t j = 0; j < kept; j++)
							if (tid + j - win0 < TCAP) tlist[tid + j - win0] = TSKIP;
						tid += kept;
					} else {
						const uint32_t e = threadIdx.x | ((uint32_t)b << 8) | (pat.m << 25) | (cord << 26);
						for (uint32_t j = 0; j < pat.ntri; j++)
							if (tid + j - win0 < TCAP) tlist[tid + j - win0] = e | ((pat.start + j) << 13);
						tid += pat.ntri;
					}
					cord += pat.centre;
				}
			}
			__syncthreads();
			const uint32_t nt = tend > win0 ? min((uint32_t)TCAP, tend - win0) : 0u;
			for (uint32_t t = threadIdx.x; t < nt; t += 256) {
				const uint32_t e = tlist[t];
				if (e & TSKIP) continue;
				const uint32_t item = e & 255u, b = (e >> 8) & 31u;
				uint64_t cell = 0;
				if (P.tcell) {
					const uint32_t ir = P.R > 1 ? fastdiv(item, P.CW, P.mCW) : 0u, iw = w0 + item - ir * P.CW, ilr = lr0 + ir;
					const uint32_t izl = fastdiv(ilr, P.NY, P.mNY);
					cell = ((uint64_t)(izl + P.zlo) * P.ny + (ilr - izl * P.NY)) * P.nx + (iw << 5) + b;
				}
				emit_triangle_task(P, tb.tri[(e >> 13) & 0xFFFu], b, (e >> 25) & 1u, cbase[item] + ((e >> 26) & 31u),
				                   pmask + item, pbase + item, 256, win0 + t, cell);
			}
			__syncthreads();
		}
	}
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
	void *pinned; size_t pinned_bytes;   // staging for row-wise uploads
	// scan state
	uint64_t *scan_status; uint32_t *scan_ticket; uint32_t nbatch, epoch;
	uint32_t nchunk;         // 256-word pieces per row in the triangle kernel
	// host mirror of totals
	Totals *h_totals;
	bool counted;
	// staging outputs for the host path
	void *oV; float *oN; int32_t *oC; uint32_t *oT; uint64_t ocapV, ocapT;
	// timing
	bool timing; cudaEvent_t ev[6]; bool ev_valid;
	uint64_t launches;
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
	uint8_t pat[MC33_NTRI_WORDS];
	for (int i = 0; i < MC33_NTRI_WORDS; i++) pat[i] = (uint8_t)(MC33_PAT_NTRI[i] | (MC33_PAT_CENTRE[i] << 7));
	CU(cudaMemcpyToSymbol(d_case256, MC33_CASE256, sizeof(MC33_CASE256)));
	CU(cudaMemcpyToSymbol(d_simple256, MC33_SIMPLE256, sizeof(MC33_SIMPLE256)));
	CU(cudaMemcpyToSymbol(d_tri, MC33_TRI, sizeof(MC33_TRI)));
	CU(cudaMemcpyToSymbol(d_pat, pat, sizeof(pat)));
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
	cudaFree(P.S); cudaFree(P.Z); cudaFree(P.rowZ); cudaFree(P.wpreV); cudaFree(P.wpreT);
	cudaFree(P.rowBV); cudaFree(P.totals);
	cudaFree(c->scan_status);
	cudaFree(c->grid_owned);
	cudaFree(c->oV); cudaFree(c->oN); cudaFree(c->oC); cudaFree(c->oT);
	if (c->pinned) cudaFreeHost(c->pinned);
	if (c->h_totals) cudaFreeHost(c->h_totals);
	for (int i = 0; i < 6; i++) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
	if (c->own_stream) cudaStreamDestroy(c->own_stream);
	free(c);
}

static int set_geom(mc33cu_ctx *c, const mc33cu_desc *d);

template <typename Sample> static int set_kernel_attrs(const ClsPlan &pl)
{
	CU(cudaFuncSetAttribute(k_classify<Sample>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(pl.stage_bytes * CLS_STAGES)));
	CU(cudaFuncSetAttribute(k_emit_triangles<Sample>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EMT_SMEM));
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
	P.WP = (P.W + 1 + 3) & ~3u;
	P.W1 = P.W + 1;
	P.Lrows = (P.zhi - P.zlo) * P.NY;
	if (P.NX > 65535) { free(c); return fail(MC33CU_ERR_ARG, "rows longer than 65535 samples are not supported"); }
	if ((uint64_t)(P.zhi - P.zlo) * P.NY > 0x7FFFFFFFull / P.WP) { free(c); return fail(MC33CU_ERR_ARG, "too many rows"); }
	P.R = P.W >= 256 ? 1 : 256 / P.W;
	P.CW = P.W < 256 ? P.W : 256;
	P.mCW = P.CW >= 2 ? (uint32_t)(0x100000000ull / P.CW) : 0u;
	P.mNY = (uint32_t)(0x100000000ull / P.NY);
	c->nchunk = (P.W + P.CW - 1) / P.CW;
	set_geom(c, d);
	static const size_t ssz[5] = {4, 8, 1, 2, 4};
	c->sample_size = ssz[d->dtype];
	c->real_size = d->dtype == MC33CU_F64 ? 8 : 4;
	c->n_samples = (uint64_t)P.Lrows * P.NX;
	c->nbatch = (P.Lrows + P.R - 1) / P.R;
	{
		// classify chunks: ~16 KB of whole rows, or 16 KB pieces of one long row
		ClsPlan &pl = c->cls;
		const size_t rowb = (size_t)P.NX * c->sample_size, target = 16384;
		if (rowb <= target) {
			pl.rows = (uint32_t)(target / rowb); pl.words = P.W; pl.nwchunk = 1;
			if (pl.rows > P.Lrows) pl.rows = P.Lrows;
			pl.nchunks = (P.Lrows + pl.rows - 1) / pl.rows;
			pl.stage_bytes = (uint32_t)(((size_t)pl.rows * rowb + 32 + 127) & ~(size_t)127);
		} else {
			pl.rows = 1; pl.words = (uint32_t)(target / (32 * c->sample_size));
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
	TRY(dalloc(&P.rowZ, (size_t)P.Lrows));
	TRYCU(cudaMemsetAsync(P.rowZ, 0, (size_t)P.Lrows, c->stream));
	TRY(dalloc(&P.wpreV, (size_t)P.Lrows * P.W1)); TRY(dalloc(&P.wpreT, (size_t)P.Lrows * P.W1));
	TRYCU(cudaMemsetAsync(P.wpreV, 0, (size_t)P.Lrows * P.W1 * 8, c->stream));
	TRYCU(cudaMemsetAsync(P.wpreT, 0, (size_t)P.Lrows * P.W1 * 8, c->stream));
	TRY(dalloc(&P.rowBV, ((size_t)P.Lrows + 1) * 3));
	P.rowBT = P.rowBV + (P.Lrows + 1); P.rowBC = P.rowBT + (P.Lrows + 1);
	{
		// Totals (32 bytes) and the batch ticket share one block so that a single
		// small memset re-arms both before every extraction
		unsigned char *blk;
		TRY(dalloc(&blk, 64));
		P.totals = (Totals *)blk;
		c->scan_ticket = (uint32_t *)(blk + sizeof(Totals));
		TRYCU(cudaMemsetAsync(blk, 0, 64, c->stream));
	}
	TRY(dalloc(&c->scan_status, (size_t)c->nbatch * 3));
	TRYCU(cudaMemsetAsync(c->scan_status, 0, (size_t)c->nbatch * 3 * 8, c->stream));
	TRYCU(cudaMallocHost((void **)&c->h_totals, sizeof(Totals)));
	for (int i = 0; i < 6; i++) TRYCU(cudaEventCreate(&c->ev[i]));
	TRYCU(cudaStreamSynchronize(c->stream));
#undef TRY
#undef TRYCU
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

extern "C" int mc33cu_grid_upload(mc33cu_ctx *c, const void *host)
{
	if (!c || !host) return fail(MC33CU_ERR_ARG, "null argument");
	int rc = ensure_grid(c);
	if (rc) return rc;
	CU(cudaMemcpyAsync(c->grid_owned, host, c->n_samples * c->sample_size, cudaMemcpyHostToDevice, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	return MC33CU_OK;
}

extern "C" int mc33cu_grid_upload_rows(mc33cu_ctx *c, const void *const *const *F)
{
	if (!c || !F) return fail(MC33CU_ERR_ARG, "null argument");
	int rc = ensure_grid(c);
	if (rc) return rc;
	const Params &P = c->P;
	const size_t rowb = (size_t)P.NX * c->sample_size;
	// fast path: grid_from_data_pointer layout (MC33_util_grd.c:600-612), one block
	const char *first = (const char *)F[P.zlo][0];
	bool contiguous = true;
	for (uint32_t z = P.zlo; z < P.zhi && contiguous; z++)
		for (uint32_t y = 0; y < P.NY; y++)
			if ((const char *)F[z][y] != first + ((size_t)(z - P.zlo) * P.NY + y) * rowb) { contiguous = false; break; }
	if (contiguous) return mc33cu_grid_upload(c, first);
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
template <typename Sample> static int launch_count_phase(mc33cu_ctx *c)
{
	Params &P = c->P;
	cudaStream_t s = c->stream;
	// re-arm totals (incl. the overflow / on-iso flags) and the batch ticket
	CU(cudaMemsetAsync(P.totals, 0, 64, s));
	if (c->timing) CU(cudaEventRecord(c->ev[0], s));
	{
		const ClsPlan &pl = c->cls;
		const size_t smem = (size_t)pl.stage_bytes * CLS_STAGES;
		uint32_t per_sm = (uint32_t)((200u << 10) / (smem + 1024));
		if (per_sm < 1) per_sm = 1;
		if (per_sm > 8) per_sm = 8;
		uint32_t grid = (uint32_t)c->n_sm * per_sm;
		if (grid > pl.nchunks) grid = pl.nchunks;
		k_classify<Sample><<<grid, CLS_THREADS, smem, s>>>(P, pl);
		c->launches++;
	}
	if (c->timing) CU(cudaEventRecord(c->ev[1], s));
	{
		c->epoch++;
		uint32_t grid = (uint32_t)c->n_sm * 8;
		if (grid > c->nbatch) grid = c->nbatch;
		const uint32_t owned_end = (P.pz1 - P.zlo) * P.NY;
		k_count<Sample><<<grid, 256, TBL_BYTES, s>>>(P, c->scan_status, c->scan_ticket, c->nbatch, c->epoch, owned_end);
		c->launches++;
	}
	if (c->timing) { CU(cudaEventRecord(c->ev[2], s)); CU(cudaEventRecord(c->ev[3], s)); }
	CU(cudaGetLastError());
	return MC33CU_OK;
}

template <typename Sample> static int launch_emit_phase(mc33cu_ctx *c)
{
	Params &P = c->P;
	cudaStream_t s = c->stream;
	{
		const uint32_t rb = (P.pz0 - P.zlo) * P.NY, re = (P.pz1 - P.zlo) * P.NY;
		const uint32_t ntiles = (re - rb + P.R - 1) / P.R;
		uint32_t grid = (uint32_t)c->n_sm * 8;
		if (grid > ntiles) grid = ntiles;
		k_emit_vertices<Sample><<<grid, 256, 0, s>>>(P, rb, re, ntiles);
		c->launches++;
	}
	if (c->timing) CU(cudaEventRecord(c->ev[4], s));
	{
		const uint32_t rb = (P.cz0 - P.zlo) * P.NY, re = (P.cz1 - P.zlo) * P.NY;
		const uint32_t ntiles = ((re - rb + P.R - 1) / P.R) * c->nchunk;
		uint32_t grid = (uint32_t)c->n_sm * 5;
		if (grid > ntiles) grid = ntiles;
		k_emit_triangles<Sample><<<grid, 256, EMT_SMEM, s>>>(P, rb, re, ntiles, c->nchunk);
		c->launches++;
	}
	if (c->timing) { CU(cudaEventRecord(c->ev[5], s)); c->ev_valid = true; }
	CU(cudaGetLastError());
	return MC33CU_OK;
}

static int dispatch_count(mc33cu_ctx *c)
{
	switch (c->d.dtype) {
	case MC33CU_F32: return launch_count_phase<float>(c);
	case MC33CU_F64: return launch_count_phase<double>(c);
	case MC33CU_U8:  return launch_count_phase<uint8_t>(c);
	case MC33CU_U16: return launch_count_phase<uint16_t>(c);
	default:         return launch_count_phase<uint32_t>(c);
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

static void set_out(mc33cu_ctx *c, const mc33cu_out *o)
{
	Params &P = c->P;
	P.V = o->V; P.N = o->N; P.color = o->color; P.T = o->T; P.vkey = o->vkey; P.tcell = o->tcell;
	P.capV = o->capV; P.capT = o->capT;
	P.vbase = o->vbase; P.vbase_next = o->vbase_next; P.dbases = o->dev_bases;
	P.color_value = o->color_value;
}

static int fetch_totals(mc33cu_ctx *c)
{
	CU(cudaMemcpyAsync(c->h_totals, c->P.totals, sizeof(Totals), cudaMemcpyDeviceToHost, c->stream));
	CU(cudaStreamSynchronize(c->stream));
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
	int rc = dispatch_count(c);
	if (rc) return rc;
	if (dev_counts4) {
		k_export_counts<<<1, 1, 0, c->stream>>>(c->P.totals, dev_counts4);
		c->launches++;
		CU(cudaGetLastError());
	}
	c->counted = true;
	return MC33CU_OK;
}

extern "C" int mc33cu_emit_device(mc33cu_ctx *c, const mc33cu_out *o)
{
	if (!c || !o) return fail(MC33CU_ERR_ARG, "null argument");
	if (!c->counted) return fail(MC33CU_ERR_STATE, "mc33cu_count has not run");
	CU(cudaSetDevice(c->device));
	set_out(c, o);
	return dispatch_emit(c);
}

extern "C" int mc33cu_extract_device(mc33cu_ctx *c, double iso, const mc33cu_out *o)
{
	if (!c || !o) return fail(MC33CU_ERR_ARG, "null argument");
	if (!c->P.data) return fail(MC33CU_ERR_STATE, "no grid bound");
	CU(cudaSetDevice(c->device));
	set_iso(c, iso);
	set_out(c, o);
	int rc = dispatch_count(c);
	if (rc) return rc;
	rc = dispatch_emit(c);
	if (rc) return rc;
	c->counted = true;
	return MC33CU_OK;
}

extern "C" int mc33cu_sync(mc33cu_ctx *c)
{
	if (!c) return fail(MC33CU_ERR_ARG, "null context");
	int rc = fetch_totals(c);
	if (rc) return rc;
	if (c->h_totals->range) return fail(MC33CU_ERR_RANGE, "more than 2^32-1 vertices or triangles");
	if (c->h_totals->overflow) return fail(MC33CU_ERR_CAPACITY, "output capacity exceeded");
	return MC33CU_OK;
}

extern "C" int mc33cu_get_counts(mc33cu_ctx *c, mc33cu_counts *k)
{
	if (!c || !k) return fail(MC33CU_ERR_ARG, "null argument");
	if (!c->counted) return fail(MC33CU_ERR_STATE, "nothing counted yet");
	fill_counts(c, k);
	return MC33CU_OK;
}

extern "C" int mc33cu_emit_host(mc33cu_ctx *c, void *V, float *N, int32_t *color, uint32_t *T, int32_t color_value)
{
	if (!c) return fail(MC33CU_ERR_ARG, "null context");
	if (!c->counted) return fail(MC33CU_ERR_STATE, "mc33cu_count has not run");
	CU(cudaSetDevice(c->device));
	mc33cu_counts k;
	fill_counts(c, &k);
	if (k.nV == 0 && k.nT == 0) return MC33CU_OK;
	if (!V || !N || !color || !T) return fail(MC33CU_ERR_ARG, "null output array");
	if (c->ocapV < k.nV) {
		cudaFree(c->oV); cudaFree(c->oN); cudaFree(c->oC);
		c->oV = nullptr; c->oN = nullptr; c->oC = nullptr; c->ocapV = 0;
		CU(cudaMalloc(&c->oV, k.nV * 3 * c->real_size));
		CU(cudaMalloc((void **)&c->oN, k.nV * 3 * sizeof(float)));
		CU(cudaMalloc((void **)&c->oC, k.nV * sizeof(int32_t)));
		c->ocapV = k.nV;
	}
	if (c->ocapT < k.nT) {
		cudaFree(c->oT);
		c->oT = nullptr; c->ocapT = 0;
		CU(cudaMalloc((void **)&c->oT, (k.nT ? k.nT : 1) * 3 * sizeof(uint32_t)));
		c->ocapT = k.nT;
	}
	mc33cu_out o;
	memset(&o, 0, sizeof o);
	o.V = c->oV; o.N = c->oN; o.color = c->oC; o.T = c->oT;
	o.capV = (uint32_t)k.nV; o.capT = (uint32_t)k.nT;
	o.color_value = color_value;
	set_out(c, &o);
	int rc = dispatch_emit(c);
	if (rc) return rc;
	cudaStream_t s = c->stream;
	CU(cudaMemcpyAsync(T, c->oT, k.nT * 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
	CU(cudaMemcpyAsync(V, c->oV, k.nV * 3 * c->real_size, cudaMemcpyDeviceToHost, s));
	CU(cudaMemcpyAsync(N, c->oN, k.nV * 3 * sizeof(float), cudaMemcpyDeviceToHost, s));
	CU(cudaMemcpyAsync(color, c->oC, k.nV * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
	CU(cudaStreamSynchronize(s));
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

extern "C" uint64_t mc33cu_launch_count(const mc33cu_ctx *c) { return c ? c->launches : 0; }
