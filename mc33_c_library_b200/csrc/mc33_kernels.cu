// mc33_kernels.cu -- sm_100a kernels and the C-ABI (include/mc33cu.h) of the
// B200 Marching Cubes 33 extractor.  See DESIGN.md for the pipeline:
//
//   K1 classify   stream the samples once -> S / Z bitmaps (+ per-row on-iso flag)
//   K2 count      per (row, 32-point word): owned vertices per plane, triangles
//                 and centre vertices of the 32 cells; row-local prefixes
//   K3 scan       single-pass decoupled look-back scan over the point rows
//   K4v emit      vertices (positions, normals, colours) of owned rows
//   K4t emit      triangles (+ centre vertices) of owned cell rows
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

#define TBL_BYTES (512 + 512 + ((MC33_NTRI_WORDS * 2 + 15) / 16 * 16) + ((MC33_NTRI_WORDS + 15) / 16 * 16))

__device__ __forceinline__ Tables load_tables(unsigned char *smem)
{
	uint16_t *c = (uint16_t *)smem;
	uint16_t *s = c + 256;
	uint16_t *t = s + 256;
	uint8_t *p = (uint8_t *)(t + ((MC33_NTRI_WORDS * 2 + 15) / 16 * 8));
	for (int i = threadIdx.x; i < 256; i += blockDim.x) { c[i] = d_case256[i]; s[i] = d_simple256[i]; }
	for (int i = threadIdx.x; i < MC33_NTRI_WORDS; i += blockDim.x) { t[i] = d_tri[i]; p[i] = d_pat[i]; }
	__syncthreads();
	Tables tb;
	tb.case256 = c; tb.simple256 = s; tb.tri = t; tb.pat = p;
	return tb;
}

// ---------------------------------------------------------------------------
// K1: classify.  One warp per point row; lane l looks at x = 32*w + l, so a warp
// load is one coalesced 32-sample segment and __ballot_sync yields bitmap word w
// directly.  Four words are in flight per lane.  (sign bit of iso - F and
// iso - F == 0: reference marching_cubes_33.c:1840-1859.)
// ---------------------------------------------------------------------------
template <typename Sample>
__global__ void __launch_bounds__(256) k_classify(Params P)
{
	typedef typename Traits<Sample>::Real Real;
	const Real iso = (Real)P.iso;
	const unsigned lane = threadIdx.x & 31;
	const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
	for (uint32_t lr = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; lr < P.Lrows; lr += warps) {
		const Sample *src = (const Sample *)P.data + (uint64_t)lr * P.NX;
		uint32_t *Sr = P.S + (uint64_t)lr * P.WP, *Zr = P.Z + (uint64_t)lr * P.WP;
		uint32_t zany = 0;
		for (uint32_t w0 = 0; w0 < P.W; w0 += 4) {
			Sample f[4];
			bool ok[4];
#pragma unroll
			for (int j = 0; j < 4; j++) {
				uint32_t x = ((w0 + j) << 5) + lane;
				ok[j] = x < P.NX;
				f[j] = ok[j] ? __ldg(src + x) : (Sample)0;
			}
			uint32_t sw = 0, zw = 0;
#pragma unroll
			for (int j = 0; j < 4; j++) {
				Real v = rsub(iso, (Real)f[j]);
				uint32_t sb = __ballot_sync(0xFFFFFFFFu, ok[j] && sgn(v));
				uint32_t zb = __ballot_sync(0xFFFFFFFFu, ok[j] && v == (Real)0);
				if (lane == (unsigned)j) { sw = sb; zw = zb; }
				zany |= zb;
			}
			if (lane < 4 && w0 + lane < P.W) { Sr[w0 + lane] = sw; Zr[w0 + lane] = zw; }
		}
		if (lane == 0) P.rowZ[lr] = zany != 0;
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
// K2: count.  A CTA takes RB whole rows (RB*W <= items per pass), one thread per
// (row, word); the row-local exclusive prefixes fall out of one block scan.
// ---------------------------------------------------------------------------
template <typename Sample>
__global__ void __launch_bounds__(256) k_count(Params P, uint32_t RB)
{
	extern __shared__ __align__(16) unsigned char smem[];
	__shared__ uint64_t sw[2][8];
	const Tables tb = load_tables(smem);
	uint64_t *preV = (uint64_t *)(smem + TBL_BYTES);
	const uint32_t n = RB * P.W;
	uint64_t *preC = preV + (n + 1);
	const uint32_t row0 = blockIdx.x * RB;
	uint64_t carryV = 0, carryC = 0;
	for (uint32_t base = 0; base < n; base += 256) {
		const uint32_t it = base + threadIdx.x;
		uint64_t cv = 0, cc = 0;
		if (it < n) {
			const uint32_t r = it / P.W, w = it - r * P.W, lr = row0 + r;
			if (lr < P.Lrows) {
				const uint32_t zl = lr / P.NY, y = lr - zl * P.NY, z = zl + P.zlo;
				const bool own_p = row_points_owned(P, z) || row_points_halo(P, z);
				const bool own_c = row_cells_owned(P, z, y);
				if (own_p || own_c) count_word<Sample>(P, tb, z, y, w, own_p, own_c, cv, cc);
			}
		}
		uint64_t tv, tc;
		block_exscan2(cv, cc, tv, tc, sw);
		if (it < n) { preV[it] = carryV + cv; preC[it] = carryC + cc; }
		carryV += tv; carryC += tc;
	}
	if (threadIdx.x == 0) { preV[n] = carryV; preC[n] = carryC; }
	__syncthreads();
	for (uint32_t it = threadIdx.x; it < n; it += 256) {
		const uint32_t r = it / P.W, w = it - r * P.W, lr = row0 + r;
		if (lr >= P.Lrows) continue;
		const uint64_t bv = preV[r * P.W], bc = preC[r * P.W];
		P.wpreV[(uint64_t)lr * P.W + w] = preV[it] - bv;
		P.wpreC[(uint64_t)lr * P.W + w] = preC[it] - bc;
		if (w == P.W - 1) {
			const uint64_t tv = preV[(r + 1) * P.W] - bv, tc = preC[(r + 1) * P.W] - bc;
			P.rowNX[lr] = (uint32_t)(tv & 0x1FFFFF);
			P.rowNY[lr] = (uint32_t)((tv >> 21) & 0x1FFFFF);
			P.rowNZ[lr] = (uint32_t)((tv >> 42) & 0x1FFFFF);
			P.rowNT[lr] = (uint32_t)(tc & 0xFFFFFFFFu);
			P.rowNC[lr] = (uint32_t)(tc >> 32);
		}
	}
}

// ---------------------------------------------------------------------------
// K3: exclusive scan over the point rows of (vertices, centres, triangles):
// single pass, tiles taken in ticket order, decoupled look-back (one warp per
// scanned quantity).  status word = flag<<62 | value; flag 1 = tile aggregate,
// 2 = inclusive prefix.
// ---------------------------------------------------------------------------
#define SCAN_ROWS_PER_THREAD 4
#define SCAN_TILE (256 * SCAN_ROWS_PER_THREAD)
#define ST_AGG (1ull << 62)
#define ST_PRE (2ull << 62)
#define ST_VAL (~(3ull << 62))

__device__ __forceinline__ uint64_t ld_status(const volatile uint64_t *p) { return *p; }

__global__ void __launch_bounds__(256) k_scan_rows(Params P, uint64_t *status, uint32_t *ticket, uint32_t owned_end_row)
{
	__shared__ uint32_t s_tile;
	__shared__ uint64_t s_w[3][8];
	__shared__ uint64_t s_excl[3], s_agg[3];
	const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
	__syncthreads();
	const uint32_t tile = s_tile;
	const uint32_t r0 = tile * SCAN_TILE + threadIdx.x * SCAN_ROWS_PER_THREAD;
	uint32_t nX[SCAN_ROWS_PER_THREAD], nY[SCAN_ROWS_PER_THREAD], nZ[SCAN_ROWS_PER_THREAD],
	         nC[SCAN_ROWS_PER_THREAD], nT[SCAN_ROWS_PER_THREAD];
	uint64_t t[3] = {0, 0, 0};
#pragma unroll
	for (int k = 0; k < SCAN_ROWS_PER_THREAD; k++) {
		const uint32_t r = r0 + k;
		const bool ok = r < P.Lrows;
		nX[k] = ok ? P.rowNX[r] : 0; nY[k] = ok ? P.rowNY[r] : 0; nZ[k] = ok ? P.rowNZ[r] : 0;
		nC[k] = ok ? P.rowNC[r] : 0; nT[k] = ok ? P.rowNT[r] : 0;
		t[0] += (uint64_t)nX[k] + nY[k] + nZ[k]; t[1] += nC[k]; t[2] += nT[k];
	}
	// block exclusive scan of the three per-thread sums
	uint64_t inc[3], exc[3];
#pragma unroll
	for (int q = 0; q < 3; q++) {
		uint64_t v = t[q];
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			uint64_t x = __shfl_up_sync(0xFFFFFFFFu, v, d);
			if (lane >= (unsigned)d) v += x;
		}
		inc[q] = v;
		if (lane == 31) s_w[q][wid] = v;
	}
	__syncthreads();
#pragma unroll
	for (int q = 0; q < 3; q++) {
		uint64_t o = 0, s = 0;
#pragma unroll
		for (int k = 0; k < 8; k++) { uint64_t v = s_w[q][k]; if ((unsigned)k < wid) o += v; s += v; }
		exc[q] = o + inc[q] - t[q];
		if (threadIdx.x == 0) s_agg[q] = s;
	}
	__syncthreads();
	// warps 0..2: publish aggregate of quantity `wid`, look back, publish prefix
	if (wid < 3) {
		const unsigned q = wid;
		volatile uint64_t *st = status;
		const uint64_t agg = s_agg[q];
		uint64_t excl = 0;
		if (tile == 0) {
			if (lane == 0) { st[3 * (uint64_t)tile + q] = ST_PRE | agg; }
		} else {
			if (lane == 0) { st[3 * (uint64_t)tile + q] = ST_AGG | agg; }
			int64_t pos = (int64_t)tile - 1;
			while (true) {
				const int64_t idx = pos - (int64_t)lane;
				uint64_t s = ST_PRE;     // tiles before the first one: prefix 0
				if (idx >= 0) {
					do { s = ld_status(st + 3 * (uint64_t)idx + q); } while ((s >> 62) == 0);
				}
				const unsigned pre_mask = __ballot_sync(0xFFFFFFFFu, (s >> 62) == 2);
				// lanes up to and including the first inclusive prefix contribute
				const unsigned first = pre_mask ? (unsigned)__ffs((int)pre_mask) - 1u : 32u;
				uint64_t v = (lane <= first) ? (s & ST_VAL) : 0;
#pragma unroll
				for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
				excl += v;
				if (pre_mask) break;
				pos -= 32;
			}
			if (lane == 0) { st[3 * (uint64_t)tile + q] = ST_PRE | (excl + agg); }
		}
		if (lane == 0) s_excl[q] = excl;
	}
	__syncthreads();
	uint64_t bv = s_excl[0] + exc[0], bc = s_excl[1] + exc[1], bt = s_excl[2] + exc[2];
#pragma unroll
	for (int k = 0; k < SCAN_ROWS_PER_THREAD; k++) {
		const uint32_t r = r0 + k;
		if (r == owned_end_row) P.totals->nShared = (uint32_t)bv;
		if (r < P.Lrows) {
			P.rowBX[r] = (uint32_t)bv; P.rowBY[r] = (uint32_t)(bv + nX[k]); P.rowBZ[r] = (uint32_t)(bv + nX[k] + nY[k]);
			P.rowBC[r] = (uint32_t)bc; P.rowBT[r] = (uint32_t)bt;
		}
		bv += (uint64_t)nX[k] + nY[k] + nZ[k]; bc += nC[k]; bt += nT[k];
	}
	// the tile holding the last row finalises the totals
	if (r0 <= P.Lrows - 1 && P.Lrows - 1 < r0 + SCAN_ROWS_PER_THREAD) {
		const uint64_t tv = bv, tc = bc, tt = bt;   // bv.. now hold the inclusive totals
		if (owned_end_row >= P.Lrows) P.totals->nShared = (uint32_t)tv;
		P.totals->nCentre = (uint32_t)tc;
		P.totals->nT = (uint32_t)tt;
		P.totals->pad_[0] = (uint32_t)tv;   // all shared vertices counted (own + halo)
		// 32-bit index range check (include/marching_cubes_33.h:140 uses unsigned int)
		P.totals->pad_[1] = (tv + tc >= 0xFFFFFFFFull || tt >= 0xFFFFFFFFull) ? 1u : 0u;
	}
}

// ---------------------------------------------------------------------------
// K4v / K4t: one thread per (row, word)
// ---------------------------------------------------------------------------
template <typename Sample>
__global__ void __launch_bounds__(256) k_emit_vertices(Params P, uint32_t row_begin, uint32_t row_end)
{
	const uint64_t it = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t r = (uint32_t)(it / P.W), w = (uint32_t)(it - (uint64_t)r * P.W), lr = row_begin + r;
	if (lr >= row_end) return;
	const uint32_t zl = lr / P.NY, y = lr - zl * P.NY;
	emit_vertices_word<Sample>(P, zl + P.zlo, y, w);
}

template <typename Sample>
__global__ void __launch_bounds__(256) k_emit_triangles(Params P, uint32_t row_begin, uint32_t row_end)
{
	extern __shared__ __align__(16) unsigned char smem[];
	const Tables tb = load_tables(smem);
	uint32_t *scr_mask = (uint32_t *)(smem + TBL_BYTES);
	uint32_t *scr_base = scr_mask + 8 * 256;
	const uint64_t it = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t r = (uint32_t)(it / P.WC), w = (uint32_t)(it - (uint64_t)r * P.WC), lr = row_begin + r;
	if (lr >= row_end) return;
	const uint32_t zl = lr / P.NY, y = lr - zl * P.NY;
	if (y >= P.ny) return;
	emit_triangles_word<Sample>(P, tb, zl + P.zlo, y, w, scr_mask + threadIdx.x, scr_base + threadIdx.x, 256);
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
	cudaStream_t own_stream, stream;
	Params P;
	size_t sample_size, real_size;
	uint64_t n_samples;
	void *grid_owned;        // device copy made by the upload calls
	void *pinned; size_t pinned_bytes;   // staging for row-wise uploads
	// scan state
	uint64_t *scan_status; uint32_t *scan_ticket; uint32_t scan_tiles;
	// host mirror of totals
	Totals *h_totals;
	bool counted;
	// staging outputs for the host path
	void *oV; float *oN; int32_t *oC; uint32_t *oT; uint64_t ocapV, ocapT;
	// timing
	bool timing; cudaEvent_t ev[6]; bool ev_valid;
	uint64_t launches;
	uint32_t RB;
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
	cudaFree(P.S); cudaFree(P.Z); cudaFree(P.rowZ); cudaFree(P.wpreV); cudaFree(P.wpreC);
	cudaFree(P.rowNX); cudaFree(P.rowBX); cudaFree(P.totals);
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
	Params &P = c->P;
	P.nx = d->nx; P.ny = d->ny; P.nz = d->nz;
	P.NX = d->nx + 1; P.NY = d->ny + 1;
	P.zlo = d->z_lo; P.zhi = d->z_hi;
	P.cz0 = d->cell_z0; P.cz1 = d->cell_z1;
	P.pz0 = d->cell_z0; P.pz1 = d->is_last ? NZ : d->cell_z1;
	P.hz = d->is_last ? 0xFFFFFFFFu : d->cell_z1;
	P.W = (P.NX + 31) / 32; P.WC = (P.nx + 31) / 32;
	P.WP = (P.W + 1 + 3) & ~3u;
	P.Lrows = (P.zhi - P.zlo) * P.NY;
	if (P.W > 2048) { free(c); return fail(MC33CU_ERR_ARG, "rows longer than 65536 samples are not supported"); }
	if ((uint64_t)(P.zhi - P.zlo) * P.NY > 0x7FFFFFFFull) { free(c); return fail(MC33CU_ERR_ARG, "too many rows"); }
	set_geom(c, d);
	static const size_t ssz[5] = {4, 8, 1, 2, 4};
	c->sample_size = ssz[d->dtype];
	c->real_size = d->dtype == MC33CU_F64 ? 8 : 4;
	c->n_samples = (uint64_t)P.Lrows * P.NX;
	c->RB = P.W >= 256 ? 1 : 256 / P.W;
	c->scan_tiles = (P.Lrows + SCAN_TILE - 1) / SCAN_TILE;

#define TRY(x) do { rc = (x); if (rc) { mc33cu_destroy(c); return rc; } } while (0)
#define TRYCU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { mc33cu_destroy(c); \
	return fail(e_ == cudaErrorMemoryAllocation ? MC33CU_ERR_NOMEM : MC33CU_ERR_CUDA, "%s: %s", #x, cudaGetErrorString(e_)); } } while (0)
	TRYCU(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
	c->stream = c->own_stream;
	const size_t bm = (size_t)P.Lrows * P.WP;
	TRY(dalloc(&P.S, bm)); TRY(dalloc(&P.Z, bm));
	TRYCU(cudaMemsetAsync(P.S, 0, bm * 4, c->stream)); TRYCU(cudaMemsetAsync(P.Z, 0, bm * 4, c->stream));
	TRY(dalloc(&P.rowZ, (size_t)P.Lrows));
	TRY(dalloc(&P.wpreV, (size_t)P.Lrows * P.W)); TRY(dalloc(&P.wpreC, (size_t)P.Lrows * P.W));
	TRY(dalloc(&P.rowNX, (size_t)P.Lrows * 5));
	P.rowNY = P.rowNX + P.Lrows; P.rowNZ = P.rowNY + P.Lrows; P.rowNC = P.rowNZ + P.Lrows; P.rowNT = P.rowNC + P.Lrows;
	TRY(dalloc(&P.rowBX, (size_t)P.Lrows * 5));
	P.rowBY = P.rowBX + P.Lrows; P.rowBZ = P.rowBY + P.Lrows; P.rowBC = P.rowBZ + P.Lrows; P.rowBT = P.rowBC + P.Lrows;
	TRY(dalloc(&P.totals, 1));
	TRYCU(cudaMemsetAsync(P.totals, 0, sizeof(Totals), c->stream));
	TRY(dalloc(&c->scan_status, (size_t)c->scan_tiles * 3 + 1));
	c->scan_ticket = (uint32_t *)(c->scan_status + (size_t)c->scan_tiles * 3);
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
	if (c->timing) CU(cudaEventRecord(c->ev[0], s));
	{
		const uint32_t rows_per_cta = 8;
		uint32_t grid = (P.Lrows + rows_per_cta - 1) / rows_per_cta;
		const uint32_t cap = 148 * 8 * 4;
		if (grid > cap) grid = cap;
		k_classify<Sample><<<grid, 256, 0, s>>>(P);
		c->launches++;
	}
	if (c->timing) CU(cudaEventRecord(c->ev[1], s));
	{
		const uint32_t n = c->RB * P.W;
		const size_t smem = TBL_BYTES + 2 * (size_t)(n + 1) * 8;
		const uint32_t grid = (P.Lrows + c->RB - 1) / c->RB;
		k_count<Sample><<<grid, 256, smem, s>>>(P, c->RB);
		c->launches++;
	}
	if (c->timing) CU(cudaEventRecord(c->ev[2], s));
	{
		CU(cudaMemsetAsync(c->scan_status, 0, ((size_t)c->scan_tiles * 3 + 1) * 8, s));
		const uint32_t owned_end = (P.pz1 - P.zlo) * P.NY;
		k_scan_rows<<<c->scan_tiles, 256, 0, s>>>(P, c->scan_status, c->scan_ticket, owned_end);
		c->launches++;
	}
	if (c->timing) CU(cudaEventRecord(c->ev[3], s));
	CU(cudaGetLastError());
	return MC33CU_OK;
}

template <typename Sample> static int launch_emit_phase(mc33cu_ctx *c)
{
	Params &P = c->P;
	cudaStream_t s = c->stream;
	{
		const uint32_t rb = (P.pz0 - P.zlo) * P.NY, re = (P.pz1 - P.zlo) * P.NY;
		const uint64_t items = (uint64_t)(re - rb) * P.W;
		k_emit_vertices<Sample><<<(unsigned)((items + 255) / 256), 256, 0, s>>>(P, rb, re);
		c->launches++;
	}
	if (c->timing) CU(cudaEventRecord(c->ev[4], s));
	{
		const uint32_t rb = (P.cz0 - P.zlo) * P.NY, re = (P.cz1 - P.zlo) * P.NY;
		const uint64_t items = (uint64_t)(re - rb) * P.WC;
		const size_t smem = TBL_BYTES + 256 * 8 * 8;
		k_emit_triangles<Sample><<<(unsigned)((items + 255) / 256), 256, smem, s>>>(P, rb, re);
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
	if (c->d.dtype == MC33CU_F64) c->P.iso = iso + 0.0;
	else c->P.iso = (double)((float)iso + 0.0f);
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
	k->nSharedHalo = t.pad_[0] - t.nShared;
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
	if (c->h_totals->pad_[1]) return fail(MC33CU_ERR_RANGE, "more than 2^32-1 vertices or triangles");
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
	CU(cudaMemsetAsync(&c->P.totals->overflow, 0, 4, c->stream));
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
	CU(cudaMemsetAsync(&c->P.totals->overflow, 0, 4, c->stream));
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
	if (c->h_totals->pad_[1]) return fail(MC33CU_ERR_RANGE, "more than 2^32-1 vertices or triangles");
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
