// mc33_pipeline.cuh -- bodies of the round-2 count and cell kernels, written against the small
// SIMT context of mc33_simt.h so that the very same code runs as sm_100a kernels
// (mc33_kernels.cu) and, lane by lane as fibers, in the CPU test harness (tests/hostemu).
//
//   count_body       K2: per (row, 32-point word) owned vertices per plane and triangles.  New in
//                    round 2: (1) the on-iso rules are applied per WORD with masks (one cold call
//                    per quad that has an on-iso sample in reach), never per cell; (2) the cells
//                    that must be looked at one by one -- complex (face / interior tests) or with
//                    an on-iso corner -- are compacted across the warp into a shared-memory queue
//                    and drained 32 at a time, neighbouring lanes holding neighbouring cells
//                    (coalesced corner loads, full lanes in select_pattern); the pattern each
//                    complex cell selects is kept in P.pcache for the cell kernel; (3) the row
//                    bases come out of a single-pass decoupled look-back over the per-CTA
//                    aggregates (status word + inclusive prefix per block, Merrill & Garland),
//                    which replaces the second kernel (k_rowscan) and the one-thread count
//                    export kernels of round 1.
//                    Replaces reference marching_cubes_33.c:1892-1940 / :1258-1724.
//   emit_cells_body  K4: a warp stages, per point row and word of its row group, one record
//                    {sign word, on-iso word, (plane mask, id of the plane's first vertex) x 3} in
//                    shared memory -- computed once per word, on-iso rules included -- and every
//                    visited cell then takes its 12 edge vertex ids from four records with 8
//                    popcounts.  Triangles are written one lane per triangle; the owner cell of a
//                    triangle is found by a 5-step search over the shuffled scan (no byte scatter
//                    in shared memory).  Cells with an on-iso corner redirect edge ids to the
//                    corner's POINT vertex and drop zero-area triangles with a keep mask, in the
//                    same dense loops.  Replaces reference marching_cubes_33.c:780-1253.
#pragma once
#include "mc33_core.cuh"
#include "mc33_simt.h"

namespace mc33 {

// ---------------------------------------------------------------------------
// shared pieces
// ---------------------------------------------------------------------------
// any row among (y .. y+n+1) x (slices z .. z+2) of the group starting at local row lr0 with an on-iso sample?
template <typename CX>
SIMT_FN bool group_oniso(const CX &cx, const Params &P, bool any, uint32_t lr0, uint32_t nrows)
{
	if (!any) return false;
	bool f = false;
	const uint32_t n = nrows + 2;
	for (uint32_t i = cx.lane(); i < 3 * n; i += 32) {
		const uint32_t dz = i / n, dy = i - dz * n;
		const uint64_t row = (uint64_t)lr0 + dy + (uint64_t)dz * P.NY;
		if (row < P.Lrows) f = f || P.rowZ[row] == P.zepoch;
	}
	return cx.any(f);
}

// exclusive scan of a 32-bit value across the warp; *total = sum
template <typename CX>
SIMT_FN uint32_t warp_exscan(const CX &cx, uint32_t v, uint32_t *total)
{
	uint32_t iv = v;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
	for (unsigned d = 1; d < 32; d <<= 1) {
		const uint32_t x = cx.shfl_up(iv, d);
		if (cx.lane() >= d) iv += x;
	}
	*total = cx.shfl(iv, 31);
	return iv - v;
}

// case index and on-iso corner mask of the cell at bit b, from the sign / on-iso words of its four point rows
// (rows 00, 10, 11, 01 = corners 0..3 at x, 4..7 at x+1; n* = the next word of the same row)
SIMT_HD unsigned corner_bits(uint32_t s00, uint32_t n00, uint32_t s10, uint32_t n10, uint32_t s11, uint32_t n11,
                             uint32_t s01, uint32_t n01, uint32_t b)
{
	const uint32_t p00 = funnel_r_clamp(s00, n00, b) & 3u, p10 = funnel_r_clamp(s10, n10, b) & 3u;
	const uint32_t p11 = funnel_r_clamp(s11, n11, b) & 3u, p01 = funnel_r_clamp(s01, n01, b) & 3u;
	// corner k -> bit 7-k
	return ((p00 & 1u) << 7) | ((p10 & 1u) << 6) | ((p11 & 1u) << 5) | ((p01 & 1u) << 4) |
	       ((p00 >> 1) << 3) | ((p10 >> 1) << 2) | ((p11 >> 1) << 1) | (p01 >> 1);
}
// index bits (corner k -> bit 7-k) to the zmask convention (corner k -> bit k)
SIMT_HD unsigned index_to_zmask(unsigned i)
{
	unsigned z = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
	for (int k = 0; k < 8; k++) z |= ((i >> (7 - k)) & 1u) << k;
	return z;
}

// which triangles of the pattern survive the zero-area drop (marching_cubes_33.c:1235) for on-iso corner mask zm:
// bit j = triangle j is kept
MC_COLD uint32_t keep_mask_walk(const Tables &tb, unsigned start, unsigned zm)
{
	uint32_t keep = 0;
	for (unsigned j = 0;; j++) {
		const unsigned tw = tb.tri[start + j];
		const unsigned k0 = vertex_key((tw >> 8) & 15, zm), k1 = vertex_key((tw >> 4) & 15, zm), k2 = vertex_key(tw & 15, zm);
		if (k0 != k1 && k0 != k2 && k1 != k2) keep |= 1u << j;
		if (!(tw >> 12)) break;
	}
	return keep;
}

// ... the same from the table built once per process (383 patterns x 256 corner masks): the walk above costs ~100
// instructions with the lanes of a warp in different iterations
SIMT_HD uint32_t keep_mask(const Tables &tb, unsigned start, unsigned zm) { return tb.keep[(uint32_t)tb.pord[start] * 256u + zm]; }

// ---------------------------------------------------------------------------
// K2: count
// ---------------------------------------------------------------------------
#define P2_CNT_WARPS 8
#define P2_CNT_CQ 256                          // queue entries per warp
#define P2_CNT_SMEM (3 * 1024 + 2 * 1024 + 128 + 64 + P2_CNT_WARPS * P2_CNT_CQ * 4)
#define P2_LB_WORDS 6                          // look-back words per block
#define P2_LB_K 1                              // predecessors per lane and look-back window (4 measured 0.005 ms slower at cfg2)

struct CountArgs {
	uint32_t nblk, GW;
	unsigned long long *lb;       // [nblk][6]: aggregate {tag<<48 | V, T, C}, inclusive prefix {tag<<48 | V, T, C}
	uint32_t *lb_ticket;          // [0] hands out block ids in launch order, [1] counts finished blocks; re-armed by the last to finish
	uint32_t tag;                 // 16-bit tag of this launch (1..65535): stale words of earlier launches do not match
	uint32_t owned_end_row;       // first row after the point rows this slab owns
	uint32_t *export4;            // optional DEVICE {nV, nT, nShared, nCentre} for the all-gather across slabs
	uint32_t dbg_noprefix;        // test hook (CPU emulation): only block 0 publishes an inclusive prefix, so every block walks
	                              // back over aggregates all the way (on the device the walk normally ends at the first window)
};

struct QuadZ { uint64_t pv[4]; uint32_t vis[4], walk[4]; uint32_t nts; };

// the words of quad q of row (z,y) with an on-iso sample in reach (bit k of slow): generic rules per WORD
template <typename Sample>
MC_COLD QuadZ count_quad_z(const Params &P, uint32_t z, uint32_t y, uint32_t q, uint32_t slow, bool own_p, bool own_c)
{
	QuadZ r;
	r.nts = 0;
	const uint32_t pm = row_points_owned(P, z) ? 0xFFFFFFFFu : 0u;
	for (int k = 0; k < 4; k++) {
		r.pv[k] = 0; r.vis[k] = 0; r.walk[k] = 0;
		const uint32_t w = 4 * q + (uint32_t)k;
		if (!((slow >> k) & 1u) || w >= P.W) continue;
		WordRec rec;
		CellWords cw;
		word_masks_generic(P, z, y, w, rec, cw);
		if (!own_c) rec.act = 0;
		r.vis[k] = rec.act | ((rec.X | rec.Y | rec.Z) & pm);
		if (!own_p) { rec.X = rec.Y = rec.Z = 0; }
		r.pv[k] = pack_planes(rec);
		uint32_t zcells = 0;
		if (cw.zany) zcells = (cw.zc[0] | cw.zc[1] | cw.zc[2] | cw.zc[3] | cw.zc[4] | cw.zc[5] | cw.zc[6] | cw.zc[7]) & rec.act;
		uint32_t cxm = 0;
		if (rec.act & ~zcells) r.nts += count_simple_cells(cw.c, rec.act & ~zcells, cxm);
		r.walk[k] = cxm | zcells;
	}
	return r;
}

// one queued cell: complex (pattern by the MC33 tests, remembered in pcache) and / or with an on-iso corner
// (triangle count after the zero-area drop) -> ntri | centre << 16
template <typename Sample>
SIMT_HD uint32_t count_walk_cell(const Params &P, const Tables &tb, uint32_t lr, uint32_t x, uint32_t y, uint32_t z, bool gz)
{
	typedef typename Traits<Sample>::Real Real;
	const uint32_t w = x >> 5, b = x & 31u;
	const uint32_t i00 = lr * P.WP + w, i10 = i00 + P.WP, i01 = i00 + P.NY * P.WP, i11 = i01 + P.WP;
	const unsigned idx = corner_bits(P.S[i00], P.S[i00 + 1], P.S[i10], P.S[i10 + 1], P.S[i11], P.S[i11 + 1], P.S[i01], P.S[i01 + 1], b);
	unsigned zm = 0;
	if (gz) zm = index_to_zmask(corner_bits(P.Z[i00], P.Z[i00 + 1], P.Z[i10], P.Z[i10 + 1], P.Z[i11], P.Z[i11 + 1], P.Z[i01], P.Z[i01 + 1], b));
	const uint32_t es = tb.cinfo[idx] & 0xFFFFu;
	unsigned start, ntri, centre = 0;
	if (es != 0xFFFFu) {
		start = es & 0xFFFu; ntri = es >> 12;
	} else {
		Real v[8];
		unsigned m;
		cell_values<Sample>(P, (Real)P.iso, x, y, z, v);
		start = select_pattern<Real>(tb, idx, v, &m);
		const unsigned pi = tb.pat[start];
		ntri = pi & 0x7Fu; centre = pi >> 7;
		P.pcache[(uint64_t)lr * (P.WP * 32u) + x] = (uint16_t)start;
	}
	if (zm) ntri = (unsigned)popc32(keep_mask(tb, start, zm));
	return ntri | (centre << 16);
}

template <typename CX>
SIMT_FN void block_exscan2(const CX &cx, uint64_t &a, uint64_t &b, uint64_t &ta, uint64_t &tb, uint64_t (*sw)[8])
{
	const unsigned lane = cx.lane(), wid = cx.warp();
	uint64_t ia = a, ib = b;
	for (unsigned d = 1; d < 32; d <<= 1) {
		const uint64_t xa = cx.shfl_up(ia, d), xb = cx.shfl_up(ib, d);
		if (lane >= d) { ia += xa; ib += xb; }
	}
	if (lane == 31) { sw[0][wid] = ia; sw[1][wid] = ib; }
	cx.syncthreads();
	uint64_t oa = 0, ob = 0, sa = 0, sb = 0;
	for (unsigned k = 0; k < P2_CNT_WARPS; k++) {
		const uint64_t va = sw[0][k], vb = sw[1][k];
		if (k < wid) { oa += va; ob += vb; }
		sa += va; sb += vb;
	}
	cx.syncthreads();
	a = oa + ia - a; b = ob + ib - b;   // exclusive
	ta = sa; tb = sb;
}

template <typename Sample, typename CX>
SIMT_FN void count_body(const CX &cx, const Params &P, const Tables &tb, const CountArgs &A)
{
	unsigned char *sm = cx.smem();
	uint32_t (*s_row)[256] = reinterpret_cast<uint32_t (*)[256]>(sm);             // [3][256] row totals: V, T of the simple cells ([2] spare)
	uint32_t (*s_cx)[256] = reinterpret_cast<uint32_t (*)[256]>(sm + 3072);       // [2][256] T, C of the queued cells
	uint64_t (*s_w)[8] = reinterpret_cast<uint64_t (*)[8]>(sm + 5120);            // [2][8]
	uint64_t *s_base = reinterpret_cast<uint64_t *>(sm + 5248);                   // [3] exclusive prefix of this block
	uint32_t *s_blk = reinterpret_cast<uint32_t *>(sm + 5280);
	uint32_t *queue = reinterpret_cast<uint32_t *>(sm + 5312) + cx.warp() * P2_CNT_CQ;
	const unsigned lane = cx.lane(), wid = cx.warp(), tid = cx.tid();
	const bool anyz = *P.anyZp != 0;
	const uint32_t RB = P2_CNT_WARPS * A.GW * P.G;       // rows per block (<= 256)
	const uint32_t npass = (P.Q + 31) / 32;

	// block ids in launch order: a block only ever waits for blocks that already run (or ran)
	if (tid == 0) *s_blk = cx.atomic_add(A.lb_ticket, 1u);
	s_row[0][tid] = 0; s_row[1][tid] = 0; s_row[2][tid] = 0;
	s_cx[0][tid] = 0; s_cx[1][tid] = 0;
	cx.syncthreads();
	const uint32_t blk = *s_blk;

	for (uint32_t sub = 0; sub < A.GW; sub++) {
		const uint32_t srow = (wid * A.GW + sub) * P.G;      // first row of the group within the block
		const uint32_t row0 = blk * RB + srow;
		if (row0 >= P.Lrows) break;                          // (warp uniform)
		const bool gz = group_oniso(cx, P, anyz, row0, P.G);
		uint64_t carryV = 0, carryT = 0;
		for (uint32_t pass = 0; pass < npass; pass++) {
			uint32_t r, q;
			if (P.Q <= 32) { r = fastdiv(lane, P.Q, P.mQ); q = lane - r * P.Q; }
			else { r = 0; q = pass * 32 + lane; }
			const uint32_t lr = row0 + r;
			const bool valid = r < P.G && q < P.Q && lr < P.Lrows;
			uint64_t pv0 = 0, pv1 = 0, pv2 = 0, pv3 = 0, tt = 0;
			uint32_t wk0 = 0, wk1 = 0, wk2 = 0, wk3 = 0;
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
					// words with an on-iso sample in reach take the generic rules (one cold call for the quad)
					const uint32_t slow = gz ? quad_oniso_mask(P.Z, i00, dY, dZ) : 0u;
					uint64_t pv[4];
					uint32_t walk[4], vis[4];
					uint32_t nts = 0;
					const uint32_t pm = row_points_owned(P, z) ? 0xFFFFFFFFu : 0u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
					for (int k = 0; k < 4; k++) {
						walk[k] = 0; pv[k] = 0; vis[k] = 0;
						if (!((slow >> k) & 1u)) {
							WordRec rec;
							uint32_t c[8];
							quad_word(P, q00, q10, q01, q11, k, 4 * q + k, own_c && hasZ, rec, c);
							vis[k] = rec.act | ((rec.X | rec.Y | rec.Z) & pm);
							if (!own_p) { rec.X = rec.Y = rec.Z = 0; }
							pv[k] = pack_planes(rec);
							// simple cells are counted 32 at a time; the complex ones go to the warp's queue
							if (rec.act) nts += count_simple_cells(c, rec.act, walk[k]);
						}
					}
					if (slow) {
						cx.atomic_add(&P.totals->nslow_acc, 1u);          // (statistic for the choice of cell kernel; rare on float data)
						const QuadZ rz = count_quad_z<Sample>(P, z, y, q, slow, own_p, own_c);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
						for (int k = 0; k < 4; k++)
							if ((slow >> k) & 1u) { pv[k] = rz.pv[k]; vis[k] = rz.vis[k]; walk[k] = rz.walk[k]; }
						nts += rz.nts;
					}
					pv0 = pv[0]; pv1 = pv[1]; pv2 = pv[2]; pv3 = pv[3];
#if defined(__CUDA_ARCH__)
					*reinterpret_cast<uint4 *>(P.A + i00) = make_uint4(vis[0], vis[1], vis[2], vis[3]);
#else
					P.A[i00] = vis[0]; P.A[i00 + 1] = vis[1]; P.A[i00 + 2] = vis[2]; P.A[i00 + 3] = vis[3];
#endif
					tt = nts;
					wk0 = walk[0]; wk1 = walk[1]; wk2 = walk[2]; wk3 = walk[3];
				}
			}
			// lane-local exclusive prefix over the four words, then the warp scan
			const uint64_t e1 = pv0, e2 = e1 + pv1, e3 = e2 + pv2, tv = e3 + pv3;
			uint64_t lv, it, iv = 0;
			if (P.Q <= 32) {
				// whole rows in one pass: a warp's 32 quads hold at most 4096 vertices per plane and 49152 triangles,
				// so the four counters scan as three 32-bit words (X | Y << 16, Z, T)
				const uint32_t a0 = fldV(tv, 0) | (fldV(tv, 1) << 16), b0 = fldV(tv, 2), c0 = (uint32_t)tt;
				uint32_t ia = a0, ib = b0, ic = c0;
				for (unsigned d = 1; d < 32; d <<= 1) {
					const uint32_t xa = cx.shfl_up(ia, d), xb = cx.shfl_up(ib, d), xc = cx.shfl_up(ic, d);
					if (lane >= d) { ia += xa; ib += xb; ic += xc; }
				}
				// exclusive, relative to the first quad of the lane's row
				const int srcl = (int)(r * P.Q < 31u ? r * P.Q : 31u);
				const uint32_t ea = ia - a0, eb = ib - b0, ec = ic - c0;
				const uint32_t ra = ea - cx.shfl(ea, srcl), rb = eb - cx.shfl(eb, srcl);
				const uint32_t rc0 = cx.shfl(ec, srcl);
				lv = (uint64_t)(ra & 0xFFFFu) | ((uint64_t)(ra >> 16) << 21) | ((uint64_t)(rb & 0xFFFFu) << 42);
				// fold the plane offsets in: Y ids follow the row's X ids, Z ids follow both
				const int lastl = (int)(r * P.Q + P.Q - 1 < 31u ? r * P.Q + P.Q - 1 : 31u);
				lv += plane_offsets(cx.shfl(lv + tv, lastl));
				it = (uint64_t)(ic - rc0);                       // triangles of the row up to and including this quad
			} else {
				uint64_t jt = tt;
				iv = tv;
				for (unsigned d = 1; d < 32; d <<= 1) {
					const uint64_t xv = cx.shfl_up(iv, d), xt = cx.shfl_up(jt, d);
					if (lane >= d) { iv += xv; jt += xt; }
				}
				lv = carryV + iv - tv;                           // row-local prefix in front of this quad
				it = jt;
			}
			if (valid) {
				uint64_t *pw = P.wpreV + (uint64_t)lr * P.WP + 4 * q;
				pw[0] = lv; pw[1] = lv + e1; pw[2] = lv + e2; pw[3] = lv + e3;
				if (q == P.Q - 1) {
					const uint64_t rowT = P.Q <= 32 ? it : carryT + it;
					pw[4] = lv + tv;
					s_row[0][srow + r] = P.Q <= 32 ? fldV(lv + tv, 2) : fldV(lv + tv, 0) + fldV(lv + tv, 1) + fldV(lv + tv, 2);
					s_row[1][srow + r] = (uint32_t)rowT;
				}
			}
			if (P.Q > 32) { carryV += cx.shfl(iv, 31); carryT += cx.shfl(it, 31); }

			// ---- the cells that have to be looked at: compacted across the warp, 32 at a time ----
			const uint32_t nw = (uint32_t)(popc32(wk0) + popc32(wk1) + popc32(wk2) + popc32(wk3));
			uint32_t nwt = 0, wpos = 0;
			if (cx.any(nw != 0u)) wpos = warp_exscan(cx, nw, &nwt);      // (smooth data: most passes have nothing to walk)
			for (uint32_t w0 = 0; w0 < nwt; w0 += P2_CNT_CQ) {
				if (nw && wpos < w0 + P2_CNT_CQ && wpos + nw > w0) {
					uint32_t slot = wpos - w0;                   // (may start "negative": wraps, compared unsigned below)
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
					for (int k = 0; k < 4; k++) {
						uint32_t m = k == 0 ? wk0 : (k == 1 ? wk1 : (k == 2 ? wk2 : wk3));
						while (m) {
							const int b = ffs32(m);
							m &= m - 1;
							if (slot < P2_CNT_CQ) queue[slot] = (((4 * q + (uint32_t)k) << 5) + (uint32_t)b) | (r << 16);
							slot++;
						}
					}
				}
				cx.syncwarp();
				const uint32_t ncw = nwt - w0 < P2_CNT_CQ ? nwt - w0 : P2_CNT_CQ;
				for (uint32_t j0 = 0; j0 < ncw; j0 += 32) {
					const bool on = j0 + lane < ncw;
					uint32_t res = 0, rr = 0;
					if (on) {
						const uint32_t e = queue[j0 + lane];
						rr = e >> 16;
						const uint32_t clr = row0 + rr, zl = fastdiv(clr, P.NY, P.mNY);
						res = count_walk_cell<Sample>(P, tb, clr, e & 0xFFFFu, clr - zl * P.NY, zl + P.zlo, gz);
					}
					// per-row sums: the queue is in row order, so a round holds one or two runs of equal rows
					uint32_t pending = cx.ballot(on);
					while (pending) {
						const int l0 = ffs32(pending);
						const uint32_t r0 = cx.shfl(rr, l0);
						const bool mine = on && rr == r0;
						const uint32_t st = cx.reduce_add(mine ? (res & 0xFFFFu) : 0u), sc = cx.reduce_add(mine ? (res >> 16) : 0u);
						if ((int)lane == l0) { s_cx[0][srow + r0] += st; s_cx[1][srow + r0] += sc; }
						pending &= ~cx.ballot(mine);
					}
				}
				cx.syncwarp();
			}
		}
		if (P.Q > 32) {
			// long rows: the plane totals are only known now; add the offsets in a second sweep
			const uint64_t add = plane_offsets(carryV);
			for (uint32_t i = lane; i <= 4 * P.Q; i += 32) P.wpreV[(uint64_t)row0 * P.WP + i] += add;
		}
	}
	cx.syncthreads();

	// ---- block-relative row bases, then the block's own prefix by decoupled look-back ----
	const uint32_t t = tid;
	uint64_t a = (uint64_t)s_row[0][t] | ((uint64_t)s_cx[1][t] << 32), b = (uint64_t)s_row[1][t] + s_cx[0][t], ta, tb2;
	block_exscan2(cx, a, b, ta, tb2, s_w);
	const uint64_t aggV = ta & 0xFFFFFFFFull, aggC = ta >> 32, aggT = tb2;
	const uint64_t FIELD = 0xFFFFFFFFFFFFull;                  // 48 value bits under the 16-bit tag
	if (wid == 0) {
		unsigned long long *me = A.lb + (uint64_t)blk * P2_LB_WORDS;
		const unsigned long long tagw = (unsigned long long)A.tag << 48;
		if (lane == 0 && blk > 0) {
			// aggregate first: successors can walk over this block while it is still looking back itself
			me[1] = aggT; me[2] = aggC;
			cx.st_release(&me[0], tagw | aggV);
		}
		uint64_t exV = 0, exT = 0, exC = 0;
		// A window is 32 x P2_LB_K predecessors: every lane has P2_LB_K independent loads in flight, so the blocks of one
		// wave -- which finish counting together and therefore all have to add up a wave's worth of aggregates --
		// need wave / 128 round trips to L2 instead of wave / 32.
		for (int64_t base = (int64_t)blk - 1; base >= 0; base -= 32 * P2_LB_K) {      // (warp uniform)
			unsigned long long v0[P2_LB_K], t1[P2_LB_K], c2[P2_LB_K];
			uint32_t dmin = 0xFFFFFFFFu;                        // distance of the nearest predecessor (of this lane's) with an inclusive prefix
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
			for (int k = 0; k < P2_LB_K; k++) {
				const int64_t j = base - (int64_t)(lane + 32u * (unsigned)k);
				v0[k] = 0; t1[k] = 0; c2[k] = 0;
				if (j >= 0) {
					const unsigned long long *o = A.lb + (uint64_t)j * P2_LB_WORDS;
					unsigned ns = 32;
					for (;;) {
						const unsigned long long p0 = cx.ld_acquire(&o[3]);
						if ((p0 >> 48) == A.tag) { v0[k] = p0; t1[k] = o[4]; c2[k] = o[5]; if (dmin == 0xFFFFFFFFu) dmin = lane + 32u * (unsigned)k; break; }
						const unsigned long long a0 = cx.ld_acquire(&o[0]);
						if ((a0 >> 48) == A.tag) { v0[k] = a0; t1[k] = o[1]; c2[k] = o[2]; break; }
						cx.backoff(ns);                              // neither published yet: that block is still counting
					}
				}
			}
			// the nearest predecessor that already has an inclusive prefix ends the walk; the ones nearer than it
			// contribute their aggregates
			const uint32_t first = cx.reduce_min(dmin);
			uint64_t sV = 0, sT = 0, sC = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
			for (int k = 0; k < P2_LB_K; k++)
				if (lane + 32u * (unsigned)k <= first) { sV += v0[k] & FIELD; sT += t1[k]; sC += c2[k]; }    // (slots before block 0 hold zeros)
			for (unsigned d = 16; d; d >>= 1) {
				sV += cx.shfl(sV, (int)(lane ^ d));
				sT += cx.shfl(sT, (int)(lane ^ d));
				sC += cx.shfl(sC, (int)(lane ^ d));
			}
			exV += sV; exT += sT; exC += sC;
			if (first != 0xFFFFFFFFu) break;
		}
		if (lane == 0) {
			me[4] = exT + aggT; me[5] = exC + aggC;
			if (!A.dbg_noprefix || blk == 0) cx.st_release(&me[3], tagw | ((exV + aggV) & FIELD));
			s_base[0] = exV; s_base[1] = exT; s_base[2] = exC;
		}
	}
	cx.syncthreads();
	{
		const uint64_t bV = s_base[0], bT = s_base[1], bC = s_base[2];
		const uint32_t lr = blk * RB + t;
		if (t < RB && lr < P.Lrows) {
			const uint32_t v = (uint32_t)(bV + (a & 0xFFFFFFFFull));
			P.rowBV[lr] = v; P.rowBC[lr] = (uint32_t)(bC + (a >> 32)); P.rowBT[lr] = (uint32_t)(bT + b);
			if (lr == A.owned_end_row) P.totals->nShared = v;
		}
		if (blk == A.nblk - 1 && t == 0) {
			const uint64_t tv = bV + aggV, tt = bT + aggT, tc = bC + aggC;
			P.rowBV[P.Lrows] = (uint32_t)tv; P.rowBT[P.Lrows] = (uint32_t)tt; P.rowBC[P.Lrows] = (uint32_t)tc;
			if (A.owned_end_row >= P.Lrows) P.totals->nShared = (uint32_t)tv;
			P.totals->nCentre = (uint32_t)tc;
			P.totals->nT = (uint32_t)tt;
			P.totals->nSharedAll = (uint32_t)tv;
			// 32-bit index range check (include/marching_cubes_33.h:140 uses unsigned int)
			P.totals->range = (tv + tc >= 0xFFFFFFFFull || tt >= 0xFFFFFFFFull) ? 1u : 0u;
		}
	}
	// the block that FINISHES last (whatever its id) re-arms the counters and exports the counts: every other
	// block's totals are visible to it (fence + atomic on the way out)
	cx.threadfence();
	cx.syncthreads();
	if (t == 0) {
		const uint32_t done = cx.atomic_add(A.lb_ticket + 1, 1u);
		if (done == A.nblk - 1) {
			cx.threadfence();
			A.lb_ticket[0] = 0; A.lb_ticket[1] = 0;
			P.totals->nslow = cx.atomic_add(&P.totals->nslow_acc, 0u); P.totals->nslow_acc = 0;
			P.totals->overflow = 0;
			P.totals->ticket = 0; P.totals->ticket2 = 0;
			if (A.export4) {
				const volatile Totals *tz = P.totals;
				const uint32_t nS = tz->nShared, nC = tz->nCentre, nT = tz->nT;
				A.export4[0] = nS + nC; A.export4[1] = nT; A.export4[2] = nS; A.export4[3] = nC;
			}
		}
	}
}

// ---------------------------------------------------------------------------
// K4: cells -> triangles, centre vertices, vertex tasks
//
// Records.  For every point row the group refers to and every word of the x-segment the warp keeps
//   A = {sign word, id of the first X / Y / Z vertex of the word}            (16 bytes)
// and, only when an on-iso sample is in reach of the group (gz),
//   B = {on-iso word, X / Y / Z plane masks with the on-iso rules applied}   (16 bytes).
// Without on-iso samples the plane masks are XORs of the sign words a cell loads anyway (X: the word
// against itself shifted by one, Y: against the row y+1, Z: against the row z+1), so B is not needed and
// twice as many rows fit: a group of Ge cell rows shares its 2 (Ge + 1) point rows.
// ---------------------------------------------------------------------------
#define P2_EM_WARPS 8
#define P2_R16 240                             // 16-byte record slots per warp
#define P2_TW 96                               // triangles per staging window
#define P2_EM_CQ 128                           // visited cells per queue window
#define P2_GEMAX 16                            // cell rows per group
#define P2_SLOTS (2 * (P2_GEMAX + 1))
#define P2_EM_WARP_BYTES (P2_R16 * 16 + P2_GEMAX * 32 + P2_SLOTS * 16 + P2_EM_CQ * 4 + (13 * 32 + 3 * P2_TW) * 4)      // 8224
#define P2_EM_UNIT 2                           // groups per ticket

struct EmitShape { uint32_t Ge, Ws, nseg, mNQ; };   // cell rows per (sub-)group, words per x-segment (multiple of 4), segments per row
struct EmitArgs {
	uint32_t pick;                // 0: this kernel always runs; else it runs only if cells_pick_records() says so
	uint32_t nquads;              // quads of the slab (for the pick)
	uint32_t row_begin, row_end;  // local rows [row_begin, row_end): cell rows + owned point rows of the slab
	EmitShape f, z;               // group shape without / with on-iso samples in reach (z.Ge <= f.Ge: a group is walked in sub-groups)
	uint32_t ngroups, nunits;
};

struct RowInfo { uint32_t o00, o10, o01, o11, y, z, flags, g0; };
struct Rec16 { uint32_t a, b, c, d; };

// shape for rows of Q quads with `cap` record slots: as many cell rows as fit in one x-segment, or one cell row in
// several segments when a row does not fit
SIMT_HD EmitShape emit_shape(uint32_t Q, uint32_t cap)
{
	EmitShape e;
	const uint32_t W4 = 4 * Q;
	const uint32_t rows = cap / (2 * (W4 + 1));                  // point rows per slice that fit
	if (rows >= 2) {
		e.Ge = rows - 1 < P2_GEMAX ? rows - 1 : P2_GEMAX; e.Ws = W4; e.nseg = 1;
	} else {
		e.Ge = 1; e.Ws = ((cap / 4 - 1) / 4) * 4; e.nseg = (W4 + e.Ws - 1) / e.Ws;
	}
	e.mNQ = e.Ws >= 8 ? (uint32_t)(0x100000000ull / (e.Ws >> 2)) : 0u;      // fastdiv by the quads of a full segment
	return e;
}
SIMT_HD void emit_geometry(uint32_t Q, EmitArgs &A)
{
	A.f = emit_shape(Q, P2_R16);
	A.z = emit_shape(Q, P2_R16 / 2);
}

// position of edge id e of the cell held by `lane` in the warp's id scratch: one row of 32 words per edge, so that the
// lanes of a warp hit 32 distinct banks whatever edges they ask for (a lane only ever reads its own column)
SIMT_HD uint32_t scr_pos(uint32_t lane, uint32_t e) { return e * 32u + lane; }

SIMT_HD uint32_t rank_of(uint32_t base, uint32_t mask, uint32_t below) { return base + (uint32_t)popc32(mask & below); }

// Which point rows the (sub-)group refers to: slot rs <= Ge is row lr0 + rs, slot Ge+1+rp the row one slice above
// row lr0 + rp.  Only the rows some visited cell refers to are staged -- the group's own rows, the row after a row
// with y < ny, and the rows one slice above those when the slice exists; anything else could reach beyond the
// slices this slab holds.  -> {local row or ~0, id base of the row (global ids), y, flags | z << 4}
// flags: 1 lower set, 2 has a row y+1, 4 has a row z+1 (inside the slab)
template <typename CX>
SIMT_FN void stage_slots(const CX &cx, const Params &P, uint32_t lr0, uint32_t nr, uint32_t Ge, uint32_t vb, uint32_t vbn, Rec16 *slot)
{
	for (uint32_t rs = cx.lane(); rs < 2 * (Ge + 1); rs += 32) {
		const bool lower = rs <= Ge;
		const uint32_t rp = lower ? rs : rs - Ge - 1;
		const uint64_t lrow = (uint64_t)lr0 + rp + (lower ? 0u : P.NY);
		Rec16 si;
		si.a = 0xFFFFFFFFu; si.b = 0; si.c = 0; si.d = 0;
		if ((uint64_t)lr0 + rp < P.Lrows) {
			const uint32_t lrl = lr0 + rp, zll = fastdiv(lrl, P.NY, P.mNY), y = lrl - zll * P.NY;
			uint32_t z = zll + P.zlo;
			bool need = rp < nr || (rp <= nr && y != 0u);
			if (!lower) { need = need && z < P.nz && lrow < P.Lrows; z += 1; }
			if (need) {
				si.a = (uint32_t)lrow;
				si.b = P.rowBV[lrow] + (z == P.hz ? vbn : vb);
				si.c = y;
				si.d = (lower ? 1u : 0u) | (y < P.ny ? 2u : 0u) | ((z < P.nz && lrow + P.NY < P.Lrows) ? 4u : 0u) | (z << 4);
			}
		}
		slot[rs] = si;
	}
}

// stage the records of one x-segment for the rows stage_slots named
template <typename CX>
SIMT_FN void stage_records(const CX &cx, const Params &P, uint32_t Ge, uint32_t w0, uint32_t nws, uint32_t Ws1, uint32_t mNQ,
                           bool gz, const Rec16 *slot, Rec16 *recA, Rec16 *recB)
{
	const uint32_t nq = nws >> 2, RS = 2 * (Ge + 1), nitems = RS * nq;
	for (uint32_t it = cx.lane(); it < nitems; it += 32) {
		const uint32_t rs = fastdiv(it, nq, mNQ), qi = it - rs * nq;
		const uint32_t o = rs * Ws1 + 4 * qi;
		const Rec16 si = slot[rs];
		Rec16 zero;
		zero.a = zero.b = zero.c = zero.d = 0;
		if (si.a == 0xFFFFFFFFu) {
			for (int k = 0; k < 4; k++) { recA[o + k] = zero; if (gz) recB[o + k] = zero; }
			if (qi == nq - 1) { recA[o + 4] = zero; if (gz) recB[o + 4] = zero; }
			continue;
		}
		const uint32_t w = w0 + 4 * qi;
		const uint64_t i0 = (uint64_t)si.a * P.WP + w;
		const Quad qs = load_quad(P.S, i0);
		const uint32_t rb = si.b;
		uint64_t pre[4];
#if defined(__CUDA_ARCH__)
		{
			const ulonglong2 p01 = *reinterpret_cast<const ulonglong2 *>(P.wpreV + i0), p23 = *reinterpret_cast<const ulonglong2 *>(P.wpreV + i0 + 2);
			pre[0] = p01.x; pre[1] = p01.y; pre[2] = p23.x; pre[3] = p23.y;
		}
#else
		for (int k = 0; k < 4; k++) pre[k] = P.wpreV[i0 + k];
#endif
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
		for (int k = 0; k < 4; k++) {
			Rec16 ra;
			ra.a = qs.s[k]; ra.b = rb + fldV(pre[k], 0); ra.c = rb + fldV(pre[k], 1); ra.d = rb + fldV(pre[k], 2);
			recA[o + k] = ra;
		}
		if (qi == nq - 1) { Rec16 ra = zero; ra.a = qs.s[4]; recA[o + 4] = ra; }
		if (gz) {
			// plane masks with the on-iso rules, once per WORD (never per cell)
			const bool lower = (si.d & 1u) != 0u, hasY = (si.d & 2u) != 0u, hasZ = (si.d & 4u) != 0u;
			const uint32_t y = si.c, z = si.d >> 4;
			const Quad qy = hasY ? load_quad(P.S, i0 + P.WP) : qs;
			const Quad qz = (lower && hasZ) ? load_quad(P.S, i0 + (uint64_t)P.NY * P.WP) : qs;
			for (int k = 0; k < 4; k++) {
				const uint32_t sw = qs.s[k];
				Rec16 rbk;
				rbk.a = P.Z[i0 + k];
				rbk.b = (sw ^ shr1(sw, qs.s[k + 1])) & mask_le(w + k, P.nx - 1); rbk.c = sw ^ qy.s[k]; rbk.d = sw ^ qz.s[k];
				uint32_t dep = rbk.a | (P.Z[i0 + k + 1] & 1u);
				if (hasY) dep |= P.Z[i0 + P.WP + k];
				if (hasZ) dep |= P.Z[i0 + (uint64_t)P.NY * P.WP + k];
				if (dep && w + k < P.W) {
					WordRec rec;
					CellWords cw;
					word_masks_generic(P, z, y, w + k, rec, cw);
					rbk.b = rec.X; rbk.c = rec.Y; rbk.d = rec.Z;
				}
				recB[o + k] = rbk;
			}
			if (qi == nq - 1) { Rec16 rbk = zero; rbk.a = P.Z[i0 + 4]; recB[o + 4] = rbk; }
		}
	}
}

// the k-th (0-based) set bit of m
SIMT_HD uint32_t nth_bit(uint32_t m, uint32_t k)
{
	for (uint32_t i = 0; i < k; i++) m &= m - 1;
	return (uint32_t)ffs32(m);
}

// Which cell kernel suits the data: the record kernel (this file) when more than about one quad in 64 has an on-iso
// sample in reach (integer grids with an integer isovalue), the direct kernel of round 1 otherwise.
SIMT_HD bool cells_pick_records(const Params &P, uint32_t nquads) { return (uint64_t)P.totals->nslow * 64u > nquads; }

template <typename Sample, bool KEYS, typename CX>
SIMT_FN void emit_cells_body(const CX &cx, const Params &P, const Tables &tb, const EmitArgs &A, unsigned char *wsm)
{
	if (A.pick && !cells_pick_records(P, A.nquads)) return;
	// per-warp shared memory
	Rec16 *recA = reinterpret_cast<Rec16 *>(wsm), *recB = recA + P2_R16 / 2;
	RowInfo *rowi = reinterpret_cast<RowInfo *>(wsm + P2_R16 * 16);
	Rec16 *slot = reinterpret_cast<Rec16 *>(wsm + P2_R16 * 16 + P2_GEMAX * 32);      // per staged point row: {row, id base, y, flags | z << 4}
	uint32_t *cq = reinterpret_cast<uint32_t *>(wsm + P2_R16 * 16 + P2_GEMAX * 32 + P2_SLOTS * 16);
	uint32_t *scr = cq + P2_EM_CQ;
	uint32_t *tst = scr + 13 * 32;                              // the round's triangles, 3 words each, before they go out as whole words
	const unsigned lane = cx.lane();
	const bool anyz = *P.anyZp == P.zepoch;
	const uint32_t nShared = P.totals->nShared;
	const uint32_t vb = P.dbases ? P.dbases[0] : P.vbase;
	const uint32_t vbn = (P.dbases ? P.dbases[1] : P.vbase_next) - nShared;   // halo slice: ids of the next slab
	const uint32_t WQ = 4 * P.Q;
	const uint32_t nwarps = cx.nblocks() * P2_EM_WARPS;
	if (cx.block() == 0 && cx.tid() == 0) P.totals->ticket2 = 0;     // re-arm the vertex kernel's counter (it is not running: stream order)

	// units of P2_EM_UNIT row groups are handed out by a ticket counter (the work of a group follows the surface);
	// every warp's first unit is its own index, the next ticket is fetched while the current unit runs
	uint32_t unit = cx.block() * P2_EM_WARPS + cx.warp(), unext = 0;
	for (; unit < A.nunits; unit = cx.shfl(unext, 0)) {
		if (lane == 0) unext = nwarps + cx.atomic_add(&P.totals->ticket, 1u);
		const uint32_t gend = (unit + 1) * P2_EM_UNIT < A.ngroups ? (unit + 1) * P2_EM_UNIT : A.ngroups;
		for (uint32_t gi = unit * P2_EM_UNIT; gi < gend; gi++) {
			const uint32_t glr0 = A.row_begin + gi * A.f.Ge;
			const uint32_t glrE = glr0 + A.f.Ge < A.row_end ? glr0 + A.f.Ge : A.row_end;
			// a group near an on-iso sample is walked in smaller sub-groups with the two-record layout
			const bool gz = group_oniso(cx, P, anyz, glr0, A.f.Ge);
			const EmitShape sh = gz ? A.z : A.f;
			const uint32_t Ge = sh.Ge, Ws1 = sh.Ws + 1;
			for (uint32_t lr0 = glr0; lr0 < glrE; lr0 += Ge) {
				const uint32_t lrE = lr0 + Ge < glrE ? lr0 + Ge : glrE, nr = lrE - lr0;
				cx.syncwarp();
				if (lane < nr) {
					const uint32_t lr = lr0 + lane, zl = fastdiv(lr, P.NY, P.mNY), y = lr - zl * P.NY, z = zl + P.zlo;
					const uint32_t uy = y < P.ny ? 1u : 0u, uz = z < P.nz ? 1u : 0u;
					RowInfo ri;
					ri.o00 = lane * Ws1; ri.o10 = (lane + uy) * Ws1;
					ri.o01 = uz ? (Ge + 1 + lane) * Ws1 : ri.o00; ri.o11 = uz ? (Ge + 1 + lane + uy) * Ws1 : ri.o10;
					ri.y = y; ri.z = z;
					ri.flags = (row_points_owned(P, z) ? 1u : 0u) | (row_cells_owned(P, z, y) ? 2u : 0u);
					ri.g0 = z == P.hz ? vbn : vb;
					rowi[lane] = ri;
				}
				stage_slots(cx, P, lr0, nr, Ge, vb, vbn, slot);
				const uint32_t tbase = P.rowBT[lr0], cloc0 = nShared + P.rowBC[lr0];
				uint32_t runT = 0, runC = 0;                             // triangles / centres of the sub-group so far
				for (uint32_t seg = 0; seg < sh.nseg; seg++) {
					const uint32_t w0 = seg * sh.Ws, nws = WQ - w0 < sh.Ws ? WQ - w0 : sh.Ws, nq = nws >> 2;
					cx.syncwarp();                                       // (the previous segment's readers are done)
					stage_records(cx, P, Ge, w0, nws, Ws1, nq == (sh.Ws >> 2) ? sh.mNQ : (nq >= 2 ? (uint32_t)(0x100000000ull / nq) : 0u), gz, slot, recA, recB);
					cx.syncwarp();
					const uint32_t nitems = nr * nq;
					const uint32_t mNQs = nq == (sh.Ws >> 2) ? sh.mNQ : (nq >= 2 ? (uint32_t)(0x100000000ull / nq) : 0u);
					for (uint32_t p0 = 0; p0 < nitems; p0 += 32) {
						// ---- fill: visited cells (active cells + grid points that own a vertex) in sweep order ----
						const uint32_t it = p0 + lane;
						uint32_t r = 0, qi = 0, act0 = 0, act1 = 0, act2 = 0, act3 = 0;
						if (it < nitems) {
							r = fastdiv(it, nq, mNQs); qi = it - r * nq;
							const uint64_t ia = (uint64_t)(lr0 + r) * P.WP + w0 + 4 * qi;
#if defined(__CUDA_ARCH__)
							const uint4 a = *reinterpret_cast<const uint4 *>(P.A + ia);
							act0 = a.x; act1 = a.y; act2 = a.z; act3 = a.w;
#else
							act0 = P.A[ia]; act1 = P.A[ia + 1]; act2 = P.A[ia + 2]; act3 = P.A[ia + 3];
#endif
						}
						const uint32_t na = (uint32_t)(popc32(act0) + popc32(act1) + popc32(act2) + popc32(act3));
						uint32_t ncp;
						const uint32_t pos0 = warp_exscan(cx, na, &ncp);
						for (uint32_t cw0 = 0; cw0 < ncp; cw0 += P2_EM_CQ) {
							if (na && pos0 < cw0 + P2_EM_CQ && pos0 + na > cw0) {
								uint32_t slot = pos0 - cw0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
								for (int k = 0; k < 4; k++) {
									uint32_t m = k == 0 ? act0 : (k == 1 ? act1 : (k == 2 ? act2 : act3));
									while (m) {
										const int b = ffs32(m);
										m &= m - 1;
										if (slot < P2_EM_CQ) cq[slot] = (((w0 + 4 * qi + (uint32_t)k) << 5) + (uint32_t)b) | (r << 16);
										slot++;
									}
								}
							}
							cx.syncwarp();
							const uint32_t ncw = ncp - cw0 < P2_EM_CQ ? ncp - cw0 : P2_EM_CQ;
							for (uint32_t j0 = 0; j0 < ncw; j0 += 32) {
								// ---- one lane per visited CELL ----
								const bool on = j0 + lane < ncw;
								uint32_t ntri = 0, centre = 0, sm = 0, keep = 0, x = 0, y = 0, z = 0;
								if (on) {
									const uint32_t e = cq[j0 + lane];
									x = e & 0xFFFFu;
									const uint32_t b = x & 31u, wl = (x >> 5) - w0, rr = e >> 16;
									const RowInfo ri = rowi[rr];
									y = ri.y; z = ri.z;
									const uint32_t a00 = ri.o00 + wl, a10 = ri.o10 + wl, a01 = ri.o01 + wl, a11 = ri.o11 + wl;
									const Rec16 R00 = recA[a00], R10 = recA[a10], R01 = recA[a01], R11 = recA[a11];
									const uint32_t n00 = recA[a00 + 1].a, n10 = recA[a10 + 1].a, n01 = recA[a01 + 1].a, n11 = recA[a11 + 1].a;
									uint32_t mX00, mY00, mZ00, mX10, mZ10, mX01, mY01, mX11, zown = 0;
									unsigned zm = 0;
									if (!gz) {
										// no on-iso sample in reach: the plane masks are XORs of the sign words (bits beyond the
										// row's last point are zero in S; the last point has no X edge)
										mX00 = (R00.a ^ shr1(R00.a, n00)); mX10 = (R10.a ^ shr1(R10.a, n10));
										mX01 = (R01.a ^ shr1(R01.a, n01)); mX11 = (R11.a ^ shr1(R11.a, n11));
										mY00 = R00.a ^ R10.a; mZ00 = R00.a ^ R01.a; mZ10 = R10.a ^ R11.a; mY01 = R01.a ^ R11.a;
										if (x >= P.nx) mX00 &= ~(1u << b);
									} else {
										const Rec16 B00 = recB[a00], B10 = recB[a10], B01 = recB[a01], B11 = recB[a11];
										mX00 = B00.b; mY00 = B00.c; mZ00 = B00.d; mX10 = B10.b; mZ10 = B10.d; mX01 = B01.b; mY01 = B01.c; mX11 = B11.b;
										zown = B00.a;
										zm = index_to_zmask(corner_bits(B00.a, recB[a00 + 1].a, B10.a, recB[a10 + 1].a, B11.a, recB[a11 + 1].a,
										                                B01.a, recB[a01 + 1].a, b));
									}
									const uint32_t below = (1u << b) - 1u;
									uint32_t id[12];
									// edges: 0:(0,1)y 1:(1,2)z 2:(3,2)y 3:(0,3)z 4:(4,5)y 5:(5,6)z 6:(7,6)y 7:(4,7)z 8..11 x
									id[0] = rank_of(R00.c, mY00, below); id[4] = id[0] + ((mY00 >> b) & 1u);
									id[1] = rank_of(R10.d, mZ10, below); id[5] = id[1] + ((mZ10 >> b) & 1u);
									id[2] = rank_of(R01.c, mY01, below); id[6] = id[2] + ((mY01 >> b) & 1u);
									id[3] = rank_of(R00.d, mZ00, below); id[7] = id[3] + ((mZ00 >> b) & 1u);
									id[8] = rank_of(R00.b, mX00, below); id[9] = rank_of(R10.b, mX10, below);
									id[10] = rank_of(R11.b, mX11, below); id[11] = rank_of(R01.b, mX01, below);
									// vertex tasks of the planes the low corner point owns (ids before any on-iso redirection)
									if ((ri.flags & 1u) && P.vtask) {
										const uint32_t lr = lr0 + rr;
										if ((mX00 >> b) & 1u) put_vertex_task(P, id[8] - ri.g0, lr, x, 0u, ((zown >> b) & 1u) != 0u);
										if ((mY00 >> b) & 1u) put_vertex_task(P, id[0] - ri.g0, lr, x, 1u, false);
										if ((mZ00 >> b) & 1u) put_vertex_task(P, id[3] - ri.g0, lr, x, 2u, false);
									}
									const unsigned idx = corner_bits(R00.a, n00, R10.a, n10, R11.a, n11, R01.a, n01, b);
									if ((ri.flags & 2u) && x < P.nx && idx != 0u && idx != 255u) {
										const uint32_t ci = tb.cinfo[idx];
										uint32_t start;
										if ((ci & 0xFFFFu) != 0xFFFFu) {
											start = ci & 0xFFFu; ntri = (ci >> 12) & 15u;
										} else {
											start = P.pcache[(uint64_t)(lr0 + rr) * (P.WP * 32u) + x];
											const uint32_t pi = tb.pat[start];
											ntri = pi & 0x7Fu; centre = pi >> 7;
										}
										sm = start | (((ci >> 16) & 1u) << 12);
										if (zm) {
											// on-iso corners: the edges that meet one refer to its POINT vertex (X plane of the corner's row at the
											// corner's x), and triangles that collapse are dropped (marching_cubes_33.c:970-988, :1235)
											uint32_t pid[8];
											pid[0] = id[8]; pid[1] = id[9]; pid[2] = id[10]; pid[3] = id[11];
											pid[4] = id[8] + ((mX00 >> b) & 1u); pid[5] = id[9] + ((mX10 >> b) & 1u);
											pid[6] = id[10] + ((mX11 >> b) & 1u); pid[7] = id[11] + ((mX01 >> b) & 1u);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
											for (unsigned ed = 0; ed < 12; ed++) {
												const unsigned ea = edge_a(ed), eb = edge_b(ed);
												if ((zm >> ea) & 1u) id[ed] = pid[ea];
												else if ((zm >> eb) & 1u) id[ed] = pid[eb];
											}
											keep = keep_mask(tb, start, zm);
											ntri = (uint32_t)popc32(keep);
											sm |= 0x2000u;
										}
									}
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
									for (uint32_t ed = 0; ed < 12; ed++) scr[scr_pos(lane, ed)] = id[ed];
								}
								// triangle / centre offsets of the round: shuffle scan in sweep order
								uint32_t tot;
								const uint32_t ex = warp_exscan(cx, ntri | (centre << 16), &tot);
								const uint32_t e0 = ex & 0xFFFFu, ntot = tot & 0xFFFFu;
								const uint32_t cl = cloc0 + runC + (ex >> 16);
								const uint64_t cell = ((uint64_t)z * P.ny + y) * P.nx + x;
								if (on && centre) {
									if (cl < P.capV) {
										emit_centre_vertex<Sample>(P, x, y, z, cl);
										if (KEYS && P.vkey) P.vkey[cl] = cell * 4 + 3;
									} else {
										P.totals->overflow = 1;
									}
								}
								if (on) scr[scr_pos(lane, 12)] = vb + cl;
								// ---- triangles: every cell lane writes its own (pattern walk, ids from its own column, winding; cells with
								// an on-iso corner skip the triangles their keep mask drops) into a shared-memory window at their positions
								// of the round, then the warp copies the window out as consecutive words (whole-sector stores) ----
								{
									const bool swp = (((sm >> 12) & 1u) != 0u) != (P.geom.normal_neg != 0);
									for (uint32_t t0 = 0; t0 < ntot; t0 += P2_TW) {
										if (on && e0 < t0 + P2_TW && e0 + ntri > t0) {
											const uint32_t jlo = t0 > e0 ? t0 - e0 : 0u, jhi = ntri < t0 + P2_TW - e0 ? ntri : t0 + P2_TW - e0;
											for (uint32_t j = jlo; j < jhi; j++) {
												const unsigned tw = tb.tri[(sm & 0xFFFu) + ((sm & 0x2000u) ? nth_bit(keep, j) : j)];
												const uint32_t i0 = scr[scr_pos(lane, (tw >> 8) & 15u)], i1 = scr[scr_pos(lane, (tw >> 4) & 15u)];
												uint32_t *d = tst + 3 * (e0 + j - t0);
												// winding: marching_cubes_33.c:1246-1250 (nibble 2, nibble 1, nibble 0; m swaps the first two)
												d[0] = swp ? i0 : i1; d[1] = swp ? i1 : i0; d[2] = scr[scr_pos(lane, tw & 15u)];
												if (KEYS && P.tcell) { const uint32_t tj = tbase + runT + e0 + j; if (tj < P.capT) P.tcell[tj] = cell; }
											}
										}
										cx.syncwarp();
										const uint32_t nw = 3u * (ntot - t0 < (uint32_t)P2_TW ? ntot - t0 : (uint32_t)P2_TW);
										const uint64_t wbase = (uint64_t)(tbase + runT + t0) * 3u, wcap = (uint64_t)P.capT * 3u;
										for (uint32_t w = lane; w < nw; w += 32) {
#if defined(__CUDA_ARCH__) && MC33_STREAM_STORES
											if (wbase + w < wcap) __stcs(P.T + (wbase + w), tst[w]);
#else
											if (wbase + w < wcap) P.T[wbase + w] = tst[w];
#endif
											else P.totals->overflow = 1;
										}
										cx.syncwarp();
									}
								}
								cx.syncwarp();
								runT += ntot; runC += tot >> 16;
							}
						}
					}
				}
			}
		}
	}
}

// ---------------------------------------------------------------------------
// K3: shared vertices (edge / on-iso point vertices), straight from the bitmaps.
//
// The vertex ids of a run of point rows are consecutive, and the count kernel left, per word and plane, the
// row-local index of the plane's first vertex (wpreV).  A warp takes a run of rows, rebuilds the three plane
// masks of every word (XORs of sign words; the on-iso rules per word where an on-iso sample is in reach) and
// drops one 4-byte entry per vertex at its own position of a shared-memory window -- no scan, no per-vertex
// task in global memory (round 1 wrote and re-read 8 bytes per vertex).  Lane t of a round then computes the
// vertex at position t; positions, normals and colours leave through a staging buffer as consecutive words,
// so that every store instruction fills whole sectors (3 x 4-byte components per lane used to touch each
// sector three times).
// ---------------------------------------------------------------------------
#define P2_VX_WARPS 8
#define P2_VQ 512                              // vertices per window
#define P2_VX_WARP_BYTES (P2_VQ * 4 + 32 * 6 * 4)     // window + staging for 32 x 3 components (float or double)

struct VertexArgs {
	uint32_t row_begin, row_end;  // local point rows whose shared vertices this slab owns
	uint32_t Gv;                  // rows per group
	uint32_t ngroups;
};

// rows per group: about two passes of 32 quads
SIMT_HD uint32_t vertex_group_rows(uint32_t Q) { return Q >= 64 ? 1u : 64u / Q; }

template <typename Sample, bool KEYS, typename CX>
SIMT_FN void emit_vertices_body(const CX &cx, const Params &P, const VertexArgs &A, unsigned char *wsm)
{
	typedef typename Traits<Sample>::Real Real;
	uint32_t *vq = reinterpret_cast<uint32_t *>(wsm);
	uint32_t *st = vq + P2_VQ;
	const unsigned lane = cx.lane();
	const bool anyz = *P.anyZp == P.zepoch;
	const uint32_t nwarps = cx.nblocks() * P2_VX_WARPS;
	const uint32_t nq = P.Q;
	if (cx.block() == 0 && cx.tid() == 0) P.totals->ticket = 0;      // re-arm the cell kernel's counter (it has finished: stream order)

	uint32_t g = cx.block() * P2_VX_WARPS + cx.warp(), gnext = 0;
	for (; g < A.ngroups; g = cx.shfl(gnext, 0)) {
		if (lane == 0) gnext = nwarps + cx.atomic_add(&P.totals->ticket2, 1u);
		const uint32_t lr0 = A.row_begin + g * A.Gv, lrE = lr0 + A.Gv < A.row_end ? lr0 + A.Gv : A.row_end, nr = lrE - lr0;
		const uint32_t id0 = P.rowBV[lr0], nvg = P.rowBV[lrE] - id0;   // the group's vertices are [id0, id0 + nvg)
		if (nvg == 0) continue;
		const bool gz = group_oniso(cx, P, anyz, lr0, A.Gv);
		const uint32_t nitems = nr * nq;
		for (uint32_t v0 = 0; v0 < nvg; v0 += P2_VQ) {
			// ---- entries of the window [v0, v0 + VQ): every (row, quad) drops its vertices at their own positions ----
			cx.syncwarp();
			for (uint32_t it = lane; it < nitems; it += 32) {
				const uint32_t r = it / nq, q = it - r * nq;
				const uint32_t lr = lr0 + r;
				const uint64_t i0 = (uint64_t)lr * P.WP + 4 * q;
#if defined(__CUDA_ARCH__)
				const uint4 av = *reinterpret_cast<const uint4 *>(P.A + i0);
				const uint32_t anyv = av.x | av.y | av.z | av.w;
#else
				const uint32_t anyv = P.A[i0] | P.A[i0 + 1] | P.A[i0 + 2] | P.A[i0 + 3];
#endif
				if (!anyv) continue;                                  // no visited point in this quad: no vertex either
				const uint32_t zl = fastdiv(lr, P.NY, P.mNY), y = lr - zl * P.NY, z = zl + P.zlo;
				const bool hasY = y < P.ny, hasZ = z < P.nz;
				const Quad qs = load_quad(P.S, i0);
				const Quad qy = hasY ? load_quad(P.S, i0 + P.WP) : qs;
				const Quad qz = hasZ ? load_quad(P.S, i0 + (uint64_t)P.NY * P.WP) : qs;
				const uint32_t rbase = P.rowBV[lr] - id0;
				for (int k = 0; k < 4; k++) {
					const uint32_t w = 4 * q + (uint32_t)k;
					if (w >= P.W) break;
					const uint32_t sw = qs.s[k];
					uint32_t m[3], zw = 0;
					m[0] = (sw ^ shr1(sw, qs.s[k + 1])) & mask_le(w, P.nx - 1); m[1] = sw ^ qy.s[k]; m[2] = sw ^ qz.s[k];
					if (gz) {
						zw = P.Z[i0 + k];
						uint32_t dep = zw | (P.Z[i0 + k + 1] & 1u);
						if (hasY) dep |= P.Z[i0 + P.WP + k];
						if (hasZ) dep |= P.Z[i0 + (uint64_t)P.NY * P.WP + k];
						if (dep) {
							WordRec rec;
							CellWords cw;
							word_masks_generic(P, z, y, w, rec, cw);
							m[0] = rec.X; m[1] = rec.Y; m[2] = rec.Z;
						}
					}
					if (!(m[0] | m[1] | m[2])) continue;
					const uint64_t pre = P.wpreV[i0 + k];
					for (int a = 0; a < 3; a++) {
						uint32_t mm = m[a];
						uint32_t slot = rbase + fldV(pre, a) - v0;          // (wraps below the window: compared unsigned)
						while (mm) {
							const int b = ffs32(mm);
							mm &= mm - 1;
							if (slot < P2_VQ)
								vq[slot] = ((w << 5) + (uint32_t)b) | ((uint32_t)a << 16) | ((a == 0 && ((zw >> b) & 1u)) ? 1u << 18 : 0u) | (r << 19);
							slot++;
						}
					}
				}
			}
			cx.syncwarp();
			// ---- one lane per vertex, consecutive ids ----
			const uint32_t nvw = nvg - v0 < P2_VQ ? nvg - v0 : P2_VQ;
			for (uint32_t j0 = 0; j0 < nvw; j0 += 32) {
				const bool on = j0 + lane < nvw;
				const uint32_t idr = id0 + v0 + j0;                        // id of lane 0's vertex
				Real Vo[3] = {0, 0, 0};
				float No[3] = {0, 0, 0};
				if (on) {
					const uint32_t e = vq[j0 + lane];
					const uint32_t x = e & 0xFFFFu, a = (e >> 16) & 3u, lr = lr0 + (e >> 19);
					const uint32_t zl = fastdiv(lr, P.NY, P.mNY), y = lr - zl * P.NY, z = zl + P.zlo;
					Real r[6];
					if ((e >> 18) & 1u) point_vertex_r<Sample>(P, x, y, z, r);
					else edge_vertex_r<Sample>(P, x, y, z, (int)a, r);
					transform_vertex<Real>(P, r, Vo, No);
					if (KEYS && P.vkey && idr + lane < P.capV) P.vkey[idr + lane] = (((uint64_t)z * P.NY + y) * P.NX + x) * 4 + a;
				}
				const uint32_t nround = nvw - j0 < 32u ? nvw - j0 : 32u;
				if (idr + nround > P.capV) { if (lane == 0) P.totals->overflow = 1; }
				const uint64_t vcap = (uint64_t)P.capV * 3u, wb = (uint64_t)idr * 3u;
				// positions
				if (sizeof(Real) == 4) {
					float *sf = reinterpret_cast<float *>(st);
					sf[3 * lane] = (float)Vo[0]; sf[3 * lane + 1] = (float)Vo[1]; sf[3 * lane + 2] = (float)Vo[2];
					cx.syncwarp();
					float *Vg = reinterpret_cast<float *>(P.V);
					for (uint32_t k = 0; k < 3; k++) {
						const uint32_t w = lane + 32u * k;
						if (w < 3u * nround && wb + w < vcap) Vg[wb + w] = sf[w];
					}
				} else {
					double *sd = reinterpret_cast<double *>(st);
					sd[3 * lane] = (double)Vo[0]; sd[3 * lane + 1] = (double)Vo[1]; sd[3 * lane + 2] = (double)Vo[2];
					cx.syncwarp();
					double *Vg = reinterpret_cast<double *>(P.V);
					for (uint32_t k = 0; k < 3; k++) {
						const uint32_t w = lane + 32u * k;
						if (w < 3u * nround && wb + w < vcap) Vg[wb + w] = sd[w];
					}
				}
				cx.syncwarp();
				// normals
				{
					float *sf = reinterpret_cast<float *>(st);
					sf[3 * lane] = No[0]; sf[3 * lane + 1] = No[1]; sf[3 * lane + 2] = No[2];
					cx.syncwarp();
					for (uint32_t k = 0; k < 3; k++) {
						const uint32_t w = lane + 32u * k;
						if (w < 3u * nround && wb + w < vcap) P.N[wb + w] = sf[w];
					}
				}
				if (on && idr + lane < P.capV) P.color[idr + lane] = P.color_value;
				cx.syncwarp();
			}
		}
	}
}

}  // namespace mc33
