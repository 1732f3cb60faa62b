// mc33_hostemu.cu -- TEST INFRASTRUCTURE ONLY.
//
// Steps the __host__ __device__ building blocks of
// mc33_c_library_b200/csrc/mc33_core.cuh on the CPU, one (row, word) at a time,
// with plain loops standing in for the kernels' thread mapping, ballots and
// scans.  It exists so that the per-word logic can be checked against the oracle
// in a container without a GPU.  It is never linked into, or reachable from,
// the product library; the GPU tests exercise the real kernels.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../mc33_c_library_b200/csrc/mc33_core.cuh"
#include "../../include/mc33cu.h"

using namespace mc33;

template <typename Sample>
static void run(Params &P, bool emit)
{
	typedef typename Traits<Sample>::Real Real;
	Tables tb;
	static uint8_t pat[MC33_NTRI_WORDS];
	for (int i = 0; i < MC33_NTRI_WORDS; i++) pat[i] = (uint8_t)(MC33_PAT_NTRI[i] | (MC33_PAT_CENTRE[i] << 7));
	tb.case256 = MC33_CASE256; tb.simple256 = MC33_SIMPLE256; tb.tri = MC33_TRI; tb.pat = pat;
	const Real iso = (Real)P.iso;
	// K1 (the kernel takes the two bits by comparison; here as the reference does)
	for (uint32_t lr = 0; lr < P.Lrows; lr++) {
		const Sample *src = (const Sample *)P.data + (uint64_t)lr * P.NX;
		bool anyz = false;
		for (uint32_t w = 0; w < P.W; w++) {
			uint32_t s = 0, z = 0;
			for (uint32_t b = 0; b < 32; b++) {
				uint32_t x = (w << 5) + b;
				if (x >= P.NX) break;
				Real v = rsub(iso, (Real)src[x]);
				if (sgn(v)) s |= 1u << b;
				if (v == (Real)0) z |= 1u << b;
			}
			P.S[(uint64_t)lr * P.WP + w] = s;
			P.Z[(uint64_t)lr * P.WP + w] = z;
			anyz |= z != 0;
		}
		P.rowZ[lr] = anyz ? P.zepoch : 0u;
		if (anyz) *P.anyZp = 1;
	}
	const bool gz = *P.anyZp != 0;
	// per-word record of row (z,y): through the quad fast path when the word has no
	// on-iso sample among the points it depends on (as the kernels do), else the generic walk
	auto word_rec = [&](uint32_t z, uint32_t y, uint32_t w, WordRec &rec, CellWords &cw) {
		if (gz && word_oniso(P, z, y, w)) { word_masks(P, z, y, w, true, rec, cw); return; }
		const uint32_t lr = (z - P.zlo) * P.NY + y, q = w >> 2;
		const bool hasY = y < P.ny, hasZ = z < P.nz;
		const uint64_t dY = hasY ? P.WP : 0u, dZ = hasZ ? (uint64_t)P.NY * P.WP : 0u, i00 = (uint64_t)lr * P.WP + 4 * q;
		const Quad q00 = load_quad(P.S, i00), q10 = load_quad(P.S, i00 + dY), q01 = load_quad(P.S, i00 + dZ),
		           q11 = load_quad(P.S, i00 + dY + dZ);
		quad_word(P, q00, q10, q01, q11, (int)(w & 3), w, hasY && hasZ, rec, cw.c);
		cw.zany = 0;
		for (int k = 0; k < 8; k++) cw.zc[k] = 0;
	};
	// K2 + K2b: word prefixes, row bases
	{
		uint64_t bv = 0, bc = 0, bt = 0;
		const uint32_t owned_end = (P.pz1 - P.zlo) * P.NY;
		for (uint32_t lr = 0; lr < P.Lrows; lr++) {
			const uint32_t zl = lr / P.NY, y = lr - zl * P.NY, z = zl + P.zlo;
			const bool own_p = row_points_owned(P, z) || row_points_halo(P, z);
			const bool own_c = row_cells_owned(P, z, y);
			uint64_t av = 0, ac = 0;
			for (uint32_t w = 0; w < 4 * P.Q; w++) {
				uint64_t cv = 0, cc = 0;
				if ((own_p || own_c) && w < P.W) {
					WordRec rec; CellWords cw;
					word_rec(z, y, w, rec, cw);
					if (!own_p) { rec.X = rec.Y = rec.Z = 0; }
					if (!own_c) rec.act = 0;
					// visit mask for the cell kernel: active cells, points owning a vertex emitted here
					P.A[(uint64_t)lr * P.WP + w] = rec.act | (row_points_owned(P, z) ? (rec.X | rec.Y | rec.Z) : 0u);
					cv = pack_planes(rec);
					if (gz && word_oniso(P, z, y, w)) {
						cc = count_cells<Sample>(P, tb, z, y, w, rec.act, cw.c, cw.zc, cw.zany);
					} else {
						// the kernel's merged walk over the quad: give it this word's cells only
						uint32_t act4[4] = {0, 0, 0, 0};
						// simple cells 32 at a time, the complex ones one by one (as k_count does)
						uint32_t cxm = 0;
						const uint32_t nts = rec.act ? count_simple_cells(cw.c, rec.act, cxm) : 0u;
						act4[w & 3] = cxm;
						const bool hasY = y < P.ny, hasZ = z < P.nz;
						const uint64_t dY = hasY ? P.WP : 0u, dZ = hasZ ? (uint64_t)P.NY * P.WP : 0u, i00 = (uint64_t)lr * P.WP + (w & ~3u);
						cc = count_cells_quad<Sample>(P, tb, z, y, w >> 2, act4[0], act4[1], act4[2], act4[3], i00, dY, dZ);
						cc += nts;
					}
				}
				P.wpreV[(uint64_t)lr * P.WP + w] = av;
				av += cv; ac += cc;
			}
			P.wpreV[(uint64_t)lr * P.WP + 4 * P.Q] = av;
			for (uint32_t w = 0; w <= 4 * P.Q; w++) P.wpreV[(uint64_t)lr * P.WP + w] += plane_offsets(av);
			if (lr == owned_end) P.totals->nShared = (uint32_t)bv;
			P.rowBV[lr] = (uint32_t)bv; P.rowBT[lr] = (uint32_t)bt; P.rowBC[lr] = (uint32_t)bc;
			bv += fldV(av, 0) + fldV(av, 1) + fldV(av, 2); bt += ac & 0xFFFFFFFFu; bc += ac >> 32;
		}
		P.rowBV[P.Lrows] = (uint32_t)bv; P.rowBT[P.Lrows] = (uint32_t)bt; P.rowBC[P.Lrows] = (uint32_t)bc;
		if (owned_end >= P.Lrows) P.totals->nShared = (uint32_t)bv;
		P.totals->nCentre = (uint32_t)bc; P.totals->nT = (uint32_t)bt; P.totals->nSharedAll = (uint32_t)bv;
	}
	if (!emit) return;
	// K4, cells: triangles (+ centre vertices) in sweep order, vertex tasks
	const uint32_t vb = P.dbases ? P.dbases[0] : P.vbase;
	const uint32_t vbn = (P.dbases ? P.dbases[1] : P.vbase_next) - P.totals->nShared;
	const uint32_t zend = P.pz1 > P.cz1 ? P.pz1 : P.cz1;
	for (uint32_t lr = (P.cz0 - P.zlo) * P.NY; lr < (zend - P.zlo) * P.NY; lr++) {
		const uint32_t z = lr / P.NY + P.zlo, y = lr % P.NY;
		const bool own_c = row_cells_owned(P, z, y), own_p = row_points_owned(P, z);
		if (!own_c && !own_p) continue;
		uint32_t tid = P.rowBT[lr], cl = P.totals->nShared + P.rowBC[lr];
		for (uint32_t w = 0; w < P.W; w++) {
			// visited: active cells and grid points that own a vertex (the mask k_count left)
			uint32_t act = P.A[(uint64_t)lr * P.WP + w];
			if (!act) continue;
			const bool slow = gz && word_oniso(P, z, y, w);
			while (act) {
				int b = ffs32(act);
				act &= act - 1;
				const uint32_t x = (w << 5) + b;
				const bool cellok = own_c && x < P.nx;
				uint32_t ids[13];
				unsigned zm = 0;
				CellPattern cpat;
				cpat.start = 0; cpat.m = 0; cpat.ntri = 0; cpat.centre = 0;
				if (slow) {
					cpat = cell_slow<Sample>(P, tb, x, y, z, own_p, cellok, ids, 1, zm);
				} else {
					unsigned own;
					const uint32_t g0 = z == P.hz ? vbn : vb;
					const unsigned idx = cell_fast(P, x, y, z, g0, z + 1 == P.hz ? vbn : vb, ids, own);
					if (own_p) {
						if (own & 1u) put_vertex_task(P, ids[8] - g0, lr, x, 0u, false);
						if (own & 2u) put_vertex_task(P, ids[0] - g0, lr, x, 1u, false);
						if (own & 4u) put_vertex_task(P, ids[3] - g0, lr, x, 2u, false);
					}
					if (cellok) cpat = cell_pattern<Sample>(P, tb, x, y, z, idx, 0u);
				}
				const uint64_t cell = ((uint64_t)z * P.ny + y) * P.nx + x;
				if (cpat.centre) {
					if (cl < P.capV) {
						emit_centre_vertex<Sample>(P, x, y, z, cl);
						if (P.vkey) P.vkey[cl] = cell * 4 + 3;
					} else {
						P.totals->overflow = 1;
					}
				}
				if (!zm) {
					ids[12] = vb + cl;
					for (uint32_t j = 0; j < cpat.ntri; j++)
						emit_triangle_fast(P, tb.tri[cpat.start + j], cpat.m, ids, 1, tid + j, cell);
				} else {
					cell_slow_triangles<Sample>(P, tb, x, y, z, cpat, zm, vb + cl, tid, cell);
				}
				tid += cpat.ntri;
				cl += cpat.centre;
			}
		}
	}
	// K3, vertices: dense over the ids, each from the task left in its slot
	for (uint32_t id = 0; id < P.totals->nShared && id < P.capV; id++) run_vertex_task<Sample>(P, id);
}

extern "C" int mc33emu_run(const mc33cu_desc *d, const void *data, double iso, const mc33cu_out *o,
                           mc33cu_counts *counts)
{
	Params P;
	memset(&P, 0, sizeof P);
	const uint32_t NZ = d->nz + 1;
	P.data = data;
	P.nx = d->nx; P.ny = d->ny; P.nz = d->nz; P.NX = d->nx + 1; P.NY = d->ny + 1;
	P.zlo = d->z_lo; P.zhi = d->z_hi; P.cz0 = d->cell_z0; P.cz1 = d->cell_z1;
	P.pz0 = d->cell_z0; P.pz1 = d->is_last ? NZ : d->cell_z1;
	P.hz = d->is_last ? 0xFFFFFFFFu : d->cell_z1;
	P.W = (P.NX + 31) / 32; P.WC = (P.nx + 31) / 32; P.Q = (P.W + 3) / 4; P.WP = 4 * P.Q + 4;
	P.G = P.Q <= 32 ? 32 / P.Q : 1;
	P.Lrows = (P.zhi - P.zlo) * P.NY;
	P.mQ = P.Q >= 2 ? (uint32_t)(0x100000000ull / P.Q) : 0u; P.mNY = (uint32_t)(0x100000000ull / P.NY);
	P.geom.store = d->store; P.geom.normal_neg = d->normal_neg; P.geom.tsa = d->tsa;
	for (int i = 0; i < 3; i++) { P.geom.O[i] = d->O[i]; P.geom.D[i] = d->D[i]; }
	P.geom.ca = d->ca; P.geom.cb = d->cb;
	for (int i = 0; i < 3; i++) { P.geom.Of[i] = (float)d->O[i]; P.geom.Df[i] = (float)d->D[i]; }
	P.geom.caf = (float)d->ca; P.geom.cbf = (float)d->cb;
	for (int i = 0; i < 9; i++) { P.geom.A[i] = d->A[i]; P.geom.Ai[i] = d->Ai[i]; }
	P.iso = d->dtype == MC33CU_F64 ? iso + 0.0 : (double)((float)iso + 0.0f);
	std::vector<uint32_t> S((size_t)P.Lrows * P.WP, 0), Z((size_t)P.Lrows * P.WP, 0), rb(((size_t)P.Lrows + 1) * 3);
	std::vector<uint32_t> rz(P.Lrows), A((size_t)P.Lrows * P.WP, 0);
	P.A = A.data();
	P.zepoch = 1;
	std::vector<uint64_t> wv((size_t)P.Lrows * P.WP);
	Totals tot;
	memset(&tot, 0, sizeof tot);
	P.S = S.data(); P.Z = Z.data(); P.rowZ = rz.data(); P.wpreV = wv.data();
	P.rowBV = rb.data(); P.rowBT = P.rowBV + (P.Lrows + 1); P.rowBC = P.rowBT + (P.Lrows + 1);
	P.totals = &tot;
	P.anyZp = &tot.anyZ;
	bool emit = o != nullptr;
	std::vector<uint64_t> vtask(emit ? (size_t)o->capV + 1 : 1);
	P.vtask = vtask.data();
	if (emit) {
		P.V = o->V; P.N = o->N; P.color = o->color; P.T = o->T; P.vkey = o->vkey; P.tcell = o->tcell;
		P.capV = o->capV; P.capT = o->capT; P.vbase = o->vbase; P.vbase_next = o->vbase_next; P.dbases = o->dev_bases;
		P.color_value = o->color_value;
	}
	switch (d->dtype) {
	case MC33CU_F32: run<float>(P, emit); break;
	case MC33CU_F64: run<double>(P, emit); break;
	case MC33CU_U8:  run<uint8_t>(P, emit); break;
	case MC33CU_U16: run<uint16_t>(P, emit); break;
	case MC33CU_U32: run<uint32_t>(P, emit); break;
	default: return -1;
	}
	if (counts) {
		counts->nShared = tot.nShared; counts->nCentre = tot.nCentre; counts->nT = tot.nT;
		counts->nSharedHalo = tot.nSharedAll - tot.nShared; counts->nV = (uint64_t)tot.nShared + tot.nCentre;
	}
	return tot.overflow ? MC33CU_ERR_CAPACITY : 0;
}

// the bit-parallel triangle count of k_count (mc33_core.cuh count_simple_cells), for
// the exhaustive check against the case table
extern "C" uint32_t mc33emu_count_simple(const uint32_t *c8, uint32_t act, uint32_t *cx)
{
	return count_simple_cells(c8, act, *cx);
}
extern "C" const uint16_t *mc33emu_simple256(void) { return MC33_SIMPLE256; }
