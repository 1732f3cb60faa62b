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
	// K1
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
		P.rowZ[lr] = anyz;
	}
	// K2
	for (uint32_t lr = 0; lr < P.Lrows; lr++) {
		const uint32_t zl = lr / P.NY, y = lr - zl * P.NY, z = zl + P.zlo;
		const bool own_p = row_points_owned(P, z) || row_points_halo(P, z);
		const bool own_c = row_cells_owned(P, z, y);
		uint64_t av = 0, ac = 0;
		for (uint32_t w = 0; w < P.W; w++) {
			uint64_t cv = 0, cc = 0;
			if (own_p || own_c) count_word<Sample>(P, tb, z, y, w, own_p, own_c, cv, cc);
			P.wpreV[(uint64_t)lr * P.W + w] = av;
			P.wpreC[(uint64_t)lr * P.W + w] = ac;
			av += cv; ac += cc;
		}
		P.rowNX[lr] = (uint32_t)(av & 0x1FFFFF);
		P.rowNY[lr] = (uint32_t)((av >> 21) & 0x1FFFFF);
		P.rowNZ[lr] = (uint32_t)((av >> 42) & 0x1FFFFF);
		P.rowNT[lr] = (uint32_t)(ac & 0xFFFFFFFFu);
		P.rowNC[lr] = (uint32_t)(ac >> 32);
	}
	// K3
	{
		uint64_t bv = 0, bc = 0, bt = 0;
		const uint32_t owned_end = (P.pz1 - P.zlo) * P.NY;
		for (uint32_t r = 0; r < P.Lrows; r++) {
			if (r == owned_end) P.totals->nShared = (uint32_t)bv;
			P.rowBX[r] = (uint32_t)bv; P.rowBY[r] = (uint32_t)(bv + P.rowNX[r]);
			P.rowBZ[r] = (uint32_t)(bv + P.rowNX[r] + P.rowNY[r]);
			P.rowBC[r] = (uint32_t)bc; P.rowBT[r] = (uint32_t)bt;
			bv += (uint64_t)P.rowNX[r] + P.rowNY[r] + P.rowNZ[r]; bc += P.rowNC[r]; bt += P.rowNT[r];
		}
		if (owned_end >= P.Lrows) P.totals->nShared = (uint32_t)bv;
		P.totals->nCentre = (uint32_t)bc; P.totals->nT = (uint32_t)bt; P.totals->pad_[0] = (uint32_t)bv;
	}
	if (!emit) return;
	// K4v
	for (uint32_t lr = (P.pz0 - P.zlo) * P.NY; lr < (P.pz1 - P.zlo) * P.NY; lr++)
		for (uint32_t w = 0; w < P.W; w++)
			emit_vertices_word<Sample>(P, lr / P.NY + P.zlo, lr % P.NY, w);
	// K4t
	uint32_t scr_mask[8], scr_base[8];
	for (uint32_t lr = (P.cz0 - P.zlo) * P.NY; lr < (P.cz1 - P.zlo) * P.NY; lr++) {
		if (lr % P.NY >= P.ny) continue;
		for (uint32_t w = 0; w < P.WC; w++)
			emit_triangles_word<Sample>(P, tb, lr / P.NY + P.zlo, lr % P.NY, w, scr_mask, scr_base, 1);
	}
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
	P.W = (P.NX + 31) / 32; P.WC = (P.nx + 31) / 32; P.WP = (P.W + 1 + 3) & ~3u;
	P.Lrows = (P.zhi - P.zlo) * P.NY;
	P.geom.store = d->store; P.geom.normal_neg = d->normal_neg; P.geom.tsa = d->tsa;
	for (int i = 0; i < 3; i++) { P.geom.O[i] = d->O[i]; P.geom.D[i] = d->D[i]; }
	P.geom.ca = d->ca; P.geom.cb = d->cb;
	for (int i = 0; i < 9; i++) { P.geom.A[i] = d->A[i]; P.geom.Ai[i] = d->Ai[i]; }
	P.iso = d->dtype == MC33CU_F64 ? iso + 0.0 : (double)((float)iso + 0.0f);
	std::vector<uint32_t> S((size_t)P.Lrows * P.WP, 0), Z((size_t)P.Lrows * P.WP, 0), rn((size_t)P.Lrows * 5), rb((size_t)P.Lrows * 5);
	std::vector<uint8_t> rz(P.Lrows);
	std::vector<uint64_t> wv((size_t)P.Lrows * P.W), wc((size_t)P.Lrows * P.W);
	Totals tot;
	memset(&tot, 0, sizeof tot);
	P.S = S.data(); P.Z = Z.data(); P.rowZ = rz.data(); P.wpreV = wv.data(); P.wpreC = wc.data();
	P.rowNX = rn.data(); P.rowNY = P.rowNX + P.Lrows; P.rowNZ = P.rowNY + P.Lrows; P.rowNC = P.rowNZ + P.Lrows; P.rowNT = P.rowNC + P.Lrows;
	P.rowBX = rb.data(); P.rowBY = P.rowBX + P.Lrows; P.rowBZ = P.rowBY + P.Lrows; P.rowBC = P.rowBZ + P.Lrows; P.rowBT = P.rowBC + P.Lrows;
	P.totals = &tot;
	bool emit = o != nullptr;
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
		counts->nSharedHalo = tot.pad_[0] - tot.nShared; counts->nV = (uint64_t)tot.nShared + tot.nCentre;
	}
	return tot.overflow ? MC33CU_ERR_CAPACITY : 0;
}
