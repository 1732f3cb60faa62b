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
#include "simt_emu.h"
#include "../../mc33_c_library_b200/csrc/mc33_pipeline.cuh"
#include "../../include/mc33cu.h"

using namespace mc33;

static uint64_t g_rt_checked = 0, g_rt_mismatch = 0;      // MC33_EMU_CHECK_RT (see run())

static const Tables &host_tables()
{
	static Tables tb;
	static uint8_t pat[MC33_NTRI_WORDS];
	static uint32_t cinfo[256];
	static uint16_t pord[MC33_NTRI_WORDS], keep[MC33_NPATTERNS * 256];
	static bool done = false;
	if (!done) {
		for (int i = 0; i < MC33_NTRI_WORDS; i++) pat[i] = (uint8_t)(MC33_PAT_NTRI[i] | (MC33_PAT_CENTRE[i] << 7));
		for (int i = 0; i < 256; i++) cinfo[i] = (uint32_t)MC33_SIMPLE256[i] | ((uint32_t)((MC33_CASE256[i] >> 11) & 1u) << 16);
		tb.case256 = MC33_CASE256; tb.simple256 = MC33_SIMPLE256; tb.tri = MC33_TRI; tb.pat = pat; tb.cinfo = cinfo;
		tb.pord = pord; tb.keep = keep;
		unsigned n = 0;
		for (int i = 0; i < MC33_NTRI_WORDS; i++) {
			pord[i] = 0;
			if (MC33_PAT_NTRI[i]) {
				pord[i] = (uint16_t)n;
				for (unsigned zm = 0; zm < 256; zm++) keep[n * 256 + zm] = (uint16_t)keep_mask_walk(tb, (unsigned)i, zm);
				n++;
			}
		}
		done = true;
	}
	return tb;
}

// The real kernel bodies (mc33_pipeline.cuh) under the fiber scheduler: K2 count (queue of complex / on-iso cells,
// decoupled look-back) and K4 cells (shared-memory records, one lane per cell / per triangle).  K1 is the plain
// comparison loop below (the TMA ring is not emulated); K3 runs the vertex tasks one after the other.
template <typename Sample>
static void run(Params &P, bool emit)
{
	typedef typename Traits<Sample>::Real Real;
	const Tables &tb = host_tables();
	const Real iso = (Real)P.iso;
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
		if (anyz) *P.anyZp = P.zepoch;
	}
	// K2
	{
		CountArgs A;
		uint32_t rpw = 32;
		A.GW = rpw / P.G ? rpw / P.G : 1;
		const uint32_t RB = P2_CNT_WARPS * A.GW * P.G;
		A.nblk = (P.Lrows + RB - 1) / RB;
		std::vector<unsigned long long> lb((size_t)A.nblk * P2_LB_WORDS, 0);
		uint32_t ticket[2] = {0, 0};
		A.lb = lb.data(); A.lb_ticket = ticket; A.tag = 7;
		A.owned_end_row = (P.pz1 - P.zlo) * P.NY;
		A.export4 = nullptr;
		A.dbg_noprefix = getenv("MC33_EMU_LB_NOPREFIX") ? 1u : 0u;
		mc33emu::launch(A.nblk, 256, P2_CNT_SMEM, [&](EmuCtx &cx) { count_body<Sample>(cx, P, tb, A); });
	}
	// The direct cell kernel (k_emit_cells in mc33_kernels.cu) is device-only; its per-cell arithmetic is not: with
	// MC33_EMU_CHECK_RT the row-table form it uses (make_rowt + cell_fast_rt) is compared with cell_fast for every
	// grid point of every row the kernel would visit, on the state the count kernel just left.
	if (getenv("MC33_EMU_CHECK_RT") && !*P.anyZp) {
		const uint32_t zend = P.pz1 > P.cz1 ? P.pz1 : P.cz1;
		const uint32_t rb = (P.cz0 - P.zlo) * P.NY, re = (zend - P.zlo) * P.NY;
		const uint32_t vb = P.vbase, vbn = P.vbase_next - P.totals->nShared;
		for (uint32_t lr = rb; lr < re; lr++) {
			const RowT rt = make_rowt(P, lr, vb, vbn);
			const uint32_t z = lr / P.NY + P.zlo, y = lr % P.NY;
			if (rt.y != y || rt.z != z) g_rt_mismatch++;
			for (uint32_t x = 0; x <= P.nx; x++) {
				uint32_t ia[12], ib[12];
				unsigned oa, ob;
				const unsigned xa = cell_fast(P, x, y, z, z == P.hz ? vbn : vb, z + 1 == P.hz ? vbn : vb, ia, oa);
				const unsigned xb = cell_fast_rt(P, rt, x, ib, ob);
				if (x >= P.nx) ob &= ~1u;               // (the kernel clears the X plane of the row's last point)
				bool same = xa == xb && oa == ob;
				for (int e = 0; e < 12; e++) same = same && ia[e] == ib[e];
				g_rt_checked++;
				if (!same) g_rt_mismatch++;
			}
		}
	}
	if (!emit) return;
	// K4
	{
		EmitArgs A;
		A.pick = 0; A.nquads = 0;
		const uint32_t zend = P.pz1 > P.cz1 ? P.pz1 : P.cz1;
		A.row_begin = (P.cz0 - P.zlo) * P.NY; A.row_end = (zend - P.zlo) * P.NY;
		emit_geometry(P.Q, A);
		const uint32_t nrows = A.row_end - A.row_begin;
		A.ngroups = (nrows + A.f.Ge - 1) / A.f.Ge;
		A.nunits = (A.ngroups + P2_EM_UNIT - 1) / P2_EM_UNIT;
		const uint32_t nblocks = 3;      // a few persistent CTAs: exercises the ticket hand-out too
		mc33emu::launch(nblocks, 256, P2_EM_WARPS * P2_EM_WARP_BYTES, [&](EmuCtx &cx) {
			emit_cells_body<Sample, true>(cx, P, tb, A, cx.smem() + cx.warp() * P2_EM_WARP_BYTES);
		});
	}
	// K3, vertices: one task per vertex left by K4 (what the product launches); MC33_EMU_VTX2=1: the task-free body
	// that rebuilds the plane masks per row group (MC33_B200_VTX=2 in the product)
	if (P.vtask) {
		for (uint32_t id = 0; id < P.totals->nShared && id < P.capV; id++) run_vertex_task<Sample>(P, id);
	} else {
		VertexArgs A;
		A.row_begin = (P.pz0 - P.zlo) * P.NY; A.row_end = (P.pz1 - P.zlo) * P.NY;
		A.Gv = vertex_group_rows(P.Q);
		A.ngroups = (A.row_end - A.row_begin + A.Gv - 1) / A.Gv;
		mc33emu::launch(2, 256, P2_VX_WARPS * P2_VX_WARP_BYTES, [&](EmuCtx &cx) {
			emit_vertices_body<Sample, true>(cx, P, A, cx.smem() + cx.warp() * P2_VX_WARP_BYTES);
		});
	}
}

extern "C" void mc33emu_rt_stats(uint64_t *checked, uint64_t *mismatch)
{
	*checked = g_rt_checked; *mismatch = g_rt_mismatch;
	g_rt_checked = 0; g_rt_mismatch = 0;
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
	std::vector<uint16_t> pcache((size_t)P.Lrows * P.WP * 32, 0xFFFF);
	P.pcache = pcache.data();
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
	P.vtask = getenv("MC33_EMU_VTX2") ? nullptr : vtask.data();
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
