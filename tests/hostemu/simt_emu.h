// simt_emu.h -- TEST INFRASTRUCTURE ONLY: a cooperative fiber scheduler that runs the kernel
// bodies of mc33_c_library_b200/csrc/mc33_pipeline.cuh on the CPU.  Every thread of a CTA is
// a ucontext fiber; a warp collective (shuffle, ballot, reduce, syncwarp) or a CTA barrier
// parks the calling fiber until all its participants have arrived, then hands every one of
// them the value the hardware would.  CTAs run one after the other in block order, so a
// decoupled look-back never has to wait (its predecessors are complete).
//
// What it catches before a GPU run: wrong lane arithmetic, divergent collectives (lanes of a
// warp arriving at different operations -> abort), lanes that exit while others still wait
// (deadlock -> abort), out-of-range shared / global indices (under valgrind / ASan), and any
// difference from the oracle's mesh.  It does not model memory ordering or bank conflicts.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>
#include <vector>

#include "../../mc33_c_library_b200/csrc/mc33_simt.h"

namespace mc33emu {

struct Warp {
	int arrived = 0, gen = 0;
	uint64_t val[2][32];
	int kind[2][32], arg[2][32];
};

struct Block {
	unsigned nthreads = 0;
	std::vector<ucontext_t> ctx;
	std::vector<char *> stacks;
	std::vector<int> state;          // 0 runnable, 1 waits for its warp, 2 waits for the CTA, 3 done
	std::vector<int> wait_gen;
	std::vector<Warp> warps;
	int cta_arrived = 0, cta_gen = 0;
	ucontext_t sched;
	std::function<void(mc33::EmuCtx &)> body;
	std::vector<mc33::EmuCtx> cx;
	static const size_t STACK = 256 << 10;

	void yield(unsigned tid) { swapcontext(&ctx[tid], &sched); }

	uint64_t collective(unsigned tid, int kind, uint64_t v, int arg, bool cta)
	{
		if (cta) {
			const int g = cta_gen;
			if (++cta_arrived == (int)nthreads) { cta_arrived = 0; cta_gen++; }
			else { state[tid] = 2; wait_gen[tid] = g; yield(tid); }
			return 0;
		}
		Warp &W = warps[tid >> 5];
		const unsigned l = tid & 31u;
		const int g = W.gen, par = g & 1;
		W.val[par][l] = v; W.kind[par][l] = kind; W.arg[par][l] = arg;
		if (++W.arrived == 32) {
			for (int i = 1; i < 32; i++)
				if (W.kind[par][i] != W.kind[par][0]) {
					fprintf(stderr, "simt_emu: divergent collective in warp %u (lane 0 op %d, lane %d op %d)\n", tid >> 5,
					        W.kind[par][0], i, W.kind[par][i]);
					abort();
				}
			W.arrived = 0; W.gen++;
		} else {
			state[tid] = 1; wait_gen[tid] = g; yield(tid);
		}
		switch (kind) {
		case 0: return W.val[par][arg & 31];
		case 1: { uint32_t m = 0; for (int i = 0; i < 32; i++) m |= (W.val[par][i] ? 1u : 0u) << i; return m; }
		case 2: { uint64_t m = 0; for (int i = 0; i < 32; i++) m |= W.val[par][i]; return m; }
		case 3: { uint64_t m = 0; for (int i = 0; i < 32; i++) m += (uint32_t)W.val[par][i]; return (uint32_t)m; }
		case 4: { uint64_t m = 0; for (int i = 0; i < 32; i++) m = W.val[par][i] > m ? W.val[par][i] : m; return m; }
		default: return 0;
		}
	}
};

inline Block *&current() { static thread_local Block *b = nullptr; return b; }
inline unsigned &current_tid() { static thread_local unsigned t = 0; return t; }

inline void trampoline()
{
	Block *b = current();
	const unsigned tid = current_tid();
	b->body(b->cx[tid]);
	b->state[tid] = 3;
	swapcontext(&b->ctx[tid], &b->sched);
}

// run `body` for every thread of every block; smem_bytes of zero-initialised "shared memory" per block
inline void launch(unsigned nblocks, unsigned nthreads, size_t smem_bytes, std::function<void(mc33::EmuCtx &)> body)
{
	Block b;
	b.nthreads = nthreads;
	b.ctx.resize(nthreads); b.state.resize(nthreads); b.wait_gen.resize(nthreads); b.cx.resize(nthreads);
	b.warps.resize((nthreads + 31) / 32);
	b.stacks.resize(nthreads);
	for (unsigned t = 0; t < nthreads; t++) b.stacks[t] = (char *)malloc(Block::STACK);
	std::vector<unsigned char> smem(smem_bytes + 128);
	b.body = body;
	for (unsigned blk = 0; blk < nblocks; blk++) {
		memset(smem.data(), 0xCD, smem.size());          // shared memory is NOT zero on the device either
		b.cta_arrived = 0; b.cta_gen = 0;
		for (auto &w : b.warps) { w.arrived = 0; w.gen = 0; }
		for (unsigned t = 0; t < nthreads; t++) {
			mc33::EmuCtx &c = b.cx[t];
			c.lane_ = t & 31u; c.warp_ = t >> 5; c.tid_ = t; c.nthreads_ = nthreads; c.block_ = blk; c.nblocks_ = nblocks;
			c.smem_ = (unsigned char *)(((uintptr_t)smem.data() + 127) & ~(uintptr_t)127);
			c.sched_ = &b;
			b.state[t] = 0;
			getcontext(&b.ctx[t]);
			b.ctx[t].uc_stack.ss_sp = b.stacks[t];
			b.ctx[t].uc_stack.ss_size = Block::STACK;
			b.ctx[t].uc_link = &b.sched;
			makecontext(&b.ctx[t], (void (*)())trampoline, 0);
		}
		unsigned done = 0;
		while (done < nthreads) {
			bool progress = false;
			for (unsigned t = 0; t < nthreads; t++) {
				if (b.state[t] == 3) continue;
				if (b.state[t] == 1 && b.warps[t >> 5].gen == b.wait_gen[t]) continue;
				if (b.state[t] == 2 && b.cta_gen == b.wait_gen[t]) continue;
				b.state[t] = 0;
				current() = &b; current_tid() = t;
				swapcontext(&b.sched, &b.ctx[t]);
				progress = true;
				if (b.state[t] == 3) done++;
			}
			if (!progress) {
				fprintf(stderr, "simt_emu: deadlock in block %u: a lane left the kernel (or took another path) while others wait at a collective\n", blk);
				for (unsigned t = 0; t < nthreads; t++)
					if (b.state[t] != 3) fprintf(stderr, "  thread %u state %d\n", t, b.state[t]);
				abort();
			}
		}
	}
	for (unsigned t = 0; t < nthreads; t++) free(b.stacks[t]);
}

}  // namespace mc33emu

inline uint64_t mc33::EmuCtx::collective(int kind, uint64_t v, int arg, bool cta) const
{
	return ((mc33emu::Block *)sched_)->collective(tid_, kind, v, arg, cta);
}
inline void mc33::EmuCtx::backoff(unsigned &) const
{
	fprintf(stderr, "simt_emu: a spin loop cannot make progress here (blocks run in order)\n");
	abort();
}
