// mc33_simt.h -- the few SIMT primitives the count / emit kernels use, behind one small
// context type, so that the SAME kernel bodies (mc33_pipeline.cuh) compile
//   * for sm_100a, where every member is a forced-inline wrapper of the intrinsic, and
//   * for the host, where tests/hostemu runs the 32 lanes of a warp (and the warps of a CTA)
//     as cooperatively scheduled fibers and the collectives exchange values between them.
// The host form is TEST INFRASTRUCTURE (it lets the warp-level logic be checked against the
// oracle in a container without a GPU); the product only ever instantiates DevCtx.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SIMT_HD __host__ __device__ __forceinline__
#else
#define SIMT_HD inline
#endif

namespace mc33 {

#if defined(__CUDACC__)
// ---------------------------------------------------------------------------------------
// device (nvcc, both passes: the members are __device__ only)
// ---------------------------------------------------------------------------------------
struct DevCtx {
	unsigned char *smem_;
	__device__ __forceinline__ explicit DevCtx(unsigned char *smem) : smem_(smem) {}
	__device__ __forceinline__ unsigned lane() const { return threadIdx.x & 31u; }
	__device__ __forceinline__ unsigned warp() const { return threadIdx.x >> 5; }
	__device__ __forceinline__ unsigned tid() const { return threadIdx.x; }
	__device__ __forceinline__ unsigned nthreads() const { return blockDim.x; }
	__device__ __forceinline__ unsigned block() const { return blockIdx.x; }
	__device__ __forceinline__ unsigned nblocks() const { return gridDim.x; }
	__device__ __forceinline__ unsigned char *smem() const { return smem_; }
	__device__ __forceinline__ uint32_t shfl(uint32_t v, int src) const { return __shfl_sync(0xFFFFFFFFu, v, src); }
	__device__ __forceinline__ uint64_t shfl(uint64_t v, int src) const { return __shfl_sync(0xFFFFFFFFu, v, src); }
	__device__ __forceinline__ uint32_t shfl_up(uint32_t v, unsigned d) const { return __shfl_up_sync(0xFFFFFFFFu, v, d); }
	__device__ __forceinline__ uint64_t shfl_up(uint64_t v, unsigned d) const { return __shfl_up_sync(0xFFFFFFFFu, v, d); }
	__device__ __forceinline__ uint32_t shfl_down(uint32_t v, unsigned d) const { return __shfl_down_sync(0xFFFFFFFFu, v, d); }
	__device__ __forceinline__ uint32_t ballot(bool p) const { return __ballot_sync(0xFFFFFFFFu, p); }
	__device__ __forceinline__ bool any(bool p) const { return __any_sync(0xFFFFFFFFu, p) != 0; }
	__device__ __forceinline__ uint32_t reduce_or(uint32_t v) const { return __reduce_or_sync(0xFFFFFFFFu, v); }
	__device__ __forceinline__ uint32_t reduce_add(uint32_t v) const { return __reduce_add_sync(0xFFFFFFFFu, v); }
	__device__ __forceinline__ uint32_t reduce_max(uint32_t v) const { return __reduce_max_sync(0xFFFFFFFFu, v); }
	__device__ __forceinline__ uint32_t reduce_min(uint32_t v) const { return __reduce_min_sync(0xFFFFFFFFu, v); }
	__device__ __forceinline__ void syncwarp() const { __syncwarp(); }
	__device__ __forceinline__ void syncthreads() const { __syncthreads(); }
	__device__ __forceinline__ uint32_t atomic_add(uint32_t *p, uint32_t v) const { return atomicAdd(p, v); }
	__device__ __forceinline__ unsigned long long atomic_add(unsigned long long *p, unsigned long long v) const { return atomicAdd(p, v); }
	__device__ __forceinline__ uint32_t atomic_or(uint32_t *p, uint32_t v) const { return atomicOr(p, v); }
	__device__ __forceinline__ void threadfence() const { __threadfence(); }
	// publication of per-CTA aggregates for the decoupled look-back (release / acquire on 64-bit words)
	__device__ __forceinline__ void st_release(unsigned long long *p, unsigned long long v) const
	{
		asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
	}
	__device__ __forceinline__ unsigned long long ld_acquire(const unsigned long long *p) const
	{
		unsigned long long v;
		asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
		return v;
	}
	// a predecessor has not published yet: sleep, longer every time (blocks over the surface take many times longer than
	// empty ones, so a fast block can wait for a slow predecessor for a long time: ncu showed a quarter of the count
	// kernel's instructions in this loop on CT-like data with a fixed 20 ns sleep)
	__device__ __forceinline__ void backoff(unsigned &ns) const { __nanosleep(ns); if (ns < 2048u) ns <<= 1; }
};
typedef DevCtx SimtCtx;
#define SIMT_FN __device__

#else
// ---------------------------------------------------------------------------------------
// host emulation (tests/hostemu only): the scheduler lives in tests/hostemu/simt_emu.h
// ---------------------------------------------------------------------------------------
struct EmuCtx {
	unsigned lane_, warp_, tid_, nthreads_, block_, nblocks_;
	unsigned char *smem_;
	void *sched_;                       // EmuBlock* (tests/hostemu/simt_emu.h)
	unsigned lane() const { return lane_; }
	unsigned warp() const { return warp_; }
	unsigned tid() const { return tid_; }
	unsigned nthreads() const { return nthreads_; }
	unsigned block() const { return block_; }
	unsigned nblocks() const { return nblocks_; }
	unsigned char *smem() const { return smem_; }
	// collectives: implemented by the fiber scheduler
	uint64_t collective(int kind, uint64_t v, int arg, bool cta) const;
	uint32_t shfl(uint32_t v, int src) const { return (uint32_t)collective(0, v, src & 31, false); }
	uint64_t shfl(uint64_t v, int src) const { return collective(0, v, src & 31, false); }
	uint32_t shfl_up(uint32_t v, unsigned d) const { return (uint32_t)collective(0, v, lane_ >= d ? (int)(lane_ - d) : (int)lane_, false); }
	uint64_t shfl_up(uint64_t v, unsigned d) const { return collective(0, v, lane_ >= d ? (int)(lane_ - d) : (int)lane_, false); }
	uint32_t shfl_down(uint32_t v, unsigned d) const { return (uint32_t)collective(0, v, lane_ + d < 32 ? (int)(lane_ + d) : (int)lane_, false); }
	uint32_t ballot(bool p) const { return (uint32_t)collective(1, p ? 1u : 0u, 0, false); }
	bool any(bool p) const { return collective(1, p ? 1u : 0u, 0, false) != 0; }
	uint32_t reduce_or(uint32_t v) const { return (uint32_t)collective(2, v, 0, false); }
	uint32_t reduce_add(uint32_t v) const { return (uint32_t)collective(3, v, 0, false); }
	uint32_t reduce_max(uint32_t v) const { return (uint32_t)collective(4, v, 0, false); }
	uint32_t reduce_min(uint32_t v) const { return ~(uint32_t)collective(4, (uint32_t)~v, 0, false); }
	void syncwarp() const { collective(5, 0, 0, false); }
	void syncthreads() const { collective(5, 0, 0, true); }
	uint32_t atomic_add(uint32_t *p, uint32_t v) const { uint32_t o = *p; *p = o + v; return o; }
	unsigned long long atomic_add(unsigned long long *p, unsigned long long v) const { unsigned long long o = *p; *p = o + v; return o; }
	uint32_t atomic_or(uint32_t *p, uint32_t v) const { uint32_t o = *p; *p = o | v; return o; }
	void threadfence() const {}
	void st_release(unsigned long long *p, unsigned long long v) const { *p = v; }
	unsigned long long ld_acquire(const unsigned long long *p) const { return *p; }
	void backoff(unsigned &ns) const;   // a spin that cannot make progress here is a bug: the scheduler aborts
};
typedef EmuCtx SimtCtx;
#define SIMT_FN
#endif

// small helpers with a device intrinsic and a host twin
SIMT_HD uint32_t funnel_r_clamp(uint32_t lo, uint32_t hi, uint32_t sh)   // ((hi:lo) >> min(sh, 32)) low word
{
#if defined(__CUDA_ARCH__)
	return __funnelshift_rc(lo, hi, sh);
#else
	if (sh >= 32) return hi;
	return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
#endif
}

}  // namespace mc33
