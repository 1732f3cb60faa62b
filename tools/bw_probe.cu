// bw_probe.cu -- read-bandwidth probes on the B200 (measurement tool, not product code):
//   (a) grid-stride LDG.128 reduction, (b) cp.async.bulk ring with a trivial consumer.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bw_probe bw_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k_ldg(const uint4 *p, size_t n, unsigned *out)
{
	unsigned acc = 0;
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
	for (; i + 3 * st < n; i += 4 * st) {
		uint4 a = __ldg(p + i), b = __ldg(p + i + st), c = __ldg(p + i + 2 * st), d = __ldg(p + i + 3 * st);
		acc += a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w ^ c.x ^ c.y ^ c.z ^ c.w ^ d.x ^ d.y ^ d.z ^ d.w;
	}
	for (; i < n; i += st) { uint4 a = __ldg(p + i); acc += a.x ^ a.y ^ a.z ^ a.w; }
	if (acc == 0x12345678u) *out = acc;
}

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int STAGES, int THREADS>
__global__ void k_bulk(const char *p, size_t bytes, uint32_t chunk, unsigned *out)
{
	extern __shared__ __align__(128) unsigned char smem[];
	__shared__ uint64_t full[STAGES];
	const uint32_t nchunks = (uint32_t)(bytes / chunk);
	if (threadIdx.x == 0) {
		for (int s = 0; s < STAGES; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[s])));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	auto issue = [&](uint32_t c, int s) {
		if (threadIdx.x == 0) {
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(chunk) : "memory");
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
			             ::"r"(s32(smem + (size_t)s * chunk)), "l"(p + (size_t)c * chunk), "r"(chunk), "r"(s32(&full[s])) : "memory");
		}
	};
	uint32_t nmine = blockIdx.x < nchunks ? (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
	for (uint32_t k = 0; k < STAGES && k < nmine; k++) issue(blockIdx.x + k * gridDim.x, (int)k);
	unsigned acc = 0;
	for (uint32_t k = 0; k < nmine; k++) {
		const int s = k % STAGES;
		uint32_t ok;
		do {
			asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}"
			             : "=r"(ok) : "r"(s32(&full[s])), "r"((k / STAGES) & 1) : "memory");
		} while (!ok);
		const uint4 *q = (const uint4 *)(smem + (size_t)s * chunk);
		for (uint32_t i = threadIdx.x; i < chunk / 16; i += THREADS) { uint4 a = q[i]; acc += a.x ^ a.y ^ a.z ^ a.w; }
		__syncthreads();
		if (k + STAGES < nmine) issue(blockIdx.x + (k + STAGES) * gridDim.x, s);
	}
	if (acc == 0x12345678u) *out = acc;
}

// classify-like consumer on the bulk ring: MODE bit0 = store S words, bit1 = store zero Z words,
// bit2 = only warp 0 polls the barrier (+ one more CTA barrier), bit3 = stores as one uint4 per group
template <int STAGES, int THREADS, int MODE>
__global__ void k_cls(const char *p, size_t bytes, uint32_t chunk, uint32_t *S, uint32_t *Z, float iso, unsigned *out)
{
	extern __shared__ __align__(128) unsigned char smem[];
	__shared__ uint64_t full[STAGES];
	const uint32_t nchunks = (uint32_t)(bytes / chunk);
	const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	if (threadIdx.x == 0) {
		for (int s = 0; s < STAGES; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[s])));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	auto issue = [&](uint32_t c, int s) {
		if (threadIdx.x == 0) {
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(chunk) : "memory");
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
			             ::"r"(s32(smem + (size_t)s * chunk)), "l"(p + (size_t)c * chunk), "r"(chunk), "r"(s32(&full[s])) : "memory");
		}
	};
	uint32_t nmine = blockIdx.x < nchunks ? (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
	for (uint32_t k = 0; k < STAGES && k < nmine; k++) issue(blockIdx.x + k * gridDim.x, (int)k);
	unsigned acc = 0;
	const uint32_t nit = chunk / 512, ipw = nit / (THREADS / 32);   // 128 samples per iteration
	for (uint32_t k = 0; k < nmine; k++) {
		const int s = k % STAGES;
		const uint32_t c = blockIdx.x + k * gridDim.x;
		if (!(MODE & 4) || wid == 0) {
			uint32_t ok;
			do {
				asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}"
				             : "=r"(ok) : "r"(s32(&full[s])), "r"((k / STAGES) & 1) : "memory");
			} while (!ok);
		}
		if (MODE & 4) __syncthreads();
		const uint4 *q = (const uint4 *)(smem + (size_t)s * chunk) + (size_t)wid * ipw * 32 + lane;
		uint32_t o = c * nit + wid * ipw;                // iteration index: 4 words each
		for (uint32_t it = 0; it < ipw; it++, q += 32, o += 1) {
			const uint4 r = *q;
			const float f0 = __uint_as_float(r.x), f1 = __uint_as_float(r.y), f2 = __uint_as_float(r.z), f3 = __uint_as_float(r.w);
			uint32_t nb = (f0 > iso) | ((f1 > iso) << 1) | ((f2 > iso) << 2) | ((f3 > iso) << 3);
			const bool eq = f0 == iso || f1 == iso || f2 == iso || f3 == iso;
			uint32_t v = nb << (4 * (lane & 7));
			v |= __shfl_xor_sync(0xFFFFFFFFu, v, 4); v |= __shfl_xor_sync(0xFFFFFFFFu, v, 2); v |= __shfl_xor_sync(0xFFFFFFFFu, v, 1);
			if (MODE & 8) {
				uint4 w;
				w.x = __shfl_sync(0xFFFFFFFFu, v, 0); w.y = __shfl_sync(0xFFFFFFFFu, v, 8); w.z = __shfl_sync(0xFFFFFFFFu, v, 16); w.w = __shfl_sync(0xFFFFFFFFu, v, 24);
				if (MODE & 1) *reinterpret_cast<uint4 *>(S + 4 * o) = w;
				if ((MODE & 2) && !__any_sync(0xFFFFFFFFu, eq)) *reinterpret_cast<uint4 *>(Z + 4 * o) = make_uint4(0, 0, 0, 0);
			} else {
				if ((MODE & 1) && (lane & 7) == 0) S[4 * o + (lane >> 3)] = v;
				if ((MODE & 2) && !__any_sync(0xFFFFFFFFu, eq) && lane < 4) Z[4 * o + lane] = 0;
			}
			acc += v;
		}
		__syncthreads();
		if (k + STAGES < nmine) issue(blockIdx.x + (k + STAGES) * gridDim.x, s);
	}
	if (acc == 0x12345678u) *out = acc;
}

int main()
{
	const size_t bytes = 512ull * 512 * 512 * 4;
	char *d; unsigned *o;
	cudaMalloc(&d, bytes); cudaMalloc(&o, 4);
	cudaMemset(d, 1, bytes);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	auto time = [&](auto launch, const char *name) {
		for (int i = 0; i < 3; i++) launch();
		cudaEventRecord(e0);
		for (int i = 0; i < 10; i++) launch();
		cudaEventRecord(e1); cudaEventSynchronize(e1);
		float ms; cudaEventElapsedTime(&ms, e0, e1);
		printf("%-40s %8.1f us  %7.1f GB/s  (%s)\n", name, ms * 100, bytes / (ms * 1e-4) * 1e-9, cudaGetErrorString(cudaGetLastError()));
	};
	for (int per = 2; per <= 8; per *= 2) {
		char nm[64]; snprintf(nm, 64, "ldg128 x4, 256 thr, %d CTA/SM", per);
		time([&] { k_ldg<<<148 * per, 256>>>((const uint4 *)d, bytes / 16, o); }, nm);
	}
	{
		char nm[64];
		for (uint32_t chunk : {8192u, 16384u, 32768u}) {
			for (int per : {2, 3, 4, 6}) {
				size_t sm = (size_t)chunk * 4;
				if (sm * per > 220 * 1024) continue;
				cudaFuncSetAttribute(k_bulk<4, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
				snprintf(nm, 64, "bulk ring 4 x %u B, %d CTA/SM", chunk, per);
				time([&] { k_bulk<4, 256><<<148 * per, 256, sm>>>(d, bytes, chunk, o); }, nm);
			}
		}
	}
	{
		uint32_t *S, *Z;
		cudaMalloc(&S, bytes / 32 + 1024); cudaMalloc(&Z, bytes / 32 + 1024);
		const uint32_t chunk = 16384; const size_t sm = (size_t)chunk * 4;
#define RUN(MODE, NAME) { cudaFuncSetAttribute(k_cls<4, 256, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); \
		time([&] { k_cls<4, 256, MODE><<<148 * 3, 256, sm>>>(d, bytes, chunk, S, Z, 0.5f, o); }, NAME); }
		RUN(0, "cls consumer, no stores");
		RUN(1, "cls consumer, S 4B x4 stores");
		RUN(3, "cls consumer, S + zero Z 4B stores");
		RUN(9, "cls consumer, S uint4 store");
		RUN(11, "cls consumer, S + Z uint4 stores");
		RUN(7, "cls, S+Z 4B, single-warp poll");
		RUN(15, "cls, S+Z uint4, single-warp poll");
	}
	return 0;
}
