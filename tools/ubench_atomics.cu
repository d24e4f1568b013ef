// Design-time microbenchmarks for the counting stage (not part of the product or the tests).
// Measures on the B200: shared-memory atomics, global atomics on L2- vs HBM-resident tables,
// gathers, and scattered 8-byte appends, so that DESIGN.md can choose between a direct HBM
// table, L2-resident partitions and shared-memory partitions with numbers instead of guesses.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench_atomics ubench_atomics.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t x)
{
	x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
	x ^= x >> 27; x *= 0x94d049bb133111ebull;
	x ^= x >> 31;
	return x;
}

// ---------------- shared memory ----------------
template <int MODE>
__global__ void __launch_bounds__(1024) smem_kernel(int iters, int slots_log2, uint32_t *sink)
{
	extern __shared__ uint64_t sm[];
	uint32_t *cnt = (uint32_t *)sm;
	const uint32_t mask = (1u << slots_log2) - 1;
	for (uint32_t i = threadIdx.x; i <= mask; i += blockDim.x) sm[i] = 0;
	__syncthreads();
	uint64_t r = mix64(blockIdx.x * 1024 + threadIdx.x + 1);
	uint32_t acc = 0;
	for (int i = 0; i < iters; ++i) {
		r = r * 6364136223846793005ull + 1442695040888963407ull;
		uint32_t s = (uint32_t)(r >> 33) & mask;
		if (MODE == 0) atomicAdd(&cnt[s], 1u);                       // ATOMS.ADD / RED
		else if (MODE == 1) acc += atomicAdd(&cnt[s], 1u);           // ATOMS.ADD with return
		else if (MODE == 2) acc += (uint32_t)atomicCAS((unsigned long long *)&sm[s], 0ull, (unsigned long long)(r | 1)); // CAS.64
		else if (MODE == 3) { uint64_t k = sm[s]; if (k != r) atomicAdd(&cnt[2 * s + 1], 1u); } // LDS.64 + compare + ATOMS
		else if (MODE == 4) { acc += (uint32_t)sm[s]; }                // LDS.64 only
	}
	if (acc == 0x12345) sink[0] = acc;
}

// ---------------- global memory ----------------
template <int MODE>
__global__ void __launch_bounds__(256) gmem_kernel(uint64_t *tab, uint64_t slot_mask, int iters, uint32_t *sink)
{
	uint64_t r = mix64((uint64_t)blockIdx.x * blockDim.x + threadIdx.x + 1);
	uint32_t acc = 0;
	uint32_t *cnt = (uint32_t *)tab;
#pragma unroll 4
	for (int i = 0; i < iters; ++i) {
		r = r * 6364136223846793005ull + 1442695040888963407ull;
		uint64_t s = (r >> 20) & slot_mask;                           // 16-byte slots: {u64 key, u32 count, u32 pad}
		if (MODE == 0) atomicAdd(&cnt[4 * s + 2], 1u);                // RED.ADD
		else if (MODE == 1) { uint64_t k = tab[2 * s]; if (k != r) atomicAdd(&cnt[4 * s + 2], 1u); } // LD.64 + RED
		else if (MODE == 2) acc += (uint32_t)tab[2 * s];              // LD.64 gather
		else if (MODE == 3) acc += (uint32_t)atomicCAS((unsigned long long *)&tab[2 * s], 0ull, (unsigned long long)(r | 1)); // CAS.64
		else if (MODE == 4) tab[2 * s] = r;                           // scattered ST.64
	}
	if (acc == 0x12345) sink[0] = acc;
}

// scattered appends to P frontiers.  Every iteration each bucket receives ~T/P writes that land in a contiguous
// frontier window [i*T/P, (i+1)*T/P) of that bucket, like cursor-reserved appends would; `run` consecutive lanes share a
// bucket and write adjacent 8-byte cells (emulates pre-sorting a tile by bucket in shared memory before the copy-out).
__global__ void __launch_bounds__(256) append_kernel(uint64_t *out, int p_log2, uint64_t cap, int iters, int run)
{
	const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const uint64_t T = (uint64_t)gridDim.x * blockDim.x;
	const uint64_t per = (T >> p_log2) ? (T >> p_log2) : 1; // writes per bucket per iteration
	uint64_t r = mix64(tid / run + 1);
	for (int i = 0; i < iters; ++i) {
		r = r * 6364136223846793005ull + 1442695040888963407ull;
		uint64_t b = (r >> 33) & (((uint64_t)1 << p_log2) - 1);
		uint64_t grp = (mix64(r) % ((per + run - 1) / run)) * run;
		uint64_t pos = b * cap + ((uint64_t)i * per + grp + tid % run) % cap;
		out[pos] = r;
	}
}

static float time_it(void (*launch)(void *), void *arg)
{
	cudaEvent_t a, b;
	CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
	launch(arg); CK(cudaDeviceSynchronize());
	CK(cudaEventRecord(a));
	launch(arg);
	CK(cudaEventRecord(b));
	CK(cudaEventSynchronize(b));
	float ms;
	CK(cudaEventElapsedTime(&ms, a, b));
	CK(cudaGetLastError());
	return ms;
}

struct SArg { int mode, iters, slots_log2, threads, blocks; uint32_t *sink; };
static void launch_s(void *p)
{
	SArg *a = (SArg *)p;
	size_t sm = ((size_t)8 << a->slots_log2);
	switch (a->mode) {
#define C(M) case M: cudaFuncSetAttribute(smem_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); smem_kernel<M><<<a->blocks, a->threads, sm>>>(a->iters, a->slots_log2, a->sink); break;
	C(0) C(1) C(2) C(3) C(4)
#undef C
	}
}
struct GArg { int mode, iters, blocks; uint64_t *tab; uint64_t mask; uint32_t *sink; };
static void launch_g(void *p)
{
	GArg *a = (GArg *)p;
	switch (a->mode) {
#define C(M) case M: gmem_kernel<M><<<a->blocks, 256>>>(a->tab, a->mask, a->iters, a->sink); break;
	C(0) C(1) C(2) C(3) C(4)
#undef C
	}
}
struct AArg { uint64_t *out; int p_log2; uint64_t cap; int iters, run, blocks; };
static void launch_a(void *p)
{
	AArg *a = (AArg *)p;
	append_kernel<<<a->blocks, 256>>>(a->out, a->p_log2, a->cap, a->iters, a->run);
}

int main()
{
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, 0));
	printf("# %s, %d SMs\n", prop.name, prop.multiProcessorCount);
	uint32_t *sink; CK(cudaMalloc(&sink, 4));
	const char *sname[] = { "ATOMS.ADD(no ret)", "ATOMS.ADD(ret)", "ATOMS.CAS64", "LDS64+cmp+ATOMS", "LDS64" };
	for (int mode = 0; mode < 5; ++mode)
		for (int slog = 10; slog <= 14; slog += 2)
			for (int threads = 256; threads <= 1024; threads *= 2) {
				SArg a = { mode, 4096, slog, threads, prop.multiProcessorCount * (1024 / threads > 2 ? 2 : 1024 / threads), sink };
				if (((size_t)8 << slog) * (a.blocks / prop.multiProcessorCount) > 200 * 1024) continue;
				float ms = time_it(launch_s, &a);
				double ops = (double)a.blocks * threads * a.iters;
				printf("smem %-18s slots=2^%d threads=%4d blocks=%d : %8.3f ms  %8.2f Gop/s  (%.2f op/clk/SM @1.9GHz)\n", sname[mode], slog, threads, a.blocks, ms, ops / ms * 1e-6, ops / ms * 1e-6 / prop.multiProcessorCount / 1.9);
			}
	const char *gname[] = { "RED.ADD", "LD64+cmp+RED", "LD64 gather", "CAS64", "ST64 scatter" };
	size_t max_bytes = (size_t)8 << 30;
	uint64_t *tab; CK(cudaMalloc(&tab, max_bytes));
	CK(cudaMemset(tab, 0, max_bytes));
	for (int mode = 0; mode < 5; ++mode)
		for (int blog = 20; blog <= 33; blog += (blog < 26 ? 2 : (blog < 28 ? 1 : 2))) {
			if (((size_t)1 << blog) > max_bytes) continue;
			GArg a = { mode, 256, prop.multiProcessorCount * 32, tab, (((uint64_t)1 << blog) / 16) - 1, sink };
			float ms = time_it(launch_g, &a);
			double ops = (double)a.blocks * 256 * a.iters;
			printf("gmem %-14s table=2^%d B : %8.3f ms  %8.2f Gop/s\n", gname[mode], blog, ms, ops / ms * 1e-6);
		}
	for (int plog = 8; plog <= 16; plog += 2)
		for (int run = 1; run <= 16; run *= 4) {
			uint64_t cap = (max_bytes / 8) >> plog;
			AArg a = { tab, plog, cap, 256, run, prop.multiProcessorCount * 32 };
			float ms = time_it(launch_a, &a);
			double ops = (double)a.blocks * 256 * a.iters;
			printf("append P=2^%d run=%2d : %8.3f ms  %8.2f Gkeys/s  %8.1f GB/s\n", plog, run, ms, ops / ms * 1e-6, ops * 8 / ms * 1e-6);
		}
	return 0;
}
