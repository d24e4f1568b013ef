// Order-independent digests of a build's result, computed on the device (outside any timed region): bench.py prints
// them so that the 1/2/4/8-GPU runs can be compared with each other and with the digest of the reference's own
// graph_k_<k>_level_0.bin (the test suite holds CPU restatements of the same function).
//
//   solid set   d(x)  = mix(lo ^ mix(hi + GOLD * (count + 1)))                       summed (mod 2^64) and xor-ed
//   edges       hw(e) = sum_i mix(word_i ^ GOLD * (i + 1))   over the 2-bit sequence words of the edge (App. C.1 layout)
//               d(e)  = mix(hw ^ mix((len << 32) ^ count * C2))                      summed and xor-ed over ALL edges
// Every edge is in the graph together with its reverse-complement twin, so the multiset over all edges does not depend
// on the numbering of nodes and edges; sums and xors are additive over ranks that hold disjoint parts of the solid set.
#pragma once
#include "tagpu_key.cuh"

constexpr unsigned long long TAGPU_DIGEST_GOLD = 0x9E3779B97F4A7C15ull;
constexpr unsigned long long TAGPU_DIGEST_C2 = 0xC2B2AE3D27D4EB4Full;

TAGPU_DI void tagpu_digest_commit(unsigned long long sum, unsigned long long x, unsigned long long *out)
{
#pragma unroll
	for (int d = 16; d; d >>= 1) {
		sum += __shfl_xor_sync(0xffffffffu, sum, d);
		x ^= __shfl_xor_sync(0xffffffffu, x, d);
	}
	if ((threadIdx.x & 31) == 0) {
		atomicAdd(out, sum);
		atomicXor(out + 1, x);
	}
}

template <int W>
__global__ void k_digest_solid(const Key<W> *__restrict__ key, const uint32_t *__restrict__ cnt, unsigned long long n, unsigned long long *out)
{
	unsigned long long sum = 0, x = 0;
	for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
		const Key<W> k = key[i];
		const unsigned long long d = tagpu_mix64(k.lo ^ tagpu_mix64(KeyOps<W>::hi(k) + TAGPU_DIGEST_GOLD * ((unsigned long long)cnt[i] + 1ull)));
		sum += d;
		x ^= d;
	}
	tagpu_digest_commit(sum, x, out);
}

__global__ void k_digest_edges(const uint32_t *__restrict__ e_len, const unsigned long long *__restrict__ e_count,
			       const unsigned long long *__restrict__ e_off, const uint32_t *__restrict__ e_seq, unsigned long long n_e,
			       unsigned long long *out)
{
	unsigned long long sum = 0, x = 0, tl = 0, tc = 0;
	for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < n_e; e += (unsigned long long)gridDim.x * blockDim.x) {
		const uint32_t len = e_len[e], nw = (len + 15u) >> 4;
		const uint32_t *w = e_seq + e_off[e];
		unsigned long long hw = 0;
		for (uint32_t i = 0; i < nw; ++i) hw += tagpu_mix64((unsigned long long)w[i] ^ (TAGPU_DIGEST_GOLD * (unsigned long long)(i + 1)));
		const unsigned long long d = tagpu_mix64(hw ^ tagpu_mix64(((unsigned long long)len << 32) ^ (e_count[e] * TAGPU_DIGEST_C2)));
		sum += d;
		x ^= d;
		tl += len;
		tc += e_count[e];
	}
	tagpu_digest_commit(sum, x, out);
#pragma unroll
	for (int d = 16; d; d >>= 1) {
		tl += __shfl_xor_sync(0xffffffffu, tl, d);
		tc += __shfl_xor_sync(0xffffffffu, tc, d);
	}
	if ((threadIdx.x & 31) == 0) {
		atomicAdd(out + 2, tl);
		atomicAdd(out + 3, tc);
	}
}

// node_mask[i] = edge mask of the i-th node k-mer (tagpu_copy_graph: gathered on the device, so only n_nodes bytes travel)
__global__ void k_gather_node_masks(const uint32_t *__restrict__ node_slot, const uint32_t *__restrict__ mask32, uint64_t n_nodes, uint8_t *__restrict__ out)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_nodes) return;
	const uint32_t s = node_slot[i];
	out[i] = (uint8_t)(mask32[s >> 2] >> ((s & 3u) << 3));
}
