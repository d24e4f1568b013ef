// Coverage recount on the GPU: the reference's `build_coverage` step,
//   kmer_count_on_edges  /root/reference/src/coverage/kmer_count.c:198-240  (construct_edges_hash :137-150, index_bin_edge :68-84,
//                                                                            kmer_count_iterator :152-197, get_and_add_kmer :86-111)
//   add_cnt_to_graph     /root/reference/src/coverage/kmer_count.c:113-135
// Every 31-mer of every edge of >= 32 bases enters a table; every 31-base window of every read of >= 32 bases adds 1 to the
// entry of the window and 1 to the entry of its "rev"; an edge's count is the sum over its 31-mers of min(entry, 999), then
// the maximum of itself and its reverse-complement edge.
//
// Two properties of the reference's arithmetic are visible in the result and are reproduced exactly:
//   * a read base that is not ACGTacgt has code 4 and is OR-ed unmasked into the 64-bit k-mer register: it contributes 00
//     for itself and sets the low bit of the base BEFORE it (if that base is inside the window); the window is not skipped.
//     So the 2-bit code of window base j is (c[j] & 3) | (j < 30 && c[j + 1] == 4);
//   * "rev" is the bit-reversed register (__reverse_bit, :17-23) shifted back by two: the reversed base string with C and G
//     swapped, not the reverse complement.
// A k-mer register holds the 31 bases in bits 63..2 (first base highest), bits 1..0 clear; the table stores reg | 1.
//
// The table lives in HBM (31-mers of all edges: a few 10^7 keys): one thread-block tile of the read stream does two random
// lookups per window, the same access pattern and cost as the reference's, minus its CPU.
#pragma once
#include "tagpu_key.cuh"

constexpr int TAGPU_COV_K = 31;           // KMER_SIZE_COVERAGE
constexpr uint32_t TAGPU_COV_MAX = 999;   // MAX_KMER_COUNT

struct CovTab {
	unsigned long long *key;   // reg | 1, 0 = empty
	unsigned long long *cnt;
	unsigned long long n_slots;
};

TAGPU_DI unsigned long long cov_home(const CovTab &t, unsigned long long reg)
{
	return (unsigned long long)__umul64hi(tagpu_mix64(reg), t.n_slots);
}

// slot of reg, or n_slots if absent (insert = false); claims a slot when insert = true
TAGPU_DI unsigned long long cov_find(const CovTab &t, unsigned long long reg, bool insert)
{
	const unsigned long long want = reg | 1ull;
	unsigned long long s = cov_home(t, reg);
	for (unsigned long long probes = 0; probes < t.n_slots; ++probes) {
		unsigned long long have = __ldcg(t.key + s);
		if (have == want) return s;
		if (!have) {
			if (!insert) return t.n_slots;
			have = atomicCAS(t.key + s, 0ull, want);
			if (!have || have == want) return s;
		}
		s = s + 1 == t.n_slots ? 0ull : s + 1;
	}
	return t.n_slots;
}

// the k-mer register of the 31 bases of an edge that start at base i (edge layout: base b at bits 2 (b & 15) of word b >> 4)
TAGPU_DI unsigned long long cov_edge_reg(const uint32_t *__restrict__ seq, uint32_t i)
{
	// 62 bits starting at bit 2 i of the little-endian word string: base i + j lands at bits 2 j + 1 .. 2 j
	const uint32_t w = i >> 4, sh = (i & 15u) << 1;
	const unsigned long long lo = (unsigned long long)seq[w] | ((unsigned long long)seq[w + 1] << 32);
	const unsigned long long hi = seq[w + 2];
	unsigned long long x = sh ? (lo >> sh) | (hi << (64 - sh)) : lo;
	x &= 0x3fffffffffffffffull;
	// reverse the base order (first base to the top) and restore the bit order inside each base; two low bits stay clear
	x = __brevll(x);
	return ((x & 0xaaaaaaaaaaaaaaaaull) >> 1) | ((x & 0x5555555555555555ull) << 1);
}

// one warp per edge: all 31-mers of edges with >= 32 bases enter the table (index_bin_edge)
__global__ void __launch_bounds__(256) k_cov_index(const uint32_t *__restrict__ e_len, const unsigned long long *__restrict__ e_off,
						    const uint32_t *__restrict__ e_seq, unsigned long long n_e, CovTab t, unsigned long long *err)
{
	const unsigned long long warp = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
	const uint32_t lane = threadIdx.x & 31u;
	for (unsigned long long e = warp; e < n_e; e += n_warps) {
		const uint32_t len = e_len[e];
		if (len < (uint32_t)TAGPU_COV_K + 1u) continue;
		const uint32_t *seq = e_seq + e_off[e];
		for (uint32_t i = lane; i + TAGPU_COV_K <= len; i += 32)
			if (cov_find(t, cov_edge_reg(seq, i), true) == t.n_slots) atomicOr(err, 1ull);
	}
}

TAGPU_DI uint32_t cov_nt4(uint32_t ch)
{
	const uint32_t u = ch & 0xdfu;                      // fold lower case
	return u == 'A' ? 0u : u == 'C' ? 1u : u == 'G' ? 2u : u == 'T' ? 3u : 4u;
}

// Reads: every thread takes TAGPU_COV_SPAN consecutive window starts of the stream and rolls the register through them with the
// reference's own update (register |= c << 2; use; register <<= 2), so code 4 behaves as it does there.  A window counts
// iff it holds no '\n' and its read has >= 32 bases, i.e. the read extends past the window on at least one side.
constexpr int TAGPU_COV_SPAN = 64;
__global__ void __launch_bounds__(256) k_cov_count(const uint8_t *__restrict__ s, unsigned long long n, CovTab t)
{
	const unsigned long long p0 = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) * TAGPU_COV_SPAN;
	if (p0 >= n) return;
	unsigned long long reg = 0;
	int run = 0;                                        // bases since the last '\n' among the positions taken in so far
	// positions p0 .. p0 + SPAN + 29: position q completes the window that starts at q - 30
	const unsigned long long q_end = p0 + TAGPU_COV_SPAN + TAGPU_COV_K - 1 < n ? p0 + TAGPU_COV_SPAN + TAGPU_COV_K - 1 : n;
	bool prev_is_base = p0 > 0 && s[p0 - 1] != '\n';      // the read extends to the left of the window that starts at p0
	for (unsigned long long q = p0; q < q_end; ++q) {
		const uint32_t ch = s[q];
		if (ch == '\n') { run = 0; reg = 0; prev_is_base = false; continue; }
		// get_km_i_str for the first 30 bases of a run and the window loop afterwards do the same thing to the register
		reg |= (unsigned long long)cov_nt4(ch) << 2;
		++run;
		if (run >= TAGPU_COV_K) {
			// window [q - 30, q]; left neighbour known from the scan, right neighbour is the next byte
			const bool left = run > TAGPU_COV_K || prev_is_base;
			const bool right = q + 1 < n && s[q + 1] != '\n';
			if ((left || right) && q - (TAGPU_COV_K - 1) >= p0) {
				unsigned long long a = cov_find(t, reg, false);
				if (a != t.n_slots) atomicAdd(t.cnt + a, 1ull);
				a = cov_find(t, __brevll(reg) << 2, false);
				if (a != t.n_slots) atomicAdd(t.cnt + a, 1ull);
			}
		}
		reg <<= 2;
	}
}

// one warp per edge: count = sum over its 31-mers of min(entry, 999)   (add_cnt_to_graph, first loop)
__global__ void __launch_bounds__(256) k_cov_sum(const uint32_t *__restrict__ e_len, const unsigned long long *__restrict__ e_off,
						  const uint32_t *__restrict__ e_seq, unsigned long long n_e, CovTab t, unsigned long long *__restrict__ raw)
{
	const unsigned long long warp = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
	const uint32_t lane = threadIdx.x & 31u;
	for (unsigned long long e = warp; e < n_e; e += n_warps) {
		const uint32_t len = e_len[e];
		unsigned long long sum = 0;
		if (len >= (uint32_t)TAGPU_COV_K + 1u) {
			const uint32_t *seq = e_seq + e_off[e];
			for (uint32_t i = lane; i + TAGPU_COV_K <= len; i += 32) {
				const unsigned long long a = cov_find(t, cov_edge_reg(seq, i), false);
				if (a != t.n_slots) {
					const unsigned long long c = t.cnt[a];
					sum += c < TAGPU_COV_MAX ? c : TAGPU_COV_MAX;
				}
			}
		}
#pragma unroll
		for (int d = 16; d; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
		if (lane == 0) raw[e] = sum;
	}
}

// count = max(own, reverse-complement edge)   (add_cnt_to_graph, second loop; rc_id is an involution, so order-free)
__global__ void __launch_bounds__(256) k_cov_symmetric(const uint32_t *__restrict__ e_rc, const unsigned long long *__restrict__ raw,
							unsigned long long n_e, unsigned long long *__restrict__ out)
{
	const unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= n_e) return;
	const unsigned long long a = raw[e], b = raw[e_rc[e]];
	out[e] = a > b ? a : b;
}
