// Counting stage, variant "direct": every canonical (k+1)-mer instance probes one HBM-resident
// open-addressing table.  Kept only as the bring-up path and as the in-GPU cross-check of the
// partitioned counter (tagpu_count.cuh); measured at ~18 G inserts/s it is HBM-transaction bound.
//
// The table is persistent and kept all-zero between runs: every claimed slot is remembered in a list,
// and k_compact_solid both harvests (count >= ci) and re-zeroes exactly those slots, so the timed
// region never pays for a memset proportional to the table size.
#pragma once
#include "tagpu_extract.cuh"
#include "tagpu_graph.cuh"

template <int W> struct CSlot;
template <> struct __align__(16) CSlot<1> { Key<1> key; uint32_t count; uint32_t pad; };
template <> struct __align__(32) CSlot<2> { Key<2> key; uint32_t count; uint32_t pad[3]; };

template <int W>
__global__ void __launch_bounds__(TAGPU_TILE_THREADS)
k_count_direct(const uint8_t *__restrict__ seq, uint64_t n, int K, CSlot<W> *tab, uint64_t slot_mask,
	       uint32_t *__restrict__ claimed_list, unsigned long long *ctr)
{
	typedef KeyOps<W> KO;
	__shared__ uint64_t pk[TAGPU_SMEM_WORDS];
	__shared__ uint32_t inv[TAGPU_SMEM_WORDS];
	__shared__ uint32_t s_list[TAGPU_TILE_BASES];
	__shared__ uint32_t s_n, s_inst;
	__shared__ unsigned long long s_base;
	if (threadIdx.x == 0) { s_n = 0; s_inst = 0; }
	tagpu_load_tile(seq, n, (uint64_t)blockIdx.x * TAGPU_TILE_BASES, pk, inv);
	__syncthreads();
	uint32_t n_valid = threadIdx.x >= TAGPU_TILE_WORDS ? 0u : tagpu_roll_word<W>(pk, inv, threadIdx.x + TAGPU_HALO_WORDS, K, [&](const Key<W> &key, int) {
		const Key<W> stored = KO::bnot(key);
		uint64_t slot = (KO::hash(key) >> 16) & slot_mask;
		for (uint64_t probes = 0;; ++probes) {
			Key<W> cur = ktab_load<W>(&tab[slot].key);
			if (KO::eq(cur, stored)) break;
			if (KO::is_zero(cur) || ktab_maybe_torn<W>(cur)) {
				Key<W> old = ktab_cas<W>(&tab[slot].key, stored);
				if (KO::is_zero(old)) { s_list[atomicAdd(&s_n, 1u)] = (uint32_t)slot; break; }
				if (KO::eq(old, stored)) break;
			}
			if (probes > slot_mask) { atomicOr(ctr + CTR_ERROR, (unsigned long long)TAGPU_ERR_TABLE_FULL); return; }
			slot = (slot + 1) & slot_mask;
		}
		atomicAdd(&tab[slot].count, 1u);
	});
	n_valid = __reduce_add_sync(0xffffffffu, n_valid);
	if ((threadIdx.x & 31) == 0) atomicAdd(&s_inst, n_valid);
	__syncthreads();
	if (threadIdx.x == 0) {
		s_base = atomicAdd(ctr + CTR_DISTINCT, (unsigned long long)s_n);
		atomicAdd(ctr + CTR_INSTANCES, (unsigned long long)s_inst);
	}
	__syncthreads();
	for (uint32_t i = threadIdx.x; i < s_n; i += blockDim.x)
		claimed_list[s_base + i] = s_list[i];
}

// One thread per claimed slot: emit (key, count) if solid, and restore the slot to all-zero.
template <int W>
__global__ void __launch_bounds__(1024)
k_compact_solid(CSlot<W> *tab, const uint32_t *__restrict__ claimed_list, const unsigned long long *ctr_in, uint32_t ci,
		Key<W> *__restrict__ solid, uint32_t *__restrict__ solid_cnt, unsigned long long *ctr)
{
	typedef KeyOps<W> KO;
	__shared__ uint32_t s_warp[32];
	__shared__ unsigned long long s_base;
	const uint64_t n_list = ctr_in[CTR_DISTINCT];
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	Key<W> key = KO::make(0, 0);
	uint32_t cnt = 0;
	if (i < n_list) {
		CSlot<W> *s = tab + claimed_list[i];
		key = KO::bnot(s->key);
		cnt = s->count;
		s->key = KO::make(0, 0);
		s->count = 0;
	}
	const bool keep = cnt >= ci;
	const uint32_t ballot = __ballot_sync(0xffffffffu, keep);
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	if (lane == 0) s_warp[warp] = __popc(ballot);
	__syncthreads();
	if (warp == 0) {
		uint32_t x = s_warp[lane], incl = x;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
			if (lane >= (uint32_t)d) incl += t;
		}
		s_warp[lane] = incl - x;
		if (lane == 31) s_base = incl ? atomicAdd(ctr + CTR_SOLID, (unsigned long long)incl) : 0ull;
	}
	__syncthreads();
	if (keep) {
		const uint64_t o = s_base + s_warp[warp] + __popc(ballot & ((1u << lane) - 1u));
		solid[o] = key;
		solid_cnt[o] = cnt;
	}
	unsigned long long sum = keep ? cnt : 0u;
#pragma unroll
	for (int d = 16; d; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
	if (lane == 0 && sum) atomicAdd(ctr + CTR_SUM_SOLID, sum);
}
