// De Bruijn graph stage on the GPU: solid (k+1)-mers -> k-mer table with 8-bit edge masks ->
// node marking -> successor links over oriented non-branching vertices -> Wyllie pointer jumping ->
// unitig (edge) emission, reverse-complement links and edge counts.
//
// Replaces the reference's split_kmer_from_kedge_multi (/root/reference/src/kmer_build.c:78-129),
// build_asm_graph_from_kmhash + build_graph_worker (:544-649, :421-542), the rc-link loop (:624-641)
// and build_edge_kmer_index_multi + assign_count_kedge_multi (:291-338, :143-157).
//
// Table vertex id tv = 2 * slot + orient, orient 0 = the canonical k-mer stored in `slot`, 1 = its reverse
// complement.  Low nibble of the mask = bases that may follow orient 0, high nibble = orient 1 (App. A.4).
// Non-branching ("chain") k-mers are renumbered densely (cid) so that every per-vertex array of the list-ranking
// phase is compact: chain vertex cv = 2 * cid + orient.  Node k-mers get ordinals; node vertex = 2 * ord + orient.
#pragma once
#include <cooperative_groups.h>

#include "tagpu_key.cuh"

constexpr uint32_t TAGPU_NONE = 0xffffffffu;
constexpr uint32_t TAGPU_TERM = 0x80000000u;
constexpr uint32_t TAGPU_CHAIN = 0x80000000u;   // kind[] entry: chain k-mer (low bits = cid); otherwise node ordinal

enum {
	CTR_INSTANCES = 0, CTR_DISTINCT, CTR_SOLID, CTR_KMERS, CTR_NODES, CTR_EDGES, CTR_SEQ_WORDS,
	CTR_KP1_ON_EDGE, CTR_ERROR, CTR_SUM_SOLID, CTR_SPARE0, CTR_SPARE1, CTR_CHAIN, CTR_JUMP_ROUNDS, CTR_GROUPS,
	CTR_BLOCKS, CTR_PATHS, CTR_PATH_WORDS, CTR_SPARE2, CTR_SPARE3,
	CTR_REC_LOCAL, CTR_REC_PEER,           // super-k-mer records pass 2 reads from this rank's own regions / from the other ranks' (NVLink)
	CTR_JUMP_FLAGS /* + 64 */, CTR_TOTAL = CTR_JUMP_FLAGS + 64
};

// One harvest of k_count_buckets = one contiguous block of the solid (k+1)-mer list: keys that share the minimizers of
// the buckets [b0, b0 + nbk).  flags != 0: the block does not hold ALL solid (k+1)-mers of those buckets (hash sub-class
// pass, or written past the staging area) and must not be contracted.
struct SolidBlock {
	unsigned long long base;
	uint32_t n, b0, nbk, flags;
};

enum { TAGPU_ERR_TABLE_FULL = 1, TAGPU_ERR_MISSING_SUCC = 2, TAGPU_ERR_CHAIN = 4, TAGPU_ERR_RC_LINK = 8, TAGPU_ERR_BUCKET_OVERFLOW = 16, TAGPU_ERR_RUN_LENGTH = 32, TAGPU_ERR_BLOCKS = 64, TAGPU_ERR_CONTRACT = 128 };

#define DEG4(x) __popc((x) & 15u)

TAGPU_DI uint32_t tagpu_only4(uint32_t nib) { return (uint32_t)(__ffs(nib & 15u) - 1); }
TAGPU_DI uint32_t tagpu_rank4(uint32_t nib, uint32_t c) { return (uint32_t)__popc(nib & ((1u << c) - 1u)); }

// All 32 lanes must call.  Returns this lane's base offset in a global bump allocation of `want` units.
TAGPU_DI unsigned long long tagpu_warp_alloc(unsigned long long *ctr, uint32_t want)
{
	const uint32_t lane = threadIdx.x & 31u;
	uint32_t incl = want;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
		if (lane >= (uint32_t)d) incl += t;
	}
	uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
	unsigned long long base = 0;
	if (lane == 31 && total) base = atomicAdd(ctr, (unsigned long long)total);
	base = __shfl_sync(0xffffffffu, base, 31);
	return base + (incl - want);
}

// ---------------------------------------------------------------- k-mer table (open addressing, linear probing)
// keys[] holds ~key (canonical k-mers are never all-ones, so 0 == empty); mask bytes are updated through 32-bit atomics.
// Any slot count works: the home slot is mulhi(hash, n_slots).

template <int W> struct KTab {
	Key<W> *keys;
	uint32_t *mask32;      // ceil(n_slots / 4) words
	uint32_t n_slots;
};

template <int W> TAGPU_DI uint32_t ktab_home(const KTab<W> &t, const Key<W> &key)
{
	return __umulhi((uint32_t)(KeyOps<W>::hash(key) >> 32), t.n_slots);
}
TAGPU_DI uint32_t ktab_next(uint32_t slot, uint32_t n_slots) { return slot + 1 == n_slots ? 0u : slot + 1; }

template <int W> TAGPU_DI Key<W> ktab_load(const Key<W> *p);
template <> TAGPU_DI Key<1> ktab_load<1>(const Key<1> *p) { Key<1> r; r.lo = __ldcg(&p->lo); return r; }
template <> TAGPU_DI Key<2> ktab_load<2>(const Key<2> *p)
{
	ulonglong2 v = __ldcg(reinterpret_cast<const ulonglong2 *>(p));
	Key<2> r; r.lo = v.x; r.hi = v.y; return r;
}

// returns the previous content of the slot (all-zero if we claimed it); works on global and shared memory
template <int W> TAGPU_DI Key<W> ktab_cas(Key<W> *p, const Key<W> &stored);
template <> TAGPU_DI Key<1> ktab_cas<1>(Key<1> *p, const Key<1> &stored)
{
	Key<1> r; r.lo = atomicCAS(&p->lo, 0ull, stored.lo); return r;
}
template <> TAGPU_DI Key<2> ktab_cas<2>(Key<2> *p, const Key<2> &stored)
{
	unsigned __int128 want = ((unsigned __int128)stored.hi << 64) | stored.lo;
	unsigned __int128 old = atomicCAS(reinterpret_cast<unsigned __int128 *>(p), (unsigned __int128)0, want);
	Key<2> r; r.lo = (unsigned long long)old; r.hi = (unsigned long long)(old >> 64); return r;
}

// A 16-byte load is not guaranteed single-copy atomic against a 128-bit CAS: a value with an all-zero half
// that is neither empty nor ours might be torn, so it is re-read through the CAS unit.
template <int W> TAGPU_DI bool ktab_maybe_torn(const Key<W> &v);
template <> TAGPU_DI bool ktab_maybe_torn<1>(const Key<1> &) { return false; }
template <> TAGPU_DI bool ktab_maybe_torn<2>(const Key<2> &v) { return v.lo == 0 || v.hi == 0; }
// is_zero(v) || ktab_maybe_torn(v) in one test
template <int W> TAGPU_DI bool ktab_empty_or_torn(const Key<W> &v);
template <> TAGPU_DI bool ktab_empty_or_torn<1>(const Key<1> &v) { return v.lo == 0; }
template <> TAGPU_DI bool ktab_empty_or_torn<2>(const Key<2> &v) { return v.lo == 0 || v.hi == 0; }

// insert-or-find; *claimed = true if this call created the entry
template <int W>
TAGPU_DI uint32_t ktab_insert(const KTab<W> &t, const Key<W> &key, bool *claimed, unsigned long long *err)
{
	typedef KeyOps<W> KO;
	const Key<W> stored = KO::bnot(key);
	uint32_t slot = ktab_home<W>(t, key);
	*claimed = false;
	for (uint32_t probes = 0; probes < t.n_slots; ++probes) {
		Key<W> cur = ktab_load<W>(t.keys + slot);
		if (KO::eq(cur, stored)) return slot;
		if (KO::is_zero(cur) || ktab_maybe_torn<W>(cur)) {
			Key<W> old = ktab_cas<W>(t.keys + slot, stored);
			if (KO::is_zero(old)) { *claimed = true; return slot; }
			if (KO::eq(old, stored)) return slot;
		}
		slot = ktab_next(slot, t.n_slots);
	}
	atomicOr(err, (unsigned long long)TAGPU_ERR_TABLE_FULL);
	return 0;
}

template <int W>
TAGPU_DI uint32_t ktab_find(const KTab<W> &t, const Key<W> &key)
{
	typedef KeyOps<W> KO;
	const Key<W> stored = KO::bnot(key);
	uint32_t slot = ktab_home<W>(t, key);
	for (uint32_t probes = 0; probes < t.n_slots; ++probes) {
		Key<W> cur = t.keys[slot];
		if (KO::eq(cur, stored)) return slot;
		if (KO::is_zero(cur)) return TAGPU_NONE;
		slot = ktab_next(slot, t.n_slots);
	}
	return TAGPU_NONE;
}

template <int W> TAGPU_DI uint32_t ktab_mask_of(const KTab<W> &t, uint32_t slot)
{
	return (t.mask32[slot >> 2] >> ((slot & 3u) * 8u)) & 0xffu;
}

// canonical form of an oriented k-mer y (with its reverse complement yr): slot lookup + orientation bit
template <int W>
TAGPU_DI uint32_t ktab_vertex_of(const KTab<W> &t, const Key<W> &y, const Key<W> &yr)
{
	typedef KeyOps<W> KO;
	bool fwd = KO::le(y, yr);
	uint32_t s = ktab_find<W>(t, fwd ? y : yr);
	return s == TAGPU_NONE ? TAGPU_NONE : (s * 2u + (fwd ? 0u : 1u));
}

// ---------------------------------------------------------------- B: masks (one thread per solid (k+1)-mer)
// Entries [first, end).  garbage != 0 (second launch of build_local_assembly_graph, after all solid entries): the entry
// is a (k+1)-mer of a flanking contig (add_garbage, /root/reference/src/kmer_build.c:847-888); if its out-bit is already
// set, the same (k+1)-mer is in the table (solid) and the entry is dropped (vL = vR = NONE) so it is not counted twice.
template <int W>
__global__ void __launch_bounds__(256) k_insert_kmers(const Key<W> *__restrict__ solid, uint64_t first, uint64_t n_solid, int garbage, int k,
						       KTab<W> t, uint32_t *__restrict__ vL, uint32_t *__restrict__ vR,
						       unsigned long long *ctr)
{
	typedef KeyOps<W> KO;
	const uint64_t i = first + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t n_new = 0;
	if (i < n_solid) {
		const Key<W> x = solid[i];
		const Key<W> km = KO::mask(k);
		const Key<W> k1 = KO::shr2(x), k2 = KO::band(x, km);         // kedge_get_left / kedge_get_right
		const uint32_t c1 = KO::last_base(x), c2 = 3u - KO::first_base(x, k + 1);
		const Key<W> r1 = KO::rc(k1, k), r2 = KO::rc(k2, k);
		const bool f1 = KO::le(k1, r1), f2 = KO::le(k2, r2);        // ties -> forward (kmer_build.c:110,120)
		bool claimed;
		uint32_t s1 = ktab_insert<W>(t, f1 ? k1 : r1, &claimed, ctr + CTR_ERROR);
		n_new += claimed;
		const uint32_t bit1 = (1u << (f1 ? c1 : c1 + 4u)) << ((s1 & 3u) * 8u);
		const uint32_t before = atomicOr(&t.mask32[s1 >> 2], bit1);
		if (garbage && (before & bit1)) {
			vL[i] = TAGPU_NONE;
			vR[i] = TAGPU_NONE;
		} else {
			uint32_t s2 = ktab_insert<W>(t, f2 ? k2 : r2, &claimed, ctr + CTR_ERROR);
			n_new += claimed;
			atomicOr(&t.mask32[s2 >> 2], (1u << (f2 ? c2 + 4u : c2)) << ((s2 & 3u) * 8u));
			vL[i] = s1 * 2u + (f1 ? 0u : 1u);   // oriented vertex whose out-base c1 spells this (k+1)-mer
			vR[i] = s2 * 2u + (f2 ? 1u : 0u);   // oriented vertex rc(k2) whose out-base c2 spells its reverse complement
		}
	}
	n_new = __reduce_add_sync(0xffffffffu, n_new);
	if ((threadIdx.x & 31) == 0 && n_new) atomicAdd(ctr + CTR_KMERS, (unsigned long long)n_new);
}

// ---------------------------------------------------------------- C1: nodes and chain k-mers (one thread per slot)
// Node ordinals, edge-id bases and chain ids are bump-allocated with ONE atomic per counter per 1024-thread block.
template <int W>
__global__ void __launch_bounds__(1024) k_classify(KTab<W> t, uint32_t *__restrict__ kind,
						    uint32_t *__restrict__ node_slot, uint32_t *__restrict__ node_ebase,
						    uint32_t *__restrict__ chain_slot, unsigned long long *ctr)
{
	typedef KeyOps<W> KO;
	__shared__ uint32_t s_w[3][32];
	__shared__ unsigned long long s_base[3];
	const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	const bool in = slot < t.n_slots;
	const bool occ = in && !KO::is_zero(t.keys[slot]);
	const uint32_t m = occ ? ktab_mask_of<W>(t, slot) : 0u;
	const uint32_t df = DEG4(m), dr = DEG4(m >> 4);
	const bool is_node = occ && !(df == 1 && dr == 1);             // kmer_build.c:453,561
	const bool is_chain = occ && !is_node;
	uint32_t v[3] = { is_node ? 1u : 0u, is_node ? df + dr : 0u, is_chain ? 1u : 0u }, incl[3];
#pragma unroll
	for (int c = 0; c < 3; ++c) {
		uint32_t x = v[c];
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
			if (lane >= (uint32_t)d) x += y;
		}
		incl[c] = x;
		if (lane == 31) s_w[c][warp] = x;
	}
	__syncthreads();
	if (warp < 3) {
		const uint32_t x = s_w[warp][lane];
		uint32_t y = x;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			uint32_t z = __shfl_up_sync(0xffffffffu, y, d);
			if (lane >= (uint32_t)d) y += z;
		}
		s_w[warp][lane] = y - x;
		if (lane == 31) s_base[warp] = y ? atomicAdd(ctr + (warp == 0 ? CTR_NODES : (warp == 1 ? CTR_EDGES : CTR_CHAIN)), (unsigned long long)y) : 0ull;
	}
	__syncthreads();
	const uint32_t ord = (uint32_t)s_base[0] + s_w[0][warp] + incl[0] - v[0];
	const uint32_t eb = (uint32_t)s_base[1] + s_w[1][warp] + incl[1] - v[1];
	const uint32_t cid = (uint32_t)s_base[2] + s_w[2][warp] + incl[2] - v[2];
	if (in) kind[slot] = is_node ? ord : (is_chain ? (TAGPU_CHAIN | cid) : TAGPU_NONE);
	if (is_node) {
		node_slot[ord] = slot;
		node_ebase[ord] = eb;
	}
	if (is_chain) chain_slot[cid] = slot;
}

TAGPU_DI unsigned long long tagpu_pack_jump(uint32_t ptr, uint32_t dist) { return ((unsigned long long)dist << 32) | ptr; }

// ---------------------------------------------------------------- C2: successor links (one thread per chain vertex)
// jump[cv] = (ptr, dist): "cv reaches ptr in dist hops".  The last chain vertex before a node points at itself with
// the TERM bit set; vsucc[cv] then holds the node vertex that follows it.
template <int W>
__global__ void __launch_bounds__(256) k_build_succ(KTab<W> t, int k, uint32_t n_chain_vertices, const uint32_t *__restrict__ kind,
						     const uint32_t *__restrict__ chain_slot, unsigned long long *__restrict__ jump,
						     uint32_t *__restrict__ vsucc, unsigned long long *ctr)
{
	typedef KeyOps<W> KO;
	const uint32_t cv = blockIdx.x * blockDim.x + threadIdx.x;
	if (cv >= n_chain_vertices) return;
	const uint32_t slot = chain_slot[cv >> 1], o = cv & 1u;
	const Key<W> key = KO::bnot(t.keys[slot]);
	const uint32_t m = ktab_mask_of<W>(t, slot);
	const uint32_t c = tagpu_only4(o ? (m >> 4) : m);
	const Key<W> krc = KO::rc(key, k);
	const Key<W> x = o ? krc : key, xr = o ? key : krc;
	const Key<W> y = KO::push(x, c, KO::mask(k)), yr = KO::push_front(xr, 3u - c, k);
	const uint32_t tv = ktab_vertex_of<W>(t, y, yr);
	unsigned long long j = tagpu_pack_jump(TAGPU_TERM | cv, 0);
	uint32_t succ_node = TAGPU_NONE;
	if (tv == TAGPU_NONE) {
		atomicOr(ctr + CTR_ERROR, (unsigned long long)TAGPU_ERR_MISSING_SUCC);  // kmer_build.c:475 assert
	} else {
		const uint32_t kd = kind[tv >> 1];
		if (kd & TAGPU_CHAIN) j = tagpu_pack_jump((kd & ~TAGPU_CHAIN) * 2u + (tv & 1u), 1);
		else succ_node = kd * 2u + (tv & 1u);
	}
	jump[cv] = j;
	vsucc[cv] = succ_node;
}

// ---------------------------------------------------------------- C3: in-place pointer jumping, all rounds in one cooperative launch
// (ptr, dist) travel as one 64-bit word, so a concurrent reader always sees a consistent pair.  Vertices on
// node-free cycles never terminate; the round cap leaves them unterminated and later stages skip them
// (the reference never emits such unitigs either).
__global__ void __launch_bounds__(512) k_jump_all(unsigned long long *jump, uint32_t n_vertices, unsigned long long *ctr, int max_rounds)
{
	namespace cg = cooperative_groups;
	cg::grid_group grid = cg::this_grid();
	const uint32_t stride = gridDim.x * blockDim.x;
	int round = 0;
	for (; round < max_rounds; ++round) {
		bool open = false;
		for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < n_vertices; v += stride) {
			const unsigned long long j = __ldcg(jump + v);
			const uint32_t p = (uint32_t)j;
			if (p & TAGPU_TERM) continue;
			const unsigned long long jp = __ldcg(jump + p);
			const uint32_t np = (uint32_t)jp;
			__stcg(jump + v, tagpu_pack_jump(np, (uint32_t)(j >> 32) + (uint32_t)(jp >> 32)));
			open |= !(np & TAGPU_TERM);
		}
		if (__any_sync(0xffffffffu, open) && (threadIdx.x & 31) == 0) ctr[CTR_JUMP_FLAGS + round] = 1;
		grid.sync();
		if (__ldcg(ctr + CTR_JUMP_FLAGS + round) == 0) break;
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) ctr[CTR_JUMP_ROUNDS] = (unsigned long long)round + 1;
}

// ---------------------------------------------------------------- C3': list ranking for large inputs (Helman-JaJa)
// Wyllie's pointer jumping above touches every vertex in each of ceil(log2(longest unitig)) rounds; when the per-vertex
// arrays are far larger than L2 (hundreds of millions of k-mers, unitigs of 10^4..10^5 k-mers) that is 15-20 random
// HBM accesses per vertex.  Work-efficient alternative, ~2 random accesses per vertex whatever the unitig lengths:
//   k_hj_mark     every ~64th chain vertex (by hash) and every chain head (its predecessor is a node, i.e. the successor
//                 of its reverse-complement twin is a node) becomes a SPLITTER; splitters are listed compactly
//   k_hj_walk     one thread per splitter walks its sublist to the next splitter / the chain end, leaving
//                 (splitter index, hops) in own[] of every vertex it passes -> reduced list jump2[splitter index]
//   k_jump_all    Wyllie on the reduced list (1/40 of the vertices: L2-resident)
//   k_hj_finish   every vertex takes terminal and distance from its splitter
// Vertices of node-free cycles end up unterminated exactly as with Wyllie (never visited, or their splitters never
// terminate within the round cap) and are skipped by the later stages.
TAGPU_DI bool tagpu_hj_sampled(uint32_t cv) { return ((cv * 0x9e3779b1u) >> 26) == 0u; }

__global__ void __launch_bounds__(256) k_hj_mark(const unsigned long long *__restrict__ jump, uint32_t n_cv, uint32_t *__restrict__ spl_bits,
						  uint32_t *__restrict__ spl_list, unsigned long long *__restrict__ own, unsigned long long *ctr)
{
	const uint32_t cv = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31u;
	bool spl = false;
	if (cv < n_cv) spl = tagpu_hj_sampled(cv) || ((uint32_t)jump[cv ^ 1u] & TAGPU_TERM);
	const uint32_t word = __ballot_sync(0xffffffffu, spl);
	if (lane == 0 && cv < n_cv) spl_bits[cv >> 5] = word;
	const unsigned long long base = tagpu_warp_alloc(ctr + CTR_CHAIN, spl ? 1u : 0u);   // CTR_CHAIN: number of splitters
	if (spl) {
		spl_list[base] = cv;
		own[cv] = ((unsigned long long)0xffffffffu << 32) | (uint32_t)base;             // a splitter remembers its own index
	}
}

__global__ void __launch_bounds__(256) k_hj_walk(const unsigned long long *__restrict__ jump, const uint32_t *__restrict__ spl_bits,
						  const uint32_t *__restrict__ spl_list, uint32_t n_spl, unsigned long long *__restrict__ own,
						  unsigned long long *__restrict__ jump2)
{
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n_spl) return;
	uint32_t cur = spl_list[t], hops = 0;                          // hops: distance walked, in link weights (1 per (k+1)-mer)
	for (;;) {
		const unsigned long long jc = jump[cur];
		const uint32_t p = (uint32_t)jc;
		if (p & TAGPU_TERM) { jump2[t] = tagpu_pack_jump(TAGPU_TERM | cur, hops); return; }   // chain ends at cur
		hops += (uint32_t)(jc >> 32);
		if ((spl_bits[p >> 5] >> (p & 31u)) & 1u) { jump2[t] = tagpu_pack_jump((uint32_t)own[p], hops); return; }  // next splitter's index
		own[p] = ((unsigned long long)hops << 32) | t;
		cur = p;
	}
}

__global__ void __launch_bounds__(256) k_hj_finish(unsigned long long *__restrict__ jump, uint32_t n_cv, const unsigned long long *__restrict__ own,
						    const unsigned long long *__restrict__ jump2)
{
	const uint32_t cv = blockIdx.x * blockDim.x + threadIdx.x;
	if (cv >= n_cv) return;
	const unsigned long long o = own[cv];
	unsigned long long out = tagpu_pack_jump(cv, 0);                // unterminated (node-free cycle) unless shown otherwise
	if (o != ~0ull) {
		const uint32_t t = (uint32_t)o, hops = (uint32_t)(o >> 32) == 0xffffffffu ? 0u : (uint32_t)(o >> 32);
		const unsigned long long j2 = jump2[t];
		if ((uint32_t)j2 & TAGPU_TERM) out = tagpu_pack_jump((uint32_t)j2, (uint32_t)(j2 >> 32) - hops);
	}
	jump[cv] = out;
}

// ---------------------------------------------------------------- flat graph in device memory
struct FlatGraph {
	uint32_t *e_src, *e_dst, *e_rc, *e_len;     // node-vertex ids = 2 * node ordinal + orient
	unsigned long long *e_count, *e_off;        // e_off: offset of the edge's first 32-bit sequence word
	uint32_t *e_seq;
};

// ---------------------------------------------------------------- C4: edge heads (one thread per oriented node)
template <int W>
__global__ void __launch_bounds__(128) k_edge_heads(KTab<W> t, int k, uint32_t n_nodes,
						     const uint32_t *__restrict__ kind, const uint32_t *__restrict__ node_slot,
						     const uint32_t *__restrict__ node_ebase, const unsigned long long *__restrict__ jump,
						     const uint32_t *__restrict__ vsucc, uint32_t *__restrict__ vedge,
						     FlatGraph g, unsigned long long *ctr)
{
	typedef KeyOps<W> KO;
	const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;       // oriented node id
	const bool live = u < 2u * n_nodes;
	uint32_t nib = 0, e0 = 0;
	Key<W> x = KO::make(0, 0), xr = KO::make(0, 0);
	if (live) {
		const uint32_t ord = u >> 1, o = u & 1u;
		const uint32_t slot = node_slot[ord];
		const uint32_t m = ktab_mask_of<W>(t, slot);
		nib = o ? (m >> 4) : (m & 15u);
		e0 = node_ebase[ord] + (o ? DEG4(m) : 0u);
		const Key<W> key = KO::bnot(t.keys[slot]), krc = KO::rc(key, k);
		x = o ? krc : key; xr = o ? key : krc;
	}
	uint32_t r = 0;
	for (uint32_t c = 0; c < 4; ++c) {                              // all lanes iterate: warp_alloc needs the full warp
		const bool have = live && ((nib >> c) & 1u);
		uint32_t len = 0, dst = 0, first = TAGPU_NONE;
		if (have) {
			const Key<W> y = KO::push(x, c, KO::mask(k)), yr = KO::push_front(xr, 3u - c, k);
			const uint32_t tv = ktab_vertex_of<W>(t, y, yr);
			const uint32_t kd = tv == TAGPU_NONE ? TAGPU_NONE : kind[tv >> 1];
			if (kd == TAGPU_NONE) {
				atomicOr(ctr + CTR_ERROR, (unsigned long long)TAGPU_ERR_MISSING_SUCC);
			} else if (!(kd & TAGPU_CHAIN)) {
				len = k + 1; dst = kd * 2u + (tv & 1u);
			} else {
				const uint32_t cv = (kd & ~TAGPU_CHAIN) * 2u + (tv & 1u);
				const unsigned long long j = jump[cv];
				if (!((uint32_t)j & TAGPU_TERM)) {
					atomicOr(ctr + CTR_ERROR, (unsigned long long)TAGPU_ERR_CHAIN); // a chain entered from a node must end at a node
				} else {
					const uint32_t vm = (uint32_t)j & ~TAGPU_TERM;
					len = k + 1 + (uint32_t)(j >> 32) + 1;
					dst = vsucc[vm];
					first = cv;
				}
			}
		}
		const unsigned long long off = tagpu_warp_alloc(ctr + CTR_SEQ_WORDS, have ? (len + 15u) >> 4 : 0u);
		if (have && len) {
			const uint32_t e = e0 + r;
			g.e_src[e] = u; g.e_dst[e] = dst; g.e_len[e] = len; g.e_off[e] = off; g.e_count[e] = 0;
			if (first != TAGPU_NONE) vedge[first] = e;
			// first k + 1 bases: the oriented node k-mer, then c (asm_init_edge + first asm_append_edge_char)
			uint32_t word = 0;
			for (int b = 0; b <= k; ++b) {
				const uint32_t base = b < k ? KO::base_at(x, k, b) : c;
				word |= base << ((b & 15) << 1);
				if ((b & 15) == 15 || b == k) {
					atomicOr(g.e_seq + off + (b >> 4), word);
					word = 0;
				}
			}
		}
		r += have ? 1u : 0u;
	}
}

// ---------------------------------------------------------------- C5: interior vertices write their base and learn their edge
template <int W>
__global__ void __launch_bounds__(256) k_interior(KTab<W> t, int k, uint32_t n_chain_vertices, const uint32_t *__restrict__ chain_slot,
						   const unsigned long long *__restrict__ jump, uint32_t *vedge, FlatGraph g)
{
	const uint32_t cv = blockIdx.x * blockDim.x + threadIdx.x;
	if (cv >= n_chain_vertices) return;
	const unsigned long long j = jump[cv], jt = jump[cv ^ 1u];
	if (!((uint32_t)j & TAGPU_TERM) || !((uint32_t)jt & TAGPU_TERM)) return;  // node-free cycle
	const uint32_t v1 = ((uint32_t)jt & ~TAGPU_TERM) ^ 1u;          // first interior vertex of cv's chain
	const uint32_t e = vedge[v1];
	if (e == TAGPU_NONE) return;
	const uint32_t pos = k + 1 + (uint32_t)(jt >> 32);
	const uint32_t m = ktab_mask_of<W>(t, chain_slot[cv >> 1]);
	const uint32_t c = tagpu_only4((cv & 1u) ? (m >> 4) : m);
	atomicOr(g.e_seq + g.e_off[e] + (pos >> 4), c << ((pos & 15u) << 1));
	vedge[cv] = e;
}

// ---------------------------------------------------------------- C6: reverse-complement links (one thread per edge)
template <int W>
__global__ void __launch_bounds__(256) k_rc_links(KTab<W> t, int k, uint32_t n_edges,
						   const uint32_t *__restrict__ node_slot, const uint32_t *__restrict__ node_ebase,
						   FlatGraph g, unsigned long long *ctr)
{
	const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= n_edges) return;
	const uint32_t len = g.e_len[e], pos = len - k - 1;
	const uint32_t b = (g.e_seq[g.e_off[e] + (pos >> 4)] >> ((pos & 15u) << 1)) & 3u;
	const uint32_t tv = g.e_dst[e] ^ 1u, ord = tv >> 1, o = tv & 1u;
	const uint32_t m = ktab_mask_of<W>(t, node_slot[ord]);
	const uint32_t nib = o ? (m >> 4) : (m & 15u), cb = 3u - b;
	if (!((nib >> cb) & 1u)) { atomicOr(ctr + CTR_ERROR, (unsigned long long)TAGPU_ERR_RC_LINK); return; } // kmer_build.c:640 assert
	g.e_rc[e] = node_ebase[ord] + (o ? DEG4(m) : 0u) + tagpu_rank4(nib, cb);
}

// ---------------------------------------------------------------- C7: edge counts (one thread per solid (k+1)-mer)
template <int W>
__global__ void __launch_bounds__(256) k_edge_counts(const Key<W> *__restrict__ solid, const uint32_t *__restrict__ solid_cnt,
						      uint64_t n_solid, int k, KTab<W> t, const uint32_t *__restrict__ vL,
						      const uint32_t *__restrict__ vR, const uint32_t *__restrict__ kind,
						      const uint32_t *__restrict__ node_ebase, const uint32_t *__restrict__ vedge,
						      FlatGraph g, unsigned long long *ctr)
{
	typedef KeyOps<W> KO;
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t on_edge = 0;
	if (i < n_solid) {
		const Key<W> x = solid[i];
		const uint32_t cnt = solid_cnt[i];
		// The (k+1)-mer lies on the edge through vL (via its last base) and its reverse complement on that edge's twin:
		// one lookup, then e_rc — exactly the reference's "edges[e].count += c; edges[edges[e].rc_id].count += c"
		// (/root/reference/src/kmer_build.c:154-156), self-rc edges included (they get 2c).
		const uint32_t v = vL[i], c1 = KO::last_base(x);
		if (v != TAGPU_NONE) {                                      // NONE: garbage entry that duplicates a solid (k+1)-mer
			const uint32_t slot = v >> 1, o = v & 1u, kd = kind[slot];
			uint32_t e;
			if (!(kd & TAGPU_CHAIN)) {
				const uint32_t m = ktab_mask_of<W>(t, slot), nib = o ? (m >> 4) : (m & 15u);
				e = node_ebase[kd] + (o ? DEG4(m) : 0u) + tagpu_rank4(nib, c1);
			} else {
				e = vedge[(kd & ~TAGPU_CHAIN) * 2u + o];
			}
			if (e != TAGPU_NONE) {
				atomicAdd(g.e_count + e, (unsigned long long)cnt);
				atomicAdd(g.e_count + g.e_rc[e], (unsigned long long)cnt);
				on_edge = 1;
			}
		}
	}
	on_edge = __reduce_add_sync(0xffffffffu, on_edge);
	if ((threadIdx.x & 31) == 0 && on_edge) atomicAdd(ctr + CTR_KP1_ON_EDGE, (unsigned long long)on_edge);
}

// ---------------------------------------------------------------- C7'': edge counts from ANOTHER solid set (contig-file mode)
// /root/reference/src/kmer_build.c:695-710,779-780: with a contig file among the inputs (n_files < 0) the graph is built from
// reads + contig, and the edge counts come from a second database counted on the reads alone — assign_count_kedge_multi over
// its (k+1)-mers: found on an edge -> the edge and its twin get the count, not found -> ignored.  The (k+1)-mers are not
// the ones the table was built from, so the edge is found by lookup: left k-mer in the full k-mer table (one-level graph
// stage), out-bit of the last base set, node -> edge slot by rank, chain vertex -> vedge.
template <int W>
__global__ void __launch_bounds__(256) k_edge_counts_lookup(const Key<W> *__restrict__ keys, const uint32_t *__restrict__ cnt, uint64_t n, int k,
							     KTab<W> t, const uint32_t *__restrict__ kind, const uint32_t *__restrict__ node_ebase,
							     const uint32_t *__restrict__ vedge, FlatGraph g, unsigned long long *ctr)
{
	typedef KeyOps<W> KO;
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t on_edge = 0;
	if (i < n) {
		const Key<W> x = keys[i];
		const Key<W> k1 = KO::shr2(x), r1 = KO::rc(k1, k);
		const uint32_t c1 = KO::last_base(x);
		const bool f1 = KO::le(k1, r1);
		const uint32_t slot = ktab_find<W>(t, f1 ? k1 : r1);
		if (slot != TAGPU_NONE) {
			const uint32_t o = f1 ? 0u : 1u, kd = kind[slot];
			const uint32_t m = ktab_mask_of<W>(t, slot), nib = o ? (m >> 4) : (m & 15u);
			if ((nib >> c1) & 1u) {
				const uint32_t e = !(kd & TAGPU_CHAIN) ? node_ebase[kd] + (o ? DEG4(m) : 0u) + tagpu_rank4(nib, c1)
								       : vedge[(kd & ~TAGPU_CHAIN) * 2u + o];
				if (e != TAGPU_NONE) {
					atomicAdd(g.e_count + e, (unsigned long long)cnt[i]);
					atomicAdd(g.e_count + g.e_rc[e], (unsigned long long)cnt[i]);
					on_edge = 1;
				}
			}
		}
	}
	on_edge = __reduce_add_sync(0xffffffffu, on_edge);
	if ((threadIdx.x & 31) == 0 && on_edge) atomicAdd(ctr + CTR_KP1_ON_EDGE, (unsigned long long)on_edge);
}

// ---------------------------------------------------------------- C8: assign_count_garbage (build_local_assembly_graph only)
// /root/reference/src/kmer_build.c:890-926, called with ksize + 1 (:1040-1041): every (k+1)-mer of the flanking contig
// EXCEPT ITS FIRST (the loop tests i + 1 > k + 1) that lies on an edge of the new graph lifts that edge (and its twin) to
// the contig's coverage when the edge's own coverage is lower.  One thread per contig position; writers of one edge all
// write the same value, and re-evaluating after a write changes nothing, so the result equals the sequential loop.
// Launched once per contig, in the reference's order.
template <int W>
__global__ void __launch_bounds__(256) k_garbage_counts(const uint8_t *__restrict__ contig, uint32_t len, int k, KTab<W> t,
							 const uint32_t *__restrict__ kind, const uint32_t *__restrict__ node_ebase,
							 const uint32_t *__restrict__ vedge, FlatGraph g, double old_cov)
{
	typedef KeyOps<W> KO;
	const int K = k + 1;
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;       // position of the window's last base
	if (i >= len || i + 1 <= (uint32_t)K) return;
	Key<W> fw = KO::make(0, 0);
	const Key<W> Km = KO::mask(K);
	for (int b = 0; b < K; ++b) {
		const uint32_t ch = contig[i - K + 1 + b];
		uint32_t c = (ch >> 1) & 3u;                                // A=0 C=1 G=3 T=2 ...
		c ^= c >> 1;                                                // ... A=0 C=1 G=2 T=3
		fw = KO::push(fw, c, Km);
	}
	const Key<W> rv = KO::rc(fw, K);
	const Key<W> x = KO::le(fw, rv) ? fw : rv;
	const Key<W> k1 = KO::shr2(x), r1 = KO::rc(k1, k);
	const uint32_t c1 = KO::last_base(x);
	const bool f1 = KO::le(k1, r1);
	const uint32_t slot = ktab_find<W>(t, f1 ? k1 : r1);
	if (slot == TAGPU_NONE) return;
	const uint32_t o = f1 ? 0u : 1u, kd = kind[slot];
	uint32_t e;
	if (!(kd & TAGPU_CHAIN)) {
		const uint32_t m = ktab_mask_of<W>(t, slot), nib = o ? (m >> 4) : (m & 15u);
		if (!((nib >> c1) & 1u)) return;
		e = node_ebase[kd] + (o ? DEG4(m) : 0u) + tagpu_rank4(nib, c1);
	} else {
		const uint32_t m = ktab_mask_of<W>(t, slot), nib = o ? (m >> 4) : (m & 15u);
		if (!((nib >> c1) & 1u)) return;
		e = vedge[(kd & ~TAGPU_CHAIN) * 2u + o];
	}
	if (e == TAGPU_NONE) return;
	const uint32_t rc = g.e_rc[e], new_e = min(e, rc);                 // the reference indexes a (k+1)-mer by min(e, e_rc), :219-223
	const uint32_t elen = g.e_len[new_e];
	const unsigned long long cnt = *(volatile unsigned long long *)(g.e_count + new_e);
	const double new_cov = cnt * 1.0 / (elen - (uint32_t)k);           // __get_edge_cov, n_holes = 0 (assembly_graph.h:191-192)
	if (new_cov < old_cov) {
		const unsigned long long v = (unsigned long long)old_cov * (elen - (uint32_t)K + 1u);
		*(volatile unsigned long long *)(g.e_count + new_e) = v;
		*(volatile unsigned long long *)(g.e_count + g.e_rc[new_e]) = v;
	}
}
