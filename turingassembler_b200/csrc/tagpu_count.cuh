// Counting stage: canonical (k+1)-mer counting with ALL hash-table traffic in shared memory.
//
// Why: on this B200 a (load + compare + atomic add) against an HBM-resident table sustains ~18 Gop/s and an
// L2-resident one ~60-95 Gop/s, while the same sequence on a shared-memory table sustains ~840 Gop/s chip-wide
// (tools/ubench_atomics.cu, profiles/r1_ubench_atomics.txt).  So the key space is cut into P buckets small enough that one
// bucket's distinct keys fit a shared-memory table, and buckets are shipped through HBM in a compact form:
//
//   pass 1  k_partition   reads the ASCII stream once; every maximal run of consecutive valid windows that share
//                         one minimizer OCCURRENCE ("super-k-mer", as in KMC 2/3 — the design of the library the
//                         reference delegates this stage to) becomes ONE fixed-size record (2-bit bases + window
//                         count) appended to the region of the minimizer's bucket.  ~2.2 B per instance instead of
//                         8/16 B for a raw key, and identical genomic sites yield identical records.
//   grouping              k_pull_cursors / k_scan_blocks / k_mark_groups: consecutive buckets are packed into groups
//                         of ~2 x table-slots windows with one grid-wide prefix sum.
//   pass 2  k_count_buckets  persistent CTAs, one group at a time: records staged in shared memory in canonical
//                         orientation, duplicate records collapsed, live windows cut into equal per-thread segments,
//                         canonical keys counted in a shared-memory open-addressing table (LDS + ATOMS only), keys
//                         with count >= ci emitted.  A group with too many distinct keys is re-run on hash
//                         sub-classes.  With several GPUs the records of a bucket are read from every rank's regions
//                         through NVLink peer loads (CountPeers).
//
// The bucket of a window is a function of its canonical key only (minimum over the hashes of the canonical m-mers it
// contains), so all instances of a key — on either strand — meet in the same bucket and counts are exact.
#pragma once
#include "tagpu_extract.cuh"
#include "tagpu_graph.cuh"

constexpr int TAGPU_MINIMIZER_M = 15;                 // m-mer length (30 bits): long enough that one m-mer value ~ one genomic site
constexpr uint32_t TAGPU_H_INVALID = 0xffffffffu;
#define HIDX(q) ((q) + ((q) >> 5))

template <int W> struct SkRec;                        // super-k-mer record: bases right-aligned, length in the top byte
template <> struct __align__(16) SkRec<1> { unsigned long long w[2]; };   // <= 60 bases
template <> struct __align__(32) SkRec<2> { unsigned long long w[4]; };   // <= 124 bases
template <int W> struct SkCap { static constexpr int bases = W == 1 ? 60 : 124; };

constexpr int TAGPU_MAX_RANKS = 8;

struct PartCfg {
	int K;
	int log2_buckets;
	uint32_t cap_records;         // records per bucket region
	uint32_t overflow_cap;        // records in the overflow area
	uint32_t world;               // GPUs sharing the key space (1 = single GPU)
	uint32_t per_rank;            // buckets owned by each rank: owner(b) = b / per_rank
	uint32_t packed;              // the read stream is in the packed tile layout (tagpu_extract.cuh), not ASCII
};

// Where pass 2 finds the records of a bucket: every rank ("source") partitions ITS slice of the reads into its own
// regions, for all buckets; the rank that owns a bucket then reads that bucket's records from every source.  With
// world > 1 the entries of the other ranks are CUDA-IPC mappings of THEIR allocations, so those reads are NVLink peer
// loads issued by the counting kernel itself: the hash-partitioned exchange of SURVEY.md §8e is fused into pass 2 and
// overlaps its shared-memory counting (records are prefetched one batch ahead).
template <int W> struct CountPeers {
	const SkRec<W> *regions[TAGPU_MAX_RANKS];          // [n_buckets x cap_records]
	const unsigned long long *cursor[TAGPU_MAX_RANKS]; // [n_buckets] low word records, high word windows
	const SkRec<W> *ext[TAGPU_MAX_RANKS];              // overflow records sorted by bucket
	const uint32_t *ext_off[TAGPU_MAX_RANKS];          // [n_buckets + 1] (valid when the source overflowed)
};

TAGPU_DI uint32_t tagpu_bucket_of(uint32_t minhash, int log2_buckets)
{
	return (minhash * 0x85ebca6bu) >> (32 - log2_buckets);
}

// The n_bases bases that end at packed-tile position end_q (inclusive), right-aligned, as a record with n_windows in its
// top byte.  pk holds 32 bases per 64-bit word, first base most significant; in 32-bit half-words taken in position order
// (high half first) the base at position p sits at bit 2 (15 - (p & 15)) of half-word p >> 4, so 32-bit word i of the record
// is one funnel shift of two neighbouring half-words.  Needs W + 1 words of history: end_q >= 32 (W + 1).
template <int W>
TAGPU_DI SkRec<W> tagpu_make_record(const uint64_t *pk, int end_q, int n_bases, int n_windows)
{
	constexpr int NOUT = 2 * W + 2;                      // 32-bit words that can hold bases: 49 / 95 of them at most
	const int wi = end_q >> 5;
	uint32_t H[2 * W + 4];
#pragma unroll
	for (int m = 0; m < W + 2; ++m) {
		const uint64_t v = pk[wi - (W + 1) + m];
		H[2 * m] = (uint32_t)(v >> 32);
		H[2 * m + 1] = (uint32_t)v;
	}
	const bool odd = (end_q & 16) != 0;                  // end_q lies in the low half of pk[wi]
	uint32_t t[2 * W + 3];
#pragma unroll
	for (int m = 0; m < 2 * W + 3; ++m) t[m] = odd ? H[m + 1] : H[m];
	const int sft = 2 * (15 - (end_q & 15)), total_bits = 2 * n_bases;
	uint32_t out[NOUT];
#pragma unroll
	for (int i = 0; i < NOUT; ++i) {
		const int keep = total_bits - 32 * i;              // bits of word i that belong to the record
		const uint32_t v = __funnelshift_r(t[2 * W + 2 - i], t[2 * W + 1 - i], sft);
		out[i] = keep >= 32 ? v : (keep > 0 ? v & ((1u << keep) - 1u) : 0u);
	}
	SkRec<W> r;
#pragma unroll
	for (int j = 0; j < 2 * W; ++j) r.w[j] = j < W + 1 ? (((unsigned long long)out[2 * j + 1] << 32) | out[2 * j]) : 0ull;
	r.w[2 * W - 1] |= (unsigned long long)n_windows << 56;
	return r;
}

// Minimizer (hash, position) of the window ending at tile position q (q >= w - 1).  B = the largest power of two <= w:
// phase A leaves prefix minima (hp) and suffix minima (hs) inside B-aligned blocks of positions, so the minimum over any B
// consecutive positions is one suffix entry and one prefix entry, and a window of w = B + d positions is the union of the
// two B-ranges at its ends (min is idempotent, the overlap does not matter; the position bits in the low end of a hash keep
// the LEFTMOST of equal hashes the minimum, whichever range it is found in).
template <int B>
TAGPU_DI uint32_t tagpu_window_min(const uint32_t *hs, const uint32_t *hp, int q, int w)
{
	uint32_t m = min(hs[HIDX(q - B + 1)], hp[HIDX(q)]);
	if (w != B) m = min(m, min(hs[HIDX(q - w + 1)], hp[HIDX(q - (w - B))]));
	return m;
}

// ---------------------------------------------------------------- pass 1
// Super-k-mer = maximal run of consecutive valid windows that share the SAME minimizer occurrence (hash and position),
// so a run never exceeds w = K - m + 1 windows and — unlike a cut at thread or tile boundaries — it is a function of the
// sequence alone: error-free reads that cover a genomic minimizer site produce bit-identical records, which pass 2
// collapses before counting (tagpu_count.cuh: "duplicate records").  A run is emitted by the tile that contains its END;
// the 96-base left halo lets it reach back to its start, the right halo word tells whether the run ends at the tile's
// last position.
template <int W, int TW, int B, bool EXACT>                    // EXACT: w == B is known at compile time (the kernel of the default k0 = 45 holds no code for w > B)
__global__ void __launch_bounds__(TileCfg<TW>::THREADS, TileCfg<TW>::MIN_CTAS)
k_partition(const uint8_t *__restrict__ seq, uint64_t n, uint32_t tile0, PartCfg cfg, SkRec<W> *__restrict__ regions,
	    unsigned long long *__restrict__ cursor, SkRec<W> *__restrict__ overflow, uint32_t *__restrict__ overflow_bucket,
	    unsigned long long *ctr)
{
	typedef TileCfg<TW> T;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	uint64_t *pk = reinterpret_cast<uint64_t *>(smem_raw);
	uint32_t *inv = reinterpret_cast<uint32_t *>(pk + T::SMEM_WORDS);
	uint32_t *vw = inv + T::SMEM_WORDS;              // per word: bit i = position i ends a valid window
	uint32_t *bw = vw + T::SMEM_WORDS;               // per word: bit i = position i starts a run
	uint32_t *hp = bw + T::SMEM_WORDS;               // m-mer (hash, position) per position, then block-wise prefix minima (in place)
	uint32_t *hs = hp + T::HM_LEN;                   // block-wise suffix minima
	const int K = cfg.K, m = TAGPU_MINIMIZER_M;

	if (cfg.packed) tagpu_load_tile_packed<TW>(seq, n, ((uint64_t)blockIdx.x + tile0) * T::BASES, pk, inv);
	else tagpu_load_tile<TW>(seq, n, ((uint64_t)blockIdx.x + tile0) * T::BASES, pk, inv);
	__syncthreads();

	// A. hash of the canonical m-mer ending at every position, with the position (mod 64) in the low bits: the minimum
	//    over a window then identifies one m-mer OCCURRENCE.  An m-mer that touches a non-ACGT byte gets a hash like any
	//    other: it can only be the minimum of a window that contains that byte, i.e. of an invalid window, and those are
	//    masked out in C1 — so no validity is tracked here.
	const int w = K - m + 1;
	for (int j = threadIdx.x; j < T::SMEM_WORDS; j += blockDim.x) {
		// No rolling state: the m-mer ending at position i is a 30-bit field of the packed stream (one funnel shift of two
		// 16-base half-words), and its reverse complement is the mirrored field of the reverse-complemented half-words.
		// Half-words in stream order: p_lo (positions -16..-1), c_hi (0..15), c_lo (16..31); reversed: r0 = rc(c_lo),
		// r1 = rc(c_hi), r2 = rc(p_lo), where the m-mer ending at i ends at reversed index 45 - i.
		const uint32_t mm = (1u << (2 * m)) - 1;
		const uint64_t cur = pk[j];
		const uint32_t p_lo = j ? (uint32_t)pk[j - 1] : 0u, c_hi = (uint32_t)(cur >> 32), c_lo = (uint32_t)cur;
		const uint32_t r0 = tagpu_rc32_full(c_lo), r1 = tagpu_rc32_full(c_hi), r2 = tagpu_rc32_full(p_lo);
		const uint32_t tag0 = (uint32_t)(j & 1) * 32u;
		auto mmer_hash = [&](int i) -> uint32_t {                  // i is a compile-time constant after unrolling
			const uint32_t fw = (i < 16 ? __funnelshift_r(c_hi, p_lo, 30 - 2 * i) : __funnelshift_r(c_lo, c_hi, 30 - 2 * (i - 16))) & mm;
			const int e = 45 - i;
			const uint32_t rv = (e >= 32 ? __funnelshift_r(r2, r1, 30 - 2 * (e - 32))
					     : e >= 16 ? __funnelshift_r(r1, r0, 30 - 2 * (e - 16)) : r0 >> (30 - 2 * e)) & mm;
			return ((min(fw, rv) * 0x9e3779b1u) & ~63u) | (tag0 + (uint32_t)i);
		};
		// hashes stay in registers; prefix minima inside the B-aligned blocks of the word go to hp, suffix minima to hs
		uint32_t h[32];
		uint32_t acc = TAGPU_H_INVALID;
#pragma unroll
		for (int i = 0; i < 32; ++i) {
			h[i] = mmer_hash(i);
			acc = i % B == 0 ? h[i] : min(acc, h[i]);
			hp[j * 33 + i] = acc;
		}
#pragma unroll
		for (int i = 31; i >= 0; --i) {
			acc = i % B == B - 1 ? h[i] : min(acc, h[i]);
			hs[j * 33 + i] = acc;
		}
	}
	__syncthreads();

	// C1. per word: which positions end a valid window (vw) and which of those start a new run (bw): the window before
	//     is invalid or has another minimizer occurrence.  With w > 32 a run could outgrow a record, so word starts cut too.
	//     The valid mask is bit-parallel: a window is valid iff no invalid base lies in the K positions it covers, i.e.
	//     the invalid masks of this word and the two before it, OR-smeared over K positions (log steps).
	uint32_t n_win = 0;
	for (int j = threadIdx.x + 1; j < T::SMEM_WORDS; j += blockDim.x) {
		uint32_t a = j >= 2 ? inv[j - 2] : 0xffffffffu, b = inv[j - 1], c = inv[j];     // position order a:b:c, first position = MSB
		auto smear = [&](int sft) {                                  // x |= x >> sft over the 96-bit sequence, 0 < sft < 32
			const uint32_t nc = __funnelshift_r(c, b, sft), nb_ = __funnelshift_r(b, a, sft), na = a >> sft;
			c |= nc; b |= nb_; a |= na;
		};
		smear(1); smear(2); smear(4); smear(8);                      // 16 positions
		int extra = K - 16;
		if (K >= 32) { smear(16); extra = K - 32; }                  // 32 positions
		if (extra == 32) { c |= b; b |= a; }
		else if (extra > 0) smear(extra);
		const uint32_t vmask = __brev(~c);                          // bit i = position i of this word ends a valid window
		const uint32_t pv = ~b & 1u;                                // ... and so does the last position of the word before
		// minimizer of the window ending at position i of this word (i = -1: the last position of the word before), see
		// tagpu_window_min: every index is the thread's base plus a compile-time constant, or — for the second B-range of a
		// window with w > B — plus a constant behind a base shifted by d = w - B resp. w - 1 positions (the padding word
		// between two words of 32 positions is stepped over where the shifted position falls into an earlier word)
		uint32_t ne = 0;
		const uint32_t *hpj = hp + j * 33, *hsj = hs + j * 33;
		const int d = w - B;
		auto rel = [](int p) { return p >= 0 ? p : (p >= -32 ? p - 1 : p - 2); };   // position relative to the word -> padded index
		auto min_b = [&](int i) -> uint32_t { return min(hsj[rel(i - B + 1)], hpj[rel(i)]); };   // (i: compile-time constant after unrolling)
		auto min_w = [&](int i) -> uint32_t {
			const int p1 = i - d, p2 = i - w + 1;
			int i2 = j * 33 + p2 - (p2 < 0 ? 1 : 0) - (p2 < -32 ? 1 : 0);
			if (B == 32) i2 = max(i2, 0);                           // (w > 33 in the first halo word: no window of the tile reaches there)
			return min(min_b(i), min(hs[i2], hpj[p1 - (p1 < 0 ? 1 : 0)]));
		};
		if (EXACT || d == 0) {                                      // w = B (the default k0 = 45: w = 32): one B-range is the window
			uint32_t pm = min_b(-1);
#pragma unroll
			for (int i = 0; i < 32; ++i) {
				const uint32_t cm = min_b(i);
				ne |= (cm != pm ? 1u : 0u) << i;
				pm = cm;
			}
		} else {
			uint32_t pm = min_w(-1);
#pragma unroll
			for (int i = 0; i < 32; ++i) {
				const uint32_t cm = min_w(i);
				ne |= (cm != pm ? 1u : 0u) << i;
				pm = cm;
			}
		}
		const uint32_t bmask = vmask & (~((vmask << 1) | pv) | ne | (w > 32 ? 1u : 0u));
		vw[j] = vmask;
		bw[j] = bmask;
		if (j >= TAGPU_HALO_WORDS && j < TAGPU_HALO_WORDS + T::WORDS) n_win += __popc(vmask);
	}
	__syncthreads();

	// C2. the runs that END in the tile are laid out as one dense list (block-wide prefix sum over the per-word end masks),
	//     and then every thread builds and appends one record per iteration — whatever the distribution of run ends over
	//     the words: one record (2-bit bases + window count) goes to the bucket of the run's minimizer.
	__shared__ uint16_t s_end[T::END_CAP];
	__shared__ uint32_t s_wsum[T::THREADS / 32 + 1], s_winsum[T::THREADS / 32 + 1];
	const int wi0 = threadIdx.x + TAGPU_HALO_WORDS;
	uint32_t ends = 0;
	if (threadIdx.x < T::WORDS) {
		const uint32_t V = vw[wi0], S = bw[wi0];
		const uint32_t Vn = (V >> 1) | (vw[wi0 + 1] << 31), Sn = (S >> 1) | (bw[wi0 + 1] << 31);
		ends = V & (~Vn | Sn);
	}
	const uint32_t n_mine = __popc(ends), lane_ = threadIdx.x & 31u, warp_ = threadIdx.x >> 5;
	uint32_t incl = n_mine;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
		if (lane_ >= (uint32_t)d) incl += t;
	}
	if (lane_ == 31) s_wsum[warp_] = incl;
	// (the instance total — the metric's numerator — rides on the same barrier: per-warp sums, one global atomic per CTA)
	n_win = __reduce_add_sync(0xffffffffu, n_win);
	if (lane_ == 0) s_winsum[warp_] = n_win;
	__syncthreads();
	uint32_t before = 0, n_ends = 0, n_inst = 0;
#pragma unroll
	for (int x = 0; x < T::THREADS / 32; ++x) {
		const uint32_t v = s_wsum[x];
		before += (uint32_t)x < warp_ ? v : 0u;
		n_ends += v;
		n_inst += s_winsum[x];
	}
	if (threadIdx.x == 0 && n_inst) atomicAdd(ctr + CTR_INSTANCES, (unsigned long long)n_inst);
	const uint32_t my_first = before + incl - n_mine;
	for (uint32_t pass0 = 0; pass0 < n_ends; pass0 += T::END_CAP) {          // (one pass unless the tile is pathological)
		if (pass0) __syncthreads();
		uint32_t rank = my_first, rest = ends;
		while (rest) {
			const int e = __ffs(rest) - 1;
			rest &= rest - 1;
			if (rank >= pass0 && rank < pass0 + T::END_CAP) s_end[rank - pass0] = (uint16_t)(wi0 * 32 + e);
			++rank;
		}
		__syncthreads();
		const uint32_t n_pass = min(n_ends - pass0, (uint32_t)T::END_CAP);
		for (uint32_t r = threadIdx.x; r < n_pass; r += blockDim.x) {
			const int end_q = s_end[r], wi = end_q >> 5, e = end_q & 31;
			const uint32_t upto = bw[wi] & (0xffffffffu >> (31 - e));
			const int st = upto ? 31 - __clz(upto) : -1 - __clz(bw[wi - 1]);
			const int nw = e - st + 1;
			if (nw > 32) { atomicOr(ctr + CTR_ERROR, (unsigned long long)TAGPU_ERR_RUN_LENGTH); continue; }
			const uint32_t b = tagpu_bucket_of(tagpu_window_min<B>(hs, hp, end_q, EXACT ? B : w) >> 6, cfg.log2_buckets);
			const SkRec<W> rec = tagpu_make_record<W>(pk, end_q, nw + K - 1, nw);
			const unsigned long long old = atomicAdd(cursor + b, 1ull | ((unsigned long long)nw << 32));
			const uint32_t idx = (uint32_t)old;
			if (idx < cfg.cap_records) {
				regions[(size_t)b * cfg.cap_records + idx] = rec;
			} else {
				const unsigned long long o = atomicAdd(ctr + CTR_SPARE0, 1ull);
				if (o < cfg.overflow_cap) { overflow[o] = rec; overflow_bucket[o] = b; }
				else atomicOr(ctr + CTR_ERROR, (unsigned long long)TAGPU_ERR_BUCKET_OVERFLOW);
			}
		}
	}
}

// ---------------------------------------------------------------- overflow handling (only launched when a bucket region filled up)
__global__ void k_overflow_hist(const uint32_t *__restrict__ overflow_bucket, uint64_t n_over, uint32_t *__restrict__ ext_count)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n_over) atomicAdd(ext_count + overflow_bucket[i], 1u);
}

// Exclusive scan of ext_count into ext_off (n_buckets + 1 entries), grid-wide in three small steps: per-block scan
// (k_overflow_scan_blocks), scan of the block totals (k_overflow_scan_tops, one block), and the add-back that also resets
// ext_count to 0 for its reuse as the scatter cursor (k_overflow_scan_finish).
__global__ void __launch_bounds__(1024) k_overflow_scan_blocks(const uint32_t *__restrict__ ext_count, uint32_t *__restrict__ ext_off,
							       uint32_t *__restrict__ tops, uint32_t n_buckets)
{
	__shared__ uint32_t s_w[32];
	const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	const uint32_t v = b < n_buckets ? ext_count[b] : 0u;
	uint32_t incl = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
		if (lane >= (uint32_t)d) incl += t;
	}
	if (lane == 31) s_w[warp] = incl;
	__syncthreads();
	if (warp == 0) {
		uint32_t x = s_w[lane], y = x;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			uint32_t t = __shfl_up_sync(0xffffffffu, y, d);
			if (lane >= (uint32_t)d) y += t;
		}
		s_w[lane] = y - x;
		if (lane == 31) tops[blockIdx.x] = y;
	}
	__syncthreads();
	if (b < n_buckets) ext_off[b] = s_w[warp] + incl - v;
}

__global__ void __launch_bounds__(1024) k_overflow_scan_tops(uint32_t *__restrict__ tops, uint32_t n_blocks, uint32_t *__restrict__ ext_off, uint32_t n_buckets)
{
	__shared__ uint32_t s_part[1024];
	const uint32_t per = (n_blocks + 1023) / 1024, lo = min(threadIdx.x * per, n_blocks), hi = min(lo + per, n_blocks);
	uint32_t sum = 0;
	for (uint32_t i = lo; i < hi; ++i) sum += tops[i];
	s_part[threadIdx.x] = sum;
	__syncthreads();
	if (threadIdx.x == 0) {
		uint32_t acc = 0;
		for (int t = 0; t < 1024; ++t) { uint32_t v = s_part[t]; s_part[t] = acc; acc += v; }
		ext_off[n_buckets] = acc;
	}
	__syncthreads();
	uint32_t acc = s_part[threadIdx.x];
	for (uint32_t i = lo; i < hi; ++i) { const uint32_t v = tops[i]; tops[i] = acc; acc += v; }
}

__global__ void __launch_bounds__(1024) k_overflow_scan_finish(uint32_t *__restrict__ ext_count, uint32_t *__restrict__ ext_off,
							       const uint32_t *__restrict__ tops, uint32_t n_buckets)
{
	const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= n_buckets) return;
	ext_off[b] += tops[blockIdx.x];
	ext_count[b] = 0;
}

template <int W>
__global__ void k_overflow_scatter(const SkRec<W> *__restrict__ overflow, const uint32_t *__restrict__ overflow_bucket,
				   uint64_t n_over, const uint32_t *__restrict__ ext_off, uint32_t *__restrict__ ext_cursor,
				   SkRec<W> *__restrict__ ext)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_over) return;
	const uint32_t b = overflow_bucket[i];
	ext[ext_off[b] + atomicAdd(ext_cursor + b, 1u)] = overflow[i];
}

// ---------------------------------------------------------------- pass 2
template <int W> struct BucketCfg {
#ifndef TAGPU_BC_THREADS
#define TAGPU_BC_THREADS 512
#define TAGPU_BC_CTAS 2
#define TAGPU_BC_SLOTS1 6752
#define TAGPU_BC_SLOTS2 3840
#endif
	static constexpr int THREADS = TAGPU_BC_THREADS;            // two CTAs per SM: one CTA's barriers / harvest overlap the other's inserts
	static constexpr int CTAS_PER_SM = TAGPU_BC_CTAS;
	static constexpr int NW = W + 1;                            // 64-bit words of a staged record (bases only: <= 49 / <= 95 of them)
	// records staged per round (<= 2 * THREADS: a thread stages up to two).  A staged record is kept in BOTH orientations
	// (see "orientation of a window" below), so the round is smaller than 2 * THREADS; a group of GROUP_TARGET windows
	// holds ~550 records of 151 bp reads and still fits one round.
#ifndef TAGPU_BC_DUAL
#define TAGPU_BC_DUAL 1
#endif
	static constexpr bool DUAL = W == 2 && TAGPU_BC_DUAL;                        // 64-bit keys: one rc64 per window is cheaper than a second staged copy
#ifndef TAGPU_BC_ROUND2
#define TAGPU_BC_ROUND2 608
#endif
	static constexpr int ROUND = DUAL ? TAGPU_BC_ROUND2 : 2 * THREADS;
	static_assert(ROUND <= 2 * THREADS && ROUND <= 1024, "a thread stages at most two records of a round; item words hold 10 index bits");
	static constexpr int ITEM_WINDOWS = 8;                      // windows of one work item: one per lane of an octet
	static constexpr int ITEMS = ROUND * 32 / ITEM_WINDOWS;     // work items of a round (a record has <= 32 windows)
	// shared-memory table slots per CTA (any number: the home slot is mulhi(hash, SLOTS)); sized so that two CTAs of
	// table + staging + static + the 1 KB the hardware reserves per CTA fit the SM's 228 KB
	static constexpr int SLOTS = W == 1 ? TAGPU_BC_SLOTS1 : TAGPU_BC_SLOTS2;
	static constexpr int MAX_PROBES = 48;                       // a longer probe sequence aborts the attempt (re-run on sub-classes)
	static constexpr int GROUP_MAX = 64;                        // buckets per group
	static constexpr int SUB_MAX = 256;                         // (bucket, source rank) pairs per group: group_max = min(GROUP_MAX, SUB_MAX / world)
	// a group is closed once it holds this many windows: ~0.3-0.4 load if a sixth of the windows are distinct keys.
	// Measured on C1/C2 with 1.5x / 2x / 2.5x / 3x SLOTS: 2x is best for 128-bit keys (3x overflows into re-runs), 2.5x-3x
	// for 64-bit keys; duplicate-record collapse made inserts cheap relative to the per-group harvest.
#ifndef TAGPU_GT2
#define TAGPU_GT2 8           /* group target of the 128-bit path in quarters of the slot count (developer sweeps) */
#endif
#ifndef TAGPU_GT1
#define TAGPU_GT1 10          /* the same for the 64-bit path */
#endif
	static constexpr uint32_t GROUP_TARGET = W == 1 ? SLOTS * TAGPU_GT1 / 4 : SLOTS * TAGPU_GT2 / 4;
	// staging area: record words, one 32-bit meta word per record, 16-bit work items; the harvest reuses it for its output
	// CTA-wide duplicate table of a round (the representatives of the warps meet in it): pays with 64-bit keys, whose groups
	// hold 2.5 x as many windows (C1: 2.76 -> 2.58 ms), not with 128-bit keys (C2: 3.39 -> 3.41 ms)
#ifndef TAGPU_BC_DD1
#define TAGPU_BC_DD1 1024
#define TAGPU_BC_DD2 0
#endif
	static constexpr int DD_SLOTS = W == 1 ? TAGPU_BC_DD1 : TAGPU_BC_DD2;
	static constexpr size_t STAGE_BYTES = (size_t)ROUND * ((DUAL ? 2 : 1) * NW * 8 + 4) + (size_t)ITEMS * 2 + 16 + (size_t)DD_SLOTS * 4;   // (+ the overflow flag)
	static constexpr size_t SMEM = SLOTS * (sizeof(Key<W>) + 4) + STAGE_BYTES;
};

// table hash: the key words are folded to 32 bits (one 32-bit multiply per extra word) and mixed by one more multiply;
// the top bits pick the slot, the next ones the sub-class
template <int W> TAGPU_DI uint32_t tagpu_table_hash(const Key<W> &k);
template <> TAGPU_DI uint32_t tagpu_table_hash<1>(const Key<1> &k)
{
	const uint32_t x = (uint32_t)k.lo ^ ((uint32_t)(k.lo >> 32) * 0x85ebca6bu);
	const uint32_t h = x * 0x9e3779b1u;
	return h ^ (h >> 15);
}
template <> TAGPU_DI uint32_t tagpu_table_hash<2>(const Key<2> &k)
{
	const uint32_t x = (uint32_t)k.lo ^ ((uint32_t)(k.lo >> 32) * 0x85ebca6bu) ^ ((uint32_t)k.hi * 0xc2b2ae35u) ^
			   ((uint32_t)(k.hi >> 32) * 0x27d4eb2fu);
	const uint32_t h = x * 0x9e3779b1u;
	return h ^ (h >> 15);
}

// K bases of a right-aligned record value, `sh` bits above its right end (sh <= 62)
TAGPU_DI Key<1> tagpu_record_window(const SkRec<1> &r, int sh, int K)
{
	const unsigned long long w1 = r.w[1] & 0x0000ffffffffffffull;   // top 16 bits: window count, multiplicity
	Key<1> k;
	k.lo = (uint64_t)((((unsigned __int128)w1 << 64) | r.w[0]) >> sh);
	if (K < 32) k.lo &= (1ull << (2 * K)) - 1;
	return k;
}
TAGPU_DI Key<2> tagpu_record_window(const SkRec<2> &r, int sh, int K)
{
	Key<2> k;                                                    // n <= 32 windows: bases never reach word 3
	k.lo = (uint64_t)((((unsigned __int128)r.w[1] << 64) | r.w[0]) >> sh);
	k.hi = (uint64_t)((((unsigned __int128)r.w[2] << 64) | r.w[1]) >> sh);
	return KeyOps<2>::band(k, KeyOps<2>::mask(K));
}

// reverse complement of the nb bases of a record, right-aligned again (no length byte)
TAGPU_DI SkRec<1> tagpu_record_rc(const SkRec<1> &r, int nb)
{
	const unsigned __int128 f = ((unsigned __int128)tagpu_rc64_full(r.w[0]) << 64) | tagpu_rc64_full(r.w[1] & 0x00ffffffffffffffull);
	const unsigned __int128 v = f >> (2 * (64 - nb));
	SkRec<1> o;
	o.w[0] = (uint64_t)v;
	o.w[1] = (uint64_t)(v >> 64);
	return o;
}
TAGPU_DI SkRec<2> tagpu_record_rc(const SkRec<2> &r, int nb)
{
	const uint64_t f0 = tagpu_rc64_full(r.w[2]), f1 = tagpu_rc64_full(r.w[1]), f2 = tagpu_rc64_full(r.w[0]); // f2:f1:f0 = rc of 96 bases
	const int s = 2 * (96 - nb), ws = s >> 6, bs = s & 63;      // drop the complemented padding: shift right by s bits
	const uint64_t a0 = ws == 0 ? f0 : (ws == 1 ? f1 : f2), a1 = ws == 0 ? f1 : (ws == 1 ? f2 : 0ull), a2 = ws == 0 ? f2 : 0ull;
	SkRec<2> o;
	o.w[0] = (uint64_t)((((unsigned __int128)a1 << 64) | a0) >> bs);
	o.w[1] = (uint64_t)((((unsigned __int128)a2 << 64) | a1) >> bs);
	o.w[2] = a2 >> bs;
	o.w[3] = 0;
	return o;
}

// ---------------------------------------------------------------- duplicate records
// Pass 1 cuts super-k-mers at minimizer occurrences only, so reads that cover the same genomic site without an error in
// it produce the same record (or its reverse complement).  Pass 2 brings every record into its canonical orientation
// (the smaller of the two base strings), finds equal records among the 32 of a chunk with one __match_any_sync on a hash
// plus a full compare, and counts the windows of one representative with the multiplicity of the class.
template <int W> TAGPU_DI bool tagpu_record_less(const SkRec<W> &a, const SkRec<W> &b);   // base words only (no length byte)
template <> TAGPU_DI bool tagpu_record_less<1>(const SkRec<1> &a, const SkRec<1> &b)
{
	return a.w[1] != b.w[1] ? a.w[1] < b.w[1] : a.w[0] < b.w[0];
}
template <> TAGPU_DI bool tagpu_record_less<2>(const SkRec<2> &a, const SkRec<2> &b)
{
	if (a.w[2] != b.w[2]) return a.w[2] < b.w[2];
	return a.w[1] != b.w[1] ? a.w[1] < b.w[1] : a.w[0] < b.w[0];
}
template <int W> TAGPU_DI bool tagpu_record_equal(const SkRec<W> &a, const SkRec<W> &b)
{
	bool eq = true;
#pragma unroll
	for (int i = 0; i < 2 * W; ++i) eq = eq && a.w[i] == b.w[i];
	return eq;
}
template <int W> TAGPU_DI uint32_t tagpu_record_hash(const SkRec<W> &a)
{
	uint32_t x = 0;
#pragma unroll
	for (int i = 0; i < 2 * W; ++i) {
		x = (x ^ (uint32_t)a.w[i]) * 0x85ebca6bu;
		x = (x ^ (uint32_t)(a.w[i] >> 32)) * 0xc2b2ae35u;
	}
	return (x ^ (x >> 15)) & 0x7fffffffu;
}

// ---------------------------------------------------------------- cursors of the owned buckets + bucket grouping
// Bucket sizes are very uneven (a bucket is a handful of minimizer sites; measured CV ~0.9), so pass 2 does not take
// buckets one by one: consecutive buckets are packed into groups of ~GROUP_TARGET windows, and one CTA counts a whole
// group in one shared-memory table.  With E[b] = windows of the owned buckets before b, bucket b belongs to group
// floor(E[b] / target); groups are therefore found with one prefix sum, computed by three small grid-wide kernels:
//   k_pull_cursors   one thread per owned bucket: its cursor at every source rank (one coalesced sweep, remote for the
//                    other ranks) -> cur_all[lb * world + s], ext_all[...]; block-level scan of the window totals
//   k_scan_blocks    exclusive scan of the block totals, number of groups
//   k_mark_groups    first / end bucket of every group (a group id nobody maps to stays empty)
constexpr int TAGPU_SCAN_BLOCK = 1024;

template <int W>
__global__ void __launch_bounds__(TAGPU_SCAN_BLOCK) k_pull_cursors(const __grid_constant__ CountPeers<W> peers, uint32_t world,
									  uint32_t first_bucket, uint32_t n_owned, uint32_t n_buckets,
									  uint32_t cap_records, unsigned long long *__restrict__ cur_all,
									  uint32_t *__restrict__ ext_all, unsigned long long *__restrict__ pex,
									  unsigned long long *__restrict__ bsum, uint32_t self, unsigned long long *ctr)
{
	__shared__ unsigned long long s_w[32];
	const uint32_t lb = blockIdx.x * blockDim.x + threadIdx.x, gb = first_bucket + lb, lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	unsigned long long wsum = 0, rec_local = 0, rec_peer = 0;
	if (lb < n_owned)
		for (uint32_t s = 0; s < world; ++s) {
			unsigned long long cur = 0;
			uint32_t eo = 0;
			if (gb < n_buckets) {
				cur = peers.cursor[s][gb];
				if ((uint32_t)cur > cap_records) eo = peers.ext_off[s][gb];
			}
			cur_all[(size_t)lb * world + s] = cur;
			ext_all[(size_t)lb * world + s] = eo;
			wsum += cur >> 32;
			if (s == self) rec_local += (uint32_t)cur; else rec_peer += (uint32_t)cur;
		}
	// what pass 2 is going to read, split by where it lives (the peer share crosses NVLink)
#pragma unroll
	for (int d = 16; d; d >>= 1) {
		rec_local += __shfl_xor_sync(0xffffffffu, rec_local, d);
		rec_peer += __shfl_xor_sync(0xffffffffu, rec_peer, d);
	}
	if (lane == 0) {
		if (rec_local) atomicAdd(ctr + CTR_REC_LOCAL, rec_local);
		if (rec_peer) atomicAdd(ctr + CTR_REC_PEER, rec_peer);
	}
	unsigned long long incl = wsum;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
		if (lane >= (uint32_t)d) incl += t;
	}
	if (lane == 31) s_w[warp] = incl;
	__syncthreads();
	if (warp == 0) {
		unsigned long long x = s_w[lane], y = x;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			unsigned long long t = __shfl_up_sync(0xffffffffu, y, d);
			if (lane >= (uint32_t)d) y += t;
		}
		s_w[lane] = y - x;
		if (lane == 31) bsum[blockIdx.x] = y;
	}
	__syncthreads();
	if (lb < n_owned) pex[lb] = s_w[warp] + incl - wsum;        // windows of this block's buckets before lb
}

// single block: bsum[i] -> windows before block i; ctr[CTR_GROUPS] = number of group ids; resets the work counter
__global__ void __launch_bounds__(1024) k_scan_blocks(unsigned long long *__restrict__ bsum, uint32_t n_blocks, uint32_t target, unsigned long long *ctr)
{
	__shared__ unsigned long long s_part[1024];
	const uint32_t per = (n_blocks + 1023) / 1024, lo = min(threadIdx.x * per, n_blocks), hi = min(lo + per, n_blocks);
	unsigned long long sum = 0;
	for (uint32_t i = lo; i < hi; ++i) sum += bsum[i];
	s_part[threadIdx.x] = sum;
	__syncthreads();
	if (threadIdx.x < 32) {                                     // 1024 partials: 32 per lane, then a warp scan
		unsigned long long mine = 0;
		for (int t = 0; t < 32; ++t) mine += s_part[threadIdx.x * 32 + t];
		unsigned long long incl = mine;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
			if (threadIdx.x >= (uint32_t)d) incl += t;
		}
		unsigned long long acc = incl - mine;
		for (int t = 0; t < 32; ++t) { const unsigned long long v = s_part[threadIdx.x * 32 + t]; s_part[threadIdx.x * 32 + t] = acc; acc += v; }
		if (threadIdx.x == 31) {
			ctr[CTR_GROUPS] = incl / target + 1;
			ctr[CTR_SPARE1] = 0;                                // work counter of k_count_buckets
		}
	}
	__syncthreads();
	unsigned long long acc = s_part[threadIdx.x];
	for (uint32_t i = lo; i < hi; ++i) { const unsigned long long v = bsum[i]; bsum[i] = acc; acc += v; }
}

// grp_first[g] = first owned bucket of group g (TAGPU_NONE if no bucket starts in its window interval), grp_end[g] = one
// past its last bucket.  grp_first must be pre-filled with 0xff.
__global__ void __launch_bounds__(TAGPU_SCAN_BLOCK) k_mark_groups(const unsigned long long *__restrict__ pex, const unsigned long long *__restrict__ bsum,
									 uint32_t n_owned, uint32_t target, uint32_t *__restrict__ grp_first,
									 uint32_t *__restrict__ grp_end)
{
	const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= n_owned) return;
	const unsigned long long g = (bsum[b / TAGPU_SCAN_BLOCK] + pex[b]) / target;
	const unsigned long long gp = b ? (bsum[(b - 1) / TAGPU_SCAN_BLOCK] + pex[b - 1]) / target : ~0ull;
	if (g != gp) {
		grp_first[g] = b;
		if (b) grp_end[gp] = b;
	}
	if (b + 1 == n_owned) grp_end[g] = n_owned;
}

// One 64-byte descriptor per group id, so that the counting kernel's group set-up is ONE (prefetched) global load instead of
// a chain of dependent ones: { first owned bucket, buckets (0 = empty group id), windows, flags, record prefix over the
// group's (bucket, source) pairs: rpre[0 .. n_pairs] }.  Groups with more than TAGPU_DESC_PAIRS pairs or more than
// group_max buckets are flagged SLOW: the kernel then walks grp_first / grp_end / cur_all itself.
constexpr uint32_t TAGPU_DESC_PAIRS = 11, TAGPU_DESC_SLOW = 1u;
struct __align__(16) GroupDesc { uint32_t b0, nbk, windows, flags, rpre[TAGPU_DESC_PAIRS + 1]; };

__global__ void __launch_bounds__(256) k_group_desc(const uint32_t *__restrict__ grp_first, const uint32_t *__restrict__ grp_end,
						    const unsigned long long *__restrict__ cur_all, uint32_t world, uint32_t group_max,
						    const unsigned long long *ctr, GroupDesc *__restrict__ desc)
{
	const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= (uint32_t)ctr[CTR_GROUPS]) return;
	GroupDesc d;
	d.b0 = 0; d.nbk = 0; d.windows = 0; d.flags = 0;
#pragma unroll
	for (int i = 0; i <= (int)TAGPU_DESC_PAIRS; ++i) d.rpre[i] = 0;
	const uint32_t bs = grp_first[g];
	if (bs != TAGPU_NONE) {
		const uint32_t nbk = grp_end[g] - bs;
		d.b0 = bs;
		d.nbk = nbk;
		if (nbk > group_max || nbk * world > TAGPU_DESC_PAIRS) d.flags = TAGPU_DESC_SLOW;
		else {
			uint32_t acc = 0, win = 0;
			for (uint32_t i = 0; i < nbk * world; ++i) {
				const unsigned long long cur = cur_all[(size_t)bs * world + i];
				d.rpre[i] = acc;
				acc += (uint32_t)cur;
				win += (uint32_t)(cur >> 32);
			}
			d.rpre[nbk * world] = acc;
			d.windows = win;
		}
	}
	desc[g] = d;
}

// ---------------------------------------------------------------- staged records
// A staged record = the bases of a super-k-mer in canonical orientation, right-aligned in NW = W + 1 64-bit words
// (W = 1: <= 49 bases, W = 2: <= 95), plus one meta word: windows in bits 0..7, multiplicity in bits 8..15.
template <int W> struct StagedRec { unsigned long long w[W + 1]; };

// window j (0 = leftmost) of a staged record with n windows: the K bases that start sh = 2 (n - 1 - j) <= 62 bits above the
// right end.  32-bit funnel shifts: sh = 32 a + b, word i of the result = (r[a + i + 1] : r[a + i]) >> b.
TAGPU_DI Key<1> tagpu_staged_window(const StagedRec<1> &r, int sh, int K)
{
	const uint32_t r0 = (uint32_t)r.w[0], r1 = (uint32_t)(r.w[0] >> 32), r2 = (uint32_t)r.w[1], r3 = (uint32_t)(r.w[1] >> 32);
	const bool a = sh >= 32;
	const uint32_t t0 = a ? r1 : r0, t1 = a ? r2 : r1, t2 = a ? r3 : r2;
	const uint32_t o0 = __funnelshift_r(t0, t1, sh), o1 = __funnelshift_r(t1, t2, sh);   // (shift taken mod 32)
	Key<1> k;
	k.lo = ((unsigned long long)o1 << 32) | o0;
	if (K < 32) k.lo &= (1ull << (2 * K)) - 1;
	return k;
}
TAGPU_DI Key<2> tagpu_staged_window(const StagedRec<2> &r, int sh, int K)
{
	const uint32_t r0 = (uint32_t)r.w[0], r1 = (uint32_t)(r.w[0] >> 32), r2 = (uint32_t)r.w[1], r3 = (uint32_t)(r.w[1] >> 32),
		       r4 = (uint32_t)r.w[2], r5 = (uint32_t)(r.w[2] >> 32);
	const bool a = sh >= 32;
	const uint32_t t0 = a ? r1 : r0, t1 = a ? r2 : r1, t2 = a ? r3 : r2, t3 = a ? r4 : r3, t4 = a ? r5 : r4;
	const uint32_t o0 = __funnelshift_r(t0, t1, sh), o1 = __funnelshift_r(t1, t2, sh), o2 = __funnelshift_r(t2, t3, sh),
		       o3 = __funnelshift_r(t3, t4, sh);
	Key<2> k;
	k.lo = ((unsigned long long)o1 << 32) | o0;
	k.hi = ((unsigned long long)o3 << 32) | o2;
	return KeyOps<2>::band(k, KeyOps<2>::mask(K));
}

// Persistent CTAs pull groups of buckets from a global counter and count one group at a time in the CTA's
// shared-memory table.  A group is processed in rounds of <= ROUND records:
//
//   stage    every thread loads two records (coalesced: consecutive lanes, consecutive records), brings them into
//            canonical orientation and stages them in shared memory; equal records among the 32 of a warp collapse into
//            one representative with a multiplicity;
//   items    every live record of n windows becomes ceil(n / 8) work items (record, c); a block-wide prefix sum lays the
//            items of the round out as one dense list;
//   insert   an OCTET of lanes takes an item: lane q extracts window 8 c + q straight from the staged record (two funnel
//            shifts — no rolling, no dependence on the neighbouring window), reverse-complements it, and inserts the
//            canonical key into the table (LDS probe, ATOMS.CAS claim, ATOMS.ADD of the multiplicity).  All lanes of a warp
//            do the same thing in every iteration; only the probe outcome diverges.
//
// Everything with a global round trip is off the critical path: the id of the next group is fetched while the current
// one is being counted, and the output offset of the harvest (one global atomic) travels while the solid keys are
// compacted into the staging area.
#ifdef TAGPU_TIMING
#define TM_DECL() long long tm_setup = 0, tm_insert = 0, tm_wait = 0, tm_harvest = 0, tm_ha = 0, tm_hb = 0, tm_iters = 0, tm_failed = 0, tm_stage = 0, tm_t = clock64()
#define TM_ADD(x) do { long long n_ = clock64(); x += n_ - tm_t; tm_t = n_; } while (0)
#else
#define TM_DECL()
#define TM_ADD(x)
#endif

template <int W>
__global__ void __launch_bounds__(BucketCfg<W>::THREADS, BucketCfg<W>::CTAS_PER_SM)
k_count_buckets(const __grid_constant__ CountPeers<W> peers, uint32_t world, uint32_t first_bucket, uint32_t cap_records,
		const unsigned long long *__restrict__ cur_all, const uint32_t *__restrict__ ext_all, const uint32_t *__restrict__ grp_first,
		const uint32_t *__restrict__ grp_end, const GroupDesc *__restrict__ desc, uint32_t group_max, int K, uint32_t ci, Key<W> *__restrict__ solid,
		uint32_t *__restrict__ solid_cnt, unsigned long long solid_cap, SolidBlock *__restrict__ blocks, uint32_t blocks_cap, unsigned long long *ctr)
{
	typedef KeyOps<W> KO;
	typedef BucketCfg<W> C;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	Key<W> *t_key = reinterpret_cast<Key<W> *>(smem_raw);
	uint32_t *t_cnt = reinterpret_cast<uint32_t *>(t_key + C::SLOTS);
	unsigned char *stage = reinterpret_cast<unsigned char *>(t_cnt + C::SLOTS);
	StagedRec<W> *s_rec = reinterpret_cast<StagedRec<W> *>(stage);       // [0, ROUND): canonical orientation, [ROUND, 2 ROUND): its reverse complement
	uint32_t *s_meta = reinterpret_cast<uint32_t *>(s_rec + (C::DUAL ? 2 : 1) * C::ROUND);
	uint16_t *s_item = reinterpret_cast<uint16_t *>(s_meta + C::ROUND);
	// "table too full" flag of the current class: in the dynamic area, whose address is one add away from a register (a
	// static __shared__ variable costs a special-register read per access, and the insert loop polls this one)
	volatile uint32_t *s_overflow_p = reinterpret_cast<volatile uint32_t *>(s_item + C::ITEMS);
	uint32_t *s_dd = reinterpret_cast<uint32_t *>(s_item + C::ITEMS) + 4;     // [DD_SLOTS] duplicate table of a round (behind the flag)
#define s_overflow (*s_overflow_p)
	// during the harvest the staging area holds the compacted solid (key, count) pairs of the group
	constexpr uint32_t OUT_CAP = (uint32_t)((C::STAGE_BYTES - 16 - (size_t)C::DD_SLOTS * 4) / (sizeof(Key<W>) + 4));    // (the overflow flag and the duplicate table at the end are not part of it)
	Key<W> *o_key = reinterpret_cast<Key<W> *>(stage);
	uint32_t *o_cnt = reinterpret_cast<uint32_t *>(o_key + OUT_CAP);
	constexpr uint32_t TOP_NONE = 0xffffffffu;
	__shared__ uint32_t s_b0, s_nb, s_bs, s_be, s_top, s_claims, s_sp, s_nsolid, s_nout, s_pf, s_pf_n, s_warp_solid[C::THREADS / 32], s_stack[64];
	__shared__ uint32_t s_rpre[C::SUB_MAX + 1];                  // per (bucket, source) pair of the group: records before it
	__shared__ unsigned long long s_out_base;
	const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
	constexpr uint32_t N_WARPS = C::THREADS / 32;
	const uint32_t n_groups = (uint32_t)ctr[CTR_GROUPS];
	// central stretch of a K-mer that decides its orientation in the table: cb = 16 (K even) or 15 (K odd) bases, c_low bits
	// above the K-mer's right end
	const uint32_t c_bases = 16u - ((uint32_t)K & 1u), c_low = (uint32_t)K - c_bases, c_drop = 32u - 2u * c_bases,
		       c_mask = 0xffffffffu >> c_drop;
	TM_DECL();

	for (uint32_t i = tid; i < C::SLOTS; i += C::THREADS) { t_key[i] = KO::make(0, 0); t_cnt[i] = 0; }
	if constexpr (C::DD_SLOTS > 0)
		for (uint32_t i = tid; i < (uint32_t)C::DD_SLOTS; i += C::THREADS) s_dd[i] = 0;
	// Group ids come from a global counter.  Lane 0 of warp 0 keeps a queue of two: id1, whose 64-byte descriptor already sits
	// in the registers of lanes 0..3 (loaded while the group before was counted), and id2, whose atomicAdd is still in
	// flight.  So the set-up of a group touches no global memory unless the group is flagged SLOW.
	uint32_t id1 = 0, id2 = 0;
	uint4 dd = make_uint4(0, 0, 0, 0);
	if (warp == 0) {
		if (lane == 0) { id1 = (uint32_t)atomicAdd(ctr + CTR_SPARE1, 1ull); id2 = (uint32_t)atomicAdd(ctr + CTR_SPARE1, 1ull); s_bs = 0; s_be = 0; s_pf = TAGPU_NONE; }
		const uint32_t g1 = __shfl_sync(0xffffffffu, id1, 0);
		if (lane < 4 && g1 < n_groups) dd = __ldg(reinterpret_cast<const uint4 *>(desc + g1) + lane);
	}

	for (;;) {
		__syncthreads();                                            // B0: previous group fully harvested, staging area free
		if (warp == 0) {
			// ---- group setup by warp 0: the next group id (skipping empty ones) or the rest of an oversized one
			uint32_t bs = s_bs, be = s_be;
			bool fast = false;
			uint4 dcur = make_uint4(0, 0, 0, 0);
			while (bs >= be) {
				const uint32_t grp = __shfl_sync(0xffffffffu, id1, 0);
				if (grp >= n_groups) { bs = be = TAGPU_NONE; break; }
				dcur = dd;
				if (lane == 0) { id1 = id2; id2 = (uint32_t)atomicAdd(ctr + CTR_SPARE1, 1ull); }    // lands during the inserts
				const uint32_t g1 = __shfl_sync(0xffffffffu, id1, 0);                                 // (requested a whole group ago)
				dd = make_uint4(0, 0, 0, 0);
				if (lane < 4 && g1 < n_groups) dd = __ldg(reinterpret_cast<const uint4 *>(desc + g1) + lane);   // consumed at the next set-up
				const uint32_t d_b0 = __shfl_sync(0xffffffffu, dcur.x, 0), d_nbk = __shfl_sync(0xffffffffu, dcur.y, 0),
					       d_flags = __shfl_sync(0xffffffffu, dcur.w, 0);
				if (!d_nbk) continue;                                                              // empty group id
				bs = d_b0; be = d_b0 + d_nbk;
				fast = !(d_flags & TAGPU_DESC_SLOW);
			}
			if (bs == TAGPU_NONE) {
				if (lane == 0) s_nb = TAGPU_NONE;
			} else if (fast) {
				// everything comes out of the descriptor registers: lanes 1..3 hold rpre[0..11]
				const uint32_t b0 = bs, nb = (be - bs) * world;
				if (lane >= 1 && lane < 4) {
					const uint32_t base = 4u * (lane - 1u);
					if (base + 0u <= nb) s_rpre[base + 0u] = dcur.x;
					if (base + 1u <= nb) s_rpre[base + 1u] = dcur.y;
					if (base + 2u <= nb) s_rpre[base + 2u] = dcur.z;
					if (base + 3u <= nb) s_rpre[base + 3u] = dcur.w;
				}
				if (lane == 0) {
					const uint32_t tot_inst = dcur.z;
					uint32_t L = 0;                                  // a single oversized bucket starts on 2^L hash classes
					while (L < 5 && (tot_inst >> L) > 2u * C::GROUP_TARGET) ++L;
					uint32_t sp = 0;
					for (uint32_t c = 1; c < (1u << L); ++c) s_stack[sp++] = (L << 24) | c;
					s_sp = sp;
					s_top = L << 24;
					s_claims = 0; s_overflow = 0; s_nsolid = 0; s_nout = 0;
					s_b0 = b0; s_nb = nb;
					s_bs = be; s_be = be;
				}
			} else {
				// SLOW: many pairs (multi-GPU, tiny buckets) or more buckets than one table takes: walk the cursors
				const uint32_t b0 = bs, nb = min(be - bs, group_max) * world;
				uint32_t tot_inst = 0, carry = 0;
				for (uint32_t i0 = 0; i0 < nb; i0 += 32) {
					const uint32_t i = i0 + lane;
					const unsigned long long cur = i < nb ? cur_all[(size_t)b0 * world + i] : 0ull;
					const uint32_t nrec = (uint32_t)cur;
					tot_inst += (uint32_t)(cur >> 32);
					uint32_t incl = nrec;
#pragma unroll
					for (int d = 1; d < 32; d <<= 1) {
						uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
						if (lane >= (uint32_t)d) incl += t;
					}
					if (i < nb) s_rpre[i] = carry + incl - nrec;
					carry += __shfl_sync(0xffffffffu, incl, 31);
				}
				tot_inst = __reduce_add_sync(0xffffffffu, tot_inst);
				if (lane == 0) {
					s_rpre[nb] = carry;
					uint32_t L = 0;                                  // a single oversized bucket starts on 2^L hash classes
					while (L < 5 && (tot_inst >> L) > 2u * C::GROUP_TARGET) ++L;
					uint32_t sp = 0;
					for (uint32_t c = 1; c < (1u << L); ++c) s_stack[sp++] = (L << 24) | c;
					s_sp = sp;
					s_top = L << 24;                                 // class 0 of level L goes first
					s_claims = 0; s_overflow = 0; s_nsolid = 0; s_nout = 0;
					s_b0 = b0; s_nb = nb;
					s_bs = b0 + nb / world; s_be = be;
				}
			}
		}
		__syncthreads();                                            // B1
		const uint32_t b0 = s_b0, nb = s_nb;                            // nb (bucket, source) pairs <= SUB_MAX
		if (nb == TAGPU_NONE) break;
		const uint32_t n_recs = s_rpre[nb];
		TM_ADD(tm_setup);
		for (;;) {                                                  // one iteration per hash class (L, cls) of the group
			const uint32_t top = s_top;
			const uint32_t L = top >> 24, cls = top & 0xffffffu;
			uint32_t n_claimed = 0, n_became_solid = 0;                 // this thread's share of the class's distinct / solid keys
			// ---- insert every window of the group that belongs to hash class (L, cls), in rounds of equal size
			const uint32_t n_rounds = (n_recs + (uint32_t)C::ROUND - 1u) / (uint32_t)C::ROUND;
			const uint32_t per_round = n_rounds ? (n_recs + n_rounds - 1u) / n_rounds : 0u;
			for (uint32_t rbase = 0; rbase < n_recs; rbase += per_round) {
				if (rbase) __syncthreads();                         // the previous round's records and items are no longer needed
				if (s_overflow) break;
				const uint32_t n_round = min(per_round, n_recs - rbase);
				// ---- stage: records tid and tid + THREADS of the round
				uint32_t n_items_mine[2];
#pragma unroll
				for (int h = 0; h < 2; ++h) {
					const uint32_t idx = tid + (uint32_t)h * C::THREADS;
					const bool have = idx < n_round;
					uint32_t my_n = 0, rhash = 0x80000000u | lane;   // lanes without a record never match anybody
					StagedRec<W> canon;
					if (have) {
						const uint32_t g_idx = rbase + idx;
						uint32_t lo = 0, hi = nb;                        // pair of record g_idx: last i with s_rpre[i] <= g_idx
						while (hi - lo > 1) {
							const uint32_t mid = (lo + hi) >> 1;
							if (s_rpre[mid] <= g_idx) lo = mid; else hi = mid;
						}
						const uint32_t i = lo, g = g_idx - s_rpre[i];
						uint32_t lb = b0 + i, src = 0;
						if (world > 1) { lb = b0 + i / world; src = i - (i / world) * world; }
						const size_t gb = (size_t)first_bucket + lb;     // the source indexes its regions by global bucket id
						// (requesting both records of the thread before looking at either was measured: slower, 3.43 -> 3.82 ms)
						SkRec<W> pre = g < cap_records ? peers.regions[src][gb * cap_records + g]
									       : peers.ext[src][ext_all[(size_t)lb * world + src] + (g - cap_records)];
						my_n = (uint32_t)(pre.w[2 * W - 1] >> 56);
						pre.w[2 * W - 1] &= 0x00ffffffffffffffull;
						const SkRec<W> rc = tagpu_record_rc(pre, (int)my_n + K - 1);
						const bool flip = tagpu_record_less<W>(rc, pre);
#pragma unroll
						for (int q = 0; q < W + 1; ++q) canon.w[q] = flip ? rc.w[q] : pre.w[q];
						s_rec[idx] = canon;
						if constexpr (C::DUAL) {
							StagedRec<W> other;
#pragma unroll
							for (int q = 0; q < W + 1; ++q) other.w[q] = flip ? pre.w[q] : rc.w[q];
							s_rec[C::ROUND + idx] = other;
						}
						uint32_t x = 0;
#pragma unroll
						for (int q = 0; q < W + 1; ++q) {
							x = (x ^ (uint32_t)canon.w[q]) * 0x85ebca6bu;
							x = (x ^ (uint32_t)(canon.w[q] >> 32)) * 0xc2b2ae35u;
						}
						rhash = ((x ^ (x >> 15)) & 0x7fffff00u) | my_n;  // equal records have equal window counts
					}
					const uint32_t peers_eq = __match_any_sync(0xffffffffu, rhash);
					const uint32_t leader = (uint32_t)__ffs(peers_eq) - 1u;
					__syncwarp();
					bool dup = false;
					if (have && leader != lane) {
						const StagedRec<W> lead = s_rec[idx - lane + leader];
						dup = true;
#pragma unroll
						for (int q = 0; q < W + 1; ++q) dup = dup && lead.w[q] == canon.w[q];
					}
					const uint32_t dup_mask = __ballot_sync(0xffffffffu, dup);
					uint32_t items = 0;
					if (have && !dup) {
						const uint32_t mult = leader == lane ? 1u + (uint32_t)__popc(peers_eq & dup_mask) : 1u;
						s_meta[idx] = my_n | (mult << 8);
						items = (my_n + (uint32_t)C::ITEM_WINDOWS - 1u) / (uint32_t)C::ITEM_WINDOWS;
						if constexpr (C::DD_SLOTS > 0) {
						// the representatives of the warps meet in a small table of the round (slot = index of the first record
						// with this hash + 1): an equal record that got there first takes this one's multiplicity
						__threadfence_block();                           // s_rec[idx] and s_meta[idx] before the claim
						uint32_t ds = (rhash >> 8) & (uint32_t)(C::DD_SLOTS - 1);   // (the low byte of rhash is the window count)
						for (;;) {
							const uint32_t old = atomicCAS(s_dd + ds, 0u, idx + 1u);
							if (!old) break;
							const uint32_t other = old - 1u;
							const StagedRec<W> o_rec = s_rec[other];
							bool same = (s_meta[other] & 0xffu) == my_n;
#pragma unroll
							for (int q = 0; q < W + 1; ++q) same = same && o_rec.w[q] == canon.w[q];
							if (same) {
								atomicAdd(s_meta + other, mult << 8);
								s_meta[idx] = 0;
								items = 0;
								break;
							}
							ds = (ds + 1u) & (uint32_t)(C::DD_SLOTS - 1);
						}
						}
					}
					n_items_mine[h] = items;                             // a duplicate is counted through its representative
				}
				// ---- items: block-wide exclusive prefix over the items of the live records
				const uint32_t mine_items = n_items_mine[0] + n_items_mine[1];
				uint32_t incl = mine_items;
#pragma unroll
				for (int d = 1; d < 32; d <<= 1) {
					uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
					if (lane >= (uint32_t)d) incl += t;
				}
				if (lane == 31) s_warp_solid[warp] = incl;
				__syncthreads();
				const uint32_t wt = lane < N_WARPS ? s_warp_solid[lane] : 0u;
				const uint32_t n_items = __reduce_add_sync(0xffffffffu, wt);
				uint32_t o_item = __reduce_add_sync(0xffffffffu, lane < warp ? wt : 0u) + incl - mine_items;
#pragma unroll
				for (int h = 0; h < 2; ++h)
					for (uint32_t c = 0; c < n_items_mine[h]; ++c)
						s_item[o_item++] = (uint16_t)((tid + (uint32_t)h * C::THREADS) | (c << 11));
				TM_ADD(tm_stage);
				__syncthreads();                                    // staged records, meta words and items visible to everybody
				if constexpr (C::DD_SLOTS > 0)
					for (uint32_t i = tid; i < (uint32_t)C::DD_SLOTS; i += C::THREADS) s_dd[i] = 0;   // (next used behind at least one barrier)
				// ---- insert: one item per octet and iteration, one window per lane
				const uint32_t q = lane & 7u;
				// (a shared item cursor instead of this static split was measured: slower, 3.82 -> 3.88 ms at C2 — the atomic's
				// latency per iteration costs more than the imbalance at the barrier behind the loop)
				// key of window q of item `it` (branch-free up to the rare central palindrome, so that the two derivations of an
				// iteration interleave): returns false if there is no such window or it belongs to another hash class
				auto derive = [&](uint32_t it, Key<W> &key, uint32_t &mult, uint32_t &h) -> bool {
					const bool in = it < n_items;
					const uint32_t item = s_item[in ? it : 0u];
					const uint32_t idx = item & 0x7ffu, j = (item >> 11) * (uint32_t)C::ITEM_WINDOWS + q;
					const uint32_t meta = s_meta[idx], n_r = meta & 0xffu;
					mult = meta >> 8;
					const bool have = in && j < n_r;
					// orientation of a window: the table key is the window x or its reverse complement, whichever has the
					// smaller CENTRAL cb bases (a symmetric stretch around the middle of the K-mer: the central bases of rc(x)
					// are the reverse complement of those of x, so x and rc(x) agree on the choice).  That is one 32-bit
					// reverse complement instead of a K-base one, and the chosen orientation is then cut out of the record
					// staged in that orientation.  Central palindromes (4^-8 of the even-K windows) compare the full keys.
					const int sh_fw = have ? 2 * (int)(n_r - 1u - j) : 0, sh_rv = have ? 2 * (int)j : 0;
					if constexpr (C::DUAL) {
						const uint32_t *rw = reinterpret_cast<const uint32_t *>(s_rec + idx);
						const uint32_t c_off = (uint32_t)sh_fw + c_low, ca = c_off >> 5;
						const uint32_t cen = __funnelshift_r(rw[ca], rw[ca + 1], c_off) & c_mask;
						const uint32_t cen_rc = tagpu_rc32_full(cen) >> c_drop;
						const bool fwd = cen < cen_rc;
						key = tagpu_staged_window(s_rec[fwd ? idx : (uint32_t)C::ROUND + idx], fwd ? sh_fw : sh_rv, K);
						if (cen == cen_rc) {
							const Key<W> rv = tagpu_staged_window(s_rec[idx], sh_fw, K);
							key = KO::le(key, rv) ? key : rv;
						}
					} else {
						const Key<W> fw = tagpu_staged_window(s_rec[idx], sh_fw, K);
						const Key<W> rv = KO::rc(fw, K);
						key = KO::le(fw, rv) ? fw : rv;
					}
					h = tagpu_table_hash<W>(key);
					return have && (!L || (h & ((1u << L) - 1u)) == cls);
				};
				// probe + count: the hit / claim decision is the only divergent part
				auto insert = [&](const Key<W> &key, uint32_t mult, uint32_t h) {
					const Key<W> stored = KO::bnot(key);
					uint32_t slot = __umulhi(h, (uint32_t)C::SLOTS);
					int probes = 0;
					// linear probing, two slots per iteration: both are loaded together, then examined in probe order — the
					// lanes of a warp leave the loop after about half as many (divergent) iterations
					for (;;) {
						const uint32_t slot1 = slot + 1 == C::SLOTS ? 0u : slot + 1;
						const Key<W> have = t_key[slot], have1 = t_key[slot1];
						if (KO::eq(have, stored)) break;
						if (ktab_empty_or_torn<W>(have)) {
							const Key<W> old = ktab_cas<W>(t_key + slot, stored);   // ATOMS.CAS.64 / .128
							if (KO::is_zero(old)) { ++n_claimed; break; }
							if (KO::eq(old, stored)) break;
						}
						slot = slot1;                                            // the first slot holds another key
						if (KO::eq(have1, stored)) break;
						if (ktab_empty_or_torn<W>(have1)) {
							const Key<W> old = ktab_cas<W>(t_key + slot, stored);
							if (KO::is_zero(old)) { ++n_claimed; break; }
							if (KO::eq(old, stored)) break;
						}
						slot = slot + 1 == C::SLOTS ? 0u : slot + 1;
						probes += 2;
						if (probes > C::MAX_PROBES) { s_overflow = 1; break; }        // table too full: re-run on sub-classes
					}
					// exactly one insert takes a key across the cutoff: the harvest knows its size before it starts
					const uint32_t before_add = atomicAdd(t_cnt + slot, mult);
					n_became_solid += (before_add < ci && before_add + mult >= ci) ? 1u : 0u;
				};
				// two items per octet and iteration: the two key derivations are independent instruction streams (measured against
				// one item per iteration: 3.43 -> 3.40 ms at C2)
				for (uint32_t ibase = warp * 8u; ibase < n_items; ibase += N_WARPS * 8u) {
					if (s_overflow) break;
					Key<W> key_a, key_b;
					uint32_t mult_a, mult_b, h_a, h_b;
					const bool ok_a = derive(ibase + (lane >> 3), key_a, mult_a, h_a), ok_b = derive(ibase + 4u + (lane >> 3), key_b, mult_b, h_b);
					if (ok_a) insert(key_a, mult_a, h_a);
					if (ok_b) insert(key_b, mult_b, h_b);
				}
			}
			n_claimed = __reduce_add_sync(0xffffffffu, n_claimed);
			n_became_solid = __reduce_add_sync(0xffffffffu, n_became_solid);
			if (lane == 0) {
				if (n_claimed) atomicAdd(&s_claims, n_claimed);
				if (n_became_solid) atomicAdd(&s_nsolid, n_became_solid);
			}
			if (tid == 0) { s_pf = dd.x; s_pf_n = dd.y; }               // next group's buckets (its descriptor has landed): see the prefetch below
			TM_ADD(tm_insert);
			__syncthreads();                                        // B2: all inserts of this class are in the table, counters complete
			TM_ADD(tm_wait);
			// ---- harvest (or discard on overflow) in ONE pass over the table, which is left zeroed.  The number of solid keys
			// is known (counted by the inserts), so the global output range is requested first and travels during the pass.
			const bool failed = s_overflow != 0;
			const uint32_t n_out = failed ? 0u : s_nsolid, n_claims = s_claims;
			const bool staged = n_out <= OUT_CAP;                       // else (huge group) write straight to the global arrays
			unsigned long long out_base = 0;
			if (tid == 0) {
				out_base = n_out ? atomicAdd(ctr + CTR_SOLID, (unsigned long long)n_out) : 0ull;
				if (!failed) atomicAdd(ctr + CTR_DISTINCT, (unsigned long long)n_claims);
				if (!staged) s_out_base = out_base;
			}
			// while the table is scanned: pull the records of the next group's buckets towards L2 (single GPU: local regions)
			if (world == 1) {
				const uint32_t pb = s_pf, pn = s_pf_n, bkt = tid >> 5;
				if (pb != TAGPU_NONE && bkt < pn) {
					const char *line = reinterpret_cast<const char *>(peers.regions[0] + ((size_t)first_bucket + pb + bkt) * cap_records) + (size_t)lane * 128u;
					asm volatile("prefetch.global.L2 [%0];" ::"l"(line));
				}
			}
			if (!staged) __syncthreads();                           // (rare) the direct path needs s_out_base now
			const unsigned long long gbase = staged ? 0ull : s_out_base;
			// every thread takes the slots tid, tid + THREADS, ...: all their counts are loaded first (independent loads), the
			// thread's solid keys get consecutive places behind one warp-wide prefix sum and ONE shared-memory atomic per warp
			unsigned long long sum = 0;
			constexpr int HR = (C::SLOTS + C::THREADS - 1) / C::THREADS;
			uint32_t hc[HR];
			uint32_t n_mine = 0;
#pragma unroll
			for (int r = 0; r < HR; ++r) {
				const uint32_t i = tid + (uint32_t)r * C::THREADS;
				hc[r] = i < (uint32_t)C::SLOTS ? t_cnt[i] : 0u;
				n_mine += (!failed && hc[r] >= ci && hc[r] != 0u) ? 1u : 0u;
			}
			uint32_t h_incl = n_mine;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const uint32_t t = __shfl_up_sync(0xffffffffu, h_incl, d);
				if (lane >= (uint32_t)d) h_incl += t;
			}
			uint32_t o = 0;
			if (lane == 31 && h_incl) o = atomicAdd(&s_nout, h_incl);
			o = __shfl_sync(0xffffffffu, o, 31) + h_incl - n_mine;
#pragma unroll
			for (int r = 0; r < HR; ++r) {
				const uint32_t i = tid + (uint32_t)r * C::THREADS, c = hc[r];
				if (c) {
					if (!failed && c >= ci) {
						const Key<W> key = KO::bnot(t_key[i]);
						if (staged) { o_key[o] = key; o_cnt[o] = c; }
						else if (gbase + o < solid_cap) {                   // (rare: huge group) canonical form right here
							const Key<W> krc = KO::rc(key, K);
							solid[gbase + o] = KO::le(key, krc) ? key : krc;
							solid_cnt[gbase + o] = c;
						}
						sum += c;
						++o;
					}
					t_key[i] = KO::make(0, 0);
					t_cnt[i] = 0;
				}
			}
#pragma unroll
			for (int d = 16; d; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
			if (lane == 0 && sum) atomicAdd(ctr + CTR_SUM_SOLID, sum);
			if (tid == 0) {
				s_out_base = out_base;
				if (n_out) {                                         // directory of the solid list, for the graph stage
					const uint32_t id = (uint32_t)atomicAdd(ctr + CTR_BLOCKS, 1ull);
					if (id < blocks_cap) {
						SolidBlock sb;
						sb.base = out_base; sb.n = n_out; sb.b0 = first_bucket + b0; sb.nbk = nb / world;
						sb.flags = (L || !staged) ? 1u : 0u;
						blocks[id] = sb;
					} else atomicOr(ctr + CTR_ERROR, (unsigned long long)TAGPU_ERR_BLOCKS);
				}
				// next hash class: children of a failed class first, then whatever is left on the stack
				uint32_t sp = s_sp;
				if (failed) {
					if (L >= 16 || sp + 2 > 64) atomicOr(ctr + CTR_ERROR, (unsigned long long)TAGPU_ERR_TABLE_FULL);
					else {
						s_stack[sp++] = ((L + 1) << 24) | cls;
						s_stack[sp++] = ((L + 1) << 24) | (cls + (1u << L));
					}
				}
				s_top = sp ? s_stack[--sp] : TOP_NONE;
				s_sp = sp;
			}
			TM_ADD(tm_hb);
#ifdef TAGPU_TIMING
			tm_iters += 1; tm_failed += failed ? 1 : 0;
#endif
			__syncthreads();                                        // B4: table zeroed, compacted output + s_out_base + s_top visible
			if (tid == 0) { s_claims = 0; s_overflow = 0; s_nsolid = 0; s_nout = 0; }   // (read by everybody before B4; next used after the next barrier)
			if (staged) {
				const unsigned long long base = s_out_base;
				for (uint32_t i = tid; i < n_out; i += C::THREADS)
					if (base + i < solid_cap) {                     // the host reports the overflow (n_solid > solid_cap)
						// the table key is the centrally-oriented representative: the solid list holds canonical (k+1)-mers,
						// min(x, rc(x)) — converted here, on the dense copy-out of the few keys that made the cutoff
						const Key<W> key = o_key[i], krc = KO::rc(key, K);
						solid[base + i] = KO::le(key, krc) ? key : krc;
						solid_cnt[base + i] = o_cnt[i];
					}
			}
			TM_ADD(tm_harvest);
			if (s_top == TOP_NONE) break;                           // group done (B0 of the next group protects the staging area)
			__syncthreads();                                        // another class of the same group: staging area must be drained first
		}
	}
#ifdef TAGPU_TIMING
	if (lane == 0) {          // cycles summed over all warps: setup / stage / insert / wait at the post-insert barrier / harvest
		atomicAdd(ctr + CTR_JUMP_FLAGS + 48, (unsigned long long)tm_setup);
		atomicAdd(ctr + CTR_JUMP_FLAGS + 49, (unsigned long long)tm_insert);
		atomicAdd(ctr + CTR_JUMP_FLAGS + 50, (unsigned long long)tm_wait);
		atomicAdd(ctr + CTR_JUMP_FLAGS + 51, (unsigned long long)tm_harvest);
		atomicAdd(ctr + CTR_JUMP_FLAGS + 52, (unsigned long long)tm_ha);
		atomicAdd(ctr + CTR_JUMP_FLAGS + 53, (unsigned long long)tm_hb);
		atomicAdd(ctr + CTR_JUMP_FLAGS + 56, (unsigned long long)tm_stage);
		if (warp == 0) { atomicAdd(ctr + CTR_JUMP_FLAGS + 54, (unsigned long long)tm_iters); atomicAdd(ctr + CTR_JUMP_FLAGS + 55, (unsigned long long)tm_failed); }
	}
#endif
}
#undef s_overflow
