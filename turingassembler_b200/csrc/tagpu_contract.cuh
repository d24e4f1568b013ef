// Graph stage, two-level variant: contract the solid (k+1)-mer list inside the bucket groups of the count stage first,
// then run the global stage on the contracted paths.
//
// Every harvest of k_count_buckets leaves a contiguous block of solid (k+1)-mers whose minimizers lie in one group of
// buckets (SolidBlock).  A k-mer z met in such a block can be HIDDEN when
//   (i)   the bucket of its own minimizer mu(z) belongs to the block's group, and
//   (ii)  none of its 8 possible one-base extensions brings an m-mer that hashes below mu(z)
// — then every (k+1)-mer containing z has minimizer mu(z) and lives in this very block, so the block-local edge mask of z
// is its global one — and
//   (iii) that mask is 1-in-1-out.
// Hidden k-mers are interior to a unitig by construction.  k_contract chains the (k+1)-mers of a block across hidden
// k-mers into PATHS (first / last (k+1)-mer, number of (k+1)-mers, summed count, packed interior bases) entirely in
// shared memory; the global kernels below are the path-driven, weighted counterparts of tagpu_graph.cuh: only the
// END k-mers of the paths enter the HBM table, and list ranking carries distances in (k+1)-mers (= bases of the unitig).
// A path and its reverse complement are one object (like a canonical (k+1)-mer); a cycle made of hidden k-mers only is a
// node-free component and is dropped here exactly as the reference never emits it.
#pragma once
#include "tagpu_count.cuh"
#include "tagpu_graph.cuh"

// (k+1)-mers of a block that k_contract handles, in three size classes with a launch each (occupancy falls with the
// shared memory a block needs, so every block runs in the smallest class that holds it).  SMALL: ~30-40 KB per CTA keep
// 5-7 CTAs of 128 threads on an SM; that covers nearly all blocks of a deep-coverage read set (a group of ~7700 windows
// holds ~130 solid (k+1)-mers at 130x).  MEDIUM / LARGE: up to 512 (W = 2 only) / 1024 entries (shallow coverage: a larger
// share of the windows is distinct and solid).  Still larger blocks, and blocks the count stage had to split by hash
// class, stay uncontracted (every (k+1)-mer a path of its own).
template <int W> struct ContractCfg {
	// (threads of the SMALL class measured at 64 / 128 / 256: 128 is best for 128-bit keys — C2 0.80 ms against 1.02 / 0.81 —,
	// 256 for 64-bit keys, whose blocks hold twice as many entries — C1 0.54 ms against 1.01 / 0.67)
	static constexpr int MAXN_SMALL = W == 1 ? 512 : 256, T_SMALL = W == 1 ? 256 : 128;
	static constexpr int MAXN_MEDIUM = 512, T_MEDIUM = 256;            // (W = 1: same size as SMALL, launch skipped)
	static constexpr int MAXN_LARGE = 1024, T_LARGE = 512;
};
template <int W, int MAXN> constexpr size_t tagpu_contract_smem()
{
	return (size_t)MAXN * sizeof(Key<W>) + 4 * (size_t)MAXN * sizeof(Key<W>) + (size_t)MAXN * 4 + 4 * (size_t)MAXN * 4 + 8 * (size_t)MAXN * 2 + 2 * (size_t)MAXN * 2;
}
constexpr uint32_t TAGPU_OE_END = 0xffffu, TAGPU_OE_PAL = 0x8000u;   // oriented entries are < 2 * 1024

template <int W> struct PathStore {
	Key<W> *first, *last;             // first / last (k+1)-mer of the path, oriented along the path
	uint32_t *n;                      // (k+1)-mers on the path
	unsigned long long *cnt;          // sum of their counts
	unsigned long long *off;          // first word of the interior bases (bases k+1 .. k+n-1 of the path, 2 bits each)
	uint32_t *interior;
	unsigned long long cap_paths, cap_words;
};

// hash (26 bits) of the canonical m-mer `fw` (m = TAGPU_MINIMIZER_M), exactly as k_partition computes it
TAGPU_DI uint32_t tagpu_mmer_hash26(uint32_t fw)
{
	const int m = TAGPU_MINIMIZER_M;
	const uint32_t rv = (uint32_t)(tagpu_rc64_full((uint64_t)fw) >> (64 - 2 * m));
	return (min(fw, rv) * 0x9e3779b1u) >> 6;
}

// the k bases of z, first base in the top bits of a 128-bit register pair (so that a 2-bit left shift yields the next base)
template <int W> TAGPU_DI void tagpu_left_align(const Key<W> &z, int k, uint64_t &hi, uint64_t &lo);
template <> TAGPU_DI void tagpu_left_align<1>(const Key<1> &z, int k, uint64_t &hi, uint64_t &lo) { hi = z.lo << (64 - 2 * k); lo = 0; }
template <> TAGPU_DI void tagpu_left_align<2>(const Key<2> &z, int k, uint64_t &hi, uint64_t &lo)
{
	const int sh = 128 - 2 * k;                                     // 2 .. 126 (k = 32..63 with two words)
	if (sh >= 64) { hi = z.lo << (sh - 64); lo = 0; }
	else { hi = (z.hi << sh) | (z.lo >> (64 - sh)); lo = z.lo << sh; }
}

// (i) + (ii) for the k-mer z: home bucket in [b0, b0 + nbk) and no potentially foreign extension.
// No rolling state: with z left-aligned in 16-base words, the m-mer ending at base 16 q + i is a 30-bit field of the word
// pair (q - 1, q), and its reverse complement the mirrored field of the pair's reverse-complemented words (the scheme of
// k_partition's phase A) — 8 operations per m-mer, all shift amounts compile-time constants.
template <int W>
TAGPU_DI bool tagpu_kmer_is_local(const Key<W> &z, int k, int log2_buckets, uint32_t b0, uint32_t nbk)
{
	const int m = TAGPU_MINIMIZER_M;
	const uint32_t mm = (1u << (2 * m)) - 1u;
	uint64_t hi, lo;
	tagpu_left_align<W>(z, k, hi, lo);
	uint32_t Z[2 * W], R[2 * W];
	Z[0] = (uint32_t)(hi >> 32); Z[1] = (uint32_t)hi;
	if constexpr (W == 2) { Z[2] = (uint32_t)(lo >> 32); Z[3] = (uint32_t)lo; }
#pragma unroll
	for (int q = 0; q < 2 * W; ++q) R[q] = tagpu_rc32_full(Z[q]);
	uint32_t mu = 0xffffffffu;
#pragma unroll
	for (int q = 0; q < 2 * W; ++q) {
		if (16 * q >= k) break;
		const uint32_t zq = Z[q], zp = q ? Z[q - 1] : 0u, rq = R[q], rp = q ? R[q - 1] : 0u;
#pragma unroll
		for (int i = 0; i < 16; ++i) {
			const int e = 16 * q + i;                                   // last base of the m-mer
			if (e < m - 1) continue;
			const uint32_t fw = __funnelshift_r(zq, zp, 30 - 2 * i) & mm;
			const uint32_t rv = (4 + 2 * i < 32 ? __funnelshift_r(rp, rq, 4 + 2 * i) : rq >> (4 + 2 * i - 32)) & mm;
			const uint32_t h = (min(fw, rv) * 0x9e3779b1u) >> 6;
			mu = e < k ? min(mu, h) : mu;
		}
	}
	const uint32_t b = tagpu_bucket_of(mu, log2_buckets);
	if (b < b0 || b >= b0 + nbk) return false;
	// right extensions: last m-1 bases of z + c; left extensions: c + first m-1 bases of z
	const uint32_t tail = (uint32_t)z.lo & (mm >> 2), head = Z[0] >> (32 - 2 * (m - 1));
	for (uint32_t c = 0; c < 4; ++c) {
		if (tagpu_mmer_hash26((tail << 2) | c) < mu) return false;
		if (tagpu_mmer_hash26((c << (2 * (m - 1))) | head) < mu) return false;
	}
	return true;
}

// ---------------------------------------------------------------- contraction of one block in shared memory
// Oriented entry oe = 2 i + o: o = 0 the stored canonical (k+1)-mer x_i, o = 1 its reverse complement.
// Output: dense path arrays.  A block first counts its paths and interior words (shared-memory bump counters give every
// path its place inside the block), reserves the range for all of them with one atomic per global counter
// (ctr[CTR_PATHS], ctr[CTR_PATH_WORDS]), then writes.  The order of the paths in the arrays is therefore arbitrary.
// A launch works through a list of block ids (src; nullptr = the whole directory), takes the contractible blocks of up to
// MAXN entries and appends the ids of all others to dst[] (counts in ctr[src_ctr] / ctr[dst_ctr]) for the next launch;
// the last launch (dst == nullptr) also takes the uncontractible ones (copied out as single-entry paths).
template <int W, int MAXN, int T>
__global__ void __launch_bounds__(T)
k_contract(const SolidBlock *__restrict__ blocks, uint32_t n_directory, const Key<W> *__restrict__ solid, const uint32_t *__restrict__ solid_cnt,
	   int k, int log2_buckets, const uint32_t *__restrict__ src, int src_ctr, uint32_t *__restrict__ dst, int dst_ctr, PathStore<W> ps,
	   unsigned long long *ctr)
{
	typedef KeyOps<W> KO;
	constexpr int TS_MAX = 4 * MAXN;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	Key<W> *e_key = reinterpret_cast<Key<W> *>(smem_raw);               // [MAXN]
	Key<W> *t_key = e_key + MAXN;                                       // [TS_MAX] canonical k-mers, ~key (0 = empty)
	uint32_t *e_cnt = reinterpret_cast<uint32_t *>(t_key + TS_MAX);     // [MAXN]
	uint32_t *t_mask = e_cnt + MAXN;                                    // [TS_MAX] local edge mask (low 8 bits), bit 8 = hidden
	uint16_t *t_out = reinterpret_cast<uint16_t *>(t_mask + TS_MAX);    // [TS_MAX][2] an oriented entry leaving (k-mer, orient)
	uint16_t *nxt = t_out + 2 * TS_MAX;                                 // [2 MAXN] next oriented entry on the path / END
	__shared__ uint32_t s_block, s_words, s_paths, s_hidden, s_cand, s_cand2;
	__shared__ unsigned long long s_pbase, s_wbase;
	const uint32_t tid = threadIdx.x;
	const int K = k + 1;
	const Key<W> kmask = KO::mask(k);
	const uint32_t n_blocks = src ? (uint32_t)ctr[src_ctr] : n_directory;   // (the launch before has finished that list)
	uint32_t next_block = 0;                                            // thread 0: id requested one block ahead
	if (tid == 0) next_block = (uint32_t)atomicAdd(ctr + CTR_SPARE1, 1ull);
#ifdef TAGPU_TIMING
	long long tc[6] = { 0, 0, 0, 0, 0, 0 }, tc_t = clock64();
#define TC(i) do { long long n_ = clock64(); tc[i] += n_ - tc_t; tc_t = n_; } while (0)
#else
#define TC(i)
#endif

	for (;;) {
		__syncthreads();
		if (tid == 0) {
			s_block = next_block;
			if (next_block < n_blocks) next_block = (uint32_t)atomicAdd(ctr + CTR_SPARE1, 1ull);
			s_words = 0; s_paths = 0; s_hidden = 0; s_cand = 0; s_cand2 = 0;
		}
		__syncthreads();
		if (s_block >= n_blocks) break;
		const uint32_t blk = src ? src[s_block] : s_block;
		const SolidBlock sb = blocks[blk];
		const uint32_t n = sb.n;
		if (dst && (sb.flags || n > (uint32_t)MAXN)) {                   // for the next launch
			if (tid == 0) dst[atomicAdd(ctr + dst_ctr, 1ull)] = blk;
			continue;
		}
		if (sb.flags || n > (uint32_t)MAXN) {
			// not contractible: every (k+1)-mer is a path of its own
			if (tid == 0) s_pbase = atomicAdd(ctr + CTR_PATHS, (unsigned long long)n);
			__syncthreads();
			const unsigned long long pb = s_pbase;
			for (uint32_t i = tid; i < n; i += T) {
				const Key<W> x = solid[sb.base + i];
				ps.first[pb + i] = x; ps.last[pb + i] = x; ps.n[pb + i] = 1u; ps.cnt[pb + i] = solid_cnt[sb.base + i];
				ps.off[pb + i] = 0ull;
			}
			continue;
		}
		uint32_t ts = 64;
		while (ts < 4u * n) ts <<= 1;                                // load <= 0.5 (at most 2 n k-mers)
		for (uint32_t i = tid; i < ts; i += T) { t_key[i] = KO::make(0, 0); t_mask[i] = 0; }
		for (uint32_t i = tid; i < n; i += T) { e_key[i] = solid[sb.base + i]; e_cnt[i] = solid_cnt[sb.base + i]; }
		__syncthreads();
		TC(0);
		// ---- local k-mer table with masks and one leaving entry per (k-mer, orientation)
		for (uint32_t oe = tid; oe < 2u * n; oe += T) {
			// one reverse complement per entry serves both orientations: with y = b0..bk, the head k-mer b0..b(k-1) is
			// shr2(y) and ITS reverse complement is the low k bases of rc(y)
			const Key<W> x = e_key[oe >> 1], xr = KO::rc(x, K);
			const Key<W> y = (oe & 1u) ? xr : x, yr = (oe & 1u) ? x : xr;
			const Key<W> q = KO::shr2(y), qr = KO::band(yr, kmask);       // head k-mer of the oriented entry
			const bool fwd = KO::le(q, qr);
			const Key<W> z = fwd ? q : qr, stored = KO::bnot(z);
			uint32_t s = (uint32_t)KO::hash(z) & (ts - 1);
			for (;;) {
				const Key<W> have = t_key[s];
				if (KO::eq(have, stored)) break;
				if (KO::is_zero(have) || ktab_maybe_torn<W>(have)) {
					const Key<W> old = ktab_cas<W>(t_key + s, stored);
					if (KO::is_zero(old) || KO::eq(old, stored)) break;
				}
				s = (s + 1) & (ts - 1);
			}
			const uint32_t oz = fwd ? 0u : 1u;
			atomicOr(t_mask + s, 1u << (oz * 4u + KO::last_base(y)));
			t_out[2u * s + oz] = (uint16_t)(oe | (KO::eq(x, xr) ? TAGPU_OE_PAL : 0u));   // bit 15: palindromic (k+1)-mer
		}
		__syncthreads();
		TC(1);
		// ---- which k-mers can be hidden.  The cheap conditions first, candidates compacted (nxt[] is free until the link
		// phase), so that the expensive minimizer test runs on full warps.
		for (uint32_t s = tid; s < ts; s += T) {
			const Key<W> st = t_key[s];
			if (KO::is_zero(st)) continue;
			const uint32_t m = t_mask[s];
			if (DEG4(m) != 1 || DEG4(m >> 4) != 1) continue;
			// the two (k+1)-mers through z must be two different, non-palindromic entries: a hairpin (z followed by its own
			// reverse complement) or a palindromic (k+1)-mer would make a path that is its own reverse complement, and
			// those stay with the (k+1)-mer-level rules of the global stage
			const uint32_t o0 = t_out[2u * s], o1 = t_out[2u * s + 1u];
			if ((o0 >> 1) == (o1 >> 1) || ((o0 | o1) & TAGPU_OE_PAL)) continue;
			nxt[atomicAdd(&s_cand, 1u)] = (uint16_t)s;
		}
		__syncthreads();
		uint32_t n_hidden = 0;
		for (uint32_t c = tid; c < s_cand; c += T) {
			const uint32_t s = nxt[c];
			if (tagpu_kmer_is_local<W>(KO::bnot(t_key[s]), k, log2_buckets, sb.b0, sb.nbk)) {
				t_mask[s] |= 0x100u;
				++n_hidden;
			}
		}
		if (n_hidden) atomicAdd(&s_hidden, n_hidden);
		__syncthreads();
		TC(2);
		// ---- links: the entry that continues an oriented entry across a hidden k-mer
		for (uint32_t oe = tid; oe < 2u * n; oe += T) {
			const Key<W> x = e_key[oe >> 1], xr = KO::rc(x, K);
			const Key<W> y = (oe & 1u) ? xr : x, yr = (oe & 1u) ? x : xr;
			const Key<W> q = KO::band(y, kmask), qr = KO::shr2(yr);       // tail k-mer b1..bk, and rc of it = first k bases of rc(y)
			const bool fwd = KO::le(q, qr);
			const Key<W> stored = KO::bnot(fwd ? q : qr);
			uint32_t s = (uint32_t)KO::hash(fwd ? q : qr) & (ts - 1), probes = 0;
			while (!KO::eq(t_key[s], stored) && probes < ts) { s = (s + 1) & (ts - 1); ++probes; }
			uint16_t nx = (uint16_t)TAGPU_OE_END;
			if (probes >= ts) atomicOr(ctr + CTR_ERROR, (unsigned long long)TAGPU_ERR_CONTRACT);
			else if (t_mask[s] & 0x100u) nx = (uint16_t)(t_out[2u * s + (fwd ? 0u : 1u)] & ~TAGPU_OE_PAL);
			nxt[oe] = nx;
		}
		__syncthreads();
		TC(3);
		// ---- heads.  One oriented entry in eight heads a path; they are compacted first (the table arrays are dead after the
		// link phase: t_out holds the list of heads, t_key the records of the paths to emit), so that the serial walks run on
		// dense lanes.  A path is emitted by its smaller end: head <= reverse complement of its last entry.
		uint16_t *h_list = t_out;                                       // [<= 2 n] oriented entries that head a path
		uint4 *p_list = reinterpret_cast<uint4 *>(t_key);                // [<= n] { oe | last << 16, len, path index, word offset }
		for (uint32_t oe = tid; oe < 2u * n; oe += T)
			if (nxt[oe ^ 1u] == TAGPU_OE_END) h_list[atomicAdd(&s_cand2, 1u)] = (uint16_t)oe;   // (else: the k-mer before it is hidden)
		__syncthreads();
		for (uint32_t hI = tid; hI < s_cand2; hI += T) {
			const uint32_t oe = h_list[hI];
			uint32_t len = 0, last = oe, cur = oe;
			for (;;) {
				++len;
				last = cur;
				const uint32_t nx = nxt[cur];
				if (nx == TAGPU_OE_END) break;
				if (len > 2u * n) { atomicOr(ctr + CTR_ERROR, (unsigned long long)TAGPU_ERR_CONTRACT); break; }
				cur = nx;
			}
			if (oe > (last ^ 1u)) continue;                               // the twin path emits
			// (a single entry is a head in both orientations and is emitted once, as orientation 0; a path that is its own
			// twin, oe == last ^ 1, has one head only and is emitted by it)
			const uint32_t words = len > 1 ? (len - 1 + 15) >> 4 : 0u;
			const uint32_t wo = words ? atomicAdd(&s_words, words) : 0u;
			const uint32_t pi = atomicAdd(&s_paths, 1u);
			p_list[pi] = make_uint4(oe | (last << 16), len, pi, wo);
		}
		__syncthreads();
		if (tid == 0) {
			s_pbase = atomicAdd(ctr + CTR_PATHS, (unsigned long long)s_paths);
			s_wbase = s_words ? atomicAdd(ctr + CTR_PATH_WORDS, (unsigned long long)s_words) : 0ull;
			if (s_hidden) atomicAdd(ctr + CTR_KMERS, (unsigned long long)s_hidden);
		}
		__syncthreads();
		const unsigned long long pbase = s_pbase, wbase = s_wbase;
		for (uint32_t pI = tid; pI < s_paths; pI += T) {
			const uint4 rec = p_list[pI];
			const uint32_t oe = rec.x & 0xffffu, last = rec.x >> 16, len = rec.y;
			const unsigned long long pi = pbase + rec.z, wo = wbase + rec.w;
			const Key<W> xf = (oe & 1u) ? KO::rc(e_key[oe >> 1], K) : e_key[oe >> 1];
			const Key<W> xl = (last & 1u) ? KO::rc(e_key[last >> 1], K) : e_key[last >> 1];
			ps.first[pi] = xf; ps.last[pi] = xl; ps.off[pi] = wo; ps.n[pi] = len;
			unsigned long long csum = e_cnt[oe >> 1];
			uint32_t c2 = nxt[oe], word = 0;
			for (uint32_t j = 0; j + 1 < len; ++j) {                     // interior base j = last base of the (j + 2)-th entry
				const Key<W> xe = e_key[c2 >> 1];
				csum += e_cnt[c2 >> 1];
				const uint32_t base = (c2 & 1u) ? 3u - KO::first_base(xe, K) : KO::last_base(xe);
				word |= base << ((j & 15u) << 1);
				if ((j & 15u) == 15u || j + 2 == len) { ps.interior[wo + (j >> 4)] = word; word = 0; }
				c2 = nxt[c2];
			}
			ps.cnt[pi] = csum;
		}
		TC(4);
	}
#ifdef TAGPU_TIMING
	if ((tid & 31u) == 0) for (int i = 0; i < 5; ++i) atomicAdd(ctr + CTR_JUMP_FLAGS + 57 + i, (unsigned long long)tc[i]);
#endif
}

// ---------------------------------------------------------------- multi-GPU: paths of every rank into one dense store
// Every rank contracts its own solid list into a PathStore inside its arena; this kernel pulls all of them over NVLink
// peer loads (coalesced, grid-strided) into this rank's dense arrays and rebases the interior-word offsets on the way.
template <int W> struct PathPeers {
	PathStore<W> src[TAGPU_MAX_RANKS];                       // as mapped on this rank
	unsigned long long n_paths[TAGPU_MAX_RANKS], n_words[TAGPU_MAX_RANKS], p0[TAGPU_MAX_RANKS], w0[TAGPU_MAX_RANKS];
	int world;
};

template <int W>
__global__ void __launch_bounds__(256) k_gather_paths(const __grid_constant__ PathPeers<W> pp, PathStore<W> dst)
{
	const unsigned long long t0 = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (unsigned long long)gridDim.x * blockDim.x;
	for (int r = 0; r < pp.world; ++r) {
		const PathStore<W> &s = pp.src[r];
		const unsigned long long p0 = pp.p0[r], w0 = pp.w0[r];
		for (unsigned long long i = t0; i < pp.n_paths[r]; i += stride) {
			dst.first[p0 + i] = s.first[i];
			dst.last[p0 + i] = s.last[i];
			dst.n[p0 + i] = s.n[i];
			dst.cnt[p0 + i] = s.cnt[i];
			dst.off[p0 + i] = s.off[i] + w0;
		}
		for (unsigned long long i = t0; i < pp.n_words[r]; i += stride) dst.interior[w0 + i] = s.interior[i];
	}
}

// base i (0 .. k + n - 1) of a path: the first k + 1 from its first (k+1)-mer, the rest from the interior words
template <int W>
TAGPU_DI uint32_t tagpu_path_base(const PathStore<W> &ps, unsigned long long p, const Key<W> &xf, int k, uint32_t i)
{
	if (i <= (uint32_t)k) return KeyOps<W>::base_at(xf, k + 1, (int)i);
	const uint32_t j = i - (uint32_t)k - 1u;
	return (ps.interior[ps.off[p] + (j >> 4)] >> ((j & 15u) << 1)) & 3u;
}

// The k + n bases of a path as a little-endian 2-bit stream (base i at bits 2 i, 2 i + 1 — the layout of the edge sequences,
// /root/reference/src/assembly_graph.h:182-187), 32 bits at a time: the k + 1 bases of the first (k+1)-mer with their order
// reversed (= rc(x) with the complement undone), followed by the interior words shifted by 2 (k + 1) bits.  The emission
// kernels cut whole output words out of it with funnel shifts instead of fetching the bases one by one.
template <int W> struct PathBits {
	uint32_t s[2 * W + 1];           // the first (k+1)-mer, base 0 in the lowest bits (one zero word behind it)
	const uint32_t *interior;        // interior words of the path
	int n_int;                       // how many
	int a, b;                        // 2 (k + 1) = 32 a + b

	TAGPU_DI PathBits(const PathStore<W> &ps, unsigned long long p, const Key<W> &xf, int k, uint32_t n)
	{
		typedef KeyOps<W> KO;
		const Key<W> r = KO::rc(xf, k + 1), m = KO::mask(k + 1);
		s[0] = (uint32_t)(r.lo ^ m.lo); s[1] = (uint32_t)((r.lo ^ m.lo) >> 32);
		if constexpr (W == 2) { s[2] = (uint32_t)(r.hi ^ m.hi); s[3] = (uint32_t)((r.hi ^ m.hi) >> 32); }
		s[2 * W] = 0;
		interior = ps.interior + ps.off[p];
		n_int = n > 1u ? (int)((n - 1u + 15u) >> 4) : 0;
		a = (2 * (k + 1)) >> 5; b = (2 * (k + 1)) & 31;
	}
	TAGPU_DI uint32_t int_word(int j) const { return j >= 0 && j < n_int ? interior[j] : 0u; }
	TAGPU_DI uint32_t key_word(int q) const
	{
		uint32_t v = 0;
#pragma unroll
		for (int i = 0; i <= 2 * W; ++i) v = q == i ? s[i] : v;
		return v;
	}
	// word q of the stream (0 outside of it)
	TAGPU_DI uint32_t word(int q) const
	{
		if (q < a) return q < 0 ? 0u : key_word(q);
		const int j = q - a;
		return (j == 0 ? key_word(a) : 0u) | __funnelshift_l(int_word(j - 1), int_word(j), b);
	}
	// the 32 bits that start at bit `lo` of the stream (any alignment, may start before the stream or end behind it)
	TAGPU_DI uint32_t bits(int lo) const
	{
		const int q = lo >> 5;                                   // (arithmetic shift: floor)
		return __funnelshift_r(word(q), word(q + 1), lo & 31);
	}
};

// 32-bit mask of the output bases [lo, hi) of the word that holds bases 16 w .. 16 w + 15
TAGPU_DI uint32_t tagpu_base_mask(int w, int lo, int hi)
{
	const int first = max(lo - 16 * w, 0), end = min(hi - 16 * w, 16);
	if (end <= first) return 0u;
	const uint32_t upto = end >= 16 ? 0xffffffffu : (1u << (2 * end)) - 1u;
	return upto & ~((1u << (2 * first)) - 1u);                   // (first < 16 here)
}

// ---------------------------------------------------------------- B': end k-mers of the paths into the HBM table
template <int W>
__global__ void __launch_bounds__(256) k_insert_paths(PathStore<W> ps, uint64_t n_paths, int k, KTab<W> t, uint32_t *__restrict__ vL,
						       uint32_t *__restrict__ vR, unsigned long long *ctr)
{
	typedef KeyOps<W> KO;
	const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t n_new = 0;
	if (p < n_paths && ps.n[p]) {
		const Key<W> xf = ps.first[p], xl = ps.last[p];
		const Key<W> km = KO::mask(k);
		const Key<W> k1 = KO::shr2(xf), k2 = KO::band(xl, km);
		const uint32_t c1 = KO::last_base(xf), c2 = 3u - KO::first_base(xl, k + 1);
		const Key<W> r1 = KO::rc(k1, k), r2 = KO::rc(k2, k);
		const bool f1 = KO::le(k1, r1), f2 = KO::le(k2, r2);
		bool claimed;
		const uint32_t s1 = ktab_insert<W>(t, f1 ? k1 : r1, &claimed, ctr + CTR_ERROR);
		n_new += claimed;
		atomicOr(&t.mask32[s1 >> 2], (1u << (f1 ? c1 : c1 + 4u)) << ((s1 & 3u) * 8u));
		const uint32_t s2 = ktab_insert<W>(t, f2 ? k2 : r2, &claimed, ctr + CTR_ERROR);
		n_new += claimed;
		atomicOr(&t.mask32[s2 >> 2], (1u << (f2 ? c2 + 4u : c2)) << ((s2 & 3u) * 8u));
		vL[p] = s1 * 2u + (f1 ? 0u : 1u);   // oriented vertex the path leaves (out-base c1)
		vR[p] = s2 * 2u + (f2 ? 1u : 0u);   // oriented vertex its reverse complement leaves (out-base c2)
	}
	n_new = __reduce_add_sync(0xffffffffu, n_new);
	if ((threadIdx.x & 31) == 0 && n_new) atomicAdd(ctr + CTR_KMERS, (unsigned long long)n_new);
}

// ---------------------------------------------------------------- C2': successor links, one thread per path and direction
// jump[cv] = (next chain vertex, bases of the connecting path); the last chain vertex before a node points at itself with
// TERM, vsucc = that node vertex, wlast = bases of the final path.
template <int W>
__global__ void __launch_bounds__(256) k_succ_paths(PathStore<W> ps, uint64_t n_paths, const uint32_t *__restrict__ vL, const uint32_t *__restrict__ vR,
						     const uint32_t *__restrict__ kind, unsigned long long *__restrict__ jump,
						     uint32_t *__restrict__ vsucc, uint32_t *__restrict__ wlast)
{
	const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= 2 * n_paths) return;
	const uint64_t p = idx >> 1;
	const uint32_t dir = (uint32_t)idx & 1u;
	const uint32_t w = ps.n[p];
	if (!w) return;                                                  // no path in this slot
	const uint32_t a = dir ? vR[p] : vL[p], b = (dir ? vL[p] : vR[p]) ^ 1u;
	const uint32_t ka = kind[a >> 1];
	if (!(ka & TAGPU_CHAIN)) return;                                 // leaves a node: k_heads_paths
	const uint32_t cva = (ka & ~TAGPU_CHAIN) * 2u + (a & 1u), kb = kind[b >> 1];
	if (kb & TAGPU_CHAIN) {
		jump[cva] = tagpu_pack_jump((kb & ~TAGPU_CHAIN) * 2u + (b & 1u), w);
		vsucc[cva] = TAGPU_NONE;
	} else {
		jump[cva] = tagpu_pack_jump(TAGPU_TERM | cva, 0);
		vsucc[cva] = kb * 2u + (b & 1u);
		wlast[cva] = w;
	}
}

// ---------------------------------------------------------------- C4': edge heads, one thread per path and direction
// Only one path end in seven starts an edge (leaves a node), and those write k + n bases each: the CTA compacts them
// in shared memory first, so that the base loop runs on full warps.
constexpr int TAGPU_HEADS_THREADS = 512;
template <int W>
__global__ void __launch_bounds__(TAGPU_HEADS_THREADS) k_heads_paths(PathStore<W> ps, uint64_t n_paths, int k, KTab<W> t, const uint32_t *__restrict__ vL,
						      const uint32_t *__restrict__ vR, const uint32_t *__restrict__ kind,
						      const uint32_t *__restrict__ node_ebase, const unsigned long long *__restrict__ jump,
						      const uint32_t *__restrict__ vsucc, const uint32_t *__restrict__ wlast, uint32_t *__restrict__ vedge,
						      FlatGraph g, unsigned long long *ctr)
{
	typedef KeyOps<W> KO;
	const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const bool in = idx < 2 * n_paths;
	const uint64_t p = in ? idx >> 1 : 0;
	const uint32_t dir = (uint32_t)idx & 1u;
	bool have = false;
	uint32_t a = 0, b = 0, w = 0, ord = 0, len = 0, dst = 0, first = TAGPU_NONE, e = 0;
	Key<W> xf = KO::make(0, 0);
	if (in && (w = ps.n[p]) != 0u) {
		a = dir ? vR[p] : vL[p];
		b = (dir ? vL[p] : vR[p]) ^ 1u;
		const uint32_t ka = kind[a >> 1];
		if (!(ka & TAGPU_CHAIN)) {                                   // the path leaves a node: it starts an edge
			have = true;
			ord = ka;
			xf = ps.first[p];
			const Key<W> xl = ps.last[p];
			const uint32_t c = dir ? 3u - KO::first_base(xl, k + 1) : KO::last_base(xf);
			const uint32_t m = ktab_mask_of<W>(t, a >> 1), o = a & 1u, nib = o ? (m >> 4) : (m & 15u);
			e = node_ebase[ord] + (o ? DEG4(m) : 0u) + tagpu_rank4(nib, c);
			const uint32_t kb = kind[b >> 1];
			if (!(kb & TAGPU_CHAIN)) {
				len = (uint32_t)k + w;
				dst = kb * 2u + (b & 1u);
			} else {
				const uint32_t cvb = (kb & ~TAGPU_CHAIN) * 2u + (b & 1u);
				const unsigned long long j = jump[cvb];
				if (!((uint32_t)j & TAGPU_TERM)) {
					atomicOr(ctr + CTR_ERROR, (unsigned long long)TAGPU_ERR_CHAIN);
					have = false;
				} else {
					const uint32_t tv = (uint32_t)j & ~TAGPU_TERM;
					len = (uint32_t)k + w + (uint32_t)(j >> 32) + wlast[tv];
					dst = vsucc[tv];
					first = cvb;
				}
			}
		}
	}
	__shared__ uint32_t s_n;
	__shared__ unsigned long long s_item[TAGPU_HEADS_THREADS], s_off[TAGPU_HEADS_THREADS];
	if (threadIdx.x == 0) s_n = 0;
	__syncthreads();
	const unsigned long long off = tagpu_warp_alloc(ctr + CTR_SEQ_WORDS, have ? (len + 15u) >> 4 : 0u);
	if (have) {
		g.e_src[e] = ord * 2u + (a & 1u); g.e_dst[e] = dst; g.e_len[e] = len; g.e_off[e] = off; g.e_count[e] = 0;
		if (first != TAGPU_NONE) vedge[first] = e;
		const uint32_t slot = atomicAdd(&s_n, 1u);
		s_item[slot] = idx;                                          // path and direction
		s_off[slot] = off;
	}
	__syncthreads();
	for (uint32_t it = threadIdx.x; it < s_n; it += blockDim.x) {
		const unsigned long long p2 = s_item[it] >> 1, off2 = s_off[it];
		const uint32_t dir2 = (uint32_t)s_item[it] & 1u, n = ps.n[p2];
		const Key<W> xf2 = ps.first[p2];
		// first k bases: the node k-mer as the edge sees it; then the n bases of this path.  Forward: stream words as they
		// are; backward: the reverse complement of the stream, i.e. output word w = rc of the 16 bases that end at base L - 16 w
		const PathBits<W> pb(ps, p2, xf2, k, n);
		const int L = k + (int)n, n_words = (L + 15) >> 4;
		for (int w = 0; w < n_words; ++w) {
			uint32_t word = dir2 ? tagpu_rc32_full(pb.bits(2 * (L - 16 * w - 16))) : pb.word(w);
			word &= tagpu_base_mask(w, 0, L);
			if (word) atomicOr(g.e_seq + off2 + w, word);
		}
	}
}

// ---------------------------------------------------------------- C5': every path that leaves a chain vertex writes its bases
template <int W>
__global__ void __launch_bounds__(256) k_interior_paths(PathStore<W> ps, uint64_t n_paths, int k, const uint32_t *__restrict__ vL,
							 const uint32_t *__restrict__ vR, const uint32_t *__restrict__ kind,
							 const unsigned long long *__restrict__ jump, const uint32_t *__restrict__ wlast,
							 uint32_t *vedge, FlatGraph g)
{
	const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= 2 * n_paths) return;
	const uint64_t p = idx >> 1;
	const uint32_t dir = (uint32_t)idx & 1u;
	if (!ps.n[p]) return;
	const uint32_t a = dir ? vR[p] : vL[p], ka = kind[a >> 1];
	if (!(ka & TAGPU_CHAIN)) return;
	const uint32_t cv = (ka & ~TAGPU_CHAIN) * 2u + (a & 1u);
	const unsigned long long j = jump[cv], jt = jump[cv ^ 1u];
	if (!((uint32_t)j & TAGPU_TERM) || !((uint32_t)jt & TAGPU_TERM)) return;  // node-free cycle
	const uint32_t tr = (uint32_t)jt & ~TAGPU_TERM;                  // last chain vertex of the reverse walk = twin of the edge's first one
	const uint32_t e = vedge[tr ^ 1u];
	if (e == TAGPU_NONE) return;
	const uint32_t pos = (uint32_t)k + wlast[tr] + (uint32_t)(jt >> 32), n = ps.n[p];
	const Key<W> xf = ps.first[p];
	const unsigned long long off = g.e_off[e];
	// output bases [pos, pos + n) of the edge: forward = stream bases k .. k + n - 1, backward = complement of bases n - 1 .. 0
	const PathBits<W> pb(ps, p, xf, k, n);
	const int lo = (int)pos, hi = (int)(pos + n);
	for (int w = lo >> 4; w <= (hi - 1) >> 4; ++w) {
		// output base o = 16 w + i: forward <- stream base k + o - pos; backward <- complement of stream base n - 1 - (o - pos)
		uint32_t word = dir ? tagpu_rc32_full(pb.bits(2 * ((int)n - 16 - 16 * w + lo))) : pb.bits(2 * (k + 16 * w - lo));
		word &= tagpu_base_mask(w, lo, hi);
		if (word) atomicOr(g.e_seq + off + w, word);
	}
	if (cv != (tr ^ 1u)) vedge[cv] = e;
}

// ---------------------------------------------------------------- C7': edge counts, one thread per path
template <int W>
__global__ void __launch_bounds__(256) k_counts_paths(PathStore<W> ps, uint64_t n_paths, int k, KTab<W> t, const uint32_t *__restrict__ vL,
						       const uint32_t *__restrict__ kind, const uint32_t *__restrict__ node_ebase,
						       const uint32_t *__restrict__ vedge, FlatGraph g, unsigned long long *ctr)
{
	typedef KeyOps<W> KO;
	const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	unsigned long long on_edge = 0;
	if (p < n_paths && ps.n[p]) {
		const uint32_t v = vL[p], slot = v >> 1, o = v & 1u, kd = kind[slot];
		uint32_t e;
		if (!(kd & TAGPU_CHAIN)) {
			const uint32_t m = ktab_mask_of<W>(t, slot), nib = o ? (m >> 4) : (m & 15u);
			e = node_ebase[kd] + (o ? DEG4(m) : 0u) + tagpu_rank4(nib, KO::last_base(ps.first[p]));
		} else {
			e = vedge[(kd & ~TAGPU_CHAIN) * 2u + o];
		}
		if (e != TAGPU_NONE) {
			const unsigned long long c = ps.cnt[p];
			atomicAdd(g.e_count + e, c);
			atomicAdd(g.e_count + g.e_rc[e], c);
			on_edge = ps.n[p];
		}
	}
#pragma unroll
	for (int d = 16; d; d >>= 1) on_edge += __shfl_xor_sync(0xffffffffu, on_edge, d);
	if ((threadIdx.x & 31) == 0 && on_edge) atomicAdd(ctr + CTR_KP1_ON_EDGE, on_edge);
}
