// Read-stream tile loader and rolling (k+1)-mer extraction shared by the counting kernels.
//
// Input model: a flat byte stream; A/C/G/T/a/c/g/t are bases (nt4_table,
// /root/reference/src/utils.c:26-43), every other byte (newline between reads, 'N', ...) breaks the
// window (cf. /root/reference/src/k63_build.c:397-410).  A tile is TILE_BASES consecutive byte
// positions; the CTA packs them (plus a 64-base halo to the left) into shared memory as 2-bit codes,
// 32 bases per 64-bit word (first base most significant) with a 32-bit invalid mask per word, and then
// every thread rolls forward and reverse-complement mers over the 32 window-end positions of one word.
#pragma once
#include "tagpu_key.cuh"

// Tile geometry of pass 1, by the number of 32-base words whose positions are window ends.  128 words (4096 bases, 160
// threads, 38 KB of shared memory: 6 CTAs per SM) is what pass 1 uses: many small CTAs hide the tile load and the barriers
// better than 256 words (288 threads, 74 KB: 3 CTAs per SM) — 1.40 -> 1.29 ms at C2 (w = 32), 1.89 -> 1.68 ms at C1 (w = 18).
// The 256-word geometry remains the tile of the packed stream layout and a developer switch (TAGPU_TILE_WORDS=256).
constexpr int TAGPU_HALO_WORDS = 3;                 // 96 bases to the left: K <= 64 of history for a window, plus the 31 windows a super-k-mer may reach back
constexpr int TAGPU_RHALO_WORDS = 1;                // one word to the right: whether a super-k-mer ends at the tile's last position depends on the next window
template <int TW> struct TileCfg {
	static constexpr int WORDS = TW;
	static constexpr int THREADS = TW == 256 ? 288 : 160;      // >= SMEM_WORDS: every per-word phase (halo words included) is ONE pass over the threads
#ifndef TAGPU_TILE_MIN_CTAS
#define TAGPU_TILE_MIN_CTAS 6
#endif
	static constexpr int MIN_CTAS = TW == 256 ? 3 : TAGPU_TILE_MIN_CTAS;         // CTAs per SM the shared memory allows: the register budget follows (launch bounds)
	static constexpr int BASES = TW * 32;
	static constexpr int SMEM_WORDS = TW + TAGPU_HALO_WORDS + TAGPU_RHALO_WORDS;
	static constexpr int HM_POS = SMEM_WORDS * 32;             // positions of the packed tile (incl. halo)
	static constexpr int HM_LEN = SMEM_WORDS * 33;             // padded: index q + q/32, so word-major and position-major accesses are both conflict-free
	static constexpr int END_CAP = 896 * TW / 256;             // run ends of a tile handled per emission pass (a 256-word tile of 151 bp reads has ~590)
	static constexpr size_t SMEM = (size_t)SMEM_WORDS * 8 + 3 * (size_t)SMEM_WORDS * 4 + 2 * (size_t)HM_LEN * 4;
	static_assert(THREADS >= SMEM_WORDS, "per-word phases assume one word per thread");
};
constexpr int TAGPU_TILE_WORDS = 256;               // the general tile, and the tile of the PACKED stream layout (include/tagpu.h)
constexpr int TAGPU_TILE_BASES = TAGPU_TILE_WORDS * 32;

// 4 ASCII bytes (byte 0 = first base) -> 8 bits of codes (first base in bits 7..6) + 4 invalid bits (first base = bit 3)
// Validity without byte-wise compares (the SIMD video instructions are emulated on sm_100): the 2-bit code of a byte picks
// the letter it would have to be out of "ACGT" (one PRMT); the byte is a base iff it equals that letter, case folded.
TAGPU_DI void tagpu_pack4(uint32_t w, uint32_t &codes, uint32_t &inv)
{
	uint32_t c = (w >> 1) & 0x03030303u;        // A=0 C=1 G=3 T=2
	c ^= (c >> 1) & 0x01010101u;                // A=0 C=1 G=2 T=3
	codes = (c * 0x40100401u) >> 24;
	const uint32_t t = c | (c >> 4);            // byte 0: c0 | c1 << 4, byte 2: c2 | c3 << 4
	const uint32_t sel = __byte_perm(t, 0u, 0x4420u);              // selector nibbles c0, c1, c2, c3
	const uint32_t expect = __byte_perm(0x54474341u, 0u, sel);    // "ACGT"[code] per byte
	const uint32_t diff = (w & 0xdfdfdfdfu) ^ expect;             // fold lower case; zero byte <=> valid base
	const uint32_t nz = (((diff & 0x7f7f7f7fu) + 0x7f7f7f7fu) | diff) & 0x80808080u;   // bit 7 of every non-zero byte
	inv = ((nz >> 7) * 0x08040201u) >> 24 & 0xfu;
}

// Packs the tile that owns window-end positions [tile_base, tile_base + TILE_BASES) into pk/inv.
// smem word j covers stream positions tile_base - 32 HALO_WORDS + 32 j .. +31.
template <int TW>
TAGPU_DI void tagpu_load_tile(const uint8_t *__restrict__ seq, uint64_t n, uint64_t tile_base,
			       uint64_t *pk, uint32_t *inv)
{
	for (int j = threadIdx.x; j < TileCfg<TW>::SMEM_WORDS; j += blockDim.x) {
		long long g0 = (long long)tile_base - 32 * TAGPU_HALO_WORDS + 32ll * j;
		uint64_t word = 0;
		uint32_t bad = 0;
		if (g0 >= 0 && (uint64_t)g0 + 32 <= n && ((reinterpret_cast<uintptr_t>(seq) + g0) & 15) == 0) {
			const uint4 *p = reinterpret_cast<const uint4 *>(seq + g0);
			uint4 a = __ldg(p), b = __ldg(p + 1);
			uint32_t w[8] = { a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w };
#pragma unroll
			for (int q = 0; q < 8; ++q) {
				uint32_t c, iv;
				tagpu_pack4(w[q], c, iv);
				word = (word << 8) | c;
				bad = (bad << 4) | iv;
			}
		} else {
			for (int q = 0; q < 32; ++q) {
				long long g = g0 + q;
				uint32_t c = 0, iv = 1;
				if (g >= 0 && (uint64_t)g < n) {
					uint32_t ch = seq[g], code, i4;
					tagpu_pack4(ch, code, i4); // byte 0 only: code in bits 7..6, invalid in bit 3
					c = code >> 6;
					iv = (i4 >> 3) & 1u;
				}
				word = (word << 2) | c;
				bad = (bad << 1) | iv;
			}
		}
		pk[j] = word;
		inv[j] = bad;
	}
}

// The same tile out of a PACKED read stream (include/tagpu.h "packed read stream"): the host (tagpu_pack_stream) has
// already done the ASCII -> 2-bit conversion in exactly the shared-memory layout above, tile by tile —
// TAGPU_TILE_WORDS 64-bit code words followed by TAGPU_TILE_WORDS 32-bit invalid masks = 3072 bytes per 8192 positions
// (0.375 bytes per base instead of 1: that is what crosses PCIe) — so loading a tile is a plain copy.
constexpr int TAGPU_PACKED_TILE_BYTES = TAGPU_TILE_WORDS * 12;
template <int TW>
TAGPU_DI void tagpu_load_tile_packed(const uint8_t *__restrict__ packed, uint64_t n, uint64_t tile_base, uint64_t *pk, uint32_t *inv)
{
	const uint64_t n_words = ((n + TAGPU_TILE_BASES - 1) / TAGPU_TILE_BASES) * TAGPU_TILE_WORDS;   // whole tiles of the packed layout
	for (int j = threadIdx.x; j < TileCfg<TW>::SMEM_WORDS; j += blockDim.x) {
		const long long w = (long long)(tile_base / 32) - TAGPU_HALO_WORDS + j;              // stream word of smem word j
		uint64_t word = 0;
		uint32_t bad = 0xffffffffu;
		if (w >= 0 && (uint64_t)w < n_words) {
			const uint8_t *tile = packed + ((uint64_t)w / TAGPU_TILE_WORDS) * TAGPU_PACKED_TILE_BYTES;
			const uint32_t o = (uint32_t)((uint64_t)w % TAGPU_TILE_WORDS);
			word = __ldg(reinterpret_cast<const uint64_t *>(tile) + o);
			bad = __ldg(reinterpret_cast<const uint32_t *>(tile + TAGPU_TILE_WORDS * 8) + o);
		}
		pk[j] = word;
		inv[j] = bad;
	}
}

// Calls f(canonical_key, pos_in_word) for every valid window of K bases ending in word `wi` (tile-local, >= HALO_WORDS).
// Returns the number of valid windows.
template <int W, typename F>
TAGPU_DI uint32_t tagpu_roll_word(const uint64_t *pk, const uint32_t *inv, int wi, int K, F &&f)
{
	typedef KeyOps<W> KO;
	typedef Key<W> KT;
	const KT m = KO::mask(K);
	KT fw = KO::band(KO::make(pk[wi - 2], pk[wi - 1]), m);
	KT rv = KO::rc(fw, K);
	uint32_t i1 = inv[wi - 1], i2 = inv[wi - 2];
	int run = i1 ? (__ffs(i1) - 1) : 32 + (i2 ? (__ffs(i2) - 1) : 32);
	uint64_t cur = pk[wi];
	uint32_t iv = inv[wi];
	uint32_t n_valid = 0;
#pragma unroll 4
	for (int i = 0; i < 32; ++i) {
		uint32_t c = (uint32_t)(cur >> 62);
		cur <<= 2;
		bool bad = (int)iv < 0;
		iv <<= 1;
		fw = KO::push(fw, c, m);
		rv = KO::push_front(rv, 3u - c, K);
		run = bad ? 0 : run + 1;
		if (run >= K) {
			++n_valid;
			f(KO::le(fw, rv) ? fw : rv, i);
		}
	}
	return n_valid;
}
