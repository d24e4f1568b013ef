/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the reference's coverage recount (`build_coverage` sub-command):
 *   kmer_count_on_edges  /root/reference/src/coverage/kmer_count.c:198-240
 *     construct_edges_hash :137-150 + index_bin_edge :68-84   every 31-mer of every edge of >= 32 bases enters a table
 *     kmer_count_iterator  :152-197 + get_and_add_kmer :86-111  every 31-base window of every read of >= 32 bases adds 1 to the
 *                                                              entry of the window and 1 to the entry of its "rev"
 *   add_cnt_to_graph     :113-135                             edge count = sum over its 31-mers of min(entry, 999), then
 *                                                              max with the count of the reverse-complement edge
 * with the reference's arithmetic kept literally, because two of its properties are visible in the result:
 *   * a base that is not ACGTacgt has code 4 (nt4_table, /root/reference/src/utils.c:26-43) and is OR-ed into the 64-bit
 *     k-mer register unmasked (`km |= c << (pad + 2)`): it contributes 00 for itself and sets the low bit of the base
 *     before it; the window is NOT skipped;
 *   * "rev" is the bit-reversed register (`__reverse_bit`, :17-23), i.e. the reversed base string with C and G swapped —
 *     not the reverse complement (the source's own comment: "why kmer count is not symetric").
 * Reads are the '\n'-separated lines of the stream (the sequence lines of the FASTQ files, as everywhere in this oracle).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "ta_oracle.h"

#define COV_K 31          /* KMER_SIZE_COVERAGE */
#define COV_MAX 999       /* MAX_KMER_COUNT */

static uint64_t cov_mix(uint64_t x)
{
	x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
	x ^= x >> 27; x *= 0x94d049bb133111ebull;
	x ^= x >> 31;
	return x;
}

struct cov_tab {
	uint64_t *key;    /* km | 1 (a k-mer register has its two low bits clear), 0 = empty */
	uint64_t *cnt;
	uint64_t mask;
};

static uint64_t *cov_slot(struct cov_tab *t, uint64_t km, int insert)
{
	uint64_t s = cov_mix(km) & t->mask;
	for (;;) {
		if (t->key[s] == (km | 1))
			return t->cnt + s;
		if (!t->key[s]) {
			if (!insert)
				return NULL;
			t->key[s] = km | 1;
			return t->cnt + s;
		}
		s = (s + 1) & t->mask;
	}
}

static uint64_t cov_bitrev(uint64_t a)       /* __reverse_bit, kmer_count.c:17-23 */
{
	uint64_t b;
	b = ((a & 0x5555555555555555ull) << 1) | ((a >> 1) & 0x5555555555555555ull);
	b = ((b & 0x3333333333333333ull) << 2) | ((b >> 2) & 0x3333333333333333ull);
	b = ((b & 0x0f0f0f0f0f0f0f0full) << 4) | ((b >> 4) & 0x0f0f0f0f0f0f0f0full);
	b = ((b & 0x00ff00ff00ff00ffull) << 8) | ((b >> 8) & 0x00ff00ff00ff00ffull);
	b = ((b & 0x0000ffff0000ffffull) << 16) | ((b >> 16) & 0x0000ffff0000ffffull);
	b = ((b & 0x00000000ffffffffull) << 32) | ((b >> 32) & 0x00000000ffffffffull);
	return b;
}

static inline uint64_t cov_nt4(uint8_t ch)
{
	switch (ch) {
	case 'A': case 'a': return 0;
	case 'C': case 'c': return 1;
	case 'G': case 'g': return 2;
	case 'T': case 't': return 3;
	default: return 4;
	}
}

#define COV_BASE(seq, i) ((uint64_t)(((seq)[(i) >> 4] >> (((i) & 15) << 1)) & 3u))

/* e_off[e] = first 32-bit word of edge e in e_seq; count_out[e] = the reference's g->edges[e].count after add_cnt_to_graph */
int ora_coverage_recount(const uint8_t *stream, uint64_t n, int64_t n_e, const uint32_t *e_len, const uint64_t *e_off,
			 const uint32_t *e_seq, const int64_t *e_rc, uint64_t *count_out)
{
	uint64_t n_km = 0;
	for (int64_t e = 0; e < n_e; ++e)
		if (e_len[e] >= COV_K + 1) n_km += e_len[e] - COV_K + 1;
	struct cov_tab t;
	uint64_t slots = 1024;
	while (slots < 2 * n_km + 16) slots <<= 1;
	t.mask = slots - 1;
	t.key = calloc(slots, 8);
	t.cnt = calloc(slots, 8);
	if (!t.key || !t.cnt) return -1;
	/* index_bin_edge: km = first 31 bases (get_km_i_bin, pad 0), then per window: OR the last base in, put, shift */
	for (int64_t e = 0; e < n_e; ++e) {
		if (e_len[e] < COV_K + 1) continue;
		const uint32_t *s = e_seq + e_off[e];
		uint64_t km = 0;
		for (int j = 0; j < COV_K; ++j) { km |= COV_BASE(s, j); km <<= 2; }
		for (uint32_t i = 0; i + COV_K <= e_len[e]; ++i) {
			km |= COV_BASE(s, i + COV_K - 1) << 2;
			cov_slot(&t, km, 1);
			km <<= 2;
		}
	}
	/* get_and_add_kmer over every read of >= 32 bases */
	uint64_t p = 0;
	while (p < n) {
		const uint8_t *nl = memchr(stream + p, '\n', n - p);
		const uint64_t len = nl ? (uint64_t)(nl - (stream + p)) : n - p;
		const uint8_t *r = stream + p;
		if (len >= COV_K + 1) {
			uint64_t km = 0;
			for (int j = 0; j < COV_K; ++j) { km |= cov_nt4(r[j]); km <<= 2; }
			for (uint64_t i = 0; i + COV_K <= len; ++i) {
				km |= cov_nt4(r[i + COV_K - 1]) << 2;
				const uint64_t rev = cov_bitrev(km) << 2;
				uint64_t *a = cov_slot(&t, km, 0);
				if (a) ++*a;
				a = cov_slot(&t, rev, 0);
				if (a) ++*a;
				km <<= 2;
			}
		}
		p += len + 1;
	}
	/* add_cnt_to_graph */
	for (int64_t e = 0; e < n_e; ++e) {
		count_out[e] = 0;
		if (e_len[e] < COV_K + 1) continue;
		const uint32_t *s = e_seq + e_off[e];
		uint64_t km = 0, sum = 0;
		for (int j = 0; j < COV_K; ++j) { km |= COV_BASE(s, j); km <<= 2; }
		for (uint32_t i = 0; i + COV_K <= e_len[e]; ++i) {
			km |= COV_BASE(s, i + COV_K - 1) << 2;
			const uint64_t *a = cov_slot(&t, km, 0);
			if (a) sum += *a < COV_MAX ? *a : COV_MAX;
			km <<= 2;
		}
		count_out[e] = sum;
	}
	for (int64_t e = 0; e < n_e; ++e)       /* in index order, like the reference (rc_id is an involution: order-free) */
		if (count_out[e_rc[e]] > count_out[e]) count_out[e] = count_out[e_rc[e]];
	free(t.key);
	free(t.cnt);
	return 0;
}
