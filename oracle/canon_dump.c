/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * Canonical form of an assembly-graph .bin (layout: save_asm_graph,
 * /root/reference/src/assembly_graph.c:1173-1248) so that two
 * graphs whose node/edge numbering differs can be compared byte for byte
 * (SURVEY.md App. D.3).  Also re-checks the structural invariants that the
 * reference's test_asm_graph (/root/reference/src/assembly_graph.c:987-1171)
 * enforces, in our own words.
 *
 * mode 0: one line per edge with e <= rc_id(e): min(seq, rc(seq)) \t count \t seq_len
 *         lines sorted bytewise, joined with '\n' (no trailing newline).
 * mode 1: one line per edge: oriented source k-mer \t oriented target k-mer \t
 *         first appended base \t seq_len \t count, sorted the same way.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ta_oracle.h"

struct cg_edge {
	int64_t src, dst, rc;
	uint64_t count;
	uint32_t len, n_holes;
	const uint32_t *seq;
};

#define BINSEQ_GET(seq, i) (((seq)[(i) >> 4] >> (((i) & 15) << 1)) & 3u)

static int cmp_str(const void *a, const void *b)
{
	return strcmp(*(char *const *)a, *(char *const *)b);
}

static void edge_string(const struct cg_edge *e, char *s, int revcomp)
{
	static const char nt[4] = { 'A', 'C', 'G', 'T' };
	for (uint32_t i = 0; i < e->len; ++i)
		s[i] = revcomp ? nt[BINSEQ_GET(e->seq, e->len - 1 - i) ^ 3] : nt[BINSEQ_GET(e->seq, i)];
	s[e->len] = 0;
}

#define FAIL(...) do { if (n_bad < 20) { fprintf(stderr, "canon_dump: " __VA_ARGS__); fputc('\n', stderr); } ++n_bad; } while (0)

int ora_canon_dump(const char *bin_path, const char *out_path, int mode)
{
	FILE *fp = fopen(bin_path, "rb");
	if (!fp) { perror(bin_path); return -1; }
	fseek(fp, 0, SEEK_END);
	long fsz = ftell(fp);
	fseek(fp, 0, SEEK_SET);
	uint8_t *buf = malloc(fsz + 8);
	if (fread(buf, 1, fsz, fp) != (size_t)fsz) { fclose(fp); return -1; }
	fclose(fp);
	int n_bad = 0;
	if (fsz < 28 || memcmp(buf, "asmg", 4)) { fprintf(stderr, "canon_dump: bad magic\n"); return -2; }
	const uint8_t *p = buf + 4;
	uint32_t aux_flag; memcpy(&aux_flag, p, 4); p += 4;
	int32_t k; memcpy(&k, p, 4); p += 4;
	int64_t n_v, n_e; memcpy(&n_v, p, 8); p += 8; memcpy(&n_e, p, 8); p += 8;
	if (aux_flag) FAIL("aux_flag = %u at level 0", aux_flag);

	int64_t *n_rc = malloc((n_v + 1) * 8), *n_deg = malloc((n_v + 1) * 8);
	const uint8_t **n_adj = malloc((n_v + 1) * sizeof(*n_adj));
	for (int64_t u = 0; u < n_v; ++u) {
		memcpy(&n_rc[u], p, 8); p += 8;
		memcpy(&n_deg[u], p, 8); p += 8;
		n_adj[u] = p;
		p += 8 * n_deg[u];
		if (p > buf + fsz) { fprintf(stderr, "canon_dump: truncated nodes\n"); return -2; }
	}
	struct cg_edge *E = calloc(n_e + 1, sizeof(*E));
	uint64_t sum_len_minus_k = 0;
	for (int64_t e = 0; e < n_e; ++e) {
		memcpy(&E[e].src, p, 8); p += 8;
		memcpy(&E[e].dst, p, 8); p += 8;
		if (E[e].src == -1) { FAIL("removed edge %ld at level 0", (long)e); continue; }
		memcpy(&E[e].rc, p, 8); p += 8;
		memcpy(&E[e].count, p, 8); p += 8;
		memcpy(&E[e].len, p, 4); p += 4;
		uint32_t alias; memcpy(&alias, p, 4); p += 4;
		E[e].seq = (const uint32_t *)p; /* 4-byte aligned: every field is a multiple of 4 */
		p += 4 * ((E[e].len + 15) >> 4);
		memcpy(&E[e].n_holes, p, 4); p += 4;
		if (alias || E[e].n_holes) FAIL("edge %ld has holes at level 0", (long)e);
		if (p > buf + fsz) { fprintf(stderr, "canon_dump: truncated edges\n"); return -2; }
		sum_len_minus_k += E[e].len - k;
	}
	if (p != buf + fsz) FAIL("%ld trailing bytes", (long)(buf + fsz - p));

	/* --- invariants (same properties as test_asm_graph, restated) --- */
	for (int64_t u = 0; u < n_v; ++u) {
		if (n_rc[u] < 0 || n_rc[u] >= n_v || n_rc[n_rc[u]] != u || n_rc[u] == u)
			FAIL("node %ld rc link broken", (long)u);
		for (int64_t j = 0; j < n_deg[u]; ++j) {
			int64_t e; memcpy(&e, n_adj[u] + 8 * j, 8);
			if (e < 0 || e >= n_e || E[e].src != u) { FAIL("node %ld adj[%ld] -> edge %ld not sourced here", (long)u, (long)j, (long)e); continue; }
			int64_t e0; memcpy(&e0, n_adj[u], 8);
			for (int b = 0; b < k; ++b)
				if (BINSEQ_GET(E[e].seq, b) != BINSEQ_GET(E[e0].seq, b)) { FAIL("node %ld out-edges disagree on k-prefix", (long)u); break; }
		}
		if (n_rc[u] >= 0 && n_rc[u] < n_v && n_deg[u] + n_deg[n_rc[u]] == 0)
			FAIL("isolated node %ld", (long)u);
	}
	int64_t *seen = calloc(n_e + 1, 8);
	for (int64_t u = 0; u < n_v; ++u)
		for (int64_t j = 0; j < n_deg[u]; ++j) {
			int64_t e; memcpy(&e, n_adj[u] + 8 * j, 8);
			if (e >= 0 && e < n_e) ++seen[e];
		}
	for (int64_t e = 0; e < n_e; ++e) {
		struct cg_edge *x = E + e;
		if (x->src < 0 || x->src >= n_v || x->dst < 0 || x->dst >= n_v) { FAIL("edge %ld endpoint out of range", (long)e); continue; }
		if (seen[e] != 1) FAIL("edge %ld listed %ld times in adj", (long)e, (long)seen[e]);
		if (x->len < (uint32_t)k + 1) FAIL("edge %ld shorter than k+1", (long)e);
		if (x->rc < 0 || x->rc >= n_e || E[x->rc].rc != e) { FAIL("edge %ld rc link broken", (long)e); continue; }
		struct cg_edge *y = E + x->rc;
		if (y->src != n_rc[x->dst] || y->dst != n_rc[x->src]) FAIL("edge %ld rc endpoints asymmetric", (long)e);
		if (y->count != x->count) FAIL("edge %ld count differs from rc", (long)e);
		if (y->len != x->len) { FAIL("edge %ld len differs from rc", (long)e); continue; }
		for (uint32_t b = 0; b < x->len; ++b)
			if (BINSEQ_GET(x->seq, b) != (BINSEQ_GET(y->seq, x->len - 1 - b) ^ 3)) { FAIL("edge %ld seq is not rc of its twin", (long)e); break; }
		/* consecutive edges overlap by k */
		if (n_deg[x->dst]) {
			int64_t f; memcpy(&f, n_adj[x->dst], 8);
			if (f >= 0 && f < n_e)
				for (int b = 0; b < k; ++b)
					if (BINSEQ_GET(x->seq, x->len - k + b) != BINSEQ_GET(E[f].seq, b)) { FAIL("edge %ld does not overlap its target by k", (long)e); break; }
		}
		if ((x->len & 15) && (x->seq[x->len >> 4] >> ((x->len & 15) << 1))) FAIL("edge %ld has garbage above seq_len", (long)e);
	}

	/* --- canonical lines --- */
	int64_t n_lines = 0;
	char **lines = malloc((n_e + 1) * sizeof(char *));
	for (int64_t e = 0; e < n_e; ++e) {
		struct cg_edge *x = E + e;
		if (x->src == -1) continue;
		if (mode == 0) {
			if (e > x->rc) continue;
			char *s = malloc(2 * (size_t)x->len + 64), *r = s + x->len + 1;
			edge_string(x, s, 0);
			edge_string(x, r, 1);
			if (strcmp(r, s) < 0) memcpy(s, r, x->len);
			sprintf(s + x->len, "\t%lu\t%u", (unsigned long)x->count, x->len);
			lines[n_lines++] = s;
		} else {
			char *full = malloc((size_t)x->len + 1);
			edge_string(x, full, 0);
			char *s = malloc(2 * (size_t)k + 80);
			memcpy(s, full, k); s[k] = '\t';
			memcpy(s + k + 1, full + x->len - k, k);
			sprintf(s + 2 * k + 1, "\t%c\t%u\t%lu", full[k], x->len, (unsigned long)x->count);
			free(full);
			lines[n_lines++] = s;
		}
	}
	qsort(lines, n_lines, sizeof(char *), cmp_str);
	FILE *fo = fopen(out_path, "wb");
	if (!fo) { perror(out_path); return -1; }
	setvbuf(fo, NULL, _IOFBF, 1 << 22);
	for (int64_t i = 0; i < n_lines; ++i) {
		if (i) fputc('\n', fo);
		fputs(lines[i], fo);
		free(lines[i]);
	}
	fclose(fo);
	fprintf(stderr, "canon_dump: k=%d n_v=%ld n_e=%ld lines=%ld sum(len-k)=%lu bad=%d\n",
		k, (long)n_v, (long)n_e, (long)n_lines, (unsigned long)sum_len_minus_k, n_bad);
	free(lines); free(seen); free(E); free(n_adj); free(n_rc); free(n_deg); free(buf);
	return n_bad;
}

/* ------------------------------------------------------------------ order-independent digest of a graph .bin
 * CPU restatement of the edge digest libtagpu computes on the device (csrc/tagpu_digest.cuh), evaluated on a .bin in
 * the save_asm_graph layout — e.g. the reference's own graph_k_<k>_level_0.bin — so that a full-size GPU result can be
 * compared with the reference without sorting 10^5..10^7 unitigs.  out = { sum, xor, sum of lengths, sum of counts, n_e }. */
static uint64_t dg_mix64(uint64_t x)
{
	x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
	x ^= x >> 27; x *= 0x94d049bb133111ebull;
	x ^= x >> 31;
	return x;
}

int ora_bin_digest(const char *bin_path, uint64_t out[5])
{
	FILE *fp = fopen(bin_path, "rb");
	if (!fp) { perror(bin_path); return -1; }
	fseek(fp, 0, SEEK_END);
	long fsz = ftell(fp);
	fseek(fp, 0, SEEK_SET);
	uint8_t *buf = malloc(fsz + 8);
	if (!buf || fread(buf, 1, fsz, fp) != (size_t)fsz) { fclose(fp); free(buf); return -1; }
	fclose(fp);
	if (fsz < 28 || memcmp(buf, "asmg", 4)) { free(buf); return -2; }
	const uint8_t *p = buf + 12, *end = buf + fsz;
	int64_t n_v, n_e;
	memcpy(&n_v, p, 8); p += 8;
	memcpy(&n_e, p, 8); p += 8;
	for (int64_t u = 0; u < n_v && p + 16 <= end; ++u) {
		int64_t deg;
		memcpy(&deg, p + 8, 8);
		p += 16 + 8 * deg;
	}
	uint64_t sum = 0, x = 0, tl = 0, tc = 0, live = 0;
	for (int64_t e = 0; e < n_e; ++e) {
		if (p + 16 > end) { free(buf); return -2; }
		int64_t src;
		memcpy(&src, p, 8);
		p += 16;
		if (src == -1) continue;
		uint64_t count, len8;
		memcpy(&count, p + 8, 8);
		memcpy(&len8, p + 16, 8);
		p += 24;
		const uint32_t len = (uint32_t)len8, nw = (len + 15) >> 4;
		if (p + 4 * (size_t)nw + 4 > end) { free(buf); return -2; }
		uint64_t hw = 0;
		for (uint32_t i = 0; i < nw; ++i) {
			uint32_t w;
			memcpy(&w, p + 4 * (size_t)i, 4);
			hw += dg_mix64((uint64_t)w ^ (0x9E3779B97F4A7C15ull * (uint64_t)(i + 1)));
		}
		p += 4 * (size_t)nw;
		uint32_t n_holes;
		memcpy(&n_holes, p, 4);
		p += 4 + 8 * (size_t)n_holes;
		const uint64_t d = dg_mix64(hw ^ dg_mix64(((uint64_t)len << 32) ^ (count * 0xC2B2AE3D27D4EB4Full)));
		sum += d; x ^= d; tl += len; tc += count; ++live;
	}
	free(buf);
	out[0] = sum; out[1] = x; out[2] = tl; out[3] = tc; out[4] = live;
	return 0;
}
