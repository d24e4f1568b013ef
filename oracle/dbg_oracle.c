/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into libtagpu.so.
 *
 * Plain-C restatement of the reference's graph stage: solid (k+1)-mers ->
 * 8-bit edge masks -> nodes -> unitig walk -> rc links -> edge counts.
 * Each step cites the reference lines it follows.  Lookup is a binary search
 * over sorted arrays instead of the reference's kmhash (slot order — and hence
 * node/edge numbering — is not a function of the input in the reference,
 * SURVEY.md §0.8, so parity is judged after canonical sorting; see canon_dump.c).
 *
 * Pinned against the unmodified reference (oracle/_ref/TA_ref build_0) by
 * tests/test_oracle_vs_ref.py and by the digests in tests/golden/.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ta_oracle.h"

typedef unsigned __int128 u128;

#define DEG4(e) (((e) & 1) + (((e) >> 1) & 1) + (((e) >> 2) & 1) + (((e) >> 3) & 1))
/* index of the only set bit of a nibble: /root/reference/src/kmer_build.c:23 */
#define ONLY4(e) ((((e) >> 1) & 1) * 1 + (((e) >> 2) & 1) * 2 + (((e) >> 3) & 1) * 3)

static inline u128 mask_of(int len) { return len >= 64 ? ~(u128)0 : (((u128)1 << (2 * len)) - 1); }

/* rc(x)_i = 3 - x_{len-1-i}: /root/reference/src/kmer.h:114-125 (km_get_rc), App. A.2 */
static u128 rc_of(u128 x, int len)
{
	u128 r = 0;
	for (int i = 0; i < len; ++i) {
		r = (r << 2) | (3 - (x & 3));
		x >>= 2;
	}
	return r;
}

struct kbit { u128 km; uint8_t bit; };

static int cmp_kbit(const void *a, const void *b)
{
	const struct kbit *x = a, *y = b;
	return x->km < y->km ? -1 : x->km > y->km;
}

struct solid { u128 key; uint32_t cnt; };
static int cmp_solid(const void *a, const void *b)
{
	const struct solid *x = a, *y = b;
	return x->key < y->key ? -1 : x->key > y->key;
}

static int64_t find_kmer(const struct ora_graph *g, u128 x)
{
	int64_t lo = 0, hi = g->n_kmer - 1;
	uint64_t xh = (uint64_t)(x >> 64), xl = (uint64_t)x;
	while (lo <= hi) {
		int64_t mid = (lo + hi) >> 1;
		if (g->khi[mid] < xh || (g->khi[mid] == xh && g->klo[mid] < xl))
			lo = mid + 1;
		else if (g->khi[mid] == xh && g->klo[mid] == xl)
			return mid;
		else
			hi = mid - 1;
	}
	return -1;
}

static int64_t find_solid(const struct solid *s, int64_t n, u128 x)
{
	int64_t lo = 0, hi = n - 1;
	while (lo <= hi) {
		int64_t mid = (lo + hi) >> 1;
		if (s[mid].key < x) lo = mid + 1;
		else if (s[mid].key == x) return mid;
		else hi = mid - 1;
	}
	return -1;
}

struct seqbuf { uint32_t *w; uint64_t n_words, cap; };

static uint64_t seq_reserve(struct seqbuf *sb, uint64_t words)
{
	uint64_t off = sb->n_words;
	if (off + words > sb->cap) {
		while (off + words > sb->cap)
			sb->cap = sb->cap ? sb->cap * 2 : 1024;
		sb->w = realloc(sb->w, sb->cap * sizeof(uint32_t));
	}
	memset(sb->w + off, 0, words * sizeof(uint32_t));
	sb->n_words += words;
	return off;
}

#define BINSEQ_GET(seq, i) (((seq)[(i) >> 4] >> (((i) & 15) << 1)) & 3u)

/* /root/reference/src/kmer_build.c:34-48 */
static int is_seq_rc(const uint32_t *s1, uint32_t l1, const uint32_t *s2, uint32_t l2)
{
	if (l1 != l2)
		return 0;
	for (uint32_t i = 0; i < l1; ++i)
		if (BINSEQ_GET(s1, i) != (BINSEQ_GET(s2, l1 - i - 1) ^ 3))
			return 0;
	return 1;
}

/* Graph of the solid (k+1)-mers plus, for build_local_assembly_graph (/root/reference/src/kmer_build.c:991-1044), the
 * "garbage" of n_contigs flanking contigs (2-bit codes, one byte per base): add_garbage (:847-888) applies App. A.4 to
 * every (k+1)-mer of a contig whether or not the reads support it — i.e. those (k+1)-mers join the edge-bearing set with
 * the count the reads gave them (0 if they are not solid) — and assign_count_garbage (:890-926) lifts the count of every
 * edge that carries a contig (k+1)-mer to the contig's coverage. */
static struct ora_graph *build_graph_impl(int k, int64_t n_solid_in, const uint64_t *hi, const uint64_t *lo, const uint32_t *count,
					  int n_contigs, const uint8_t *const *contig, const uint32_t *contig_len, const double *old_cov)
{
	struct ora_graph *g = calloc(1, sizeof(*g));
	const int K = k + 1;
	const u128 kmask = mask_of(k), Kmask = mask_of(K);
	g->ksize = k;

	int64_t n_garbage = 0;
	for (int c = 0; c < n_contigs; ++c)
		if (contig_len[c] > (uint32_t)k) n_garbage += contig_len[c] - k;
	int64_t n_solid = n_solid_in;
	struct solid *sol = malloc((n_solid + n_garbage ? n_solid + n_garbage : 1) * sizeof(*sol));
	for (int64_t i = 0; i < n_solid; ++i) {
		sol[i].key = ((u128)hi[i] << 64) | lo[i];
		sol[i].cnt = count[i];
	}
	qsort(sol, n_solid, sizeof(*sol), cmp_solid);
	if (n_garbage) {
		/* add_garbage: at base i >= k the pair (k-mer ending at i - 1, k-mer ending at i) gets the bits of the
		 * (k+1)-mer ending at i (:861-887); a (k+1)-mer the reads did not make solid enters with count 0 */
		int64_t n_all = n_solid;
		for (int c = 0; c < n_contigs; ++c) {
			u128 fw = 0, rv = 0;
			for (uint32_t i = 0; i < contig_len[c]; ++i) {
				const uint32_t b = contig[c][i] & 3;
				fw = ((fw << 2) | b) & Kmask;
				rv = (rv >> 2) | ((u128)(b ^ 3) << (2 * (K - 1)));
				if (i + 1 > (uint32_t)k) {                      /* :865 */
					const u128 x = fw <= rv ? fw : rv;
					if (find_solid(sol, n_solid, x) < 0) { sol[n_all].key = x; sol[n_all].cnt = 0; ++n_all; }
				}
			}
		}
		qsort(sol, n_all, sizeof(*sol), cmp_solid);
		int64_t o = 0;
		for (int64_t i = 0; i < n_all; ++i)
			if (i == 0 || sol[i].key != sol[o - 1].key) sol[o++] = sol[i];
		n_solid = o;
	}

	/* ---- masks: split_kmer_from_kedge_multi, /root/reference/src/kmer_build.c:78-129 (App. A.4) */
	struct kbit *kb = malloc((n_solid ? 2 * n_solid : 1) * sizeof(*kb));
	for (int64_t i = 0; i < n_solid; ++i) {
		u128 x = sol[i].key;
		u128 k1 = x >> 2, k2 = x & kmask;       /* kedge_get_left / kedge_get_right */
		int c1 = (int)(x & 3);                   /* :100 */
		int c2 = (int)((x >> (2 * k)) & 3) ^ 3;  /* :101 */
		u128 k1rc = rc_of(k1, k), k2rc = rc_of(k2, k);
		if (k1 <= k1rc) { kb[2 * i].km = k1; kb[2 * i].bit = c1; }          /* :110 */
		else { kb[2 * i].km = k1rc; kb[2 * i].bit = c1 + 4; }
		if (k2 <= k2rc) { kb[2 * i + 1].km = k2; kb[2 * i + 1].bit = c2 + 4; } /* :120 */
		else { kb[2 * i + 1].km = k2rc; kb[2 * i + 1].bit = c2; }
	}
	qsort(kb, 2 * n_solid, sizeof(*kb), cmp_kbit);
	int64_t nk = 0;
	for (int64_t i = 0; i < 2 * n_solid; ++i)
		if (i == 0 || kb[i].km != kb[i - 1].km)
			++nk;
	g->n_kmer = nk;
	g->khi = malloc((nk ? nk : 1) * sizeof(uint64_t));
	g->klo = malloc((nk ? nk : 1) * sizeof(uint64_t));
	g->mask = calloc(nk ? nk : 1, 1);
	nk = 0;
	for (int64_t i = 0; i < 2 * n_solid; ++i) {
		if (i == 0 || kb[i].km != kb[i - 1].km) {
			g->khi[nk] = (uint64_t)(kb[i].km >> 64);
			g->klo[nk] = (uint64_t)kb[i].km;
			++nk;
		}
		g->mask[nk - 1] |= (uint8_t)(1u << kb[i].bit);
	}
	free(kb);

	/* ---- nodes: build_asm_graph_from_kmhash, /root/reference/src/kmer_build.c:553-610 (App. A.5) */
	int64_t *ord = malloc((nk ? nk : 1) * sizeof(int64_t));
	int64_t n_nodes = 0, n_e = 0;
	for (int64_t i = 0; i < nk; ++i) {
		int df = DEG4(g->mask[i] & 15), dr = DEG4(g->mask[i] >> 4);
		if (df == 1 && dr == 1) { ord[i] = -1; continue; }
		ord[i] = n_nodes++;
		n_e += df + dr;
	}
	g->n_v = 2 * n_nodes;
	g->n_e = n_e;
	g->node_rc = malloc((g->n_v ? g->n_v : 1) * sizeof(int64_t));
	g->node_deg = malloc((g->n_v ? g->n_v : 1) * sizeof(int64_t));
	g->node_adj_off = malloc((g->n_v + 1) * sizeof(int64_t));
	g->node_adj = malloc((n_e ? n_e : 1) * sizeof(int64_t));
	g->e_src = malloc((n_e ? n_e : 1) * sizeof(int64_t));
	g->e_dst = malloc((n_e ? n_e : 1) * sizeof(int64_t));
	g->e_rc = malloc((n_e ? n_e : 1) * sizeof(int64_t));
	g->e_count = calloc(n_e ? n_e : 1, sizeof(uint64_t));
	g->e_len = malloc((n_e ? n_e : 1) * sizeof(uint32_t));
	g->e_seq_off = malloc((n_e ? n_e : 1) * sizeof(uint64_t));
	struct seqbuf sb = { NULL, 0, 0 };

	/* ---- unitig walk: build_graph_worker, /root/reference/src/kmer_build.c:421-542 (App. A.6) */
	int64_t e = 0;
	for (int64_t i = 0; i < nk; ++i) {
		if (ord[i] < 0)
			continue;
		u128 knum = ((u128)g->khi[i] << 64) | g->klo[i];
		u128 krev = rc_of(knum, k);
		for (int orient = 0; orient < 2; ++orient) {
			int64_t u = 2 * ord[i] + orient;
			uint8_t adj = orient ? (g->mask[i] >> 4) : (g->mask[i] & 15);
			g->node_rc[u] = u ^ 1;
			g->node_deg[u] = DEG4(adj);
			g->node_adj_off[u] = e;
			for (int c = 0; c < 4; ++c) {
				if (!((adj >> c) & 1))
					continue;
				u128 cur = orient ? krev : knum, cur_rc = orient ? knum : krev;
				/* growable per-edge base list */
				uint32_t len = 0, cap = 256;
				uint8_t *bases = malloc(cap);
				for (int b = 0; b < k; ++b) /* asm_init_edge :398-409: first base first */
					bases[len++] = (uint8_t)((cur >> (2 * (k - 1 - b))) & 3);
				int cur_c = c, d1, d2;
				int64_t j;
				do {
					if (len == cap) { cap *= 2; bases = realloc(bases, cap); }
					bases[len++] = (uint8_t)cur_c;                       /* :470 */
					cur = ((cur << 2) | (u128)cur_c) & kmask;             /* :471 km_shift_append */
					cur_rc = (cur_rc >> 2) | ((u128)(cur_c ^ 3) << (2 * (k - 1))); /* :472 */
					uint8_t m;
					if (cur <= cur_rc) {                                  /* :473 */
						j = find_kmer(g, cur);
						if (j < 0) { fprintf(stderr, "oracle: successor missing\n"); exit(1); }
						m = g->mask[j];
						d1 = DEG4(m & 15); d2 = DEG4(m >> 4);
						if (d1 == 1 && d2 == 1) cur_c = ONLY4(m & 15);    /* :481 */
					} else {
						j = find_kmer(g, cur_rc);
						if (j < 0) { fprintf(stderr, "oracle: successor missing\n"); exit(1); }
						m = g->mask[j];
						d1 = DEG4(m & 15); d2 = DEG4(m >> 4);
						if (d1 == 1 && d2 == 1) cur_c = ONLY4(m >> 4);    /* :490 */
					}
				} while (d1 == 1 && d2 == 1);
				g->e_src[e] = u;
				g->e_dst[e] = 2 * ord[j] + (cur <= cur_rc ? 0 : 1);       /* :494-497 */
				g->e_len[e] = len;
				g->e_seq_off[e] = seq_reserve(&sb, (len + 15) >> 4);
				for (uint32_t b = 0; b < len; ++b)                        /* __binseq_set, assembly_graph.h:182 */
					sb.w[g->e_seq_off[e] + (b >> 4)] |= (uint32_t)bases[b] << ((b & 15) << 1);
				free(bases);
				g->node_adj[e] = e;
				++e;
			}
		}
	}
	g->node_adj_off[g->n_v] = e;
	g->e_seq = sb.w ? sb.w : calloc(1, sizeof(uint32_t));

	/* ---- rc links: /root/reference/src/kmer_build.c:624-641 */
	for (e = 0; e < n_e; ++e)
		g->e_rc[e] = -1;
	for (e = 0; e < n_e; ++e) {
		int64_t v_rc = g->node_rc[g->e_dst[e]];
		for (int64_t a = g->node_adj_off[v_rc]; a < g->node_adj_off[v_rc + 1]; ++a) {
			int64_t e_rc = g->node_adj[a];
			if (g->e_dst[e_rc] == g->node_rc[g->e_src[e]] &&
			    is_seq_rc(g->e_seq + g->e_seq_off[e], g->e_len[e],
				      g->e_seq + g->e_seq_off[e_rc], g->e_len[e_rc])) {
				g->e_rc[e] = e_rc;
				g->e_rc[e_rc] = e;
				break;
			}
		}
		if (g->e_rc[e] < 0) { fprintf(stderr, "oracle: rc edge not found\n"); exit(1); }
	}

	/* ---- (k+1)-mer -> min(e, e_rc): build_edge_index_worker, /root/reference/src/kmer_build.c:202-242 */
	int64_t *idx = malloc((n_solid ? n_solid : 1) * sizeof(int64_t));
	for (int64_t i = 0; i < n_solid; ++i)
		idx[i] = -1;
	for (e = 0; e < n_e; ++e) {
		int64_t e_id = e > g->e_rc[e] ? g->e_rc[e] : e;
		const uint32_t *seq = g->e_seq + g->e_seq_off[e];
		u128 fw = 0, rv = 0;
		for (uint32_t b = 0; b < g->e_len[e]; ++b) {
			uint32_t c = BINSEQ_GET(seq, b);
			fw = ((fw << 2) | c) & Kmask;
			rv = (rv >> 2) | ((u128)(c ^ 3) << (2 * (K - 1)));
			if (b + 1 < (uint32_t)K)
				continue;
			int64_t s = find_solid(sol, n_solid, fw <= rv ? fw : rv);
			if (s < 0) { fprintf(stderr, "oracle: edge (k+1)-mer not solid\n"); exit(1); }
			if (idx[s] < 0)
				++g->n_kp1_on_edge;
			idx[s] = e_id;
		}
	}
	/* ---- counts: assign_count_kedge_multi, /root/reference/src/kmer_build.c:143-157 (App. A.7) */
	for (int64_t i = 0; i < n_solid; ++i) {
		if (idx[i] < 0)
			continue;
		g->e_count[idx[i]] += sol[i].cnt;
		g->e_count[g->e_rc[idx[i]]] += sol[i].cnt;
	}
	/* ---- assign_count_garbage(ksize + 1, ...), /root/reference/src/kmer_build.c:890-926, contig by contig in call order
	 * (:1040-1041).  Its loop tests i + 1 > ksize with ksize = k + 1, so the FIRST (k+1)-mer of a contig is never looked up. */
	for (int c = 0; c < n_contigs; ++c) {
		u128 fw = 0, rv = 0;
		for (uint32_t i = 0; i < contig_len[c]; ++i) {
			const uint32_t b = contig[c][i] & 3;
			fw = ((fw << 2) | b) & Kmask;
			rv = (rv >> 2) | ((u128)(b ^ 3) << (2 * (K - 1)));
			if (i + 1 > (uint32_t)K) {                              /* :905 */
				const int64_t s = find_solid(sol, n_solid, fw <= rv ? fw : rv);
				if (s < 0 || idx[s] < 0)
					continue;                               /* :913 not on any edge */
				const int64_t new_e = idx[s];
				const double new_cov = g->e_count[new_e] * 1.0 / (g->e_len[new_e] - (uint32_t)k);   /* __get_edge_cov, n_holes = 0 */
				if (new_cov < old_cov[c]) {                     /* :918 */
					const uint64_t v = (uint64_t)old_cov[c] * (g->e_len[new_e] - (uint32_t)K + 1);
					g->e_count[new_e] = g->e_count[g->e_rc[new_e]] = v;
				}
			}
		}
	}
	free(idx); free(ord); free(sol);
	return g;
}

struct ora_graph *ora_build_graph(int k, int64_t n_solid, const uint64_t *hi,
				  const uint64_t *lo, const uint32_t *count)
{
	return build_graph_impl(k, n_solid, hi, lo, count, 0, NULL, NULL, NULL);
}

struct ora_graph *ora_build_graph_local(int k, int64_t n_solid, const uint64_t *hi, const uint64_t *lo, const uint32_t *count,
					int n_contigs, const uint8_t *const *contig, const uint32_t *contig_len, const double *old_cov)
{
	return build_graph_impl(k, n_solid, hi, lo, count, n_contigs, contig, contig_len, old_cov);
}

void ora_graph_free(struct ora_graph *g)
{
	if (!g) return;
	free(g->khi); free(g->klo); free(g->mask);
	free(g->node_rc); free(g->node_deg); free(g->node_adj_off); free(g->node_adj);
	free(g->e_src); free(g->e_dst); free(g->e_rc); free(g->e_count); free(g->e_len);
	free(g->e_seq_off); free(g->e_seq);
	free(g);
}

/* save_asm_graph layout: /root/reference/src/assembly_graph.c:1173-1248 (App. C.1) */
int ora_graph_save_bin(const struct ora_graph *g, const char *path)
{
	FILE *fp = fopen(path, "wb");
	if (!fp) { perror(path); return -1; }
	setvbuf(fp, NULL, _IOFBF, 1 << 22);
	uint32_t aux_flag = 0;
	int32_t ksize = g->ksize;
	fwrite("asmg", 1, 4, fp);
	fwrite(&aux_flag, 4, 1, fp);
	fwrite(&ksize, 4, 1, fp);
	fwrite(&g->n_v, 8, 1, fp);
	fwrite(&g->n_e, 8, 1, fp);
	for (int64_t u = 0; u < g->n_v; ++u) {
		fwrite(&g->node_rc[u], 8, 1, fp);
		fwrite(&g->node_deg[u], 8, 1, fp);
		if (g->node_deg[u])
			fwrite(g->node_adj + g->node_adj_off[u], 8, g->node_deg[u], fp);
	}
	for (int64_t e = 0; e < g->n_e; ++e) {
		uint64_t len8 = g->e_len[e]; /* seq_len + aliased n_holes = 0, App. F.2 */
		uint32_t n_holes = 0;
		fwrite(&g->e_src[e], 8, 1, fp);
		fwrite(&g->e_dst[e], 8, 1, fp);
		fwrite(&g->e_rc[e], 8, 1, fp);
		fwrite(&g->e_count[e], 8, 1, fp);
		fwrite(&len8, 8, 1, fp);
		fwrite(g->e_seq + g->e_seq_off[e], 4, (g->e_len[e] + 15) >> 4, fp);
		fwrite(&n_holes, 4, 1, fp);
	}
	fclose(fp);
	return 0;
}
