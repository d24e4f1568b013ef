/*
 * ORACLE — TEST INFRASTRUCTURE ONLY. Never linked or called from the product path.
 *
 * Type-generic body of the CPU (k+1)-mer counter; included twice by kmc_cpu.c,
 * once with KEYT = uint64_t (K <= 32) and once with KEYT = unsigned __int128
 * (K <= 64).  SUF is the symbol suffix.
 *
 * Algorithm (no reference source exists for it because
 * the arithmetic lives in the absent libs/KMC/libkmc.a; in-tree corroboration:
 * /root/reference/src/k63_build.c:381-422 for the rolling window with reset on
 * non-ACGT, /root/reference/src/test_hash_count.c:24-72 for canonical-by-min):
 *   stream of bytes -> nt4 codes (src/utils.c:26-43); every run of >= K codes
 *   < 4 yields one window per position; key = min(fwd, rc) as integers with
 *   the first base most significant; multiset -> (key, count); keep count >= ci.
 *
 * Parallel shape mirrors KMC 2/3: phase 1 scatters keys into bins by their top
 * BIN_BITS bits, phase 2 radix-sorts + run-length-counts each bin; bins are in
 * key order, so concatenating them gives the sorted database.
 */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

struct FN(binvec) {
	KEYT *v;
	size_t n, cap;
};

struct FN(p1_arg) {
	struct ora_job *job;
	struct FN(binvec) *bins; /* [n_bins] private to the thread */
	uint64_t n_inst;
};

static inline void FN(bin_push)(struct FN(binvec) *b, KEYT x)
{
	if (b->n == b->cap) {
		b->cap = b->cap ? b->cap * 2 : 256;
		b->v = realloc(b->v, b->cap * sizeof(KEYT));
		if (!b->v) { fprintf(stderr, "oracle: out of memory\n"); exit(1); }
	}
	b->v[b->n++] = x;
}

static void *FN(p1_worker)(void *raw)
{
	struct FN(p1_arg) *a = raw;
	struct ora_job *job = a->job;
	const int K = job->K;
	const KEYT kmask = (K * 2 == (int)sizeof(KEYT) * 8) ? ~(KEYT)0 : (((KEYT)1 << (2 * K)) - 1);
	const int bin_shift = 2 * K - ORA_BIN_BITS;
	uint64_t n_inst = 0;
	for (;;) {
		size_t c = __sync_fetch_and_add(&job->next_chunk, 1);
		if (c >= job->n_chunks)
			break;
		size_t lo = c * ORA_CHUNK, hi = lo + ORA_CHUNK;
		if (hi > job->n)
			hi = job->n;
		/* windows ENDING in [lo, hi) belong to this chunk */
		size_t start = lo >= (size_t)(K - 1) ? lo - (K - 1) : 0;
		KEYT fw = 0, rv = 0;
		int run = 0;
		for (size_t p = start; p < hi; ++p) {
			int c4 = ora_nt4[job->seq[p]];
			if (c4 > 3) {
				run = 0;
				continue;
			}
			fw = ((fw << 2) | (KEYT)c4) & kmask;
			rv = (rv >> 2) | ((KEYT)(3 - c4) << (2 * (K - 1)));
			if (++run >= K && p >= lo) {
				KEYT key = fw <= rv ? fw : rv;
				FN(bin_push)(a->bins + (size_t)(key >> bin_shift), key);
				++n_inst;
			}
		}
	}
	a->n_inst = n_inst;
	return NULL;
}

struct FN(p2_arg) {
	struct ora_job *job;
	struct FN(p1_arg) *p1; /* [n_threads] */
	struct FN(binvec) *out_keys; /* [n_bins] */
	uint32_t **out_cnt; /* [n_bins] */
	uint64_t n_distinct;
};

static void FN(radix_sort)(KEYT *a, KEYT *tmp, size_t n, int n_bits)
{
	size_t cnt[256];
	KEYT *src = a, *dst = tmp;
	for (int sh = 0; sh < n_bits; sh += 8) {
		memset(cnt, 0, sizeof(cnt));
		for (size_t i = 0; i < n; ++i)
			++cnt[(unsigned)(src[i] >> sh) & 0xff];
		size_t s = 0;
		int trivial = 0;
		for (int d = 0; d < 256; ++d) {
			size_t t = cnt[d];
			if (t == n)
				trivial = 1;
			cnt[d] = s;
			s += t;
		}
		if (trivial)
			continue;
		for (size_t i = 0; i < n; ++i)
			dst[cnt[(unsigned)(src[i] >> sh) & 0xff]++] = src[i];
		KEYT *t2 = src; src = dst; dst = t2;
	}
	if (src != a)
		memcpy(a, src, n * sizeof(KEYT));
}

static void *FN(p2_worker)(void *raw)
{
	struct FN(p2_arg) *a = raw;
	struct ora_job *job = a->job;
	const int n_bins = 1 << ORA_BIN_BITS;
	uint64_t n_distinct = 0;
	for (;;) {
		int b = (int)__sync_fetch_and_add(&job->next_bin, 1);
		if (b >= n_bins)
			break;
		size_t tot = 0;
		for (int t = 0; t < job->n_threads; ++t)
			tot += a->p1[t].bins[b].n;
		if (!tot)
			continue;
		KEYT *buf = malloc(tot * sizeof(KEYT) * 2);
		if (!buf) { fprintf(stderr, "oracle: out of memory\n"); exit(1); }
		size_t o = 0;
		for (int t = 0; t < job->n_threads; ++t) {
			struct FN(binvec) *bv = a->p1[t].bins + b;
			memcpy(buf + o, bv->v, bv->n * sizeof(KEYT));
			o += bv->n;
			free(bv->v);
			bv->v = NULL;
			bv->n = bv->cap = 0;
		}
		FN(radix_sort)(buf, buf + tot, tot, 2 * job->K - ORA_BIN_BITS);
		/* run-length count, keep >= ci, in place at the front of buf */
		size_t w = 0, i = 0;
		uint32_t *cnt = malloc((tot + 1) * sizeof(uint32_t));
		while (i < tot) {
			size_t j = i + 1;
			while (j < tot && buf[j] == buf[i])
				++j;
			++n_distinct;
			if (j - i >= (size_t)job->ci) {
				buf[w] = buf[i];
				cnt[w] = (uint32_t)(j - i);
				++w;
			}
			i = j;
		}
		a->out_keys[b].v = realloc(buf, (w ? w : 1) * sizeof(KEYT));
		a->out_keys[b].n = w;
		a->out_cnt[b] = realloc(cnt, (w ? w : 1) * sizeof(uint32_t));
	}
	a->n_distinct = n_distinct;
	return NULL;
}

/* Counts the stream; returns malloc'ed sorted keys (as hi/lo pairs) + counts. */
static void FN(count_stream)(struct ora_job *job, struct ora_result *res)
{
	const int n_bins = 1 << ORA_BIN_BITS;
	const int T = job->n_threads;
	struct FN(p1_arg) *p1 = calloc(T, sizeof(*p1));
	pthread_t *th = calloc(T, sizeof(pthread_t));
	job->next_chunk = 0;
	job->n_chunks = (job->n + ORA_CHUNK - 1) / ORA_CHUNK;
	for (int t = 0; t < T; ++t) {
		p1[t].job = job;
		p1[t].bins = calloc(n_bins, sizeof(struct FN(binvec)));
		pthread_create(th + t, NULL, FN(p1_worker), p1 + t);
	}
	uint64_t n_inst = 0;
	for (int t = 0; t < T; ++t) {
		pthread_join(th[t], NULL);
		n_inst += p1[t].n_inst;
	}
	struct FN(binvec) *out_keys = calloc(n_bins, sizeof(*out_keys));
	uint32_t **out_cnt = calloc(n_bins, sizeof(uint32_t *));
	struct FN(p2_arg) *p2 = calloc(T, sizeof(*p2));
	job->next_bin = 0;
	for (int t = 0; t < T; ++t) {
		p2[t].job = job;
		p2[t].p1 = p1;
		p2[t].out_keys = out_keys;
		p2[t].out_cnt = out_cnt;
		pthread_create(th + t, NULL, FN(p2_worker), p2 + t);
	}
	uint64_t n_distinct = 0;
	for (int t = 0; t < T; ++t) {
		pthread_join(th[t], NULL);
		n_distinct += p2[t].n_distinct;
	}
	size_t n_solid = 0;
	for (int b = 0; b < n_bins; ++b)
		n_solid += out_keys[b].n;
	res->n_solid = n_solid;
	res->n_instances = n_inst;
	res->n_distinct = n_distinct;
	res->hi = malloc((n_solid ? n_solid : 1) * sizeof(uint64_t));
	res->lo = malloc((n_solid ? n_solid : 1) * sizeof(uint64_t));
	res->count = malloc((n_solid ? n_solid : 1) * sizeof(uint32_t));
	size_t o = 0;
	for (int b = 0; b < n_bins; ++b) {
		for (size_t i = 0; i < out_keys[b].n; ++i, ++o) {
			KEYT x = out_keys[b].v[i];
			res->lo[o] = (uint64_t)x;
			res->hi[o] = KEY_HI(x);
			res->count[o] = out_cnt[b][i];
		}
		free(out_keys[b].v);
		free(out_cnt[b]);
	}
	for (int t = 0; t < T; ++t)
		free(p1[t].bins);
	free(out_keys); free(out_cnt); free(p1); free(p2); free(th);
}

#undef FN
#undef CAT
#undef CAT_
