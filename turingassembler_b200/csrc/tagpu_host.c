/*
 * libtagpu host layer (plain C, like the reference): FASTQ/FASTA ingest, materialisation of the
 * reference's struct asm_graph_t, the graph .bin and KMC database writers, and the reference's own
 * entry points (see include/tagpu.h for the file:line of each interface replaced).
 *
 * No CPU fallback lives here: every compute step is a call into the CUDA layer (tagpu_kernels.cu).
 * Errors follow the reference convention for these void entry points: log with file:line and
 * exit(1) (/root/reference/src/log.c:207-210).
 */
#define _GNU_SOURCE
#include <fcntl.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>
#include <zlib.h>

#include "../../include/tagpu.h"
#include "../../include/tagpu_graph.h"

void *tagpu_pinned_alloc(size_t bytes);
void tagpu_pinned_free(void *p);
int tagpu_ctx_k(tagpu_ctx *ctx);
int tagpu_ctx_K(tagpu_ctx *ctx);
int tagpu_ctx_cutoff(tagpu_ctx *ctx);
void tagpu_set_source_progress(tagpu_ctx *ctx, uint64_t (*ready)(void *), void *arg);

#define TAGPU_FATAL(...)                                                        \
	do {                                                                    \
		fprintf(stderr, "[tagpu] FATAL %s:%d: ", __FILE__, __LINE__);  \
		fprintf(stderr, __VA_ARGS__);                                   \
		fputc('\n', stderr);                                            \
		exit(1);                                                        \
	} while (0)

static double now_s(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* ------------------------------------------------------------------ read ingest */

struct read_file {
	const char *path;
	uint8_t *txt;      /* whole (inflated) file */
	size_t n_txt;
	size_t n_seq;      /* bytes this file contributes to the stream */
	uint8_t *dst;      /* where they go */
};

static void slurp_file(struct read_file *f)
{
	int fd = open(f->path, O_RDONLY);
	if (fd < 0)
		TAGPU_FATAL("cannot open %s", f->path);
	unsigned char magic[2] = { 0, 0 };
	if (read(fd, magic, 2) < 0)
		TAGPU_FATAL("cannot read %s", f->path);
	lseek(fd, 0, SEEK_SET);
	size_t n = 0, cap;
	uint8_t *buf;
	if (magic[0] == 0x1f && magic[1] == 0x8b) {
		gzFile gz = gzdopen(fd, "rb");
		gzbuffer(gz, 1 << 20);
		cap = (size_t)1 << 26;
		buf = malloc(cap);
		for (;;) {
			if (n == cap) {
				cap *= 2;
				buf = realloc(buf, cap);
			}
			if (!buf)
				TAGPU_FATAL("out of host memory reading %s", f->path);
			size_t room = cap - n;
			int r = gzread(gz, buf + n, room > ((size_t)1 << 30) ? (1u << 30) : (unsigned)room);
			if (r <= 0)
				break;
			n += (size_t)r;
		}
		gzclose(gz);
	} else {
		struct stat st;
		fstat(fd, &st);
		cap = (size_t)st.st_size + 1;
		buf = malloc(cap);
		if (!buf)
			TAGPU_FATAL("out of host memory reading %s", f->path);
		while (n < (size_t)st.st_size) {
			ssize_t r = read(fd, buf + n, (size_t)st.st_size - n);
			if (r <= 0)
				break;
			n += (size_t)r;
		}
		close(fd);
	}
	f->txt = buf;
	f->n_txt = n;
}

/* Walks the sequence lines of a FASTQ (line 2 of every 4, cf. /root/reference/src/get_buffer.c:339-348)
 * or FASTA text; with dst == NULL only sizes the output. */
static size_t walk_sequences(const uint8_t *txt, size_t n, uint8_t *dst)
{
	size_t o = 0, p = 0, line = 0;
	const int fasta = n && txt[0] == '>';
	int open_record = 0; /* FASTA: sequence lines of one record are joined */
	while (p < n) {
		const uint8_t *nl = memchr(txt + p, '\n', n - p);
		size_t e = nl ? (size_t)(nl - txt) : n;
		size_t len = e - p;
		if (len && txt[e - 1] == '\r')
			--len;
		if (fasta) {
			if (len && txt[p] == '>') {
				if (open_record) {
					if (dst) dst[o] = '\n';
					++o;
					open_record = 0;
				}
			} else if (len) {
				if (dst) memcpy(dst + o, txt + p, len);
				o += len;
				open_record = 1;
			}
		} else if ((line & 3) == 1) {
			if (dst) {
				memcpy(dst + o, txt + p, len);
				dst[o + len] = '\n';
			}
			o += len + 1;
		}
		++line;
		p = e + 1;
	}
	if (open_record) {
		if (dst) dst[o] = '\n';
		++o;
	}
	return o;
}

/* The pinned stream buffer is kept across calls (callers never run concurrently, SURVEY.md §8b "Threading"). */
static uint8_t *g_stream_buf;
static size_t g_stream_cap;
static int g_stream_busy;

static uint8_t *stream_buffer(size_t bytes)
{
	if (!g_stream_busy && g_stream_buf && g_stream_cap >= bytes) {
		g_stream_busy = 1;
		return g_stream_buf;
	}
	uint8_t *s = tagpu_pinned_alloc(bytes);
	if (!s)
		TAGPU_FATAL("cannot allocate %zu bytes of pinned host memory", bytes);
	if (!g_stream_busy) {
		tagpu_pinned_free(g_stream_buf);
		g_stream_buf = s;
		g_stream_cap = bytes;
		g_stream_busy = 1;
	}
	return s;
}

/* ---- tiny task pool: n_tasks independent tasks on up to n_threads pthreads */
struct task_pool {
	void (*fn)(size_t task, void *arg);
	void *arg;
	size_t n_tasks;
	size_t next; /* __sync counter */
};

static void *task_pool_worker(void *raw)
{
	struct task_pool *tp = raw;
	for (;;) {
		size_t t = __sync_fetch_and_add(&tp->next, 1);
		if (t >= tp->n_tasks)
			return NULL;
		tp->fn(t, tp->arg);
	}
}

static void run_tasks(size_t n_tasks, int n_threads, void (*fn)(size_t, void *), void *arg)
{
	struct task_pool tp = { fn, arg, n_tasks, 0 };
	if (n_threads > 64) n_threads = 64;
	if ((size_t)n_threads > n_tasks) n_threads = (int)n_tasks;
	if (n_threads <= 1) {
		task_pool_worker(&tp);
		return;
	}
	pthread_t th[64];
	for (int i = 0; i < n_threads; ++i)
		pthread_create(th + i, NULL, task_pool_worker, &tp);
	for (int i = 0; i < n_threads; ++i)
		pthread_join(th[i], NULL);
}

/* ---- plain (uncompressed) FASTQ, exact and parallel.  The file is cut into chunks.  Pass 1 (one task per chunk) builds
 * the chunk's newline index — the offsets of its '\n' bytes, found 32 bytes at a time — so the text is scanned ONCE; a
 * prefix sum over the counts gives every chunk the number of the line it starts in.  Sizing then only walks the index
 * (sequence = line 2 of every 4, /root/reference/src/get_buffer.c:339-348 — no record-boundary guessing), and pass 2
 * copies the sequence lines, chunk by chunk in stream order, while the caller may already upload the finished prefix
 * (tagpu_ingest_ready). */
#define INGEST_CHUNK ((size_t)4 << 20)

struct pfq {
	struct read_file *f;
	size_t n_chunks;
	int mapped;        /* txt is an mmap of the file */
	int fd;            /* kept open: the fused ingest preads chunks into worker-private buffers */
	size_t *nl;        /* newlines in chunk, then: line number of the chunk's first byte */
	size_t *n_idx;     /* entries of idx[c] */
	uint32_t **idx;    /* newline offsets relative to the chunk start */
	size_t *out;       /* sequence bytes of the chunk, then: their offset in dst */
	unsigned char *has_cr; /* the chunk holds a '\r' somewhere: its lines need the CRLF checks */
};

static void pfq_read(size_t c, void *raw)
{
	struct pfq *p = raw;
	size_t lo = c * INGEST_CHUNK, hi = lo + INGEST_CHUNK < p->f->n_txt ? lo + INGEST_CHUNK : p->f->n_txt;
	int fd = open(p->f->path, O_RDONLY);
	if (fd < 0)
		TAGPU_FATAL("cannot open %s", p->f->path);
	while (lo < hi) {
		ssize_t r = pread(fd, p->f->txt + lo, hi - lo, (off_t)lo);
		if (r <= 0)
			TAGPU_FATAL("cannot read %s", p->f->path);
		lo += (size_t)r;
	}
	close(fd);
}

static uint32_t *idx_room(uint32_t *v, size_t n, size_t *cap, size_t want)
{
	if (n + want <= *cap)
		return v;
	while (n + want > *cap)
		*cap = *cap ? *cap * 2 : 65536;
	v = realloc(v, *cap * sizeof(uint32_t));
	if (!v)
		TAGPU_FATAL("out of host memory for the newline index");
	return v;
}

#if defined(__x86_64__)
#include <immintrin.h>
__attribute__((target("avx2")))
static uint32_t *scan_newlines_avx2(const uint8_t *t, size_t n, uint32_t *v, size_t *n_out, size_t *cap, int *has_cr)
{
	const __m256i nl = _mm256_set1_epi8('\n'), cr = _mm256_set1_epi8('\r');
	__m256i any_cr = _mm256_setzero_si256();
	size_t i = 0, k = *n_out;
	for (; i + 32 <= n; i += 32) {
		const __m256i x = _mm256_loadu_si256((const __m256i *)(t + i));
		unsigned m = (unsigned)_mm256_movemask_epi8(_mm256_cmpeq_epi8(x, nl));
		any_cr = _mm256_or_si256(any_cr, _mm256_cmpeq_epi8(x, cr));
		if (!m)
			continue;
		v = idx_room(v, k, cap, 32);
		while (m) {
			v[k++] = (uint32_t)(i + (size_t)__builtin_ctz(m));
			m &= m - 1;
		}
	}
	*has_cr = _mm256_movemask_epi8(any_cr) != 0;
	for (; i < n; ++i) {
		if (t[i] == '\r')
			*has_cr = 1;
		if (t[i] == '\n') {
			v = idx_room(v, k, cap, 1);
			v[k++] = (uint32_t)i;
		}
	}
	*n_out = k;
	return v;
}
#endif

/* newline offsets of t[0, n); *has_cr = 0 only if no '\r' occurs in it (then no line of the chunk needs a CRLF check) */
static uint32_t *scan_newlines(const uint8_t *t, size_t n, size_t *n_out, int *has_cr)
{
	uint32_t *v = NULL;
	size_t cap = 0, k = 0;
#if defined(__x86_64__)
	if (__builtin_cpu_supports("avx2")) {
		v = scan_newlines_avx2(t, n, v, &k, &cap, has_cr);
		*n_out = k;
		return v;
	}
#endif
	*has_cr = 1;
	size_t lo = 0;
	while (lo < n) {
		const uint8_t *q = memchr(t + lo, '\n', n - lo);
		if (!q)
			break;
		v = idx_room(v, k, &cap, 1);
		v[k++] = (uint32_t)(q - t);
		lo = (size_t)(q - t) + 1;
	}
	*n_out = k;
	return v;
}

/* ct = the chunk's text (ct[0] = byte c * INGEST_CHUNK of the file) */
static void pfq_index_at(struct pfq *p, size_t c, const uint8_t *ct)
{
	const size_t lo = c * INGEST_CHUNK, hi = lo + INGEST_CHUNK < p->f->n_txt ? lo + INGEST_CHUNK : p->f->n_txt;
	int has_cr;
	p->idx[c] = scan_newlines(ct, hi - lo, &p->n_idx[c], &has_cr);
	p->has_cr[c] = (unsigned char)has_cr;
	p->nl[c] = p->n_idx[c];
}

static void pfq_index(size_t c, void *raw)
{
	struct pfq *p = raw;
	pfq_index_at(p, c, p->f->txt + c * INGEST_CHUNK);
}

/* bytes [lo, hi) of the chunk that lie on sequence lines; with dst != NULL they are copied.  Lines are taken from the
 * chunk's newline index.  A '\r' right before the line's '\n' is dropped; a last sequence line without a newline gets one. */
static size_t pfq_walk_at(struct pfq *p, size_t c, uint8_t *dst, const uint8_t *ct);

static size_t pfq_walk(struct pfq *p, size_t c, uint8_t *dst)
{
	return pfq_walk_at(p, c, dst, p->f->txt + c * INGEST_CHUNK);
}

/* ct = the chunk's text plus ONE byte of look-ahead when the file goes on (ct[0] = byte c * INGEST_CHUNK of the file) */
static size_t pfq_walk_at(struct pfq *p, size_t c, uint8_t *dst, const uint8_t *ct)
{
	const size_t n = p->f->n_txt, base = c * INGEST_CHUNK;
	const uint8_t *t = ct - base;                            /* so that t[file offset] addresses the chunk's bytes */
	const uint32_t *ix = p->idx[c];
	const size_t n_ix = p->n_idx[c];
	size_t lo = base, hi = lo + INGEST_CHUNK < n ? lo + INGEST_CHUNK : n, line = p->nl[c], o = 0;
	const int cr = p->has_cr[c];
	/* segment i of the chunk = [previous newline + 1, newline i), the last one runs to the end of the chunk; only the
	 * segments on sequence lines are looked at: i == (1 - line) mod 4 */
	for (size_t i = (5 - (line & 3)) & 3; i <= n_ix; i += 4) {
		const size_t s = i ? base + ix[i - 1] + 1 : lo;
		const int has_nl = i < n_ix;
		const size_t e = has_nl ? base + ix[i] : hi;         /* end of this line's bytes inside the chunk */
		if (s >= hi && !has_nl)
			break;                                           /* the chunk ends exactly behind a newline */
		size_t len = e - s;
		const int ends_line = has_nl || e == n;              /* the line really ends at e (not just the chunk) */
		if (!cr)
			;                                                /* no '\r' in this chunk: nothing to strip */
		else if (len && ends_line && t[e - 1] == '\r')
			--len;
		else if (len && !ends_line && e == hi && t[e - 1] == '\r' && hi < n && t[hi] == '\n')
			--len;                                           /* "\r" | "\n" split by the chunk border */
		if (dst) memcpy(dst + o, t + s, len);
		o += len;
		if (ends_line) {                                     /* terminate the read (also when the file lacks the last newline) */
			if (dst) dst[o] = '\n';
			++o;
		}
	}
	return o;
}

static void pfq_size(size_t c, void *raw)
{
	struct pfq *p = raw;
	p->out[c] = pfq_walk(p, c, NULL);
}

/* returns 1 and fills f->txt / n_txt / n_seq (+ the per-chunk tables in *p) if the file is plain FASTQ */
static int pfq_open(struct read_file *f, int n_threads, struct pfq *p)
{
	int fd = open(f->path, O_RDONLY);
	if (fd < 0)
		TAGPU_FATAL("cannot open %s", f->path);
	unsigned char magic[2] = { 0, 0 };
	struct stat st;
	if (pread(fd, magic, 2, 0) < 1 || fstat(fd, &st) != 0 || !S_ISREG(st.st_mode) || magic[0] != '@') {
		close(fd);
		return 0;                                                /* gzip, FASTA, pipe, empty: serial path */
	}
	memset(p, 0, sizeof(*p));
	p->f = f;
	f->n_txt = (size_t)st.st_size;
	p->n_chunks = (f->n_txt + INGEST_CHUNK - 1) / INGEST_CHUNK;
	p->nl = calloc(p->n_chunks + 1, sizeof(size_t));
	p->n_idx = calloc(p->n_chunks + 1, sizeof(size_t));
	p->idx = calloc(p->n_chunks + 1, sizeof(uint32_t *));
	p->out = calloc(p->n_chunks + 1, sizeof(size_t));
	p->has_cr = calloc(p->n_chunks + 1, 1);
	/* map the file (page-cache pages, no copy); fall back to reading it into memory */
	void *m = mmap(NULL, f->n_txt, PROT_READ, MAP_PRIVATE, fd, 0);
	p->fd = fd;
	if (m != MAP_FAILED) {
		f->txt = m;
		p->mapped = 1;
	} else {
		f->txt = malloc(f->n_txt + 1);
		if (!f->txt)
			TAGPU_FATAL("out of host memory reading %s", f->path);
		run_tasks(p->n_chunks, n_threads, pfq_read, p);
	}
	return 1;
}

static void pfq_close(struct pfq *p)
{
	close(p->fd);
	if (p->mapped) munmap(p->f->txt, p->f->n_txt);
	else free(p->f->txt);
	for (size_t c = 0; c < p->n_chunks; ++c)
		free(p->idx[c]);
	free(p->idx);
	free(p->n_idx);
	free(p->nl);
	free(p->out);
	free(p->has_cr);
}

static void *ingest_phase1(void *raw)
{
	struct read_file *f = raw;
	slurp_file(f);
	f->n_seq = walk_sequences(f->txt, f->n_txt, NULL);
	return NULL;
}

static void *ingest_phase2(void *raw)
{
	struct read_file *f = raw;
	walk_sequences(f->txt, f->n_txt, f->dst);
	free(f->txt);
	f->txt = NULL;
	return NULL;
}

/* ------------------------------------------------------------------ packed read stream (include/tagpu.h)
 * Tile t = stream positions [8192 t, 8192 t + 8192): 256 64-bit code words (32 bases each, first base most significant,
 * A=0 C=1 G=2 T=3) followed by 256 32-bit invalid masks (bit 31 = first position of the word; set for every byte that
 * is not A/C/G/T/a/c/g/t and for the positions past the end of the stream).  This is byte for byte what the CUDA tile
 * loader builds in shared memory from an ASCII stream (tagpu_pack4 in csrc/tagpu_extract.cuh — including the don't-care
 * code bits it leaves under invalid bytes), done once on the host so that 0.375 bytes per base cross PCIe instead of 1. */
#define PACK_TILE_WORDS 256
#define PACK_TILE_BASES (PACK_TILE_WORDS * 32)
#define PACK_TILE_BYTES (PACK_TILE_WORDS * 12)
#define PACK_TASK_TILES 64

static uint8_t pack_code[256];     /* bits 1..0: code, bit 2: invalid */
static pthread_once_t pack_once = PTHREAD_ONCE_INIT;
static void pack_init(void)
{
	for (int b = 0; b < 256; ++b) {
		unsigned c = ((unsigned)b >> 1) & 3u;       /* A=0 C=1 G=3 T=2 ... */
		c ^= c >> 1;                                /* ... A=0 C=1 G=2 T=3, and whatever falls out for other bytes */
		const int u = b & 0xdf;                     /* fold lower case */
		const int ok = u == 'A' || u == 'C' || u == 'G' || u == 'T';
		pack_code[b] = (uint8_t)(c | (ok ? 0u : 4u));
	}
}

struct pack_job {
	const uint8_t *stream;
	uint64_t n, n_tiles;
	uint8_t *packed;
};

static void pack_task(size_t task, void *arg)
{
	struct pack_job *j = arg;
	uint64_t t0 = (uint64_t)task * PACK_TASK_TILES, t1 = t0 + PACK_TASK_TILES;
	if (t1 > j->n_tiles) t1 = j->n_tiles;
	for (uint64_t t = t0; t < t1; ++t) {
		uint64_t *pk = (uint64_t *)(j->packed + t * PACK_TILE_BYTES);
		uint32_t *inv = (uint32_t *)(j->packed + t * PACK_TILE_BYTES + PACK_TILE_WORDS * 8);
		for (int w = 0; w < PACK_TILE_WORDS; ++w) {
			const uint64_t g0 = t * PACK_TILE_BASES + (uint64_t)w * 32;
			uint64_t word = 0;
			uint32_t bad = 0;
			if (g0 + 32 <= j->n) {
				const uint8_t *p = j->stream + g0;
				for (int q = 0; q < 32; ++q) {
					const uint32_t c = pack_code[p[q]];
					word = (word << 2) | (c & 3u);
					bad = (bad << 1) | (c >> 2);
				}
			} else {
				for (int q = 0; q < 32; ++q) {
					uint32_t c = 4;                         /* past the end: code 0, invalid */
					if (g0 + (uint64_t)q < j->n) c = pack_code[j->stream[g0 + q]];
					word = (word << 2) | (c & 3u);
					bad = (bad << 1) | (c >> 2);
				}
			}
			pk[w] = word;
			inv[w] = bad;
		}
	}
}

uint64_t tagpu_packed_bytes(uint64_t n_positions);

int tagpu_pack_stream(const uint8_t *stream, uint64_t n_bytes, uint8_t *packed, int n_threads)
{
	if (n_threads < 1) n_threads = 1;
	pthread_once(&pack_once, pack_init);
	struct pack_job j = { stream, n_bytes, (n_bytes + PACK_TILE_BASES - 1) / PACK_TILE_BASES, packed };
	if (j.n_tiles * (uint64_t)PACK_TILE_BYTES != tagpu_packed_bytes(n_bytes)) return -1;   /* host and device layouts disagree */
	run_tasks((size_t)((j.n_tiles + PACK_TASK_TILES - 1) / PACK_TASK_TILES), n_threads, pack_task, &j);
	return 0;
}

/* ---- ingest as an object: open (index + sizes: the stream length is known), start (copy workers fill the caller's buffer
 * in stream order), ready (bytes of the stream PREFIX that are complete — the upload chases the parser with it), finish. */
struct ing_task {
	int file;
	size_t chunk;      /* plain FASTQ: chunk of the file; other formats: the whole file */
	volatile size_t end;   /* stream offset behind this task's bytes */
	/* fused mode: published by the task before (line number of the chunk's first byte, stream offset of its bytes) */
	volatile size_t line0, out0;
	volatile unsigned char ready;
};

struct tagpu_ingest {
	int n_files, n_threads, n_workers;
	struct read_file *f;
	struct pfq *pq;
	int *plain;
	size_t total, n_tasks, next, cursor;
	struct ing_task *task;
	volatile unsigned char *done;
	pthread_t *th;
	/* fused mode (tagpu_ingest_open_fused): `total` is an upper bound; the true length is known when the last chunk is sized */
	int fused;
	uint8_t *dst;
	volatile size_t true_total;
	volatile int overflowed;   /* a file held more sequence bytes than the bound allows: the stream is unusable */
};

static void *ingest_worker(void *raw)
{
	struct tagpu_ingest *g = raw;
	uint8_t *chunk_buf = NULL;
	for (;;) {
		const size_t t = __sync_fetch_and_add(&g->next, 1);
		if (t >= g->n_tasks) {
			free(chunk_buf);
			return NULL;
		}
		struct ing_task *k = g->task + t;
		if (g->fused) {
			/* one pass over the text: index the chunk, wait for the line number / stream offset the chunk before publishes,
			 * size the chunk from its index, publish for the next one, then copy */
			struct pfq *p = g->plain[k->file] ? g->pq + k->file : NULL;
			const uint8_t *ct = NULL;
			if (p) {
				/* the chunk is read into this worker's own buffer (one kernel copy out of the page cache, no page fault per
				 * 4 KB of a fresh mapping) and stays cache-warm for the sizing and the copy that follow */
				const size_t lo = k->chunk * INGEST_CHUNK, want = (lo + INGEST_CHUNK + 1 < p->f->n_txt ? lo + INGEST_CHUNK + 1 : p->f->n_txt) - lo;
				if (!chunk_buf && !(chunk_buf = malloc(INGEST_CHUNK + 64)))
					TAGPU_FATAL("out of host memory for an ingest buffer");
				size_t got = 0;
				while (got < want) {
					const ssize_t r = pread(p->fd, chunk_buf + got, want - got, (off_t)(lo + got));
					if (r <= 0)
						TAGPU_FATAL("cannot read %s", p->f->path);
					got += (size_t)r;
				}
				ct = chunk_buf;
				pfq_index_at(p, k->chunk, ct);
			}
			for (unsigned spins = 0; !k->ready; ++spins) {
				if (spins < 1024) __builtin_ia32_pause();
				else sched_yield();
			}
			__sync_synchronize();
			size_t size;
			if (p) {
				const size_t n_nl = p->nl[k->chunk];
				p->nl[k->chunk] = k->line0;
				size = pfq_walk_at(p, k->chunk, NULL, ct);
				if (t + 1 < g->n_tasks) g->task[t + 1].line0 = g->task[t + 1].file == k->file ? k->line0 + n_nl : 0;
			} else {
				size = g->f[k->file].n_seq;
				if (t + 1 < g->n_tasks) g->task[t + 1].line0 = 0;
			}
			const size_t end = k->out0 + size;
			if (end > g->total) g->overflowed = 1;
			k->end = end;
			if (t + 1 < g->n_tasks) {
				g->task[t + 1].out0 = end;
				__sync_synchronize();
				g->task[t + 1].ready = 1;
			} else {
				g->true_total = end;
			}
			if (!g->overflowed) {
				if (p) pfq_walk_at(p, k->chunk, g->dst + k->out0, ct);
				else { g->f[k->file].dst = g->dst + k->out0; ingest_phase2(g->f + k->file); }
				/* the last task pads the buffer up to the announced length: '\n' positions hold no window */
				if (t + 1 == g->n_tasks && end < g->total) memset(g->dst + end, '\n', g->total - end);
			}
			if (t + 1 == g->n_tasks) k->end = g->total;
		} else if (g->plain[k->file]) {
			struct pfq *p = g->pq + k->file;
			pfq_walk(p, k->chunk, p->f->dst + p->out[k->chunk]);
		} else {
			ingest_phase2(g->f + k->file);
		}
		__sync_synchronize();
		g->done[t] = 1;
	}
}

struct tagpu_ingest *tagpu_ingest_open(int n_files, char **files, int n_threads)
{
	if (n_threads < 1) n_threads = 1;
	if (n_threads > 64) n_threads = 64;
	struct tagpu_ingest *g = calloc(1, sizeof(*g));
	g->n_files = n_files;
	g->n_threads = n_threads;
	g->f = calloc(n_files ? n_files : 1, sizeof(*g->f));
	g->pq = calloc(n_files ? n_files : 1, sizeof(*g->pq));
	g->plain = calloc(n_files ? n_files : 1, sizeof(int));
	pthread_t *th = calloc(n_files ? n_files : 1, sizeof(pthread_t));
	/* plain FASTQ files: parallel inside the file; everything else (gzip, FASTA): one thread per file */
	for (int i = 0; i < n_files; ++i) {
		g->f[i].path = files[i];
		g->plain[i] = pfq_open(g->f + i, n_threads, g->pq + i);
		if (!g->plain[i]) pthread_create(th + i, NULL, ingest_phase1, g->f + i);
	}
	const int trace = getenv("TAGPU_TRACE_INGEST") != NULL;
	double t_a = now_s(), t_idx = 0, t_size = 0;
	for (int i = 0; i < n_files; ++i) {
		if (!g->plain[i]) continue;
		struct pfq *p = g->pq + i;
		run_tasks(p->n_chunks, n_threads, pfq_index, p);
		t_idx += now_s() - t_a; t_a = now_s();
		size_t acc = 0;
		for (size_t c = 0; c < p->n_chunks; ++c) { size_t v = p->nl[c]; p->nl[c] = acc; acc += v; }
		run_tasks(p->n_chunks, n_threads, pfq_size, p);
		acc = 0;
		for (size_t c = 0; c < p->n_chunks; ++c) { size_t v = p->out[c]; p->out[c] = acc; acc += v; }
		p->out[p->n_chunks] = acc;
		g->f[i].n_seq = acc;
		t_size += now_s() - t_a; t_a = now_s();
	}
	if (trace) fprintf(stderr, "[tagpu] ingest: newline index %.1f ms, sizing %.1f ms\n", t_idx * 1e3, t_size * 1e3);
	for (int i = 0; i < n_files; ++i) {
		if (!g->plain[i]) pthread_join(th[i], NULL);
		g->total += g->f[i].n_seq;
		g->n_tasks += g->plain[i] ? g->pq[i].n_chunks : 1;
	}
	free(th);
	g->task = calloc(g->n_tasks ? g->n_tasks : 1, sizeof(*g->task));
	g->done = calloc(g->n_tasks ? g->n_tasks : 1, 1);
	size_t t = 0, o = 0;
	for (int i = 0; i < n_files; ++i) {
		if (g->plain[i]) {
			for (size_t c = 0; c < g->pq[i].n_chunks; ++c, ++t) {
				g->task[t].file = i;
				g->task[t].chunk = c;
				g->task[t].end = o + g->pq[i].out[c + 1];
			}
		} else {
			g->task[t].file = i;
			g->task[t].end = o + g->f[i].n_seq;
			++t;
		}
		o += g->f[i].n_seq;
	}
	return g;
}

/* Fused variant: nothing is scanned here.  The stream length returned by tagpu_ingest_bytes is an UPPER BOUND (half of a
 * plain FASTQ file: a record holds at least as many quality as sequence bytes); the workers index, size and copy every
 * chunk in one pass over the text and pad the buffer behind the true end with '\n' (positions that hold no window), so a
 * consumer simply processes `bound` bytes.  tagpu_ingest_finish_fused reports the true length, or -1 if a file turned out
 * to hold more sequence than the bound (not FASTQ-shaped: use tagpu_ingest_open, which sizes exactly). */
struct tagpu_ingest *tagpu_ingest_open_fused(int n_files, char **files, int n_threads)
{
	if (n_threads < 1) n_threads = 1;
	if (n_threads > 64) n_threads = 64;
	struct tagpu_ingest *g = calloc(1, sizeof(*g));
	g->n_files = n_files;
	g->n_threads = n_threads;
	g->fused = 1;
	g->f = calloc(n_files ? n_files : 1, sizeof(*g->f));
	g->pq = calloc(n_files ? n_files : 1, sizeof(*g->pq));
	g->plain = calloc(n_files ? n_files : 1, sizeof(int));
	pthread_t *th = calloc(n_files ? n_files : 1, sizeof(pthread_t));
	for (int i = 0; i < n_files; ++i) {
		g->f[i].path = files[i];
		g->plain[i] = pfq_open(g->f + i, n_threads, g->pq + i);
		if (!g->plain[i]) pthread_create(th + i, NULL, ingest_phase1, g->f + i);
	}
	for (int i = 0; i < n_files; ++i) {
		if (!g->plain[i]) pthread_join(th[i], NULL);
		g->total += g->plain[i] ? g->f[i].n_txt / 2 + 2 : g->f[i].n_seq;
		g->n_tasks += g->plain[i] ? g->pq[i].n_chunks : 1;
	}
	free(th);
	g->task = calloc(g->n_tasks ? g->n_tasks : 1, sizeof(*g->task));
	g->done = calloc(g->n_tasks ? g->n_tasks : 1, 1);
	size_t t = 0;
	for (int i = 0; i < n_files; ++i) {
		if (g->plain[i]) {
			for (size_t c = 0; c < g->pq[i].n_chunks; ++c, ++t) {
				g->task[t].file = i;
				g->task[t].chunk = c;
			}
		} else {
			g->task[t].file = i;
			++t;
		}
	}
	if (g->n_tasks) g->task[0].ready = 1;          /* line 0, offset 0 */
	else g->total = 0;
	return g;
}

uint64_t tagpu_ingest_bytes(const struct tagpu_ingest *g) { return g->total; }

void tagpu_ingest_start(struct tagpu_ingest *g, uint8_t *dst)
{
	g->dst = dst;
	size_t o = 0;
	for (int i = 0; i < g->n_files; ++i) {
		g->f[i].dst = dst + o;
		o += g->f[i].n_seq;
	}
	g->n_workers = g->n_tasks < (size_t)g->n_threads ? (int)g->n_tasks : g->n_threads;
	const long n_cpu = sysconf(_SC_NPROCESSORS_ONLN);             /* the fused workers wait for each other in task order: */
	if (g->fused && n_cpu > 0 && g->n_workers > n_cpu) g->n_workers = (int)n_cpu;   /* never more of them than cores */
	g->th = calloc(g->n_workers ? g->n_workers : 1, sizeof(pthread_t));
	for (int i = 0; i < g->n_workers; ++i)
		pthread_create(g->th + i, NULL, ingest_worker, g);
}

/* bytes of the stream prefix that are complete; monotone; to be polled by ONE thread (void * so that it can serve as the
 * progress callback of tagpu_set_source_progress) */
uint64_t tagpu_ingest_ready(void *raw)
{
	struct tagpu_ingest *g = raw;
	while (g->cursor < g->n_tasks && g->done[g->cursor])
		++g->cursor;
	__sync_synchronize();
	return g->cursor == g->n_tasks ? g->total : (g->cursor ? g->task[g->cursor - 1].end : 0);
}

static void ingest_release(struct tagpu_ingest *g);

void tagpu_ingest_finish(struct tagpu_ingest *g)
{
	for (int i = 0; i < g->n_workers; ++i)
		pthread_join(g->th[i], NULL);
	ingest_release(g);
}

/* joins the workers of a fused ingest: true stream length, or -1 (see tagpu_ingest_open_fused) */
int64_t tagpu_ingest_finish_fused(struct tagpu_ingest *g)
{
	for (int i = 0; i < g->n_workers; ++i)
		pthread_join(g->th[i], NULL);
	const int64_t n = g->overflowed ? -1 : (int64_t)(g->n_tasks ? g->true_total : 0);
	ingest_release(g);
	return n;
}

static void ingest_release(struct tagpu_ingest *g)
{
	for (int i = 0; i < g->n_files; ++i)
		if (g->plain[i]) pfq_close(g->pq + i);
	free(g->th);
	free(g->task);
	free((void *)g->done);
	free(g->f);
	free(g->pq);
	free(g->plain);
	free(g);
}

int64_t tagpu_load_reads(int n_files, char **files, int n_threads, uint8_t **stream)
{
	/* one pass over the text (fused); a file that is not FASTQ-shaped enough for the size bound is re-read with exact sizes */
	int64_t total = -1;
	uint8_t *s = NULL;
	struct tagpu_ingest *g;
	if (!getenv("TAGPU_INGEST_TWO_PASS")) {                    /* (developer knob: compare the two ingest schedules) */
		g = tagpu_ingest_open_fused(n_files, files, n_threads);
		s = stream_buffer(g->total + 64);
		tagpu_ingest_start(g, s);
		total = tagpu_ingest_finish_fused(g);
	}
	if (total < 0) {
		if (s) tagpu_free_reads(s);
		g = tagpu_ingest_open(n_files, files, n_threads);
		total = (int64_t)g->total;
		s = stream_buffer((size_t)total + 64);
		tagpu_ingest_start(g, s);
		tagpu_ingest_finish(g);
	}
	*stream = s;
	return total;
}

void tagpu_free_reads(uint8_t *stream)
{
	if (stream && stream == g_stream_buf)
		g_stream_busy = 0;      /* kept for the next call (pinning 600 MB costs more than parsing it) */
	else
		tagpu_pinned_free(stream);
}

/* Rank's share of a read stream: nominal byte split, each cut moved forward to just after the next newline, so every
 * read belongs to exactly one rank (windows never span reads: App. A.1). */
static uint64_t shard_cut(const uint8_t *s, uint64_t n, uint64_t pos)
{
	if (pos == 0 || pos >= n)
		return pos > n ? n : pos;
	if (s[pos - 1] == '\n')
		return pos;
	const uint8_t *nl = memchr(s + pos, '\n', n - pos);
	return nl ? (uint64_t)(nl - s) + 1 : n;
}

void tagpu_dist_shard_range(const uint8_t *h_seq, uint64_t n_bytes, int rank, int world, uint64_t *begin, uint64_t *end)
{
	*begin = shard_cut(h_seq, n_bytes, n_bytes / world * rank);
	*end = rank + 1 == world ? n_bytes : shard_cut(h_seq, n_bytes, n_bytes / world * (rank + 1));
}

/* ------------------------------------------------------------------ flat graph on the host */

/* The flat arrays land in ONE pinned scratch block that is kept across calls (like the read stream buffer: pinning costs
 * more than the copy), so the device-to-host copies run at PCIe speed and touch no fresh pages. */
static uint8_t *g_flat_buf;
static size_t g_flat_cap;

static int fetch_graph(tagpu_ctx *ctx, struct tagpu_flat_graph *h, struct tagpu_stats *st)
{
	tagpu_get_stats(ctx, st);
	memset(h, 0, sizeof(*h));
	const uint64_t nn = st->n_v / 2, ne = st->n_e, nw = st->n_seq_words;
	size_t off = 0, o_mask, o_ebase, o_src, o_dst, o_rc, o_len, o_count, o_off, o_seq;
#define TAKE(var, bytes) do { var = off; off += ((size_t)(bytes) + 63) & ~(size_t)63; } while (0)
	TAKE(o_count, (ne + 1) * 8); TAKE(o_off, (ne + 1) * 8); TAKE(o_src, (ne + 1) * 4); TAKE(o_dst, (ne + 1) * 4);
	TAKE(o_rc, (ne + 1) * 4); TAKE(o_len, (ne + 1) * 4); TAKE(o_ebase, (nn + 1) * 4); TAKE(o_seq, (nw + 1) * 4); TAKE(o_mask, nn + 1);
#undef TAKE
	if (off > g_flat_cap) {
		tagpu_pinned_free(g_flat_buf);
		g_flat_cap = off + off / 8;
		g_flat_buf = tagpu_pinned_alloc(g_flat_cap);
		if (!g_flat_buf) {
			g_flat_cap = 0;
			return -1;
		}
	}
	uint8_t *b = g_flat_buf;
	h->node_mask = b + o_mask; h->node_ebase = (uint32_t *)(b + o_ebase);
	h->e_src = (uint32_t *)(b + o_src); h->e_dst = (uint32_t *)(b + o_dst); h->e_rc = (uint32_t *)(b + o_rc); h->e_len = (uint32_t *)(b + o_len);
	h->e_count = (uint64_t *)(b + o_count); h->e_off = (uint64_t *)(b + o_off); h->e_seq = (uint32_t *)(b + o_seq);
	return tagpu_copy_graph(ctx, h);
}

static void free_flat(struct tagpu_flat_graph *h)
{
	(void)h;      /* the arrays live in the persistent pinned scratch block */
}

static inline int popc4(unsigned x) { return __builtin_popcount(x & 15u); }

#define FILL_CHUNK 8192

struct fill_job {
	struct tagpu_flat_graph *h;
	struct asm_graph_t *g;
	int failed;
};

static void fill_nodes_task(size_t c, void *raw)
{
	struct fill_job *j = raw;
	const struct tagpu_flat_graph *h = j->h;
	const int64_t lo = (int64_t)c * FILL_CHUNK, hi = lo + FILL_CHUNK < (int64_t)h->n_nodes ? lo + FILL_CHUNK : (int64_t)h->n_nodes;
	for (int64_t i = lo; i < hi; ++i) {
		const unsigned m = h->node_mask[i];
		int64_t e = h->node_ebase[i];
		for (int o = 0; o < 2; ++o) {
			struct asm_node_t *nd = j->g->nodes + 2 * i + o;
			const int deg = popc4(o ? m >> 4 : m);
			nd->rc_id = 2 * i + (o ^ 1);
			nd->deg = deg;
			nd->adj = malloc(deg * sizeof(gint_t)); /* malloc(0) for dead ends, like kmer_build.c:605-606 */
			if (deg && !nd->adj) {
				j->failed = 1;
				return;
			}
			for (int a = 0; a < deg; ++a)
				nd->adj[a] = e++;
		}
	}
}

static void fill_edges_task(size_t c, void *raw)
{
	struct fill_job *j = raw;
	const struct tagpu_flat_graph *h = j->h;
	const int64_t lo = (int64_t)c * FILL_CHUNK, hi = lo + FILL_CHUNK < (int64_t)h->n_e ? lo + FILL_CHUNK : (int64_t)h->n_e;
	memset(j->g->edges + lo, 0, (size_t)(hi - lo) * sizeof(struct asm_edge_t));   /* calloc semantics, first touch in parallel */
	for (int64_t e = lo; e < hi; ++e) {
		struct asm_edge_t *ed = j->g->edges + e;
		const size_t words = ((size_t)h->e_len[e] + 15) >> 4;
		ed->count = h->e_count[e];
		ed->seq_len = h->e_len[e];
		ed->seq = malloc(words * sizeof(uint32_t));
		if (!ed->seq) {
			j->failed = 1;
			return;
		}
		memcpy(ed->seq, h->e_seq + h->e_off[e], words * sizeof(uint32_t));
		ed->source = h->e_src[e];
		ed->target = h->e_dst[e];
		ed->rc_id = h->e_rc[e];
		/* n_holes, p_holes, l_holes, barcodes, lock: zero from calloc (kmer_build.c:567) */
	}
}

/* Fills a caller-owned, uninitialised struct asm_graph_t exactly as build_asm_graph_from_kmhash leaves it
 * (/root/reference/src/kmer_build.c:567-575): nodes/edges are single calloc blocks, every adj and every seq is its
 * own allocation because later stages realloc/free them one by one. g->candidates is not touched. */
/* after a failed fill: some chunks of nodes / edges were never written (malloc'ed, not cleared), so they cannot be walked;
 * the failed run is repeated as a marking pass instead: everything a task allocated is found through the flat arrays */
static void tagpu_free_partial_graph(struct asm_graph_t *g, const struct tagpu_flat_graph *h)
{
	(void)h;
	/* the blocks themselves; the individually allocated adj / seq of the chunks that did complete are leaked on this
	 * out-of-memory path rather than guessed at (the process is about to exit(1) through TAGPU_FATAL anyway) */
	free(g->nodes);
	free(g->edges);
	g->nodes = NULL;
	g->edges = NULL;
	g->n_v = g->n_e = 0;
}

/* host half of tagpu_fill_asm_graph: flat arrays -> the reference's pointer-rich struct */
/* The node and edge blocks of a large graph are ~100 MB of fresh memory that the fill tasks touch for the first time:
 * 2 MB-aligned and advised as huge pages, so that the first touch costs a few dozen page faults instead of tens of
 * thousands (still one free()-able block each, like the reference's calloc). */
static void *big_block(size_t bytes)
{
	void *p = NULL;
	if (bytes >= ((size_t)8 << 20) && posix_memalign(&p, (size_t)2 << 20, bytes) == 0) {
#ifdef MADV_HUGEPAGE
		madvise(p, bytes, MADV_HUGEPAGE);
#endif
		return p;
	}
	return malloc(bytes);
}

int tagpu_fill_asm_graph_from_flat(const struct tagpu_flat_graph *h, int ksize, struct asm_graph_t *g)
{
	const int64_t n_nodes = (int64_t)h->n_nodes, n_e = (int64_t)h->n_e;
	g->ksize = ksize;
	g->aux_flag = 0;
	g->bin_size = 0;
	g->n_v = 2 * n_nodes;
	g->n_e = n_e;
	/* single blocks like the reference's calloc (kmer_build.c:567-575); every field of a node is written by the fill
	 * tasks and every edge is zero-filled there, in parallel, so the blocks need no serial clearing here */
	g->nodes = big_block((g->n_v ? g->n_v : 1) * sizeof(struct asm_node_t));
	g->edges = big_block((n_e ? n_e : 1) * sizeof(struct asm_edge_t));
	if (!g->nodes || !g->edges) {
		free(g->nodes);
		free(g->edges);
		g->nodes = NULL;
		g->edges = NULL;
		g->n_v = g->n_e = 0;
		return -1;
	}
	if (!n_e) memset(g->edges, 0, sizeof(struct asm_edge_t));
	if (!g->n_v) memset(g->nodes, 0, sizeof(struct asm_node_t));
	/* millions of small allocations: spread over the host threads (glibc malloc keeps one arena per thread) */
	struct fill_job job = { (struct tagpu_flat_graph *)h, g, 0 };
	int n_threads = (int)sysconf(_SC_NPROCESSORS_ONLN);
	if (n_threads > 16) n_threads = 16;
	run_tasks((size_t)((n_nodes + FILL_CHUNK - 1) / FILL_CHUNK), n_threads, fill_nodes_task, &job);
	if (!job.failed)
		run_tasks((size_t)((n_e + FILL_CHUNK - 1) / FILL_CHUNK), n_threads, fill_edges_task, &job);
	if (job.failed) {
		/* out of memory half-way: release what exists (untouched nodes / edges must look empty to the release loop) */
		tagpu_free_partial_graph(g, h);
		return -1;
	}
	return 0;
}

/* Fills a caller-owned, uninitialised struct asm_graph_t exactly as build_asm_graph_from_kmhash leaves it
 * (/root/reference/src/kmer_build.c:567-575): nodes/edges are single blocks, every adj and every seq is its
 * own allocation because later stages realloc/free them one by one. g->candidates is not touched. */
static pthread_mutex_t g_flat_lock = PTHREAD_MUTEX_INITIALIZER;   /* the pinned scratch block is shared by all contexts */

int tagpu_fill_asm_graph(tagpu_ctx *ctx, struct asm_graph_t *g)
{
	struct tagpu_flat_graph h;
	struct tagpu_stats st;
	const int trace = getenv("TAGPU_TRACE_FILL") != NULL;
	const double t0 = now_s();
	pthread_mutex_lock(&g_flat_lock);
	if (fetch_graph(ctx, &h, &st)) {
		pthread_mutex_unlock(&g_flat_lock);
		return -1;
	}
	const double t1 = now_s();
	const int rc = tagpu_fill_asm_graph_from_flat(&h, tagpu_ctx_k(ctx), g);
	if (trace)
		fprintf(stderr, "[tagpu] fill: device -> pinned host %.2f ms, nodes + edges %.2f ms\n", (t1 - t0) * 1e3, (now_s() - t1) * 1e3);
	free_flat(&h);
	pthread_mutex_unlock(&g_flat_lock);
	return rc;
}

/* Releases what tagpu_fill_asm_graph (or one of the reference entry points above it) allocated, block by block like the
 * reference's asm_graph_destroy (/root/reference/src/assembly_graph.c:1459-1480); the struct itself stays the caller's. */
void tagpu_free_asm_graph(struct asm_graph_t *g)
{
	if (!g) return;
	for (gint_t u = 0; g->nodes && u < g->n_v; ++u)
		free(g->nodes[u].adj);
	for (gint_t e = 0; g->edges && e < g->n_e; ++e) {
		free(g->edges[e].seq);
		free(g->edges[e].p_holes);
		free(g->edges[e].l_holes);
	}
	free(g->nodes);
	free(g->edges);
	g->nodes = NULL;
	g->edges = NULL;
	g->n_v = g->n_e = 0;
}

/* save_asm_graph layout (/root/reference/src/assembly_graph.c:1173-1248), streamed straight from the
 * flat arrays without building the pointer-rich struct. */
int tagpu_write_graph_bin(tagpu_ctx *ctx, const char *path)
{
	struct tagpu_flat_graph h;
	struct tagpu_stats st;
	if (fetch_graph(ctx, &h, &st))
		return -1;
	FILE *fp = fopen(path, "wb");
	if (!fp) {
		perror(path);
		free_flat(&h);
		return -1;
	}
	setvbuf(fp, NULL, _IOFBF, 1 << 22);
	const uint32_t aux_flag = 0;
	const int32_t ksize = tagpu_ctx_k(ctx);
	const int64_t n_v = 2 * (int64_t)h.n_nodes, n_e = (int64_t)h.n_e;
	fwrite("asmg", 1, 4, fp);
	fwrite(&aux_flag, 4, 1, fp);
	fwrite(&ksize, 4, 1, fp);
	fwrite(&n_v, 8, 1, fp);
	fwrite(&n_e, 8, 1, fp);
	for (int64_t i = 0; i < (int64_t)h.n_nodes; ++i) {
		const unsigned m = h.node_mask[i];
		int64_t e = h.node_ebase[i];
		for (int o = 0; o < 2; ++o) {
			const int64_t rc_id = 2 * i + (o ^ 1), deg = popc4(o ? m >> 4 : m);
			fwrite(&rc_id, 8, 1, fp);
			fwrite(&deg, 8, 1, fp);
			for (int64_t j = 0; j < deg; ++j, ++e)
				fwrite(&e, 8, 1, fp);
		}
	}
	for (int64_t e = 0; e < n_e; ++e) {
		const int64_t src = h.e_src[e], dst = h.e_dst[e], rc = h.e_rc[e];
		const uint64_t len8 = h.e_len[e]; /* seq_len plus the aliased, zero n_holes (App. F.2) */
		const uint32_t n_holes = 0;
		fwrite(&src, 8, 1, fp);
		fwrite(&dst, 8, 1, fp);
		fwrite(&rc, 8, 1, fp);
		fwrite(&h.e_count[e], 8, 1, fp);
		fwrite(&len8, 8, 1, fp);
		fwrite(h.e_seq + h.e_off[e], 4, ((size_t)h.e_len[e] + 15) >> 4, fp);
		fwrite(&n_holes, 4, 1, fp);
	}
	int rc = ferror(fp) ? -1 : 0;
	fclose(fp);
	free_flat(&h);
	return rc;
}

/* ------------------------------------------------------------------ KMC database writer */

struct solid_rec { uint64_t hi, lo; uint32_t cnt; };

static int cmp_solid_rec(const void *a, const void *b)
{
	const struct solid_rec *x = a, *y = b;
	if (x->hi != y->hi) return x->hi < y->hi ? -1 : 1;
	if (x->lo != y->lo) return x->lo < y->lo ? -1 : 1;
	return 0;
}

struct kmc_trailer {
	uint32_t kmer_length, mode, counter_size, lut_prefix_length, signature_length, min_count, max_count;
	uint64_t total_kmers;
	uint8_t both_strands, pad8[3];
	uint32_t pad32[6];
	uint32_t version;
} __attribute__((packed));

/* Writes what /root/reference/src/KMC_reader.c:22-74 (prefix file) and :204-256 (suffix records) read back:
 * records sorted by value, prefix LUT over the top lut_prefix_length bases, 4-byte little-endian counters. */
/* The records must be sorted by value, and the prefix file needs the number of records per lut_prefix_length-base prefix
 * anyway: so the sort is a parallel bucket sort on exactly that prefix (per-chunk histograms -> offsets -> scatter, all
 * chunks in parallel), followed by an independent qsort inside every prefix bucket, buckets in parallel. */
#define KMC_SORT_CHUNKS 64

struct kmc_sort {
	const uint64_t *hi, *lo;
	const uint32_t *cnt;
	uint64_t n, n_lut;
	int shift;                 /* 2 * suffix bases: key >> shift = prefix */
	uint64_t *hist;            /* [KMC_SORT_CHUNKS][n_lut]: counts, then scatter cursors */
	const uint64_t *lut;       /* [n_lut + 1] first record of every prefix */
	struct solid_rec *rec;
};

static inline uint64_t kmc_prefix(const struct kmc_sort *k, uint64_t i)
{
	const unsigned __int128 x = ((unsigned __int128)k->hi[i] << 64) | k->lo[i];
	return (uint64_t)(x >> k->shift);
}

static void kmc_hist_task(size_t c, void *raw)
{
	struct kmc_sort *k = raw;
	const uint64_t lo = k->n * c / KMC_SORT_CHUNKS, hi = k->n * (c + 1) / KMC_SORT_CHUNKS;
	uint64_t *h = k->hist + c * k->n_lut;
	for (uint64_t i = lo; i < hi; ++i)
		++h[kmc_prefix(k, i)];
}

static void kmc_scatter_task(size_t c, void *raw)
{
	struct kmc_sort *k = raw;
	const uint64_t lo = k->n * c / KMC_SORT_CHUNKS, hi = k->n * (c + 1) / KMC_SORT_CHUNKS;
	uint64_t *cur = k->hist + c * k->n_lut;
	for (uint64_t i = lo; i < hi; ++i) {
		struct solid_rec *r = k->rec + cur[kmc_prefix(k, i)]++;
		r->hi = k->hi[i]; r->lo = k->lo[i]; r->cnt = k->cnt[i];
	}
}

static void kmc_bucket_sort_task(size_t t, void *raw)
{
	struct kmc_sort *k = raw;
	/* task t sorts the prefixes [t * 64, t * 64 + 64) */
	for (uint64_t b = t * 64; b < (t + 1) * 64 && b < k->n_lut; ++b)
		if (k->lut[b + 1] - k->lut[b] > 1)
			qsort(k->rec + k->lut[b], k->lut[b + 1] - k->lut[b], sizeof(*k->rec), cmp_solid_rec);
}

int tagpu_write_kmc_db(tagpu_ctx *ctx, const char *working_dir)
{
	struct tagpu_stats st;
	tagpu_get_stats(ctx, &st);
	const int K = tagpu_ctx_K(ctx);
	const uint64_t n = st.n_solid;
	const int p = (K % 4) + 4, suf_bases = K - p, suf_bytes = suf_bases / 4;
	const uint64_t n_lut = (uint64_t)1 << (2 * p);
	int rc = -1, n_threads = (int)sysconf(_SC_NPROCESSORS_ONLN);
	if (n_threads > 32) n_threads = 32;
	if (n_threads < 1) n_threads = 1;
	FILE *fs = NULL, *fp = NULL;
	uint64_t *hi = malloc((n + 1) * 8), *lo = malloc((n + 1) * 8), *lut = calloc(n_lut + 1, 8);
	uint64_t *hist = calloc((size_t)KMC_SORT_CHUNKS * n_lut, 8);
	uint32_t *cnt = malloc((n + 1) * 4);
	struct solid_rec *rec = malloc((n + 1) * sizeof(*rec));
	char path[4096];
	if (!hi || !lo || !cnt || !rec || !lut || !hist || tagpu_copy_solid(ctx, hi, lo, cnt))
		goto done;
	struct kmc_sort ks = { hi, lo, cnt, n, n_lut, 2 * suf_bases, hist, lut, rec };
	run_tasks(KMC_SORT_CHUNKS, n_threads, kmc_hist_task, &ks);
	/* lut[b] = records before prefix b; hist[c][b] becomes the place of chunk c's first record with prefix b */
	uint64_t acc = 0;
	for (uint64_t b = 0; b < n_lut; ++b) {
		lut[b] = acc;
		for (int c = 0; c < KMC_SORT_CHUNKS; ++c) {
			const uint64_t v = hist[(size_t)c * n_lut + b];
			hist[(size_t)c * n_lut + b] = acc;
			acc += v;
		}
	}
	lut[n_lut] = acc;
	run_tasks(KMC_SORT_CHUNKS, n_threads, kmc_scatter_task, &ks);
	run_tasks((size_t)((n_lut + 63) / 64), n_threads, kmc_bucket_sort_task, &ks);

	snprintf(path, sizeof(path), "%s/KMC_%d_count.kmc_suf", working_dir, K);
	fs = fopen(path, "wb");
	if (!fs) { perror(path); goto done; }
	setvbuf(fs, NULL, _IOFBF, 1 << 22);
	fwrite("KMCS", 1, 4, fs);
	for (uint64_t i = 0; i < n; ++i) {
		unsigned __int128 x = ((unsigned __int128)rec[i].hi << 64) | rec[i].lo;
		uint8_t out[40];
		for (int j = 0; j < suf_bytes; ++j)
			out[j] = (uint8_t)(x >> (8 * (suf_bytes - 1 - j)));
		memcpy(out + suf_bytes, &rec[i].cnt, 4);
		fwrite(out, 1, suf_bytes + 4, fs);
	}
	fwrite("KMCS", 1, 4, fs);
	if (ferror(fs)) goto done;
	snprintf(path, sizeof(path), "%s/KMC_%d_count.kmc_pre", working_dir, K);
	fp = fopen(path, "wb");
	if (!fp) { perror(path); goto done; }
	const uint32_t sigmap[2] = { 0, 0 };
	struct kmc_trailer t;
	memset(&t, 0, sizeof(t));
	t.kmer_length = K; t.counter_size = 4; t.lut_prefix_length = p; t.signature_length = 0;
	t.min_count = tagpu_ctx_cutoff(ctx); t.max_count = 0xffffffffu; t.total_kmers = n; t.both_strands = 1;
	t.version = 0x200;
	const uint32_t trailer_size = sizeof(t);
	fwrite("KMCP", 1, 4, fp);
	fwrite(lut, 8, n_lut + 1, fp);
	fwrite(sigmap, 4, 2, fp);
	fwrite(&t, sizeof(t), 1, fp);
	fwrite(&trailer_size, 4, 1, fp);
	fwrite("KMCP", 1, 4, fp);
	rc = ferror(fp) ? -1 : 0;
done:	/* one exit: nothing leaks on the early returns */
	if (fs) fclose(fs);
	if (fp) fclose(fp);
	free(hi); free(lo); free(cnt); free(rec); free(lut); free(hist);
	return rc;
}

/* ------------------------------------------------------------------ the reference's entry points */

static tagpu_ctx *g_ctx;

static tagpu_ctx *global_ctx(void)
{
	if (!g_ctx) {
		g_ctx = tagpu_create(-1);
		if (!g_ctx)
			TAGPU_FATAL("no usable CUDA device; libtagpu has no CPU path");
		const char *ci = getenv("TAGPU_CUTOFF");
		if (ci)
			tagpu_set_cutoff(g_ctx, atoi(ci));
	}
	return g_ctx;
}

static int64_t gather_files(int n_files, char **files_1, char **files_2, int n_threads, uint8_t **stream)
{
	/* files_1[0..n) ++ files_2[0..n): /root/reference/src/kmer_build.c:733-735.  (n_files < 0, the contig-file mode, exists for
	 * the global stage entry points only — stage_entry — not for the callers of this function.) */
	if (n_files < 0)
		TAGPU_FATAL("n_files < 0 (contig-file mode) is a mode of build_graph_from_scratch only");
	char **all = malloc(2 * (size_t)n_files * sizeof(char *));
	memcpy(all, files_1, n_files * sizeof(char *));
	memcpy(all + n_files, files_2, n_files * sizeof(char *));
	int64_t n = tagpu_load_reads(2 * n_files, all, n_threads, stream);
	free(all);
	return n;
}

/* files_1[0..n) ++ files_2[0..n) (/root/reference/src/kmer_build.c:733-735) opened for ingest: index + sizes done, the
 * stream length known; the copy into the pinned stream buffer then runs while the GPU layer already uploads the finished
 * prefix (tagpu_set_source_progress) */
static struct tagpu_ingest *open_pairs(int n_files, char **files_1, char **files_2, int n_threads, int fused)
{
	if (n_files < 0)
		TAGPU_FATAL("n_files < 0 (contig-file mode) is not supported by the GPU path");
	char **all = malloc(2 * (size_t)n_files * sizeof(char *) + 1);
	memcpy(all, files_1, n_files * sizeof(char *));
	memcpy(all + n_files, files_2, n_files * sizeof(char *));
	struct tagpu_ingest *ing = fused ? tagpu_ingest_open_fused(2 * n_files, all, n_threads) : tagpu_ingest_open(2 * n_files, all, n_threads);
	free(all);
	return ing;
}

/* ------------------------------------------------------------------ plain FASTQ files, parsed on the GPU
 * The host only moves bytes: reader threads pread the files, chunk by chunk in file order, into the slots of a pinned ring;
 * the calling thread sends every slot up as soon as it is full (one H2D copy per chunk, the slot is free again when that
 * copy has completed) and then asks the device to parse the records and build (tagpu_kernels.cu:tagpu_build_fastq_device,
 * tagpu_fastq.cuh).  Anything that is not a regular, uncompressed FASTQ file takes the host parser (returns 1). */
#define RAW_CHUNK ((size_t)8 << 20)
#define RAW_SLOTS 24

struct raw_chunk { int file; uint64_t off, len; };
struct raw_job {
	int *fd;
	char **path;
	size_t n_chunks;
	struct raw_chunk *chunk;
	volatile unsigned char *ready;
	volatile size_t next, freed;       /* next chunk to hand out; chunks whose upload has completed (in order) */
	uint8_t *ring;
};

static void *raw_reader(void *raw)
{
	struct raw_job *j = raw;
	for (;;) {
		const size_t c = __sync_fetch_and_add(&j->next, 1);
		if (c >= j->n_chunks)
			return NULL;
		for (unsigned spins = 0; j->freed + RAW_SLOTS <= c; ++spins) {   /* the slot still holds chunk c - RAW_SLOTS */
			if (spins < 1024) __builtin_ia32_pause();
			else sched_yield();
		}
		uint8_t *dst = j->ring + (c % RAW_SLOTS) * RAW_CHUNK;
		const struct raw_chunk *k = j->chunk + c;
		uint64_t got = 0;
		while (got < k->len) {
			const ssize_t r = pread(j->fd[k->file], dst + got, k->len - got, (off_t)(k->off + got));
			if (r <= 0)
				TAGPU_FATAL("cannot read %s", j->path[k->file]);
			got += (uint64_t)r;
		}
		__sync_synchronize();
		j->ready[c] = 1;
	}
}

/* 0 = built, 1 = not applicable (take the host parser), exits on errors like the other entry points */
static int build_files_on_device(tagpu_ctx *ctx, int n_files, char **files, int n_threads, int ksize, int with_graph)
{
	/* opt-in (TAGPU_DEVICE_PARSE=1): on the boxes measured, reading the files out of the page cache is what bounds both ingest
	 * paths (~33 GB/s over 16 threads), and the host parser sends up only the sequence bytes and overlaps pass 1 with the
	 * reading — 47 ms against 53 ms for two 638 MB files */
	if (!getenv("TAGPU_DEVICE_PARSE") || n_files <= 0 || n_files > 4096)
		return 1;
	struct raw_job j;
	memset(&j, 0, sizeof(j));
	j.fd = malloc(n_files * sizeof(int));
	j.path = files;
	uint64_t *len = calloc(n_files, 8), *dev_off = calloc(n_files, 8), total = 0;
	uint8_t *ends_nl = calloc(n_files, 1);
	int ok = 1, n_open = 0;
	const char *max_env = getenv("TAGPU_RAW_MAX_BYTES");
	const uint64_t max_total = max_env ? strtoull(max_env, NULL, 10) : (uint64_t)48 << 30;
	for (int i = 0; i < n_files && ok; ++i) {
		struct stat st;
		unsigned char first = 0, last = 0;
		j.fd[i] = open(files[i], O_RDONLY);
		if (j.fd[i] < 0)
			TAGPU_FATAL("cannot open %s", files[i]);
		++n_open;
		if (fstat(j.fd[i], &st) != 0 || !S_ISREG(st.st_mode) || st.st_size < 1 || pread(j.fd[i], &first, 1, 0) != 1 || first != '@' ||
		    pread(j.fd[i], &last, 1, st.st_size - 1) != 1) {
			ok = 0;
			break;
		}
		len[i] = (uint64_t)st.st_size;
		ends_nl[i] = last == '\n';
		dev_off[i] = total;
		total += (len[i] + 255) & ~(uint64_t)255;
		j.n_chunks += (len[i] + RAW_CHUNK - 1) / RAW_CHUNK;
	}
	if (ok && total > max_total) ok = 0;
	if (ok && !(j.ring = tagpu_raw_ring(ctx, RAW_SLOTS * RAW_CHUNK))) ok = 0;
	if (ok && tagpu_raw_begin(ctx, total)) ok = 0;
	if (!ok) {
		for (int i = 0; i < n_open; ++i) close(j.fd[i]);
		free(j.fd); free(len); free(dev_off); free(ends_nl);
		return 1;
	}
	j.chunk = malloc((j.n_chunks ? j.n_chunks : 1) * sizeof(*j.chunk));
	j.ready = calloc(j.n_chunks ? j.n_chunks : 1, 1);
	size_t c = 0;
	for (int i = 0; i < n_files; ++i)
		for (uint64_t o = 0; o < len[i]; o += RAW_CHUNK, ++c) {
			j.chunk[c].file = i;
			j.chunk[c].off = o;
			j.chunk[c].len = len[i] - o < RAW_CHUNK ? len[i] - o : RAW_CHUNK;
		}
	if (n_threads < 1) n_threads = 1;
	if (n_threads > 64) n_threads = 64;
	if ((size_t)n_threads > j.n_chunks) n_threads = (int)j.n_chunks;
	pthread_t th[64];
	for (int t = 0; t < n_threads; ++t)
		pthread_create(th + t, NULL, raw_reader, &j);
	for (size_t issued = 0; issued < j.n_chunks; ++issued) {
		for (unsigned spins = 0; !j.ready[issued]; ++spins) {
			if (spins < 1024) __builtin_ia32_pause();
			else sched_yield();
		}
		__sync_synchronize();
		const struct raw_chunk *k = j.chunk + issued;
		if (tagpu_raw_put(ctx, dev_off[k->file] + k->off, j.ring + (issued % RAW_SLOTS) * RAW_CHUNK, k->len, (int)(issued % RAW_SLOTS)))
			TAGPU_FATAL("upload of %s failed: %s", files[k->file], tagpu_last_error(ctx));
		while (issued + 1 - j.freed > RAW_SLOTS / 2) {       /* keep half of the ring in flight, hand the rest back */
			if (tagpu_raw_slot_wait(ctx, (int)(j.freed % RAW_SLOTS)))
				TAGPU_FATAL("upload failed: %s", tagpu_last_error(ctx));
			__sync_synchronize();
			++j.freed;
		}
	}
	for (int t = 0; t < n_threads; ++t)
		pthread_join(th[t], NULL);
	const int rc = tagpu_build_fastq_device(ctx, n_files, dev_off, len, ends_nl, ksize, with_graph);
	for (int i = 0; i < n_files; ++i) close(j.fd[i]);
	free(j.fd); free(len); free(dev_off); free(ends_nl); free(j.chunk); free((void *)j.ready);
	if (rc)
		TAGPU_FATAL("GPU graph build failed: %s", tagpu_last_error(ctx));
	return 0;
}

static void stage_entry(int ksize, int n_threads, int n_files, char **files_1, char **files_2, char *work_dir,
			struct asm_graph_t *g, int skip_counts)
{
	tagpu_ctx *ctx = global_ctx();
	double t0 = now_s(), t1 = t0;
	tagpu_set_skip_counts(ctx, skip_counts);
	/* TAGPU_DEVICE_PARSE=1: plain FASTQ files are parsed on the GPU (everything else — gzip, FASTA, pipes — by the host parser) */
	int on_device = 0;
	if (n_files > 0) {
		char **all = malloc(2 * (size_t)n_files * sizeof(char *) + 1);
		memcpy(all, files_1, n_files * sizeof(char *));
		memcpy(all + n_files, files_2, n_files * sizeof(char *));
		on_device = build_files_on_device(ctx, 2 * n_files, all, n_threads, ksize, 1) == 0;
		free(all);
	}
	/* n_files < 0: the reference's contig-file mode (/root/reference/src/kmer_build.c:677-679,722-731,779-781).  files_2 holds
	 * the R2 files, then ONE contig file, then the 2 |n_files| read files the edge counts are taken from.  The graph comes from
	 * files_1 ++ files_2[0 .. |n_files|] (the contig included); without counts that is all, with counts a second count pass
	 * over the read files alone supplies them. */
	if (n_files < 0) {
		const int n = -n_files;
		char **set_a = malloc((2 * (size_t)n + 1) * sizeof(char *));
		memcpy(set_a, files_1, n * sizeof(char *));
		memcpy(set_a + n, files_2, ((size_t)n + 1) * sizeof(char *));
		uint8_t *sa = NULL, *sb = NULL;
		const int64_t na = tagpu_load_reads(2 * n + 1, set_a, n_threads, &sa);
		t1 = now_s();
		if (skip_counts) {
			if (tagpu_build_host(ctx, sa, (uint64_t)na, ksize))
				TAGPU_FATAL("GPU graph build failed: %s", tagpu_last_error(ctx));
		} else {
			const int64_t nb = tagpu_load_reads(2 * n, files_2 + n + 1, n_threads, &sb);
			if (tagpu_build_host_counts_from(ctx, sa, (uint64_t)na, sb, (uint64_t)nb, ksize))
				TAGPU_FATAL("GPU graph build failed: %s", tagpu_last_error(ctx));
			tagpu_free_reads(sb);
		}
		tagpu_free_reads(sa);
		free(set_a);
		on_device = -1;
	}
	for (int fused = getenv("TAGPU_INGEST_TWO_PASS") ? 0 : 1; fused >= 0 && !on_device; --fused) {
		/* fused ingest: the workers index, size and copy each chunk in one pass while pass 1 on the GPU chases them; the
		 * announced stream length is an upper bound and the buffer is padded with '\n' (no windows there).  Files that do
		 * not respect the bound (not FASTQ-shaped) take the two-pass ingest with exact sizes. */
		struct tagpu_ingest *ing = open_pairs(n_files, files_1, files_2, n_threads, fused);
		const uint64_t n = tagpu_ingest_bytes(ing);
		uint8_t *stream = stream_buffer(n + 64);
		t1 = now_s();
		tagpu_ingest_start(ing, stream);
		tagpu_set_source_progress(ctx, tagpu_ingest_ready, ing);
		if (tagpu_build_host(ctx, stream, n, ksize))
			TAGPU_FATAL("GPU graph build failed: %s", tagpu_last_error(ctx));
		int64_t true_n = (int64_t)n;
		if (fused) true_n = tagpu_ingest_finish_fused(ing);
		else tagpu_ingest_finish(ing);
		tagpu_free_reads(stream);
		if (true_n >= 0) break;
	}
	double t2 = now_s();
	if (getenv("TAGPU_WRITE_KMC_DB") && tagpu_write_kmc_db(ctx, work_dir))
		TAGPU_FATAL("cannot write the KMC database into %s", work_dir);
	struct tagpu_stats st;
	tagpu_get_stats(ctx, &st);
	/* the reference's three known-answer log lines: kmer_build.c:758,763,772 */
	fprintf(stderr, "[tagpu] Number of kmer: %lu\n", (unsigned long)st.n_kmers);
	fprintf(stderr, "[tagpu] Number of nodes: %ld; Number of edges: %ld\n", (long)st.n_v, (long)st.n_e);
	if (!skip_counts)
		fprintf(stderr, "[tagpu] Number of (k+1)-mer on edge: %lu\n", (unsigned long)st.n_kp1_on_edge);
	if (tagpu_fill_asm_graph(ctx, g))
		TAGPU_FATAL("cannot materialise the assembly graph: %s", tagpu_last_error(ctx));
	double t3 = now_s();
	fprintf(stderr, "[tagpu] k=%d: %lu (k+1)-mer instances, %lu solid; open %.3f s, %s + H2D + GPU %.3f s (device build %.3f ms), "
			"graph materialisation %.3f s\n", ksize, (unsigned long)st.n_instances, (unsigned long)st.n_solid,
		t1 - t0, on_device > 0 ? "read (records parsed on the GPU)" : "parse", t2 - t1, st.ms_total, t3 - t2);
}

void build_graph_from_scratch(int ksize, int n_threads, int mmem, int n_files, char **files_1, char **files_2,
			      char *work_dir, struct asm_graph_t *g)
{
	(void)mmem;
	stage_entry(ksize, n_threads, n_files, files_1, files_2, work_dir, g, 0);
}

void build_graph_from_scratch_without_count(int ksize, int n_threads, int mmem, int n_files, char **files_1,
					    char **files_2, char *work_dir, struct asm_graph_t *g)
{
	(void)mmem;
	stage_entry(ksize, n_threads, n_files, files_1, files_2, work_dir, g, 1);
}

/* /root/reference/src/assembly_graph.h:160-162, body /root/reference/src/kmer_build.c:991-1044: the stage re-entered per
 * gap by local assembly (get_local_assembly, /root/reference/src/barcode_resolve2.c:2100) — reads of the gap plus the
 * "garbage" of the two flanking edges e1, e2 of the global graph g0 (add_garbage / assign_count_garbage). */
void build_local_assembly_graph(int ksize, int n_threads, int mmem, int n_files, char **files_1, char **files_2,
				char *work_dir, struct asm_graph_t *g, struct asm_graph_t *g0, gint_t e1, gint_t e2)
{
	(void)mmem; (void)work_dir;
	tagpu_ctx *ctx = global_ctx();
	uint8_t *stream;
	int64_t n = gather_files(n_files, files_1, files_2, n_threads, &stream);
	/* the two contigs as text, each followed by a newline (add_garbage walks the 2-bit sequence base by base) */
	const gint_t ee[2] = { e1, e2 };
	uint64_t off[2], total = 0;
	uint32_t len[2];
	double cov[2];
	for (int c = 0; c < 2; ++c) {
		const struct asm_edge_t *ed = g0->edges + ee[c];
		off[c] = total;
		len[c] = ed->seq_len;
		/* __get_edge_cov(g0->edges + e, g0->ksize), /root/reference/src/assembly_graph.h:191-192 */
		cov[c] = ed->count * 1.0 / (ed->seq_len - (ed->n_holes + 1) * (g0->ksize));
		total += (uint64_t)ed->seq_len + 1;
	}
	uint8_t *txt = malloc(total + 1);
	if (!txt)
		TAGPU_FATAL("out of host memory for the flanking contigs");
	for (int c = 0; c < 2; ++c) {
		const struct asm_edge_t *ed = g0->edges + ee[c];
		for (uint32_t i = 0; i < ed->seq_len; ++i)
			txt[off[c] + i] = "ACGT"[(ed->seq[i >> 4] >> ((i & 15) << 1)) & 3];
		txt[off[c] + ed->seq_len] = '\n';
	}
	tagpu_set_skip_counts(ctx, 0);
	if (tagpu_build_local_host(ctx, stream, (uint64_t)n, ksize, txt, total, 2, off, len, cov))
		TAGPU_FATAL("GPU local graph build failed: %s", tagpu_last_error(ctx));
	free(txt);
	tagpu_free_reads(stream);
	struct tagpu_stats st;
	tagpu_get_stats(ctx, &st);
	fprintf(stderr, "[tagpu] Number of kmer: %lu\n", (unsigned long)st.n_kmers);                    /* kmer_build.c:1020 */
	fprintf(stderr, "[tagpu] Number of nodes: %ld; Number of edges: %ld\n", (long)st.n_v, (long)st.n_e); /* :1026 */
	fprintf(stderr, "[tagpu] Number of (k+1)-mer on edge: %lu\n", (unsigned long)st.n_kp1_on_edge);   /* :1032 */
	if (tagpu_fill_asm_graph(ctx, g))
		TAGPU_FATAL("cannot materialise the assembly graph: %s", tagpu_last_error(ctx));
}

/* ------------------------------------------------------------------ many local builds in flight (row f1)
 * The reference calls build_local_assembly_graph in a sequential loop over thousands of gaps
 * (/root/reference/src/build_bridge.c:1036-1062), each a build of a few thousand reads: on a GPU one such build is a chain
 * of ~30 tiny kernels and a handful of read-backs, i.e. launch latency.  The key spaces of different gaps must stay
 * separate, so the gaps are not merged into one build; instead n_ctx contexts — each with its own CUDA stream, buffers
 * and host thread — work through the job list concurrently and the small kernels of different gaps overlap on the GPU. */
struct local_batch {
	struct tagpu_local_job *jobs;
	int n_jobs, next, device;
};

static tagpu_ctx *g_pool[64];
static pthread_mutex_t g_pool_lock = PTHREAD_MUTEX_INITIALIZER;

static void *local_batch_worker(void *raw)
{
	struct local_batch *b = raw;
	tagpu_ctx *ctx = NULL;
	int slot = -1;
	pthread_mutex_lock(&g_pool_lock);                       /* contexts are kept across calls: creating one allocates streams and events */
	for (int i = 0; i < 64 && slot < 0; ++i)
		if (g_pool[i]) { ctx = g_pool[i]; g_pool[i] = NULL; slot = i; }
	pthread_mutex_unlock(&g_pool_lock);
	if (!ctx) ctx = tagpu_create(b->device);
	for (;;) {
		const int j = __sync_fetch_and_add(&b->next, 1);
		if (j >= b->n_jobs)
			break;
		struct tagpu_local_job *job = b->jobs + j;
		if (!ctx) { job->rc = -1; continue; }
		tagpu_set_cutoff(ctx, job->cutoff > 0 ? job->cutoff : 2);
		tagpu_set_skip_counts(ctx, 0);
		job->rc = tagpu_build_local_host(ctx, job->reads, job->n_bytes, job->k, job->contigs, job->n_contig_bytes, job->n_contigs,
						 job->contig_off, job->contig_len, job->contig_cov);
		if (!job->rc) {
			tagpu_get_stats(ctx, &job->stats);
			if (job->g) job->rc = tagpu_fill_asm_graph(ctx, job->g);
		}
	}
	if (ctx) {
		pthread_mutex_lock(&g_pool_lock);
		for (int i = 0; i < 64; ++i)
			if (!g_pool[i]) { g_pool[i] = ctx; ctx = NULL; break; }
		pthread_mutex_unlock(&g_pool_lock);
		if (ctx) tagpu_destroy(ctx);
	}
	return NULL;
}

int tagpu_build_local_batch(int device, int n_ctx, struct tagpu_local_job *jobs, int n_jobs)
{
	if (n_ctx < 1) n_ctx = 1;
	if (n_ctx > 64) n_ctx = 64;
	if (n_ctx > n_jobs) n_ctx = n_jobs;
	struct local_batch b = { jobs, n_jobs, 0, device };
	pthread_t th[64];
	for (int i = 0; i < n_ctx; ++i)
		pthread_create(th + i, NULL, local_batch_worker, &b);
	for (int i = 0; i < n_ctx; ++i)
		pthread_join(th[i], NULL);
	int bad = 0;
	for (int j = 0; j < n_jobs; ++j)
		bad += jobs[j].rc != 0;
	return bad ? -1 : 0;
}

void build_initial_graph(struct opt_proc_t *opt, int ksize, struct asm_graph_t *g)
{
	double t0 = now_s();
	build_graph_from_scratch(ksize, opt->n_threads, opt->mmem, opt->n_files, opt->files_1, opt->files_2, opt->out_dir, g);
	fprintf(stderr, "[tagpu] Building graph time: %.3f\n", now_s() - t0); /* kmer_build.c:844 */
}

int KMC_build_kmer_database(int ksize, const char *working_dir, int n_threads, int mmem, int n_files, char **files)
{
	(void)mmem;
	tagpu_ctx *ctx = global_ctx();
	for (int fused = 1; fused >= 0; --fused) {
		struct tagpu_ingest *ing = fused ? tagpu_ingest_open_fused(n_files, files, n_threads) : tagpu_ingest_open(n_files, files, n_threads);
		const uint64_t n = tagpu_ingest_bytes(ing);
		uint8_t *stream = stream_buffer(n + 64);
		tagpu_ingest_start(ing, stream);
		tagpu_set_source_progress(ctx, tagpu_ingest_ready, ing);
		if (tagpu_count_host(ctx, stream, n, ksize))
			TAGPU_FATAL("GPU k-mer counting failed: %s", tagpu_last_error(ctx));
		int64_t true_n = (int64_t)n;
		if (fused) true_n = tagpu_ingest_finish_fused(ing);
		else tagpu_ingest_finish(ing);
		tagpu_free_reads(stream);
		if (true_n >= 0) break;
	}
	if (tagpu_write_kmc_db(ctx, working_dir))
		TAGPU_FATAL("cannot write the KMC database into %s", working_dir);
	return 0;
}

/* /root/reference/src/coverage/kmer_count.c:198-240 (kmer_count_on_edges) and :113-135 (add_cnt_to_graph): the coverage recount
 * of build_coverage_process (/root/reference/src/process.c:823-835).  The "table" handed from the first to the second is opaque
 * to every caller; ours carries the finished per-edge counts. */
struct cov_result {
	gint_t n_e;
	uint64_t *count;
};

struct mini_hash_t *kmer_count_on_edges(struct opt_proc_t *opt, struct asm_graph_t *g)
{
	tagpu_ctx *ctx = global_ctx();
	uint8_t *stream;
	const uint64_t n = (uint64_t)gather_files(opt->n_files, opt->files_1, opt->files_2, opt->n_threads, &stream);
	/* flatten the edges */
	const gint_t n_e = g->n_e;
	uint32_t *len = malloc((n_e + 1) * 4), *rc = malloc((n_e + 1) * 4);
	uint64_t *off = malloc((n_e + 1) * 8), n_words = 0;
	for (gint_t e = 0; e < n_e; ++e) {
		const int live = g->edges[e].source != -1 && g->edges[e].seq != NULL;
		len[e] = live ? g->edges[e].seq_len : 0;      /* (a removed edge has nothing to index; the reference would crash on it) */
		rc[e] = (uint32_t)(live ? g->edges[e].rc_id : e);
		off[e] = n_words;
		n_words += ((uint64_t)len[e] + 15) >> 4;
	}
	uint32_t *words = malloc((n_words + 1) * 4);
	struct cov_result *res = calloc(1, sizeof(*res));
	if (!len || !rc || !off || !words || !res || !(res->count = calloc(n_e + 1, 8)))
		TAGPU_FATAL("out of host memory for the coverage recount");
	for (gint_t e = 0; e < n_e; ++e)
		if (len[e]) memcpy(words + off[e], g->edges[e].seq, (((size_t)len[e] + 15) >> 4) * 4);
	if (tagpu_coverage_recount_host(ctx, stream, n, (uint64_t)n_e, len, off, words, n_words, rc, res->count))
		TAGPU_FATAL("GPU coverage recount failed: %s", tagpu_last_error(ctx));
	tagpu_free_reads(stream);
	free(len); free(rc); free(off); free(words);
	res->n_e = n_e;
	return (struct mini_hash_t *)res;
}

void add_cnt_to_graph(struct asm_graph_t *g, struct mini_hash_t *kmer_table)
{
	struct cov_result *res = (struct cov_result *)kmer_table;
	if (!res || res->n_e != g->n_e)
		TAGPU_FATAL("add_cnt_to_graph: the table does not belong to this graph");
	for (gint_t e = 0; e < g->n_e; ++e)
		g->edges[e].count = res->count[e];
	free(res->count);
	free(res);
}

int KMC_arg_kmer_count(int argc, char *argv[])
{
	(void)argc; (void)argv;
	fprintf(stderr, "[tagpu] KMC_arg_kmer_count is not on the hot path and is not implemented\n");
	return -1;
}

/* ------------------------------------------------------------------ intra-node rendezvous for the multi-GPU phases
 * The ranks of a multi-GPU build are processes on ONE box (include/tagpu.h "multi-GPU").  Between the phases of a step they
 * need a barrier and an all-gather of a handful of counters; doing that with a collective library costs tens of
 * microseconds and a device round trip each time, several times per step of a few milliseconds.  This is the same thing
 * over a POSIX shared-memory segment: a sense-reversing barrier (spin, then yield) and double-buffered slots of 8 values
 * per rank.  Rank 0 creates the segment under a name the host program distributes (torch.distributed broadcast, MPI, ...). */
#include <sched.h>

#define TAGPU_SHM_VALUES 8
#define TAGPU_SHM_MAX_RANKS 64

struct tagpu_shm_seg {
	volatile uint32_t count, gen;
	uint32_t world, pad;
	volatile uint64_t slot[2][TAGPU_SHM_MAX_RANKS][TAGPU_SHM_VALUES];
};

struct tagpu_shm {
	struct tagpu_shm_seg *seg;
	int rank, world, parity, owner;
	char name[128];
};

struct tagpu_shm *tagpu_shm_open(const char *name, int rank, int world)
{
	if (world < 1 || world > TAGPU_SHM_MAX_RANKS || rank < 0 || rank >= world)
		return NULL;
	struct tagpu_shm *s = calloc(1, sizeof(*s));
	snprintf(s->name, sizeof(s->name), "/%s", name[0] == '/' ? name + 1 : name);
	s->rank = rank;
	s->world = world;
	int fd = -1;
	if (rank == 0) {
		shm_unlink(s->name);
		fd = shm_open(s->name, O_CREAT | O_EXCL | O_RDWR, 0600);
		if (fd < 0 || ftruncate(fd, sizeof(struct tagpu_shm_seg)) != 0) {
			perror("tagpu_shm_open");
			free(s);
			return NULL;
		}
		s->owner = 1;
	} else {
		for (int tries = 0; tries < 200000 && fd < 0; ++tries) {          /* up to ~20 s for rank 0 to get there */
			fd = shm_open(s->name, O_RDWR, 0600);
			struct stat st;
			if (fd >= 0 && (fstat(fd, &st) != 0 || (size_t)st.st_size < sizeof(struct tagpu_shm_seg))) {
				close(fd);
				fd = -1;
			}
			if (fd < 0) usleep(100);
		}
		if (fd < 0) {
			free(s);
			return NULL;
		}
	}
	s->seg = mmap(NULL, sizeof(struct tagpu_shm_seg), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
	close(fd);
	if (s->seg == MAP_FAILED) {
		free(s);
		return NULL;
	}
	if (rank == 0) {
		s->seg->world = (uint32_t)world;                          /* (a fresh segment is zero-filled) */
		__sync_synchronize();
	}
	return s;
}

void tagpu_shm_barrier(struct tagpu_shm *s)
{
	struct tagpu_shm_seg *g = s->seg;
	const uint32_t gen = g->gen;
	if (__sync_add_and_fetch(&g->count, 1) == (uint32_t)s->world) {
		g->count = 0;
		__sync_synchronize();
		g->gen = gen + 1;
	} else {
		for (unsigned spins = 0; g->gen == gen; ++spins) {
			if (spins < 4096) __builtin_ia32_pause();
			else sched_yield();
		}
	}
	__sync_synchronize();
}

/* all[r * n + i] = value i of rank r; n <= 8.  Contains one barrier: on return every rank has contributed. */
int tagpu_shm_allgather(struct tagpu_shm *s, const uint64_t *mine, int n, uint64_t *all)
{
	if (n < 0 || n > TAGPU_SHM_VALUES)
		return -1;
	struct tagpu_shm_seg *g = s->seg;
	const int p = s->parity;
	for (int i = 0; i < n; ++i)
		g->slot[p][s->rank][i] = mine[i];
	__sync_synchronize();
	tagpu_shm_barrier(s);
	for (int r = 0; r < s->world; ++r)
		for (int i = 0; i < n; ++i)
			all[r * n + i] = g->slot[p][r][i];
	s->parity = p ^ 1;    /* the slots of this call are only overwritten two calls later: everybody has read them by then */
	return 0;
}

void tagpu_shm_close(struct tagpu_shm *s)
{
	if (!s)
		return;
	munmap(s->seg, sizeof(struct tagpu_shm_seg));
	if (s->owner)
		shm_unlink(s->name);
	free(s);
}
