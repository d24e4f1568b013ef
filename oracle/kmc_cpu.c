/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's
 * cpu_baseline / --impl reference leg). Never linked into libtagpu.so.
 *
 * CPU restatement of the KMC stage of TuringAssembler's build_graph_from_scratch
 * (/root/reference/src/kmer_build.c:733-737): the two symbols the reference
 * imports from the absent libs/KMC/libkmc.a (/root/reference/include/kmc_skipping.h:8-11).
 *
 * PARITY STATUS: "parity unpinned" at THIS boundary only — the reference tree
 * holds neither the KMC source nor any golden vector for the cutoff / counter
 * width.  The decision recorded there is implemented here:
 * ci = 2, no upper cutoff, exact u32 counts, counter_size = 4, canonical =
 * min(fwd, rc).  Everything downstream of the database files IS pinned: the
 * files written here are parsed by the unmodified reference reader
 * (/root/reference/src/KMC_reader.c:22-150,204-334) in oracle/_ref/TA_ref.
 *
 * Database layout written: SURVEY.md App. B (KMC_VER 0x200 branch of
 * /root/reference/src/KMC_reader.c:50-74), 1 bin, lut_prefix_length = (K%4)+4,
 * signature_length = 0, counter_size = 4.
 */
#define _GNU_SOURCE
#include <fcntl.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include "ta_oracle.h"

#define ORA_BIN_BITS 12
#define ORA_CHUNK ((size_t)1 << 20)

/* A/a=0 C/c=1 G/g=2 T/t=3, else 4: /root/reference/src/utils.c:26-43 */
static uint8_t ora_nt4[256];
static pthread_once_t nt4_once = PTHREAD_ONCE_INIT;
static void init_nt4(void)
{
	memset(ora_nt4, 4, sizeof(ora_nt4));
	ora_nt4['A'] = ora_nt4['a'] = 0;
	ora_nt4['C'] = ora_nt4['c'] = 1;
	ora_nt4['G'] = ora_nt4['g'] = 2;
	ora_nt4['T'] = ora_nt4['t'] = 3;
}

struct ora_job {
	const uint8_t *seq;
	size_t n;
	int K, ci, n_threads;
	size_t next_chunk, n_chunks;
	int next_bin;
};

struct ora_result {
	uint64_t *hi, *lo;
	uint32_t *count;
	size_t n_solid;
	uint64_t n_instances, n_distinct;
};

#define KEYT uint64_t
#define SUF _64
#define KEY_HI(x) 0
#include "kmc_core_impl.h"
#undef KEYT
#undef SUF
#undef KEY_HI

#define KEYT unsigned __int128
#define SUF _128
#define KEY_HI(x) ((uint64_t)((x) >> 64))
#include "kmc_core_impl.h"
#undef KEYT
#undef SUF
#undef KEY_HI

static int g_cutoff = 2;
void ora_set_cutoff(int ci) { g_cutoff = ci < 1 ? 1 : ci; }

int64_t ora_count_stream(const uint8_t *seq, uint64_t n, int K, int ci, int n_threads,
			 uint64_t **hi, uint64_t **lo, uint32_t **count,
			 uint64_t *n_instances, uint64_t *n_distinct)
{
	if (K < 7 || K > 64)
		return -1;
	pthread_once(&nt4_once, init_nt4);
	struct ora_job job = { seq, n, K, ci, n_threads < 1 ? 1 : n_threads, 0, 0, 0 };
	struct ora_result res;
	if (K <= 32)
		count_stream_64(&job, &res);
	else
		count_stream_128(&job, &res);
	*hi = res.hi; *lo = res.lo; *count = res.count;
	if (n_instances) *n_instances = res.n_instances;
	if (n_distinct) *n_distinct = res.n_distinct;
	return (int64_t)res.n_solid;
}

void ora_free(void *p) { free(p); }

/* ---------------------------------------------------------------- FASTQ/FASTA */

struct file_job {
	const char *path;
	uint8_t *stream; /* sequence lines, each followed by '\n' */
	size_t n;
};

static uint8_t *slurp(const char *path, size_t *n_out)
{
	int fd = open(path, O_RDONLY);
	if (fd < 0) { perror(path); exit(1); }
	unsigned char magic[2] = { 0, 0 };
	ssize_t got = read(fd, magic, 2);
	(void)got;
	lseek(fd, 0, SEEK_SET);
	size_t cap, n = 0;
	uint8_t *buf;
	if (magic[0] == 0x1f && magic[1] == 0x8b) {
		gzFile gz = gzdopen(fd, "rb");
		gzbuffer(gz, 1 << 20);
		cap = (size_t)1 << 26;
		buf = malloc(cap);
		for (;;) {
			if (n == cap) { cap *= 2; buf = realloc(buf, cap); }
			size_t want = cap - n > ((size_t)1 << 30) ? ((size_t)1 << 30) : cap - n;
			int r = gzread(gz, buf + n, (unsigned)want);
			if (r <= 0) break;
			n += (size_t)r;
		}
		gzclose(gz);
	} else {
		struct stat st;
		fstat(fd, &st);
		cap = (size_t)st.st_size + 1;
		buf = malloc(cap);
		while (n < (size_t)st.st_size) {
			ssize_t r = read(fd, buf + n, (size_t)st.st_size - n);
			if (r <= 0) break;
			n += (size_t)r;
		}
		close(fd);
	}
	*n_out = n;
	return buf;
}

/* 4-line FASTQ records, sequence = 2nd line (cf. /root/reference/src/get_buffer.c:339-348);
 * FASTA: every non-'>' line, lines of one record joined. */
static void *file_worker(void *raw)
{
	struct file_job *fj = raw;
	size_t n;
	uint8_t *txt = slurp(fj->path, &n);
	uint8_t *out = malloc(n + 2);
	size_t o = 0, p = 0;
	if (n && txt[0] == '>') {
		while (p < n) {
			uint8_t *nl = memchr(txt + p, '\n', n - p);
			size_t e = nl ? (size_t)(nl - txt) : n;
			if (txt[p] == '>') {
				if (o && out[o - 1] != '\n') out[o++] = '\n';
			} else {
				size_t len = e - p;
				if (len && txt[e - 1] == '\r') --len;
				memcpy(out + o, txt + p, len);
				o += len;
			}
			p = e + 1;
		}
		if (o && out[o - 1] != '\n') out[o++] = '\n';
	} else {
		size_t line = 0;
		while (p < n) {
			uint8_t *nl = memchr(txt + p, '\n', n - p);
			size_t e = nl ? (size_t)(nl - txt) : n;
			if ((line & 3) == 1) {
				size_t len = e - p;
				if (len && txt[e - 1] == '\r') --len;
				memcpy(out + o, txt + p, len);
				o += len;
				out[o++] = '\n';
			}
			++line;
			p = e + 1;
		}
	}
	free(txt);
	fj->stream = out;
	fj->n = o;
	return NULL;
}

int64_t ora_load_reads(int n_files, char **files, uint8_t **stream)
{
	struct file_job *fj = calloc(n_files, sizeof(*fj));
	pthread_t *th = calloc(n_files, sizeof(pthread_t));
	for (int i = 0; i < n_files; ++i) {
		fj[i].path = files[i];
		pthread_create(th + i, NULL, file_worker, fj + i);
	}
	size_t tot = 0;
	for (int i = 0; i < n_files; ++i) {
		pthread_join(th[i], NULL);
		tot += fj[i].n;
	}
	uint8_t *s = malloc(tot + 1);
	size_t o = 0;
	for (int i = 0; i < n_files; ++i) {
		memcpy(s + o, fj[i].stream, fj[i].n);
		o += fj[i].n;
		free(fj[i].stream);
	}
	free(fj); free(th);
	*stream = s;
	return (int64_t)tot;
}

/* ---------------------------------------------------------------- KMC database writer (App. B) */

struct kmc_hdr {
	uint32_t kmer_length, mode, counter_size, lut_prefix_length, signature_length;
	uint32_t min_count, max_count;
	uint64_t total_kmers;
	uint8_t both_strands, pad8[3];
	uint32_t pad32[6];
	uint32_t kmc_ver;
} __attribute__((packed));

int ora_write_kmc_db(const char *working_dir, int K, int ci, int64_t n,
		     const uint64_t *hi, const uint64_t *lo, const uint32_t *count)
{
	char path[4096];
	const int p = (K % 4) + 4;
	const int suf_bases = K - p, suf_bytes = suf_bases / 4;
	const uint64_t n_lut = (uint64_t)1 << (2 * p);
	uint64_t *lut = calloc(n_lut + 1, sizeof(uint64_t));

	snprintf(path, sizeof(path), "%s/KMC_%d_count.kmc_suf", working_dir, K);
	FILE *fs = fopen(path, "wb");
	if (!fs) { perror(path); return -1; }
	setvbuf(fs, NULL, _IOFBF, 1 << 22);
	fwrite("KMCS", 1, 4, fs);
	uint8_t rec[32];
	for (int64_t i = 0; i < n; ++i) {
		unsigned __int128 x = ((unsigned __int128)hi[i] << 64) | lo[i];
		uint64_t prefix = (uint64_t)(x >> (2 * suf_bases));
		++lut[prefix + 1];
		for (int j = 0; j < suf_bytes; ++j) /* most significant suffix byte first */
			rec[j] = (uint8_t)(x >> (8 * (suf_bytes - 1 - j)));
		memcpy(rec + suf_bytes, count + i, 4);
		fwrite(rec, 1, suf_bytes + 4, fs);
	}
	fwrite("KMCS", 1, 4, fs);
	fclose(fs);
	for (uint64_t i = 0; i < n_lut; ++i)
		lut[i + 1] += lut[i];

	snprintf(path, sizeof(path), "%s/KMC_%d_count.kmc_pre", working_dir, K);
	FILE *fp = fopen(path, "wb");
	if (!fp) { perror(path); return -1; }
	fwrite("KMCP", 1, 4, fp);
	fwrite(lut, sizeof(uint64_t), n_lut + 1, fp);
	uint32_t sigmap[2] = { 0, 0 }; /* 4^0 + 1 entries */
	fwrite(sigmap, sizeof(uint32_t), 2, fp);
	struct kmc_hdr h;
	memset(&h, 0, sizeof(h));
	h.kmer_length = K;
	h.mode = 0;
	h.counter_size = 4;
	h.lut_prefix_length = p;
	h.signature_length = 0;
	h.min_count = ci;
	h.max_count = 0xffffffffu;
	h.total_kmers = (uint64_t)n;
	h.both_strands = 1;
	h.kmc_ver = 0x200;
	fwrite(&h, sizeof(h), 1, fp);
	uint32_t header_offset = sizeof(h);
	fwrite(&header_offset, 4, 1, fp);
	fwrite("KMCP", 1, 4, fp);
	fclose(fp);
	free(lut);
	return 0;
}

/* ---------------------------------------------------------------- the libkmc.a entry points */

/* /root/reference/include/kmc_skipping.h:8-9; called with ksize = k + 1 and
 * files = files_1 ++ files_2 (/root/reference/src/kmer_build.c:733-737). */
int KMC_build_kmer_database(int ksize, const char *working_dir, int n_threads,
			    int mmem, int n_files, char **files)
{
	(void)mmem;
	uint8_t *stream;
	int64_t n = ora_load_reads(n_files, files, &stream);
	uint64_t *hi, *lo, n_inst, n_dist;
	uint32_t *cnt;
	int64_t n_solid = ora_count_stream(stream, (uint64_t)n, ksize, g_cutoff, n_threads,
					   &hi, &lo, &cnt, &n_inst, &n_dist);
	free(stream);
	if (n_solid < 0)
		return -1;
	fprintf(stderr, "[oracle-kmc] K=%d instances=%lu distinct=%lu solid=%ld (ci=%d)\n",
		ksize, (unsigned long)n_inst, (unsigned long)n_dist, (long)n_solid, g_cutoff);
	int rc = ora_write_kmc_db(working_dir, ksize, g_cutoff, n_solid, hi, lo, cnt);
	free(hi); free(lo); free(cnt);
	return rc;
}

/* /root/reference/include/kmc_skipping.h:11 — never called by the reference. */
int KMC_arg_kmer_count(int argc, char *argv[])
{
	(void)argc; (void)argv;
	fprintf(stderr, "KMC_arg_kmer_count: not part of the hot path\n");
	return -1;
}
