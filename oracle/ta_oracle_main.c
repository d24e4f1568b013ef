/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Command-line front end of the CPU oracle.
 *
 *   ta_oracle build_0 -1 R1.fq[,..] -2 R2.fq[,..] -k0 <k> [-t n] [-ci c] -o <dir>
 *       same file products as the reference's build_0 sub-command
 *       (/root/reference/src/process.c:703-709): KMC_<k+1>_count.kmc_{pre,suf}
 *       and graph_k_<k>_level_0.bin, all from the restatement.
 *   ta_oracle canon <graph.bin> <out.txt> [mode]
 *   ta_oracle count -1 .. -2 .. -k0 <k> [-t n]    (timing only: KMC stage)
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>

#include "ta_oracle.h"

static double now(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static int split_list(char *s, char **out, int max)
{
	int n = 0;
	for (char *t = strtok(s, ","); t && n < max; t = strtok(NULL, ","))
		out[n++] = t;
	return n;
}

int main(int argc, char **argv)
{
	if (argc >= 4 && !strcmp(argv[1], "canon")) {
		int mode = argc > 4 ? atoi(argv[4]) : 0;
		return ora_canon_dump(argv[2], argv[3], mode) ? 1 : 0;
	}
	if (argc < 2 || (strcmp(argv[1], "build_0") && strcmp(argv[1], "count"))) {
		fprintf(stderr, "usage: ta_oracle build_0|count -1 R1 -2 R2 -k0 k [-t n] [-ci c] -o dir | canon in.bin out.txt [mode]\n");
		return 2;
	}
	char *files[128];
	int n1 = 0, n2 = 0, k = 45, t = 4, ci = 2;
	const char *out = ".";
	char *f1[64], *f2[64];
	for (int i = 2; i + 1 < argc; i += 2) {
		if (!strcmp(argv[i], "-1")) n1 = split_list(argv[i + 1], f1, 64);
		else if (!strcmp(argv[i], "-2")) n2 = split_list(argv[i + 1], f2, 64);
		else if (!strcmp(argv[i], "-k0")) k = atoi(argv[i + 1]);
		else if (!strcmp(argv[i], "-t")) t = atoi(argv[i + 1]);
		else if (!strcmp(argv[i], "-ci")) ci = atoi(argv[i + 1]);
		else if (!strcmp(argv[i], "-o")) out = argv[i + 1];
		else if (!strcmp(argv[i], "-l")) ; /* accepted and ignored like the reference's build_0 */
	}
	int nf = 0;
	for (int i = 0; i < n1; ++i) files[nf++] = f1[i];
	for (int i = 0; i < n2; ++i) files[nf++] = f2[i];
	if (!nf) { fprintf(stderr, "no input files\n"); return 2; }
	mkdir(out, 0777);

	double t0 = now();
	uint8_t *stream;
	int64_t n = ora_load_reads(nf, files, &stream);
	double t1 = now();
	uint64_t *hi, *lo, n_inst, n_dist;
	uint32_t *cnt;
	int64_t n_solid = ora_count_stream(stream, (uint64_t)n, k + 1, ci, t, &hi, &lo, &cnt, &n_inst, &n_dist);
	double t2 = now();
	ora_free(stream);
	if (n_solid < 0) { fprintf(stderr, "unsupported k\n"); return 2; }
	printf("K=%d instances=%lu distinct=%lu solid=%ld load_s=%.3f count_s=%.3f\n", k + 1,
	       (unsigned long)n_inst, (unsigned long)n_dist, (long)n_solid, t1 - t0, t2 - t1);
	if (!strcmp(argv[1], "count"))
		return 0;
	ora_write_kmc_db(out, k + 1, ci, n_solid, hi, lo, cnt);
	uint64_t sum = 0;
	for (int64_t i = 0; i < n_solid; ++i) sum += cnt[i];
	struct ora_graph *g = ora_build_graph(k, n_solid, hi, lo, cnt);
	double t3 = now();
	printf("Number of kmer: %ld\nNumber of nodes: %ld; Number of edges: %ld\n"
	       "Number of (k+1)-mer on edge: %lu\nsum_solid_count = %lu graph_s=%.3f\n",
	       (long)g->n_kmer, (long)g->n_v, (long)g->n_e, (unsigned long)g->n_kp1_on_edge,
	       (unsigned long)sum, t3 - t2);
	char path[4096];
	snprintf(path, sizeof(path), "%s/graph_k_%d_level_0.bin", out, k);
	ora_graph_save_bin(g, path);
	ora_graph_free(g);
	ora_free(hi); ora_free(lo); ora_free(cnt);
	return 0;
}
