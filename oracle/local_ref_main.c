/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Driver that calls the reference's build_local_assembly_graph
 * (/root/reference/src/kmer_build.c:991-1044; declared /root/reference/src/assembly_graph.h:160-162) the way
 * get_local_assembly does (/root/reference/src/barcode_resolve2.c:2100-2102): n_files = 1, one R1 / R2 pair, the
 * global graph g0 and the two flanking edges.  The reference has no sub-command that reaches this function on its
 * own, so this 40-line main() is linked with the UNMODIFIED reference objects (oracle/build_ref.sh):
 *
 *   TA_local_ref   every reference object (+ oracle/kmc_cpu.c for the absent libkmc.a)  -> golden vectors
 *   TA_local_gpu   the same objects with kmer_build.o's build_local_assembly_graph localised, so the call binds to
 *                  libtagpu.so                                                           -> drop-in test
 *
 *   usage: TA_local_* <g0.bin> <e1> <e2> <lk> <R1.fq> <R2.fq> <work_dir> <out.bin> [n_threads]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "assembly_graph.h"

int main(int argc, char **argv)
{
	if (argc < 9) {
		fprintf(stderr, "usage: %s g0.bin e1 e2 lk R1.fq R2.fq work_dir out.bin [n_threads]\n", argv[0]);
		return 2;
	}
	struct asm_graph_t g0, lg;
	load_asm_graph(&g0, argv[1]);
	gint_t e1 = atol(argv[2]), e2 = atol(argv[3]);
	int lk = atoi(argv[4]), n_threads = argc > 9 ? atoi(argv[9]) : 4;
	char *r1 = argv[5], *r2 = argv[6];
	if (e1 < 0 || e1 >= g0.n_e || e2 < 0 || e2 >= g0.n_e) {
		fprintf(stderr, "edge ids out of range (n_e = %ld)\n", (long)g0.n_e);
		return 2;
	}
	build_local_assembly_graph(lk, n_threads, 32, 1, &r1, &r2, argv[7], &lg, &g0, e1, e2);
	test_asm_graph(&lg);
	save_asm_graph(&lg, argv[8]);
	printf("local graph: k=%d n_v=%ld n_e=%ld\n", lg.ksize, (long)lg.n_v, (long)lg.n_e);
	return 0;
}
