/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Driver that calls the reference's build_graph_from_scratch in its contig-file mode
 * (/root/reference/src/kmer_build.c:714-786 with n_files < 0: have_contig_file :677-679, the two databases :722-731, the
 * count pass over the second one :695-710,779-780).  No sub-command of the reference sets n_files < 0, so this main() is
 * linked with the UNMODIFIED reference objects (oracle/build_ref.sh):
 *
 *   TA_contig_ref   every reference object (+ oracle/kmc_cpu.c for the absent libkmc.a)  -> golden vectors
 *   TA_contig_gpu   the same objects with kmer_build.o's stage entry points localised, so the call binds to libtagpu.so
 *
 *   usage: TA_contig_* <k> <R1.fq> <R2.fq> <contigs.fa> <work_dir> <out.bin> [n_threads] [without_count]
 * files_1 = { R1 }, files_2 = { R2, contigs, R1, R2 }: the graph from R1 + R2 + contigs, the counts from R1 + R2.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "assembly_graph.h"
#include "kmer_build.h"

int main(int argc, char **argv)
{
	if (argc < 7) {
		fprintf(stderr, "usage: %s k R1.fq R2.fq contigs.fa work_dir out.bin [n_threads] [without_count]\n", argv[0]);
		return 2;
	}
	const int k = atoi(argv[1]), n_threads = argc > 7 ? atoi(argv[7]) : 4, without = argc > 8 && atoi(argv[8]);
	char *files_1[1] = { argv[2] };
	char *files_2[4] = { argv[3], argv[4], argv[2], argv[3] };
	struct asm_graph_t g;
	if (without)
		build_graph_from_scratch_without_count(k, n_threads, 32, -1, files_1, files_2, argv[5], &g);
	else
		build_graph_from_scratch(k, n_threads, 32, -1, files_1, files_2, argv[5], &g);
	test_asm_graph(&g);
	save_asm_graph(&g, argv[6]);
	printf("contig-mode graph: k=%d n_v=%ld n_e=%ld\n", g.ksize, (long)g.n_v, (long)g.n_e);
	return 0;
}
