/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's
 * global-assembly k-mer stage, used as the parity checker for the CUDA path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this; libtagpu.so never does.
 */
#ifndef TA_ORACLE_H
#define TA_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- KMC stage (kmc_cpu.c) ---- */
void ora_set_cutoff(int ci);
/* Flat byte stream in, sorted canonical solid K-mers out (malloc'ed; free with ora_free).
 * Any byte outside ACGTacgt breaks the window.  Returns n_solid or -1. */
int64_t ora_count_stream(const uint8_t *seq, uint64_t n, int K, int ci, int n_threads,
			 uint64_t **hi, uint64_t **lo, uint32_t **count,
			 uint64_t *n_instances, uint64_t *n_distinct);
/* FASTQ/FASTA(.gz) files -> sequence lines joined by '\n' (malloc'ed). */
int64_t ora_load_reads(int n_files, char **files, uint8_t **stream);
int ora_write_kmc_db(const char *working_dir, int K, int ci, int64_t n,
		     const uint64_t *hi, const uint64_t *lo, const uint32_t *count);
void ora_free(void *p);
int KMC_build_kmer_database(int ksize, const char *working_dir, int n_threads,
			    int mmem, int n_files, char **files);
int KMC_arg_kmer_count(int argc, char *argv[]);

/* ---- graph stage (dbg_oracle.c) ---- */
struct ora_graph {
	int ksize;
	int64_t n_kmer;   /* canonical k-mers */
	uint64_t *khi, *klo; /* sorted */
	uint8_t *mask;    /* App. A.4 */
	int64_t n_v, n_e;
	int64_t *node_rc, *node_deg, *node_adj_off, *node_adj; /* adj flattened */
	int64_t *e_src, *e_dst, *e_rc;
	uint64_t *e_count;
	uint32_t *e_len;
	uint64_t *e_seq_off; /* offset in u32 words into e_seq */
	uint32_t *e_seq;
	uint64_t n_kp1_on_edge;
};
/* solid (k+1)-mers (sorted or not) -> masks -> nodes -> unitigs -> counts */
struct ora_graph *ora_build_graph(int k, int64_t n_solid, const uint64_t *hi,
				  const uint64_t *lo, const uint32_t *count);
/* build_local_assembly_graph (/root/reference/src/kmer_build.c:991-1044): same, plus the "garbage" of the flanking contigs
 * (one byte per base, codes 0-3) and their coverages old_cov[c] = __get_edge_cov(g0->edges + e_c, g0->ksize) */
struct ora_graph *ora_build_graph_local(int k, int64_t n_solid, const uint64_t *hi, const uint64_t *lo, const uint32_t *count,
					int n_contigs, const uint8_t *const *contig, const uint32_t *contig_len, const double *old_cov);
void ora_graph_free(struct ora_graph *g);
int ora_graph_save_bin(const struct ora_graph *g, const char *path);

/* ---- canonical form of a graph .bin (canon_dump.c), App. D ---- */
/* Loads an App. C .bin, checks its structural invariants, writes the sorted canonical
 * lines to out_path.  Returns 0 on success, >0 = number of invariant violations, <0 I/O. */
int ora_canon_dump(const char *bin_path, const char *out_path, int with_topology);
/* Order-independent digest of the edges of a .bin (same function as libtagpu's tagpu_digest, csrc/tagpu_digest.cuh):
 * out = { sum, xor, sum of lengths, sum of counts, live edges }.  Returns 0 or <0 on I/O / format errors. */
int ora_bin_digest(const char *bin_path, uint64_t out[5]);

/* ---- coverage recount (cov_oracle.c): kmer_count_on_edges + add_cnt_to_graph, /root/reference/src/coverage/kmer_count.c ---- */
int ora_coverage_recount(const uint8_t *stream, uint64_t n, int64_t n_e, const uint32_t *e_len, const uint64_t *e_off,
			 const uint32_t *e_seq, const int64_t *e_rc, uint64_t *count_out);

#ifdef __cplusplus
}
#endif
#endif
