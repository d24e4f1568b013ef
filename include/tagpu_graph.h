/*
 * Layout-compatible re-declaration of the reference types that cross the stage boundary, for callers
 * (and for libtagpu's own host code) that are built without the reference headers.  When compiling
 * INSIDE the reference tree include its own assembly_graph.h / attribute.h instead — the field order,
 * types and sizes below mirror
 *   struct asm_node_t / asm_edge_t / asm_graph_t   /root/reference/src/assembly_graph.h:52-95
 *   struct barcode_hash_t                          /root/reference/src/barcode_hash.h:10-17
 *   struct opt_proc_t                              /root/reference/src/attribute.h:49-71
 *   gint_t                                         /root/reference/src/attribute.h:38
 * tests/test_abi.py checks sizeof/offsetof against the reference headers when they are present.
 */
#ifndef TAGPU_GRAPH_H
#define TAGPU_GRAPH_H

#include <pthread.h>
#include <stdint.h>

typedef int64_t gint_t;

struct barcode_hash_t {
	uint32_t size;
	uint32_t n_item;
	uint32_t n_unique;
	uint64_t *keys;
	uint32_t *cnts;
};

struct asm_node_t {
	gint_t rc_id;   /* id of the reverse-complement node */
	gint_t deg;     /* out degree */
	gint_t *adj;    /* out edges, individually malloc'ed */
};

struct asm_edge_t {
	uint64_t count;     /* sum of (k+1)-mer counts on the edge */
	uint32_t *seq;      /* 2-bit bases, 16 per word, individually malloc'ed */
	uint32_t seq_len;
	uint32_t n_holes;   /* must stay adjacent to seq_len: save_asm_graph writes both as one 8-byte field */
	uint32_t *p_holes;
	uint32_t *l_holes;
	gint_t source;
	gint_t target;
	gint_t rc_id;
	pthread_mutex_t lock;
	struct barcode_hash_t *barcodes;
	struct barcode_hash_t barcodes_scaf;
	struct barcode_hash_t barcodes_cov;
};

struct asm_graph_t {
	int ksize;
	int bin_size;
	uint32_t aux_flag;
	gint_t n_v, n_e;
	struct asm_node_t *nodes;
	struct asm_edge_t *edges;
	void *candidates;   /* khash_t(pair_contig_count) * in the reference; left untouched by the builder */
};

struct opt_proc_t {
	int n_threads;
	int hash_size;
	int k0;
	int k1;
	int k2;
	int split_len;
	int lib_type;
	int n_files;
	char **files_1, **files_2, **files_I, **var;
	int metagenomics;
	char *out_dir;
	char *in_file;
	char *in_fasta;
	char *in_fastg;
	char *in_contig_file;
	int mmem;
	char *lc;
	int lk;
	int log_level;
	char *bx_str;
	int thresh;
};

#endif /* TAGPU_GRAPH_H */
