/*
 * tagpu — B200-native k-mer counting + de Bruijn graph construction behind TuringAssembler's own
 * C entry points.  C ABI of libtagpu.so (plain pointers and sizes; no CUDA or torch types).
 *
 * Level 1/2 symbols are the reference's own names and signatures, so the reference links against
 * libtagpu.so instead of libs/KMC/libkmc.a + its own kmer_build.c versions.
 * There is no CPU fallback: every entry point needs a CUDA device and exits loudly without one.
 */
#ifndef TAGPU_H
#define TAGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ level 1: library boundary
 * replaces libs/KMC/libkmc.a — /root/reference/include/kmc_skipping.h:8-11
 * (call sites /root/reference/src/kmer_build.c:726,731,736,800,807,943,999; src/resolve_big.c:240).
 * Counts canonical ksize-mers of all files on the GPU and writes
 * working_dir/KMC_<ksize>_count.kmc_pre/.kmc_suf in the layout /root/reference/src/KMC_reader.c:22-150
 * parses (KMC_VER 0x200, counter_size 4).  Returns 0. */
int KMC_build_kmer_database(int ksize, const char *working_dir, int n_threads,
			    int mmem, int n_files, char **files);
/* /root/reference/include/kmc_skipping.h:11 — exported for link compatibility; never called by the reference. */
int KMC_arg_kmer_count(int argc, char *argv[]);

/* ------------------------------------------------------------------ level 2: stage boundary
 * struct asm_graph_t / struct opt_proc_t are the reference's types
 * (/root/reference/src/assembly_graph.h:52-95, /root/reference/src/attribute.h:49-71); tagpu_graph.h
 * re-declares them layout-compatibly for callers that do not include the reference headers. */
struct asm_graph_t;
struct opt_proc_t;

/* /root/reference/src/kmer_build.h:17-19, body /root/reference/src/kmer_build.c:714-786 */
void build_graph_from_scratch(int ksize, int n_threads, int mmem, int n_files,
			      char **files_1, char **files_2, char *work_dir,
			      struct asm_graph_t *g);
/* /root/reference/src/kmer_build.h:20-22, body /root/reference/src/kmer_build.c:788-837 (edge counts left 0) */
void build_graph_from_scratch_without_count(int ksize, int n_threads, int mmem, int n_files,
					    char **files_1, char **files_2, char *work_dir,
					    struct asm_graph_t *g);
/* /root/reference/src/assembly_graph.h:141, body /root/reference/src/kmer_build.c:839-845 */
void build_initial_graph(struct opt_proc_t *opt, int ksize, struct asm_graph_t *g);

/* /root/reference/src/assembly_graph.h:160-162, body /root/reference/src/kmer_build.c:991-1044 — the same stage re-entered per
 * gap by local assembly, with the two flanking edges of the global graph forced in */
void build_local_assembly_graph(int ksize, int n_threads, int mmem, int n_files,
				char **files_1, char **files_2, char *work_dir, struct asm_graph_t *g,
				struct asm_graph_t *g0, int64_t e1, int64_t e2);

/* ------------------------------------------------------------------ coverage recount
 * /root/reference/src/coverage/kmer_count.h:9-10, bodies /root/reference/src/coverage/kmer_count.c:198-240 and :113-135, called
 * back to back by build_coverage_process (/root/reference/src/process.c:823-835).  kmer_count_on_edges counts the 31-mers of
 * the reads of opt->files_1/files_2 that occur on the edges of g, on the GPU; what it returns is opaque to the caller (the
 * reference's struct mini_hash_t * is never looked into outside these two functions) and is consumed — and released — by
 * add_cnt_to_graph, which leaves the same counts in g->edges[].count as the reference. */
struct mini_hash_t;
struct mini_hash_t *kmer_count_on_edges(struct opt_proc_t *opt, struct asm_graph_t *g);
void add_cnt_to_graph(struct asm_graph_t *g, struct mini_hash_t *kmer_table);

/* ------------------------------------------------------------------ native API (what bench.py, the tests and the
 * level-1/2 wrappers call) */
typedef struct tagpu_ctx tagpu_ctx;

struct tagpu_stats {
	uint64_t n_instances;    /* valid (k+1)-mer windows (SURVEY.md §8d metric numerator) */
	uint64_t n_distinct;     /* distinct canonical (k+1)-mers */
	uint64_t n_solid;        /* ... with count >= cutoff */
	uint64_t sum_solid;      /* sum of their counts */
	uint64_t n_kmers;        /* canonical k-mers ("Number of kmer", kmer_build.c:758) */
	uint64_t n_v, n_e;       /* "Number of nodes / edges" (kmer_build.c:763): n_v = 2 * #node k-mers */
	uint64_t n_seq_words;    /* total 32-bit words of edge sequence */
	uint64_t n_kp1_on_edge;  /* "Number of (k+1)-mer on edge" (kmer_build.c:772) */
	uint64_t error;          /* 0 = ok; bit set = internal invariant violated (TAGPU_ERR_*) */
	uint64_t jump_rounds;    /* pointer-jumping rounds executed */
	uint64_t gpu_launches;   /* kernels launched by the last build */
	float ms_count, ms_graph, ms_total; /* CUDA-event times of the last build (device side) */
	uint32_t record_bytes;   /* size of one super-k-mer record of the last count stage (16 / 32) */
	uint64_t n_records_local, n_records_peer; /* records the counting kernel of THIS rank read from its own regions / from the
	                          * other ranks' regions over NVLink peer loads (multi-GPU builds; peer = 0 on one GPU) */
};

/* device < 0: current device.  Returns NULL (after printing the reason) if CUDA is unusable. */
tagpu_ctx *tagpu_create(int device);
void tagpu_destroy(tagpu_ctx *ctx);
/* cuda_stream: a cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream); 0 = the context's own stream */
void tagpu_set_stream(tagpu_ctx *ctx, void *cuda_stream);
/* min count of a solid (k+1)-mer; default 2 (SURVEY.md §8c decision) */
void tagpu_set_cutoff(tagpu_ctx *ctx, int ci);
/* 0 (default) = build edge counts; 1 = build_graph_from_scratch_without_count behaviour */
void tagpu_set_skip_counts(tagpu_ctx *ctx, int skip);
const char *tagpu_last_error(tagpu_ctx *ctx);
/* 1 = two-level graph stage: the solid (k+1)-mers are first contracted into paths inside the bucket groups of the count stage
 * (csrc/tagpu_contract.cuh); same graph (tagpu_copy_kmers rebuilds the full k-mer table on demand: hidden k-mers never reach
 * the stage's own table).  Builds with contig garbage (local assembly) always use the one-level stage. */
void tagpu_set_contract(tagpu_ctx *ctx, int on);
/* per-kernel CUDA-event timing of the next builds: JSON {"kernel": {"ms": total, "launches": n}, ...} */
void tagpu_set_profile(tagpu_ctx *ctx, int on);
const char *tagpu_profile_json(tagpu_ctx *ctx);

/* Count + build from a flat byte stream already in device memory ('\n' or any non-ACGT byte between reads).
 * Results stay on the device until copied.  k = node k-mer size (17..63); returns 0 on success. */
int tagpu_build_device(tagpu_ctx *ctx, const uint8_t *d_seq, uint64_t n_bytes, int k);
/* Same from host memory: the stream is uploaded inside the call, in chunks overlapped with the partition pass
 * (pass pinned memory — e.g. from tagpu_load_reads — for the copy to be asynchronous). */
int tagpu_build_host(tagpu_ctx *ctx, const uint8_t *h_seq, uint64_t n_bytes, int k);

/* Packed read stream: the same stream with the ASCII -> 2-bit conversion already done on the host, in the tile layout the
 * CUDA kernels use in shared memory (per 8192 positions: 256 x u64 codes, 32 bases each, first base most significant,
 * then 256 x u32 invalid masks; 0.375 bytes per position).  tagpu_pack_stream(stream, n, packed, threads) fills a buffer
 * of tagpu_packed_bytes(n) bytes; the *_packed calls take it with n_positions = n and give bit-identical results to the
 * ASCII calls at 3/8 of the host-to-device traffic. */
uint64_t tagpu_packed_bytes(uint64_t n_positions);
int tagpu_pack_stream(const uint8_t *stream, uint64_t n_bytes, uint8_t *packed, int n_threads);
int tagpu_build_host_packed(tagpu_ctx *ctx, const uint8_t *h_packed, uint64_t n_positions, int k);
int tagpu_count_host_packed(tagpu_ctx *ctx, const uint8_t *h_packed, uint64_t n_positions, int K);
int tagpu_build_device_packed(tagpu_ctx *ctx, const uint8_t *d_packed, uint64_t n_positions, int k);
/* Counting stage only (what KMC_build_kmer_database needs); ksize_plus_1 = K = k + 1 */
int tagpu_count_device(tagpu_ctx *ctx, const uint8_t *d_seq, uint64_t n_bytes, int ksize_plus_1);
int tagpu_count_host(tagpu_ctx *ctx, const uint8_t *h_seq, uint64_t n_bytes, int ksize_plus_1);

/* build_local_assembly_graph on host buffers: reads as above; h_contigs = the flanking contigs as ACGT text, contig c at
 * [contig_off[c], contig_off[c] + contig_len[c]) and followed by a newline; contig_cov[c] = its coverage in the global
 * graph (__get_edge_cov).  n_contigs <= 4. */
int tagpu_build_local_host(tagpu_ctx *ctx, const uint8_t *h_reads, uint64_t n_bytes, int k, const uint8_t *h_contigs,
			   uint64_t n_contig_bytes, int n_contigs, const uint64_t *contig_off, const uint32_t *contig_len,
			   const double *contig_cov);

/* Coverage recount (the kernels behind kmer_count_on_edges / add_cnt_to_graph): reads as a host stream; edges either as flat
 * host arrays (e_off in 32-bit words, layout of struct tagpu_flat_graph) or, with e_len == NULL, the graph of the last
 * build, which is still on the device.  count_out[e] = the count the reference leaves in g->edges[e].count. */
int tagpu_coverage_recount_host(tagpu_ctx *ctx, const uint8_t *h_seq, uint64_t n_bytes, uint64_t n_e, const uint32_t *e_len,
				const uint64_t *e_off, const uint32_t *e_seq, uint64_t n_seq_words, const uint32_t *e_rc,
				uint64_t *count_out);

int tagpu_get_stats(tagpu_ctx *ctx, struct tagpu_stats *out);

/* Many local builds in flight (row f1; the reference's caller is a sequential loop over thousands of gaps,
 * /root/reference/src/build_bridge.c:1036-1062): n_jobs independent tagpu_build_local_host builds worked through by n_ctx
 * contexts (<= 64; kept across calls), each with its own CUDA stream and host thread, so that the small kernels of
 * different gaps overlap on the device.  Per job: inputs as for tagpu_build_local_host; cutoff (0 = 2); g = NULL or a
 * caller-owned struct to fill like build_local_assembly_graph does; rc and stats are set on return.  Returns 0 if every
 * job succeeded. */
struct tagpu_local_job {
	const uint8_t *reads;
	uint64_t n_bytes;
	int k, cutoff;
	const uint8_t *contigs;
	uint64_t n_contig_bytes;
	int n_contigs;
	const uint64_t *contig_off;
	const uint32_t *contig_len;
	const double *contig_cov;
	struct asm_graph_t *g;
	int rc;
	struct tagpu_stats stats;
};
int tagpu_build_local_batch(int device, int n_ctx, struct tagpu_local_job *jobs, int n_jobs);

/* Device -> host copies of the last build; caller allocates from tagpu_stats sizes.
 * Keys are (hi, lo) pairs of the 2-bit packed mer, first base most significant. Order is unspecified. */
int tagpu_copy_solid(tagpu_ctx *ctx, uint64_t *hi, uint64_t *lo, uint32_t *count);         /* n_solid each */
int tagpu_copy_kmers(tagpu_ctx *ctx, uint64_t *hi, uint64_t *lo, uint8_t *mask);           /* n_kmers each */
struct tagpu_flat_graph {
	uint64_t n_nodes;         /* node k-mers; node ids 2i (canonical) / 2i+1 (reverse complement) */
	uint64_t n_e, n_seq_words;
	uint8_t *node_mask;       /* [n_nodes]  low nibble: out-bases of 2i, high nibble: out-bases of 2i+1 */
	uint32_t *node_ebase;     /* [n_nodes]  first edge id of node 2i; edges of 2i+1 follow those of 2i */
	uint32_t *e_src, *e_dst, *e_rc, *e_len; /* [n_e] */
	uint64_t *e_count, *e_off;              /* [n_e] e_off = first word of the edge in e_seq */
	uint32_t *e_seq;                        /* [n_seq_words] base i at bits 2(i&15) of word i>>4 (assembly_graph.h:182-187) */
};
int tagpu_copy_graph(tagpu_ctx *ctx, struct tagpu_flat_graph *host_arrays);

/* Order-independent digests of the last build, computed on the device (csrc/tagpu_digest.cuh; same function on the CPU:
 * oracle/canon_dump.c ora_bin_digest, tests/_digest.py).  out[0..1] sum / xor over the solid (k+1)-mers this context holds,
 * out[2] how many, out[3] 1 = the whole set / 0 = this rank's share of a sharded multi-GPU build (shares add up);
 * out[4..5] sum / xor over all edges (length, count, 2-bit sequence), out[6] sum of lengths, out[7] sum of counts, out[8] n_e. */
int tagpu_digest(tagpu_ctx *ctx, uint64_t out[9]);

/* Host-side materialisation of the last build (tagpu_host.c) */
int tagpu_fill_asm_graph(tagpu_ctx *ctx, struct asm_graph_t *g);   /* individually malloc'ed seq/adj, SURVEY.md §8b */
/* the host half alone: flat arrays (e.g. from tagpu_copy_graph) -> struct asm_graph_t; no device involved */
int tagpu_fill_asm_graph_from_flat(const struct tagpu_flat_graph *h, int ksize, struct asm_graph_t *g);
void tagpu_free_asm_graph(struct asm_graph_t *g);                  /* frees a graph filled by tagpu_fill_asm_graph / the level-2 entry points */
int tagpu_write_graph_bin(tagpu_ctx *ctx, const char *path);       /* save_asm_graph layout, assembly_graph.c:1173-1248 */
int tagpu_write_kmc_db(tagpu_ctx *ctx, const char *working_dir);   /* KMC_<K>_count.kmc_pre/.kmc_suf of the last count */

/* ------------------------------------------------------------------ multi-GPU: one process per GPU
 * The (k+1)-mer space is hash-partitioned: every rank owns a contiguous range of minimizer buckets.  The reference has no
 * counterpart (single process, pthreads only); this is the north-star extension "each k-mer is hash-partitioned to an
 * owner GPU".  The host program (MPI, torch.distributed, ...) supplies rendezvous and barriers; per build the order is
 *
 *   once:   tagpu_dist_plan -> exchange the 64-byte handles of all ranks -> tagpu_dist_connect -> BARRIER
 *   step:   tagpu_dist_partition  -> BARRIER ->      (pass 1 over this rank's slice of the reads, into its OWN bucket regions)
 *           tagpu_dist_count      -> ALL-GATHER of the 4 stats values (doubles as the barrier) ->
 *                                                    (pass 2 over the buckets this rank owns; the counting kernel reads the
 *                                                     records of every rank through NVLink peer loads, overlapped with counting)
 *           tagpu_dist_contract   -> ALL-GATHER of the 4 path values (barrier again) ->
 *                                                    (two-level graph stage, level 1: every rank contracts ITS solid list
 *                                                     into unbranched paths, left in its arena)
 *           tagpu_dist_graph_paths -> BARRIER        (pulls all paths over NVLink peer loads, global stage on every rank;
 *                                                     the barrier keeps the next partition off regions still being read)
 *     or    tagpu_dist_graph                         (one-level: pulls all solid sets over NVLink, builds the graph on every
 *                                                     rank; also the fallback when a rank reports paths_out[3] == 0)
 *
 * All ranks must pass the same n_total_bytes (size of the WHOLE read stream), k and cutoff.  After tagpu_dist_graph the
 * stats / copy / write calls above describe the global result on every rank. */
#define TAGPU_IPC_HANDLE_BYTES 64
int tagpu_dist_plan(tagpu_ctx *ctx, int rank, int world, uint64_t n_total_bytes, int k, void *handle_out);
int tagpu_dist_connect(tagpu_ctx *ctx, const void *all_handles /* world x TAGPU_IPC_HANDLE_BYTES, rank order */);
int tagpu_dist_partition(tagpu_ctx *ctx, const uint8_t *d_seq_local, uint64_t n_local_bytes);
/* same, with this rank's slice still in (pinned) host memory: the upload is overlapped with pass 1 */
int tagpu_dist_partition_host(tagpu_ctx *ctx, const uint8_t *h_seq_local, uint64_t n_local_bytes);
/* same, with this rank's slice as a packed read stream (packed by the rank itself, positions from the slice start) */
int tagpu_dist_partition_host_packed(tagpu_ctx *ctx, const uint8_t *h_packed_local, uint64_t n_local_positions);
int tagpu_dist_count(tagpu_ctx *ctx, uint64_t stats_out[4]);
int tagpu_dist_graph(tagpu_ctx *ctx, const uint64_t *all_stats /* world x 4, rank order */, int with_graph);
/* paths_out = { paths, interior words, k-mers hidden inside paths, 1 if this rank contracted (0: use tagpu_dist_graph) } */
int tagpu_dist_contract(tagpu_ctx *ctx, uint64_t paths_out[4]);
/* with_graph: 1 = graph only (the solid set stays sharded over its owner ranks: tagpu_copy_solid / tagpu_copy_kmers /
 * tagpu_write_kmc_db fail); 3 = graph, and every rank also pulls the whole solid set */
int tagpu_dist_graph_paths(tagpu_ctx *ctx, const uint64_t *all_stats /* world x 4 */, const uint64_t *all_paths /* world x 4 */,
			   int with_graph);
/* The whole step above in ONE call, with the barriers and counter exchanges taken from a tagpu_shm segment (below): no
 * host-language round trip between the kernels of a step.  src_kind: 0 = device stream, 1 = pinned host ASCII stream,
 * 2 = pinned host packed stream (n = positions).  flags: 1 = build the graph, 2 = also gather the solid sets on every rank,
 * 4 = the previous step used the two-level stage (*used_paths was 1): barrier first, peers may still be pulling paths. */
struct tagpu_shm;
int tagpu_dist_step(tagpu_ctx *ctx, struct tagpu_shm *shm, const uint8_t *src, uint64_t n, int src_kind, int flags, int *used_paths);
/* Teardown / re-plan: every rank calls tagpu_dist_disconnect (unmaps the peers' arenas) -> BARRIER -> tagpu_dist_close or a
 * new tagpu_dist_plan (frees its own arena, which nobody maps any more). */
void tagpu_dist_disconnect(tagpu_ctx *ctx);
void tagpu_dist_close(tagpu_ctx *ctx);
/* [begin, end) of rank's share of a host read stream, cut at read boundaries ('\n') so no window is lost or doubled */
void tagpu_dist_shard_range(const uint8_t *h_seq, uint64_t n_bytes, int rank, int world, uint64_t *begin, uint64_t *end);

/* Intra-node rendezvous for the phases above (tagpu_host.c): a barrier and an all-gather of <= 8 values per rank over a
 * POSIX shared-memory segment — microseconds instead of a collective launch plus a device round trip.  Rank 0 creates the
 * segment; `name` is any string all ranks agree on (the host program distributes it once). */
struct tagpu_shm *tagpu_shm_open(const char *name, int rank, int world);
void tagpu_shm_barrier(struct tagpu_shm *s);
int tagpu_shm_allgather(struct tagpu_shm *s, const uint64_t *mine, int n, uint64_t *all /* world x n */);
void tagpu_shm_close(struct tagpu_shm *s);

/* FASTQ/FASTA(.gz) files -> pinned host stream of sequence lines joined by '\n' (free with tagpu_free_reads) */
int64_t tagpu_load_reads(int n_files, char **files, int n_threads, uint8_t **stream);
void tagpu_free_reads(uint8_t *stream);
/* The same ingest as an object, so that the upload can chase the parser: open (newline index + sizes: the stream length is
 * known), start (copy workers fill dst in stream order), ready (bytes of the stream prefix that are final; pass it with the
 * object to tagpu_set_source_progress before tagpu_build_host / tagpu_count_host), finish (join + release). */
/* Contig-file mode of the stage entry points (n_files < 0, /root/reference/src/kmer_build.c:722-731,779-781): the graph from
 * stream A (reads + contig file), without counts, then the edge counts from the solid (k+1)-mers of stream B (the reads alone,
 * assign_count_kedge_multi over a second database: on an edge -> the edge and its twin get the count, else ignored). */
int tagpu_build_host_counts_from(tagpu_ctx *ctx, const uint8_t *h_a, uint64_t n_a, const uint8_t *h_b, uint64_t n_b, int k);

/* Raw FASTQ files on the device (the files entry points use it for plain FASTQ when TAGPU_DEVICE_PARSE=1): the host reads the files into a pinned ring
 * (tagpu_raw_ring), sends the slots up (tagpu_raw_put; tagpu_raw_slot_wait tells when a slot may be overwritten) into one
 * device buffer (tagpu_raw_begin; file f at byte off[f], 256-aligned), and the device parses the records — the sequence is
 * line 2 of every 4, /root/reference/src/get_buffer.c:339-348 — and builds.  ends_nl[f]: the file's last byte is a newline.
 * tagpu_parse_fastq_device only parses and returns the stream length (h_out, if not NULL, receives the stream). */
void *tagpu_raw_ring(tagpu_ctx *ctx, size_t bytes);
int tagpu_raw_begin(tagpu_ctx *ctx, uint64_t total_bytes);
int tagpu_raw_put(tagpu_ctx *ctx, uint64_t dev_off, const void *host, uint64_t bytes, int slot);
int tagpu_raw_slot_wait(tagpu_ctx *ctx, int slot);
int64_t tagpu_parse_fastq_device(tagpu_ctx *ctx, int n_files, const uint64_t *off, const uint64_t *len, const uint8_t *ends_nl, uint8_t *h_out);
int tagpu_build_fastq_device(tagpu_ctx *ctx, int n_files, const uint64_t *off, const uint64_t *len, const uint8_t *ends_nl, int k, int with_graph);

struct tagpu_ingest;
struct tagpu_ingest *tagpu_ingest_open(int n_files, char **files, int n_threads);
uint64_t tagpu_ingest_bytes(const struct tagpu_ingest *ing);
void tagpu_ingest_start(struct tagpu_ingest *ing, uint8_t *dst);
uint64_t tagpu_ingest_ready(void *ing);
void tagpu_ingest_finish(struct tagpu_ingest *ing);
/* Fused variant (what the entry points use): open scans nothing, tagpu_ingest_bytes is an UPPER BOUND of the stream length
 * (half of a plain FASTQ file), the workers index + size + copy each chunk in one pass and pad the buffer behind the true
 * end with '\n' — positions that hold no window — so the consumer processes `bound` bytes.  finish_fused returns the true
 * length, or -1 if a file holds more sequence than the bound (then use tagpu_ingest_open: exact sizes, two passes). */
struct tagpu_ingest *tagpu_ingest_open_fused(int n_files, char **files, int n_threads);
int64_t tagpu_ingest_finish_fused(struct tagpu_ingest *ing);
void tagpu_set_source_progress(tagpu_ctx *ctx, uint64_t (*ready)(void *), void *arg);

#ifdef __cplusplus
}
#endif
#endif /* TAGPU_H */
