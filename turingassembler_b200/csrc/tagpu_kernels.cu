// libtagpu device layer: owns the CUDA context state, launches the sm_100a kernels and exposes the
// native half of include/tagpu.h.  The reference-facing entry points and all file I/O live in
// tagpu_host.c (plain C), which only calls the functions declared in include/tagpu.h.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include <cuda_runtime.h>
#include <unistd.h>

#include "../../include/tagpu.h"
#include "tagpu_contract.cuh"
#include "tagpu_count.cuh"
#include "tagpu_coverage.cuh"
#include "tagpu_digest.cuh"
#include "tagpu_extract.cuh"
#include "tagpu_fastq.cuh"
#include "tagpu_graph.cuh"
#include "tagpu_key.cuh"

constexpr int TAGPU_UPLOAD_CHUNKS = 32;      // pieces the host read stream is uploaded in (copy overlapped with pass 1)
constexpr int TAGPU_UPLOAD_CHUNKS_MAX = 32;

struct Buf {
	void *p = nullptr;
	size_t cap = 0;
};

struct DistState;
constexpr int TAGPU_RAW_SLOTS_MAX = 64;

struct tagpu_ctx {
	int device = 0;
	cudaStream_t stream = nullptr, own_stream = nullptr, copy_stream = nullptr;
	cudaEvent_t ev_chunk[TAGPU_UPLOAD_CHUNKS_MAX];
	const uint8_t *h_src = nullptr;    // host source of the read stream while its upload is pending (tagpu_*_host calls)
	uint64_t (*src_ready)(void *) = nullptr;   // optional: bytes of the host stream's prefix that are final (the ingest may still
	void *src_ready_arg = nullptr;             // be writing behind it); consulted before every chunk upload of the NEXT host build
	bool src_packed = false;           // the read stream of the current build is in the packed tile layout (tagpu_extract.cuh)
	int ci = 2, skip_counts = 0;
	int contract = 1;                  // two-level graph stage (tagpu_contract.cuh)
	bool contracted = false;           // the last graph was built that way (hidden k-mers are not in the table)
	bool solid_sharded = false;        // multi-GPU build that left the solid set with its owners (only the paths travelled)
	int log2_buckets = 0;              // of the last count stage
	int k = 0, K = 0, W = 0;
	char err[512] = { 0 };
	unsigned long long *d_ctr = nullptr, *h_ctr = nullptr;
	Buf regions, cursor, overflow, overflow_bucket, ext, ext_off, ext_count, cur_all, ext_all, pex, bsum, grp_end;
	uint64_t count_stream_bytes = 0;   // bytes of the WHOLE read stream of the current build (all ranks)
	int n_sm = 0, jump_grid = 0;
	// per-device launch state (a process may hold contexts on several devices)
	bool attr_done[3] = { false, false, false };
	bool attr_done_part[32] = {};                               // k_partition<W, TW, B, EXACT>: [(((W - 1) * 2 + (TW == 128)) * 4 + log2(B) - 2) * 2 + EXACT]
	int grid_s[3] = { 0, 0, 0 }, grid_m[3] = { 0, 0, 0 }, grid_l[3] = { 0, 0, 0 };
	uint64_t budget_n = 0;
	size_t budget = 0;                 // count_budget(): memory the count stage planned with for a stream of budget_n bytes
	uint64_t n_solid_local = 0;        // entries of the solid list THIS context holds (= st.n_solid unless the set is sharded)
	Buf chain_slot, grp_start, grp_desc, node_mask;
	// raw FASTQ on the device (tagpu_fastq.cuh): file bytes, per-block newline counts / bases, per-record bounds and offsets
	Buf raw, fq_cnt, fq_base, fq_tot, fq_lo, fq_hi, fq_len, fq_off;
	void *raw_ring = nullptr;          // pinned ring the host reads the files into (tagpu_raw_ring)
	size_t raw_ring_bytes = 0;
	cudaEvent_t ev_slot[TAGPU_RAW_SLOTS_MAX], ev_raw = nullptr;
	bool ev_slot_made = false;
	Buf seq, solid_key, solid_cnt, kt_keys, kt_mask, node_ord, node_slot, node_ebase, vL, vR, jump, vsucc,
		vedge, e_src, e_dst, e_rc, e_len, e_count, e_off, e_seq;
	uint32_t kt_slots = 0;
	bool have_count = false, have_graph = false;
	void *cur_solid_key = nullptr, *cur_solid_cnt = nullptr; // solid set the graph stage reads (local, or gathered from all ranks)
	// build_local_assembly_graph: (k+1)-mers of the flanking contigs appended behind the solid ones (count 0), and the contigs
	uint64_t n_garbage = 0, n_blocks = 0;                    // n_blocks: directory entries of the local solid list (0 = no directory)
	bool local_mode = false;
	Buf blocks, blk_defer, d_first, d_last, d_n, d_cnt, d_off, d_int, wlast, g_key, comb_key, comb_cnt, g_seq, hj_own, hj_bits, hj_list, hj_jump2;
	int n_contigs = 0;
	uint64_t contig_off[4] = { 0 };
	uint32_t contig_len[4] = { 0 };
	double contig_cov[4] = { 0 };
	struct DistState *dist = nullptr;
	tagpu_stats st;
	cudaEvent_t ev[4];
	uint64_t launches = 0;
	// optional per-kernel timing (tagpu_set_profile): one CUDA event pair around every launch
	int profile = 0;
	std::vector<cudaEvent_t> ev_pool;
	size_t ev_used = 0;
	struct ProfRec { const char *name; cudaEvent_t a, b; };
	std::vector<ProfRec> prof;
	std::string prof_json;
};

static cudaEvent_t prof_event(tagpu_ctx *ctx)
{
	if (ctx->ev_used == ctx->ev_pool.size()) {
		cudaEvent_t e;
		cudaEventCreate(&e);
		ctx->ev_pool.push_back(e);
	}
	return ctx->ev_pool[ctx->ev_used++];
}

struct ProfScope {
	tagpu_ctx *ctx;
	cudaEvent_t a = nullptr;
	const char *name;
	ProfScope(tagpu_ctx *c, const char *n) : ctx(c), name(n)
	{
		if (ctx->profile) { a = prof_event(ctx); cudaEventRecord(a, ctx->stream); }
	}
	~ProfScope()
	{
		if (ctx->profile) { cudaEvent_t b = prof_event(ctx); cudaEventRecord(b, ctx->stream); ctx->prof.push_back({ name, a, b }); }
	}
};

static void prof_finish(tagpu_ctx *ctx)
{
	ctx->prof_json = "{";
	if (ctx->profile) {
		std::map<std::string, std::pair<double, int>> agg;
		std::vector<std::string> order;
		for (auto &r : ctx->prof) {
			float ms = 0;
			cudaEventElapsedTime(&ms, r.a, r.b);
			if (!agg.count(r.name)) order.push_back(r.name);
			agg[r.name].first += ms;
			agg[r.name].second += 1;
		}
		bool first = true;
		for (auto &n : order) {
			char buf[256];
			snprintf(buf, sizeof(buf), "%s\"%s\": {\"ms\": %.6f, \"launches\": %d}", first ? "" : ", ", n.c_str(), agg[n].first, agg[n].second);
			ctx->prof_json += buf;
			first = false;
		}
	}
	ctx->prof_json += "}";
	ctx->prof.clear();
	ctx->ev_used = 0;
}

static int fail(tagpu_ctx *c, const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(c->err, sizeof(c->err), fmt, ap);
	va_end(ap);
	fprintf(stderr, "[tagpu] ERROR: %s\n", c->err);
	return -1;
}

#define CU(call)                                                                                         \
	do {                                                                                             \
		cudaError_t e_ = (call);                                                                 \
		if (e_ != cudaSuccess)                                                                   \
			return fail(ctx, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
	} while (0)

static int ensure(tagpu_ctx *ctx, Buf &b, size_t bytes, bool *grew = nullptr)
{
	if (grew) *grew = false;
	if (bytes <= b.cap && b.p) return 0;
	static const bool trace = getenv("TAGPU_TRACE_ALLOC") != nullptr;   // developer aid: a steady-state step must not allocate
	if (trace) fprintf(stderr, "tagpu: buffer at ctx+%ld grows %zu -> %zu bytes\n", (long)((char *)&b - (char *)ctx), b.cap, bytes);
	if (b.p) CU(cudaFree(b.p));
	b.p = nullptr;
	b.cap = 0;
	size_t want = bytes < 256 ? 256 : bytes;
	CU(cudaMalloc(&b.p, want));
	b.cap = want;
	if (grew) *grew = true;
	return 0;
}

// For buffers whose size follows a quantity that varies a little from build to build on the same input (the number of
// paths of the two-level graph stage and what derives from it): grow with head-room, so that a steady-state step never
// reallocates (a cudaFree in the middle of a step costs milliseconds).
static int ensure_slack(tagpu_ctx *ctx, Buf &b, size_t bytes)
{
	if (bytes <= b.cap && b.p) return 0;
	return ensure(ctx, b, bytes + bytes / 8 + 65536);
}

static int ensure_exact(tagpu_ctx *ctx, Buf &b, size_t bytes) { return ensure(ctx, b, bytes); }


extern "C" tagpu_ctx *tagpu_create(int device)
{
	int n_dev = 0;
	cudaError_t e = cudaGetDeviceCount(&n_dev);
	if (e != cudaSuccess || n_dev == 0) {
		fprintf(stderr, "[tagpu] ERROR: no CUDA device (%s); libtagpu has no CPU fallback\n",
			e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
		return nullptr;
	}
	tagpu_ctx *ctx = new tagpu_ctx();
	if (device < 0) cudaGetDevice(&device);
	ctx->device = device;
	if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
	    cudaMalloc(&ctx->d_ctr, CTR_TOTAL * sizeof(unsigned long long)) != cudaSuccess ||
	    cudaMallocHost(&ctx->h_ctr, CTR_TOTAL * sizeof(unsigned long long)) != cudaSuccess) {
		fprintf(stderr, "[tagpu] ERROR: cannot initialise device %d: %s\n", device, cudaGetErrorString(cudaGetLastError()));
		delete ctx;
		return nullptr;
	}
	ctx->stream = ctx->own_stream;
	cudaDeviceGetAttribute(&ctx->n_sm, cudaDevAttrMultiProcessorCount, device);
	for (int i = 0; i < 4; ++i) cudaEventCreate(&ctx->ev[i]);
	cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
	for (int i = 0; i < TAGPU_UPLOAD_CHUNKS_MAX; ++i) cudaEventCreateWithFlags(&ctx->ev_chunk[i], cudaEventDisableTiming);
	memset(&ctx->st, 0, sizeof(ctx->st));
	const char *ct = getenv("TAGPU_CONTRACT");
	if (ct) ctx->contract = atoi(ct);
	return ctx;
}

static void dist_release(tagpu_ctx *ctx);

extern "C" void tagpu_destroy(tagpu_ctx *ctx)
{
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	cudaDeviceSynchronize();
	dist_release(ctx);
	Buf *bufs[] = { &ctx->regions, &ctx->cursor, &ctx->overflow, &ctx->overflow_bucket, &ctx->ext, &ctx->ext_off, &ctx->ext_count, &ctx->cur_all, &ctx->ext_all, &ctx->pex, &ctx->bsum, &ctx->grp_end, &ctx->blocks, &ctx->blk_defer, &ctx->d_first, &ctx->d_last, &ctx->d_n, &ctx->d_cnt, &ctx->d_off, &ctx->d_int, &ctx->wlast, &ctx->g_key, &ctx->comb_key, &ctx->comb_cnt, &ctx->g_seq, &ctx->hj_own, &ctx->hj_bits, &ctx->hj_list, &ctx->hj_jump2, &ctx->chain_slot, &ctx->grp_start, &ctx->grp_desc, &ctx->node_mask, &ctx->seq, &ctx->solid_key, &ctx->solid_cnt, &ctx->kt_keys, &ctx->kt_mask,
			&ctx->node_ord, &ctx->node_slot, &ctx->node_ebase, &ctx->vL, &ctx->vR, &ctx->jump, &ctx->vsucc, &ctx->vedge,
			&ctx->e_src, &ctx->e_dst, &ctx->e_rc, &ctx->e_len, &ctx->e_count, &ctx->e_off, &ctx->e_seq,
			&ctx->raw, &ctx->fq_cnt, &ctx->fq_base, &ctx->fq_tot, &ctx->fq_lo, &ctx->fq_hi, &ctx->fq_len, &ctx->fq_off };
	for (Buf *b : bufs)
		if (b->p) cudaFree(b->p);
	cudaFree(ctx->d_ctr);
	cudaFreeHost(ctx->h_ctr);
	if (ctx->raw_ring) cudaFreeHost(ctx->raw_ring);
	if (ctx->ev_slot_made) {
		for (int i = 0; i < TAGPU_RAW_SLOTS_MAX; ++i) cudaEventDestroy(ctx->ev_slot[i]);
		cudaEventDestroy(ctx->ev_raw);
	}
	for (int i = 0; i < 4; ++i) cudaEventDestroy(ctx->ev[i]);
	for (int i = 0; i < TAGPU_UPLOAD_CHUNKS_MAX; ++i) cudaEventDestroy(ctx->ev_chunk[i]);
	cudaStreamDestroy(ctx->copy_stream);
	for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
	cudaStreamDestroy(ctx->own_stream);
	delete ctx;
}

// Progress callback for the NEXT host build / count: ready(arg) = bytes at the front of the host stream that are final.
// The chunked upload of pass 1 waits for it chunk by chunk, so parsing the read files and uploading them overlap
// (tagpu_host.c: tagpu_ingest_ready).  Only the ASCII host calls consult it; it is cleared after one use.
extern "C" void tagpu_set_source_progress(tagpu_ctx *ctx, uint64_t (*ready)(void *), void *arg)
{
	ctx->src_ready = ready;
	ctx->src_ready_arg = arg;
}

extern "C" void tagpu_set_stream(tagpu_ctx *ctx, void *s) { ctx->stream = s ? (cudaStream_t)s : ctx->own_stream; }
extern "C" void tagpu_set_cutoff(tagpu_ctx *ctx, int ci) { ctx->ci = ci < 1 ? 1 : ci; }
extern "C" void tagpu_set_skip_counts(tagpu_ctx *ctx, int skip) { ctx->skip_counts = skip; }
extern "C" const char *tagpu_last_error(tagpu_ctx *ctx) { return ctx->err; }
extern "C" void tagpu_set_profile(tagpu_ctx *ctx, int on) { ctx->profile = on; }
extern "C" void tagpu_set_contract(tagpu_ctx *ctx, int on) { ctx->contract = on; }
extern "C" const char *tagpu_profile_json(tagpu_ctx *ctx) { return ctx->prof_json.c_str(); }

// allow: error bits the caller handles itself (capacity overflows it can recover from by re-running a pass)
static int read_counters(tagpu_ctx *ctx, unsigned long long allow = 0)
{
	CU(cudaMemcpyAsync(ctx->h_ctr, ctx->d_ctr, CTR_TOTAL * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	if (ctx->h_ctr[CTR_ERROR] & ~allow)
		return fail(ctx, "device-side invariant violated (error bits 0x%llx)", ctx->h_ctr[CTR_ERROR]);
	return 0;
}

#define LAUNCH(kernel, grid, block, ...)                                                   \
	do {                                                                               \
		ProfScope ps_(ctx, #kernel);                                               \
		kernel<<<(grid), (block), 0, ctx->stream>>>(__VA_ARGS__);                  \
		++ctx->launches;                                                           \
		CU(cudaGetLastError());                                                    \
	} while (0)

// ------------------------------------------------------------------------------------------------ count stage (partitioned)
#define LAUNCH_SMEM(kernel, grid, block, smem, ...) LAUNCH_SMEM_NAMED(#kernel, kernel, grid, block, smem, __VA_ARGS__)
#define LAUNCH_SMEM_NAMED(name, kernel, grid, block, smem, ...)                            \
	do {                                                                               \
		ProfScope ps_(ctx, name);                                                  \
		kernel<<<(grid), (block), (smem), ctx->stream>>>(__VA_ARGS__);             \
		++ctx->launches;                                                           \
		CU(cudaGetLastError());                                                    \
	} while (0)

// Sizing of the bucket space for a read stream of n_total bytes shared by `world` ranks.  Every rank computes the same
// plan, so bucket -> owner and all region geometry agree without communication.  Capacities are ESTIMATES: a record
// that does not fit its bucket's region goes to the overflow list, and a single-GPU build whose overflow list or solid
// buffer turns out too small re-runs the pass in question with buffers sized from what the first attempt counted
// (partition_local / count_owned), so no input is refused for its shape.
static PartCfg plan_cfg(uint64_t n_total, int K, int world, uint32_t group_target, size_t rec_bytes, size_t budget_bytes)
{
	// bucket count: buckets are packed into groups of ~GROUP_TARGET windows by k_group_buckets; fine buckets keep the
	// packing tight although bucket sizes are skewed (CV ~0.9).  Guess: windows ~ stream bytes, a fifth of them distinct.
	uint64_t want = n_total / 8 / (group_target / 5 / 4) + 1;   // mean bucket ~ a quarter of a group
	int log2p = 10;
	while ((1ull << log2p) < want && log2p < 22) ++log2p;
	const uint64_t n_buckets = 1ull << log2p;
	const uint64_t n_source = n_total / world + 1;              // bytes each rank partitions
	PartCfg cfg;
	cfg.K = K;
	cfg.log2_buckets = log2p;
	// records per source: a super-k-mer holds ~(w + 1) / 2 of the w = K - 14 windows that can share a minimizer, and a
	// window takes >= 1 stream byte: n_source / max(2, (w + 1) / 2) overestimates by ~1.4x on 151 bp reads.
	const int w = K - TAGPU_MINIMIZER_M + 1;
	const uint64_t rec_est = n_source / (uint64_t)(w + 1 >= 4 ? (w + 1) / 2 : 2) + 1;
	// region capacity at each source: 4x the estimated mean bucket (CV ~0.9: a fraction of a percent of the records
	// spills), less if the regions would take more than the memory budget — the overflow list absorbs the difference
	uint64_t cap = rec_est * 4 / n_buckets + 64;
	if (budget_bytes && n_buckets * cap * rec_bytes > budget_bytes) {
		const uint64_t fit = budget_bytes / rec_bytes / n_buckets, floor_cap = rec_est * 3 / 2 / n_buckets + 64;
		cap = fit > floor_cap ? fit : floor_cap;
	}
	cfg.cap_records = (uint32_t)cap;
	// TAGPU_REGION_CAP overrides it (tests use a tiny value to drive everything through the overflow path)
	static const char *cap_env = getenv("TAGPU_REGION_CAP");
	if (cap_env && atoi(cap_env) > 0) cfg.cap_records = (uint32_t)atoi(cap_env);
	cfg.world = (uint32_t)world;
	cfg.per_rank = (uint32_t)((n_buckets + world - 1) / world);
	cfg.packed = 0;
	cfg.overflow_cap = (uint32_t)(rec_est / 8 + 4096);
	if (cap_env && atoi(cap_env) > 0) cfg.overflow_cap = (uint32_t)(rec_est + 4096);
	static const char *ov_env = getenv("TAGPU_OVERFLOW_CAP");       // tests: force the re-run with a grown overflow list
	if (ov_env && atoi(ov_env) > 0) cfg.overflow_cap = (uint32_t)atoi(ov_env);
	return cfg;
}

// device memory the count stage may plan with: what is free plus what this context's regions already hold.  Asked once
// per stream size (cudaMemGetInfo is a synchronous driver call of a few hundred microseconds).
static size_t count_budget(tagpu_ctx *ctx, uint64_t n)
{
	if (ctx->budget_n == n && ctx->budget) return ctx->budget;
	size_t free_b = 0, total_b = 0;
	if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); return 0; }
	ctx->budget_n = n;
	ctx->budget = (free_b + ctx->regions.cap) / 2;
	return ctx->budget;
}

// One pass-1 sweep over the stream with tiles of TW words (TileCfg).  Reads that are still in host memory (ctx->h_src; d_seq
// is our staging buffer then) are uploaded in chunks on a second stream and partitioned chunk by chunk, so that the PCIe
// copy hides the pass.  Chunks are counted in 256-word tiles, the unit of the packed stream layout; a launch covers the
// kernel tiles whose right halo word is already on the device: with an ASCII stream the copy runs 32 bytes ahead, with a
// packed stream (whole 3072-byte tiles) the last 256-word tile of a chunk waits for the next chunk.
template <int W, int TW, int B, bool EXACT>
static int partition_sweep(tagpu_ctx *ctx, const uint8_t *d_seq, uint64_t n, const PartCfg &cfg)
{
	typedef TileCfg<TW> T;
	const size_t smem1 = T::SMEM;
	bool *attr_done = ctx->attr_done_part + (((W - 1) * 2 + (TW == 256 ? 0 : 1)) * 4 + (B == 4 ? 0 : B == 8 ? 1 : B == 16 ? 2 : 3)) * 2 + (EXACT ? 1 : 0);
	if (!*attr_done) {
		CU((cudaFuncSetAttribute(k_partition<W, TW, B, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1)));
		*attr_done = true;
	}
	constexpr uint64_t F = TAGPU_TILE_WORDS / TW;                 // kernel tiles per 256-word tile
	const uint64_t n_tiles = (n + T::BASES - 1) / T::BASES, n_big = (n + TAGPU_TILE_BASES - 1) / TAGPU_TILE_BASES;
	if (n_tiles && !ctx->h_src) {
		LAUNCH_SMEM_NAMED("k_partition<W>", (k_partition<W, TW, B, EXACT>), (unsigned)n_tiles, T::THREADS, smem1, d_seq, n, 0u, cfg, (SkRec<W> *)ctx->regions.p,
			    (unsigned long long *)ctx->cursor.p, (SkRec<W> *)ctx->overflow.p, (uint32_t *)ctx->overflow_bucket.p, ctx->d_ctr);
	} else if (n_tiles) {
		const bool packed = cfg.packed != 0;
		uint64_t chunk_big = n_big / TAGPU_UPLOAD_CHUNKS + 1;
		const uint64_t min_big = packed ? 1366 : 512;              // >= 4 MB per copy: small inputs go up in one piece
		if (chunk_big < min_big) chunk_big = min_big;
		const uint8_t *h_src = ctx->h_src;
		ctx->h_src = nullptr;
		const uint64_t total = packed ? n_big * (uint64_t)TAGPU_PACKED_TILE_BYTES : n;
		uint64_t copied = 0, tile0 = 0, up_big = 0;                 // up_big: 256-word tiles whose upload has been issued
		for (int c = 0; tile0 < n_tiles; ++c) {
			up_big = up_big + chunk_big < n_big ? up_big + chunk_big : n_big;
			uint64_t want, tile1;
			if (packed) {
				want = up_big * (uint64_t)TAGPU_PACKED_TILE_BYTES;
				tile1 = up_big == n_big ? n_tiles : (up_big - 1) * F;
			} else {
				want = up_big == n_big ? n : up_big * TAGPU_TILE_BASES + 32 * TAGPU_RHALO_WORDS;
				tile1 = up_big == n_big ? n_tiles : up_big * F;
			}
			if (want > total) want = total;
			if (ctx->src_ready && want > copied)                    // the parser is still filling the host stream: wait for this chunk
				while (ctx->src_ready(ctx->src_ready_arg) < want) usleep(20);
			if (want > copied) CU(cudaMemcpyAsync((uint8_t *)ctx->seq.p + copied, h_src + copied, want - copied, cudaMemcpyHostToDevice, ctx->copy_stream));
			copied = want > copied ? want : copied;
			CU(cudaEventRecord(ctx->ev_chunk[c % TAGPU_UPLOAD_CHUNKS_MAX], ctx->copy_stream));
			CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_chunk[c % TAGPU_UPLOAD_CHUNKS_MAX], 0));
			if (tile1 > tile0)
				LAUNCH_SMEM_NAMED("k_partition<W>", (k_partition<W, TW, B, EXACT>), (unsigned)(tile1 - tile0), T::THREADS, smem1, d_seq, n, (uint32_t)tile0, cfg, (SkRec<W> *)ctx->regions.p,
					    (unsigned long long *)ctx->cursor.p, (SkRec<W> *)ctx->overflow.p, (uint32_t *)ctx->overflow_bucket.p, ctx->d_ctr);
			tile0 = tile1;
		}
	}
	return 0;
}

// Pass 1 over this rank's reads + the bucket sort of the records that overflowed their region.  Purely local.
template <int W>
static int partition_local(tagpu_ctx *ctx, const uint8_t *d_seq, uint64_t n, const PartCfg &cfg_in)
{
	typedef BucketCfg<W> BC;
	PartCfg cfg = cfg_in;
	cfg.packed = ctx->src_packed ? 1u : 0u;
	const uint32_t n_buckets = 1u << cfg.log2_buckets;
	bool *attr_done = ctx->attr_done;
	if (!attr_done[W]) {
		CU(cudaFuncSetAttribute(k_count_buckets<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BC::SMEM));
		attr_done[W] = true;
	}
	// 128-word tiles (TileCfg) unless TAGPU_TILE_WORDS=256 asks for the large ones (developer A/B)
	static const char *tile_env = getenv("TAGPU_TILE_WORDS");
	const bool small_tile = tile_env ? atoi(tile_env) != 256 : true;
	for (int attempt = 0;; ++attempt) {
	{ ProfScope ps_(ctx, "memset"); CU(cudaMemsetAsync(ctx->cursor.p, 0, (size_t)n_buckets * 8, ctx->stream)); }
	// B = the largest power of two <= w (tagpu_window_min): 64-bit keys have w = 4 .. 18, 128-bit keys w = 19 .. 50
	const int w = cfg.K - TAGPU_MINIMIZER_M + 1;
	int rc;
	if constexpr (W == 1) {
		if (w >= 16) rc = small_tile ? partition_sweep<W, 128, 16, false>(ctx, d_seq, n, cfg) : partition_sweep<W, 256, 16, false>(ctx, d_seq, n, cfg);
		else if (w >= 8) rc = small_tile ? partition_sweep<W, 128, 8, false>(ctx, d_seq, n, cfg) : partition_sweep<W, 256, 8, false>(ctx, d_seq, n, cfg);
		else rc = small_tile ? partition_sweep<W, 128, 4, false>(ctx, d_seq, n, cfg) : partition_sweep<W, 256, 4, false>(ctx, d_seq, n, cfg);
	} else {
		if (w == 32) rc = small_tile ? partition_sweep<W, 128, 32, true>(ctx, d_seq, n, cfg) : partition_sweep<W, 256, 32, true>(ctx, d_seq, n, cfg);
		else if (w > 32) rc = small_tile ? partition_sweep<W, 128, 32, false>(ctx, d_seq, n, cfg) : partition_sweep<W, 256, 32, false>(ctx, d_seq, n, cfg);
		else rc = small_tile ? partition_sweep<W, 128, 16, false>(ctx, d_seq, n, cfg) : partition_sweep<W, 256, 16, false>(ctx, d_seq, n, cfg);
	}
	if (rc) return -1;
	ctx->h_src = nullptr;                                       // (a second attempt reads the stream from the device)
	ctx->src_ready = nullptr;
	if (read_counters(ctx, TAGPU_ERR_BUCKET_OVERFLOW)) return -1;
	if (!(ctx->h_ctr[CTR_ERROR] & TAGPU_ERR_BUCKET_OVERFLOW)) break;
	// the overflow list was too small: CTR_SPARE0 counted every record that wanted a place in it.  Grow and repeat the pass.
	const uint64_t need = ctx->h_ctr[CTR_SPARE0] + ctx->h_ctr[CTR_SPARE0] / 8 + 4096;
	if (ctx->dist || attempt || need > 0xffffffffull)
		return fail(ctx, "bucket overflow list too small (%llu records wanted, room for %u)", (unsigned long long)ctx->h_ctr[CTR_SPARE0], cfg.overflow_cap);
	if (ensure(ctx, ctx->overflow, need * sizeof(SkRec<W>)) || ensure(ctx, ctx->overflow_bucket, need * 4)) return -1;
	cfg.overflow_cap = (uint32_t)need;
	{ ProfScope ps_(ctx, "memset"); CU(cudaMemsetAsync(ctx->d_ctr, 0, CTR_TOTAL * sizeof(unsigned long long), ctx->stream)); }
	}
	ctx->st.n_instances = ctx->h_ctr[CTR_INSTANCES];
	const uint64_t n_over = ctx->h_ctr[CTR_SPARE0];
	if (n_over) {
		if (n_over * sizeof(SkRec<W>) > ctx->ext.cap) {
			if (ctx->dist) return fail(ctx, "bucket overflow area too small (%llu records)", (unsigned long long)n_over);
			if (ensure(ctx, ctx->ext, n_over * sizeof(SkRec<W>))) return -1;
		}
		{ ProfScope ps_(ctx, "memset"); CU(cudaMemsetAsync(ctx->ext_count.p, 0, (size_t)n_buckets * 4, ctx->stream)); }
		LAUNCH(k_overflow_hist, (unsigned)((n_over + 255) / 256), 256, (const uint32_t *)ctx->overflow_bucket.p, n_over, (uint32_t *)ctx->ext_count.p);
		const uint32_t n_ob = (n_buckets + 1023) / 1024;
		if (ensure(ctx, ctx->bsum, ((size_t)n_ob + 1) * 8)) return -1;       // scratch: block totals of the scan
		LAUNCH(k_overflow_scan_blocks, n_ob, 1024, (const uint32_t *)ctx->ext_count.p, (uint32_t *)ctx->ext_off.p, (uint32_t *)ctx->bsum.p, n_buckets);
		LAUNCH(k_overflow_scan_tops, 1, 1024, (uint32_t *)ctx->bsum.p, n_ob, (uint32_t *)ctx->ext_off.p, n_buckets);
		LAUNCH(k_overflow_scan_finish, n_ob, 1024, (uint32_t *)ctx->ext_count.p, (uint32_t *)ctx->ext_off.p, (const uint32_t *)ctx->bsum.p, n_buckets);
		LAUNCH(k_overflow_scatter<W>, (unsigned)((n_over + 255) / 256), 256, (const SkRec<W> *)ctx->overflow.p,
		       (const uint32_t *)ctx->overflow_bucket.p, n_over, (const uint32_t *)ctx->ext_off.p, (uint32_t *)ctx->ext_count.p,
		       (SkRec<W> *)ctx->ext.p);
	}
	return 0;
}

template <int W>
static int count_owned_once(tagpu_ctx *ctx, const PartCfg &cfg, const CountPeers<W> &peers, uint32_t first_bucket, uint64_t solid_cap);

// Pass 2 over the buckets [first_bucket, first_bucket + n_owned) — all of them when world == 1 — reading every source's
// records through `peers`.  solid_cap bounds the output arrays.
template <int W>
static int count_owned(tagpu_ctx *ctx, const PartCfg &cfg, const CountPeers<W> &peers, uint32_t first_bucket, uint64_t solid_cap)
{
	for (int attempt = 0;; ++attempt) {
		const int rc = count_owned_once<W>(ctx, cfg, peers, first_bucket, solid_cap);
		if (rc <= 0) return rc;
		// rc == 1: more solid (k+1)-mers than the buffers hold.  The records are still in the regions: size the buffers
		// from the count of this attempt and run pass 2 again (single GPU; the arena of a multi-GPU build is fixed).
		if (ctx->dist || attempt)
			return fail(ctx, "solid (k+1)-mer buffer too small (%llu > %llu)", (unsigned long long)ctx->st.n_solid, (unsigned long long)solid_cap);
		solid_cap = ctx->st.n_solid + ctx->st.n_solid / 16 + 4096;
		if (ensure(ctx, ctx->solid_key, solid_cap * sizeof(Key<W>)) || ensure(ctx, ctx->solid_cnt, solid_cap * 4)) return -1;
		const int reset[] = { CTR_DISTINCT, CTR_SOLID, CTR_SUM_SOLID, CTR_BLOCKS, CTR_REC_LOCAL, CTR_REC_PEER };
		for (int c : reset) CU(cudaMemsetAsync(ctx->d_ctr + c, 0, 8, ctx->stream));
	}
}

// one attempt: 0 = done, -1 = error, 1 = solid_cap exceeded (nothing else wrong)
template <int W>
static int count_owned_once(tagpu_ctx *ctx, const PartCfg &cfg, const CountPeers<W> &peers, uint32_t first_bucket, uint64_t solid_cap)
{
	typedef BucketCfg<W> BC;
	const uint32_t n_buckets = 1u << cfg.log2_buckets, world = cfg.world;
	const uint32_t n_owned = first_bucket >= n_buckets ? 0u : (n_buckets - first_bucket < cfg.per_rank ? n_buckets - first_bucket : cfg.per_rank);
	const uint32_t group_max = BC::SUB_MAX / world < BC::GROUP_MAX ? BC::SUB_MAX / world : BC::GROUP_MAX;
	const uint32_t n_scan_blocks = (n_owned + TAGPU_SCAN_BLOCK - 1) / TAGPU_SCAN_BLOCK;
	// owned windows <= windows of the whole stream <= stream bytes: upper bound of the group ids
	const uint64_t n_groups_cap = ctx->count_stream_bytes / BC::GROUP_TARGET + 2;
	if (ensure(ctx, ctx->cur_all, ((size_t)cfg.per_rank * world + 1) * 8) || ensure(ctx, ctx->ext_all, ((size_t)cfg.per_rank * world + 1) * 4) ||
	    ensure(ctx, ctx->pex, ((size_t)cfg.per_rank + 1) * 8) || ensure(ctx, ctx->bsum, ((size_t)n_scan_blocks + 1) * 8) ||
	    ensure(ctx, ctx->grp_start, n_groups_cap * 4) || ensure(ctx, ctx->grp_end, n_groups_cap * 4) ||
	    ensure(ctx, ctx->grp_desc, n_groups_cap * sizeof(GroupDesc)) ||
	    ensure(ctx, ctx->blocks, (2 * n_groups_cap + 4096) * sizeof(SolidBlock)))
		return -1;
	const uint32_t blocks_cap = (uint32_t)(2 * n_groups_cap + 4096);
	if (n_owned) {
		{ ProfScope ps_(ctx, "memset"); CU(cudaMemsetAsync(ctx->grp_start.p, 0xff, n_groups_cap * 4, ctx->stream)); }
		LAUNCH(k_pull_cursors<W>, n_scan_blocks, TAGPU_SCAN_BLOCK, peers, world, first_bucket, n_owned, n_buckets, cfg.cap_records,
		       (unsigned long long *)ctx->cur_all.p, (uint32_t *)ctx->ext_all.p, (unsigned long long *)ctx->pex.p, (unsigned long long *)ctx->bsum.p,
		       first_bucket / (cfg.per_rank ? cfg.per_rank : 1u), ctx->d_ctr);
		LAUNCH(k_scan_blocks, 1, 1024, (unsigned long long *)ctx->bsum.p, n_scan_blocks, (uint32_t)BC::GROUP_TARGET, ctx->d_ctr);
		LAUNCH(k_mark_groups, n_scan_blocks, TAGPU_SCAN_BLOCK, (const unsigned long long *)ctx->pex.p, (const unsigned long long *)ctx->bsum.p, n_owned,
		       (uint32_t)BC::GROUP_TARGET, (uint32_t *)ctx->grp_start.p, (uint32_t *)ctx->grp_end.p);
		LAUNCH(k_group_desc, (unsigned)((n_groups_cap + 255) / 256), 256, (const uint32_t *)ctx->grp_start.p, (const uint32_t *)ctx->grp_end.p,
		       (const unsigned long long *)ctx->cur_all.p, world, group_max, (const unsigned long long *)ctx->d_ctr, (GroupDesc *)ctx->grp_desc.p);
		// persistent CTAs, but never more than there can be groups: a tiny build (local assembly: a few hundred KB of reads)
		// then occupies a few SMs and the builds of other contexts overlap with it (tagpu_build_local_batch)
		const uint64_t full_grid = (uint64_t)BC::CTAS_PER_SM * ctx->n_sm;
		static const bool full_grids = getenv("TAGPU_FULL_GRIDS") != nullptr;     // developer A/B: always the whole device
		const unsigned count_grid = (unsigned)(!full_grids && n_groups_cap < full_grid ? n_groups_cap : full_grid);
		LAUNCH_SMEM(k_count_buckets<W>, count_grid, BC::THREADS, BC::SMEM, peers, world, first_bucket, cfg.cap_records,
			    (const unsigned long long *)ctx->cur_all.p, (const uint32_t *)ctx->ext_all.p, (const uint32_t *)ctx->grp_start.p,
			    (const uint32_t *)ctx->grp_end.p, (const GroupDesc *)ctx->grp_desc.p, group_max, cfg.K, (uint32_t)ctx->ci, (Key<W> *)ctx->solid_key.p, (uint32_t *)ctx->solid_cnt.p,
			    (unsigned long long)solid_cap, (SolidBlock *)ctx->blocks.p, blocks_cap, ctx->d_ctr);
	}
	if (read_counters(ctx)) return -1;
	ctx->st.n_distinct = ctx->h_ctr[CTR_DISTINCT];
	ctx->st.n_solid = ctx->h_ctr[CTR_SOLID];
	ctx->st.sum_solid = ctx->h_ctr[CTR_SUM_SOLID];
	ctx->n_blocks = ctx->h_ctr[CTR_BLOCKS];
	ctx->n_solid_local = ctx->st.n_solid;
	ctx->st.n_records_local = ctx->h_ctr[CTR_REC_LOCAL];
	ctx->st.n_records_peer = ctx->h_ctr[CTR_REC_PEER];
	ctx->st.record_bytes = sizeof(SkRec<W>);
	ctx->log2_buckets = cfg.log2_buckets;
#ifdef TAGPU_TIMING
	{
		const unsigned long long *t = ctx->h_ctr + CTR_JUMP_FLAGS;
		const double tot = (double)(t[48] + t[49] + t[50] + t[51] + t[52] + t[53] + t[56]);
		fprintf(stderr, "[tagpu timing] k_count_buckets warp-cycles: setup %.1f%%  stage+items %.1f%%  insert %.1f%%  wait at barrier %.1f%%  "
				"harvest A %.1f%%  B %.1f%%  copy-out %.1f%%; class iterations %llu (failed %llu), groups %llu\n",
			100.0 * t[48] / tot, 100.0 * t[56] / tot, 100.0 * t[49] / tot, 100.0 * t[50] / tot, 100.0 * t[52] / tot, 100.0 * t[53] / tot,
			100.0 * t[51] / tot, (unsigned long long)t[54], (unsigned long long)t[55], (unsigned long long)ctx->h_ctr[CTR_GROUPS]);
	}
#endif
	if (ctx->st.n_solid > solid_cap) return 1;
	ctx->cur_solid_key = ctx->solid_key.p;
	ctx->cur_solid_cnt = ctx->solid_cnt.p;
	ctx->have_count = true;
	return 0;
}

template <int W>
static int count_stage_partitioned(tagpu_ctx *ctx, const uint8_t *d_seq, uint64_t n)
{
	typedef BucketCfg<W> BC;
	const PartCfg cfg = plan_cfg(n, ctx->K, 1, BC::GROUP_TARGET, sizeof(SkRec<W>), count_budget(ctx, n));
	const uint32_t n_buckets = 1u << cfg.log2_buckets;
	ctx->count_stream_bytes = n;
	if (ensure(ctx, ctx->regions, (size_t)n_buckets * cfg.cap_records * sizeof(SkRec<W>)) || ensure(ctx, ctx->cursor, (size_t)n_buckets * 8) ||
	    ensure(ctx, ctx->overflow, (size_t)cfg.overflow_cap * sizeof(SkRec<W>)) ||
	    ensure(ctx, ctx->overflow_bucket, (size_t)cfg.overflow_cap * 4) || ensure(ctx, ctx->ext_off, (size_t)(n_buckets + 1) * 4) ||
	    ensure(ctx, ctx->ext_count, (size_t)n_buckets * 4))
		return -1;
	// an estimate (a twelfth of the windows solid, where windows <= stream bytes): count_owned re-runs pass 2 with larger
	// buffers if the input is shallower than that.  TAGPU_SOLID_CAP: tests force the re-run.
	uint64_t solid_cap = n / 12 + (1u << 20);
	if (solid_cap > n / (uint64_t)ctx->ci + 1) solid_cap = n / (uint64_t)ctx->ci + 1;      // (never more than can exist)
	static const char *sc_env = getenv("TAGPU_SOLID_CAP");
	if (sc_env && atoll(sc_env) > 0) solid_cap = (uint64_t)atoll(sc_env);
	if (ctx->solid_key.cap / sizeof(Key<W>) > solid_cap && ctx->solid_cnt.cap / 4 > solid_cap && !sc_env)
		solid_cap = ctx->solid_key.cap / sizeof(Key<W>) < ctx->solid_cnt.cap / 4 ? ctx->solid_key.cap / sizeof(Key<W>) : ctx->solid_cnt.cap / 4;
	if (ensure(ctx, ctx->solid_key, solid_cap * sizeof(Key<W>)) || ensure(ctx, ctx->solid_cnt, solid_cap * 4)) return -1;
	if (partition_local<W>(ctx, d_seq, n, cfg)) return -1;
	CountPeers<W> peers;
	memset(&peers, 0, sizeof(peers));
	peers.regions[0] = (const SkRec<W> *)ctx->regions.p;
	peers.cursor[0] = (const unsigned long long *)ctx->cursor.p;
	peers.ext[0] = (const SkRec<W> *)ctx->ext.p;
	peers.ext_off[0] = (const uint32_t *)ctx->ext_off.p;
	return count_owned<W>(ctx, cfg, peers, 0, solid_cap);
}

// ------------------------------------------------------------------------------------------------ list ranking
// jump[cv] = (successor | TAGPU_TERM at a chain end, link weight) -> (terminal, distance to it) for every chain vertex
static int rank_lists(tagpu_ctx *ctx, unsigned long long *jump, uint32_t n_cv)
{
	unsigned long long *ctr = ctx->d_ctr;
	// all pointer-jumping rounds in one cooperative launch (grid-wide sync between rounds)
	if (!ctx->jump_grid) {
		int per_sm = 0;
		CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_jump_all, 512, 0));
		ctx->jump_grid = ctx->n_sm * (per_sm > 0 ? per_sm : 1);
	}
	int max_rounds = 40;
	// small inputs: plain pointer jumping (the arrays sit in L2); large ones: work-efficient list ranking
	static const char *rank_env = getenv("TAGPU_LIST_RANKING");      // "hj" / "wyllie" force one of them (tests)
	const bool hj = rank_env ? !strcmp(rank_env, "hj") : n_cv >= (48u << 20);
	if (!hj) {
		uint32_t n_cv_arg = n_cv;
		void *args[] = { &jump, &n_cv_arg, &ctr, &max_rounds };
		ProfScope ps_(ctx, "k_jump_all");
		// (a cooperative launch needs all its CTAs co-resident: a small list takes a small grid, so that the launches of
		// several contexts fit on the device side by side)
		static const bool full_grids = getenv("TAGPU_FULL_GRIDS") != nullptr;
		const unsigned need = (n_cv + 511u) / 512u, jg = !full_grids && need < (unsigned)ctx->jump_grid ? (need ? need : 1u) : (unsigned)ctx->jump_grid;
		CU(cudaLaunchCooperativeKernel((void *)k_jump_all, dim3(jg), dim3(512), args, 0, ctx->stream));
		++ctx->launches;
		return 0;
	}
	if (ensure_slack(ctx, ctx->hj_own, ((size_t)n_cv + 1) * 8) || ensure_slack(ctx, ctx->hj_bits, ((size_t)n_cv / 32 + 2) * 4) ||
	    ensure_slack(ctx, ctx->hj_list, ((size_t)n_cv + 1) * 4))
		return -1;
	unsigned long long *own = (unsigned long long *)ctx->hj_own.p;
	uint32_t *bits = (uint32_t *)ctx->hj_bits.p, *list = (uint32_t *)ctx->hj_list.p;
	{ ProfScope ps_(ctx, "memset"); CU(cudaMemsetAsync(own, 0xff, ((size_t)n_cv + 1) * 8, ctx->stream)); }
	{ ProfScope ps_(ctx, "memset"); CU(cudaMemsetAsync(ctr + CTR_CHAIN, 0, 8, ctx->stream)); }
	LAUNCH(k_hj_mark, (n_cv + 255) / 256, 256, jump, n_cv, bits, list, own, ctr);
	if (read_counters(ctx)) return -1;
	const uint32_t n_spl = (uint32_t)ctx->h_ctr[CTR_CHAIN];
	// (sized with head-room: the splitter count varies a little from build to build with the table layout)
	const size_t spl_cap = (size_t)n_spl + 1 > (size_t)n_cv / 16 ? (size_t)n_spl + 1 : (size_t)n_cv / 16;
	if (ensure(ctx, ctx->hj_jump2, spl_cap * 8)) return -1;
	unsigned long long *jump2 = (unsigned long long *)ctx->hj_jump2.p;
	if (n_spl) {
		LAUNCH(k_hj_walk, (n_spl + 255) / 256, 256, jump, bits, list, n_spl, own, jump2);
		uint32_t n_arg = n_spl;
		void *args[] = { &jump2, &n_arg, &ctr, &max_rounds };
		ProfScope ps_(ctx, "k_jump_all");
		CU(cudaLaunchCooperativeKernel((void *)k_jump_all, dim3(ctx->jump_grid), dim3(512), args, 0, ctx->stream));
		++ctx->launches;
	}
	LAUNCH(k_hj_finish, (n_cv + 255) / 256, 256, jump, n_cv, own, jump2);
	return 0;
}

// ------------------------------------------------------------------------------------------------ graph stage
template <int W>
static int graph_stage(tagpu_ctx *ctx)
{
	const int k = ctx->k;
	const uint64_t n_reads_solid = ctx->st.n_solid;
	const uint64_t n_solid = n_reads_solid + ctx->n_garbage;        // entries of the (k+1)-mer list: solid ones, then contig garbage
	// every solid (k+1)-mer touches two k-mers; in practice #k-mers ~ #solid, so 2.5x keeps the load near 0.4
	const uint64_t slots64 = (n_solid * 5) / 2 + 1024;
	if (slots64 > (1ull << 30)) return fail(ctx, "k-mer table would need %llu slots (> 2^30)", (unsigned long long)slots64);
	const uint32_t n_slots = (uint32_t)slots64;
	ctx->kt_slots = n_slots;
	const size_t mask_bytes = ((size_t)n_slots + 3) / 4 * 4;
	if (ensure(ctx, ctx->kt_keys, (size_t)n_slots * sizeof(Key<W>)) || ensure(ctx, ctx->kt_mask, mask_bytes) ||
	    ensure(ctx, ctx->node_ord, (size_t)n_slots * 4) || ensure(ctx, ctx->vL, (n_solid + 1) * 4) ||
	    ensure(ctx, ctx->vR, (n_solid + 1) * 4) || ensure(ctx, ctx->node_slot, (2 * n_solid + 1) * 4) ||
	    ensure(ctx, ctx->node_ebase, (2 * n_solid + 1) * 4) || ensure(ctx, ctx->chain_slot, (2 * n_solid + 1) * 4))
		return -1;
	KTab<W> t;
	t.keys = (Key<W> *)ctx->kt_keys.p;
	t.mask32 = (uint32_t *)ctx->kt_mask.p;
	t.n_slots = n_slots;
	{ ProfScope ps_(ctx, "memset"); CU(cudaMemsetAsync(t.keys, 0, (size_t)n_slots * sizeof(Key<W>), ctx->stream)); }
	{ ProfScope ps_(ctx, "memset"); CU(cudaMemsetAsync(t.mask32, 0, mask_bytes, ctx->stream)); }
	uint32_t *kind = (uint32_t *)ctx->node_ord.p, *node_slot = (uint32_t *)ctx->node_slot.p,
		 *node_ebase = (uint32_t *)ctx->node_ebase.p, *chain_slot = (uint32_t *)ctx->chain_slot.p,
		 *vL = (uint32_t *)ctx->vL.p, *vR = (uint32_t *)ctx->vR.p;
	unsigned long long *ctr = ctx->d_ctr;
	const Key<W> *solid = (const Key<W> *)ctx->cur_solid_key;

	if (n_reads_solid) LAUNCH(k_insert_kmers<W>, (unsigned)((n_reads_solid + 255) / 256), 256, solid, 0ull, n_reads_solid, 0, k, t, vL, vR, ctr);
	if (ctx->n_garbage)
		LAUNCH(k_insert_kmers<W>, (unsigned)((ctx->n_garbage + 255) / 256), 256, solid, n_reads_solid, n_solid, 1, k, t, vL, vR, ctr);
	LAUNCH(k_classify<W>, (n_slots + 1023) / 1024, 1024, t, kind, node_slot, node_ebase, chain_slot, ctr);
	if (read_counters(ctx)) return -1;
	const uint64_t n_nodes = ctx->h_ctr[CTR_NODES], n_e = ctx->h_ctr[CTR_EDGES], n_chain = ctx->h_ctr[CTR_CHAIN];
	ctx->st.n_kmers = ctx->h_ctr[CTR_KMERS];
	ctx->st.n_v = 2 * n_nodes;
	ctx->st.n_e = n_e;
	if (n_e > 0xfffffff0ull) return fail(ctx, "too many edges (%llu)", (unsigned long long)n_e);
	const uint32_t n_cv = (uint32_t)(2 * n_chain);
	const uint64_t seq_cap = (n_e * (uint64_t)k + 2 * n_solid) / 16 + n_e + 16;
	if (ensure(ctx, ctx->jump, ((size_t)n_cv + 1) * 8) || ensure(ctx, ctx->vsucc, ((size_t)n_cv + 1) * 4) ||
	    ensure(ctx, ctx->vedge, ((size_t)n_cv + 1) * 4) || ensure(ctx, ctx->e_src, (n_e + 1) * 4) ||
	    ensure(ctx, ctx->e_dst, (n_e + 1) * 4) || ensure(ctx, ctx->e_rc, (n_e + 1) * 4) || ensure(ctx, ctx->e_len, (n_e + 1) * 4) ||
	    ensure(ctx, ctx->e_count, (n_e + 1) * 8) || ensure(ctx, ctx->e_off, (n_e + 1) * 8) || ensure(ctx, ctx->e_seq, seq_cap * 4))
		return -1;
	unsigned long long *jump = (unsigned long long *)ctx->jump.p;
	uint32_t *vsucc = (uint32_t *)ctx->vsucc.p, *vedge = (uint32_t *)ctx->vedge.p;
	FlatGraph g;
	g.e_src = (uint32_t *)ctx->e_src.p; g.e_dst = (uint32_t *)ctx->e_dst.p; g.e_rc = (uint32_t *)ctx->e_rc.p;
	g.e_len = (uint32_t *)ctx->e_len.p; g.e_count = (unsigned long long *)ctx->e_count.p;
	g.e_off = (unsigned long long *)ctx->e_off.p; g.e_seq = (uint32_t *)ctx->e_seq.p;
	{ ProfScope ps_(ctx, "memset"); CU(cudaMemsetAsync(vedge, 0xff, ((size_t)n_cv + 1) * 4, ctx->stream)); }
	{ ProfScope ps_(ctx, "memset"); CU(cudaMemsetAsync(g.e_seq, 0, seq_cap * 4, ctx->stream)); }
	if (n_cv) {
		LAUNCH(k_build_succ<W>, (n_cv + 255) / 256, 256, t, k, n_cv, kind, chain_slot, jump, vsucc, ctr);
		if (rank_lists(ctx, jump, n_cv)) return -1;
	}
	if (n_nodes)
		LAUNCH(k_edge_heads<W>, (unsigned)((2 * n_nodes + 127) / 128), 128, t, k, (uint32_t)n_nodes, kind, node_slot, node_ebase,
		       jump, vsucc, vedge, g, ctr);
	if (n_cv) LAUNCH(k_interior<W>, (n_cv + 255) / 256, 256, t, k, n_cv, chain_slot, jump, vedge, g);
	if (n_e) LAUNCH(k_rc_links<W>, (unsigned)((n_e + 255) / 256), 256, t, k, (uint32_t)n_e, node_slot, node_ebase, g, ctr);
	if (n_solid && !ctx->skip_counts)
		LAUNCH(k_edge_counts<W>, (unsigned)((n_solid + 255) / 256), 256, solid, (const uint32_t *)ctx->cur_solid_cnt, n_solid, k, t,
		       vL, vR, kind, node_ebase, vedge, g, ctr);
	// assign_count_garbage, contig by contig in the reference's call order (kmer_build.c:1040-1041)
	if (!ctx->skip_counts)
		for (int c = 0; c < ctx->n_contigs; ++c)
			if (ctx->contig_len[c] > (uint32_t)(k + 1) && n_e)
				LAUNCH(k_garbage_counts<W>, (ctx->contig_len[c] + 255) / 256, 256, (const uint8_t *)ctx->g_seq.p + ctx->contig_off[c],
				       ctx->contig_len[c], k, t, kind, node_ebase, vedge, g, ctx->contig_cov[c]);
	CU(cudaEventRecord(ctx->ev[2], ctx->stream));
	if (read_counters(ctx)) return -1;
	ctx->st.jump_rounds = ctx->h_ctr[CTR_JUMP_ROUNDS];
	ctx->st.n_seq_words = ctx->h_ctr[CTR_SEQ_WORDS];
	ctx->st.n_kp1_on_edge = ctx->h_ctr[CTR_KP1_ON_EDGE];
	if (ctx->st.n_seq_words > seq_cap) return fail(ctx, "edge sequence buffer overflow (%llu > %llu words)",
						       (unsigned long long)ctx->st.n_seq_words, (unsigned long long)seq_cap);
	ctx->have_graph = true;
	return 0;
}

// ------------------------------------------------------------------------------------------------ graph stage, two-level (paths)
// A PathStore for up to cap paths laid out in one flat buffer (the multi-GPU build keeps it in the arena, where the
// other ranks read it): first | last | cnt | off | n | interior, every part 256-byte aligned.
template <int W>
static size_t path_store_bytes(uint64_t cap)
{
	auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
	return 2 * al((cap + 1) * sizeof(Key<W>)) + 2 * al((cap + 1) * 8) + al((cap + 1) * 4) + al((cap + 16) * 4);
}

template <int W>
static PathStore<W> path_store_at(char *base, uint64_t cap)
{
	auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
	PathStore<W> ps;
	char *p = base;
	ps.first = (Key<W> *)p; p += al((cap + 1) * sizeof(Key<W>));
	ps.last = (Key<W> *)p; p += al((cap + 1) * sizeof(Key<W>));
	ps.cnt = (unsigned long long *)p; p += al((cap + 1) * 8);
	ps.off = (unsigned long long *)p; p += al((cap + 1) * 8);
	ps.n = (uint32_t *)p; p += al((cap + 1) * 4);
	ps.interior = (uint32_t *)p;
	ps.cap_paths = cap + 1;
	ps.cap_words = cap + 16;
	return ps;
}

// dense path arrays in this context's own buffers
template <int W>
static int path_store_own(tagpu_ctx *ctx, uint64_t cap, PathStore<W> *out, bool slack)
{
	PathStore<W> ps;
	ps.cap_paths = cap + 1;
	ps.cap_words = cap + 16;
	int (*grow)(tagpu_ctx *, Buf &, size_t) = slack ? ensure_slack : ensure_exact;
	if (grow(ctx, ctx->d_first, ps.cap_paths * sizeof(Key<W>)) || grow(ctx, ctx->d_last, ps.cap_paths * sizeof(Key<W>)) ||
	    grow(ctx, ctx->d_n, ps.cap_paths * 4) || grow(ctx, ctx->d_cnt, ps.cap_paths * 8) || grow(ctx, ctx->d_off, ps.cap_paths * 8) ||
	    grow(ctx, ctx->d_int, ps.cap_words * 4))
		return -1;
	ps.first = (Key<W> *)ctx->d_first.p; ps.last = (Key<W> *)ctx->d_last.p; ps.n = (uint32_t *)ctx->d_n.p;
	ps.cnt = (unsigned long long *)ctx->d_cnt.p; ps.off = (unsigned long long *)ctx->d_off.p; ps.interior = (uint32_t *)ctx->d_int.p;
	*out = ps;
	return 0;
}

// ---- level 1: contraction inside the blocks of the local solid list -> dense paths in ps (at most one per solid entry).
// Leaves the counters in ctx->h_ctr: CTR_PATHS, CTR_PATH_WORDS, and CTR_KMERS = the k-mers hidden inside the paths.
template <int W>
static int contract_local(tagpu_ctx *ctx, const PathStore<W> &ps)
{
	const uint32_t n_blocks = (uint32_t)ctx->n_blocks;
	typedef ContractCfg<W> CC;
	constexpr size_t smem_s = tagpu_contract_smem<W, CC::MAXN_SMALL>(), smem_m = tagpu_contract_smem<W, CC::MAXN_MEDIUM>(),
			 smem_l = tagpu_contract_smem<W, CC::MAXN_LARGE>();
	constexpr bool medium = CC::MAXN_MEDIUM > CC::MAXN_SMALL;
	auto k_small = k_contract<W, CC::MAXN_SMALL, CC::T_SMALL>;
	auto k_medium = k_contract<W, CC::MAXN_MEDIUM, CC::T_MEDIUM>;
	auto k_large = k_contract<W, CC::MAXN_LARGE, CC::T_LARGE>;
	int *grid_s = ctx->grid_s, *grid_m = ctx->grid_m, *grid_l = ctx->grid_l;
	if (!grid_s[W]) {
		CU(cudaFuncSetAttribute(k_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
		CU(cudaFuncSetAttribute(k_medium, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_m));
		CU(cudaFuncSetAttribute(k_large, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_l));
		int a = 0, b = 0, c = 0;
		CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_small, CC::T_SMALL, smem_s));
		CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_medium, CC::T_MEDIUM, smem_m));
		CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c, k_large, CC::T_LARGE, smem_l));
		grid_s[W] = ctx->n_sm * (a > 0 ? a : 1);
		grid_m[W] = ctx->n_sm * (b > 0 ? b : 1);
		grid_l[W] = ctx->n_sm * (c > 0 ? c : 1);
	}
	// two lists of deferred block ids (SMALL -> MEDIUM -> LARGE), each at most n_blocks long
	if (ensure_slack(ctx, ctx->blk_defer, 2 * ((size_t)n_blocks + 1) * 4)) return -1;
	uint32_t *list1 = (uint32_t *)ctx->blk_defer.p, *list2 = list1 + n_blocks + 1;
	if (n_blocks) {
		const SolidBlock *blocks = (const SolidBlock *)ctx->blocks.p;
		const Key<W> *solid = (const Key<W> *)ctx->cur_solid_key;
		const uint32_t *solid_cnt = (const uint32_t *)ctx->cur_solid_cnt;
		const uint32_t *none = nullptr;
		{ ProfScope ps_(ctx, "memset"); CU(cudaMemsetAsync(ctx->d_ctr + CTR_SPARE1, 0, 8, ctx->stream)); }
		{ ProfScope ps_(ctx, "memset"); CU(cudaMemsetAsync(ctx->d_ctr + CTR_SPARE2, 0, 16, ctx->stream)); }   // and CTR_SPARE3
		{
			ProfScope ps_(ctx, "k_contract<W>");
			k_small<<<(n_blocks < (uint32_t)grid_s[W] ? n_blocks : (uint32_t)grid_s[W]), CC::T_SMALL, smem_s, ctx->stream>>>(blocks, n_blocks, solid, solid_cnt, ctx->k, ctx->log2_buckets, none, 0, list1,
										 (int)CTR_SPARE2, ps, ctx->d_ctr);
			++ctx->launches;
			CU(cudaGetLastError());
		}
		const uint32_t *rest = list1;
		int rest_ctr = CTR_SPARE2;
		if (medium) {
			{ ProfScope ps_(ctx, "memset"); CU(cudaMemsetAsync(ctx->d_ctr + CTR_SPARE1, 0, 8, ctx->stream)); }
			ProfScope ps_(ctx, "k_contract_medium<W>");
			k_medium<<<(n_blocks < (uint32_t)grid_m[W] ? n_blocks : (uint32_t)grid_m[W]), CC::T_MEDIUM, smem_m, ctx->stream>>>(blocks, n_blocks, solid, solid_cnt, ctx->k, ctx->log2_buckets, list1,
										   (int)CTR_SPARE2, list2, (int)CTR_SPARE3, ps, ctx->d_ctr);
			++ctx->launches;
			CU(cudaGetLastError());
			rest = list2;
			rest_ctr = CTR_SPARE3;
		}
		{ ProfScope ps_(ctx, "memset"); CU(cudaMemsetAsync(ctx->d_ctr + CTR_SPARE1, 0, 8, ctx->stream)); }
		{
			ProfScope ps_(ctx, "k_contract_large<W>");
			k_large<<<(n_blocks < (uint32_t)grid_l[W] ? n_blocks : (uint32_t)grid_l[W]), CC::T_LARGE, smem_l, ctx->stream>>>(blocks, n_blocks, solid, solid_cnt, ctx->k, ctx->log2_buckets, rest, rest_ctr,
										 (uint32_t *)nullptr, 0, ps, ctx->d_ctr);
			++ctx->launches;
			CU(cudaGetLastError());
		}
	}
	if (read_counters(ctx)) return -1;
#ifdef TAGPU_TIMING
	{
		const unsigned long long *t = ctx->h_ctr + CTR_JUMP_FLAGS + 57;
		const double tot = (double)(t[0] + t[1] + t[2] + t[3] + t[4]);
		fprintf(stderr, "[tagpu timing] k_contract warp-cycles: load %.1f%%  table %.1f%%  hide test %.1f%%  links %.1f%%  walk + output %.1f%%\n",
			100.0 * t[0] / tot, 100.0 * t[1] / tot, 100.0 * t[2] / tot, 100.0 * t[3] / tot, 100.0 * t[4] / tot);
	}
#endif
	static const bool trace = getenv("TAGPU_TRACE_CONTRACT") != nullptr;
	if (trace)
		fprintf(stderr, "tagpu: contraction of %llu solid (k+1)-mers in %u blocks -> %llu paths, %llu interior words, %llu k-mers hidden\n",
			(unsigned long long)ctx->st.n_solid, n_blocks, (unsigned long long)ctx->h_ctr[CTR_PATHS], (unsigned long long)ctx->h_ctr[CTR_PATH_WORDS],
			(unsigned long long)ctx->h_ctr[CTR_KMERS]);
	if (trace && n_blocks) {
		std::vector<SolidBlock> hb(n_blocks);
		CU(cudaMemcpy(hb.data(), ctx->blocks.p, (size_t)n_blocks * sizeof(SolidBlock), cudaMemcpyDeviceToHost));
		unsigned long long nb[4] = { 0, 0, 0, 0 }, ne[4] = { 0, 0, 0, 0 };
		for (const SolidBlock &b : hb) {
			const int c = b.flags ? 3 : b.n <= (uint32_t)CC::MAXN_MEDIUM ? 0 : b.n <= (uint32_t)CC::MAXN_LARGE ? 1 : 2;
			++nb[c];
			ne[c] += b.n;
		}
		fprintf(stderr, "tagpu: blocks (entries): up to medium %llu (%llu), large %llu (%llu), too large %llu (%llu), split by hash class %llu (%llu)\n",
			nb[0], ne[0], nb[1], ne[1], nb[2], ne[2], nb[3], ne[3]);
	}
	if (ctx->h_ctr[CTR_PATHS] >= ps.cap_paths || ctx->h_ctr[CTR_PATH_WORDS] >= ps.cap_words)
		return fail(ctx, "contraction produced more paths (%llu, %llu words) than there is room for", (unsigned long long)ctx->h_ctr[CTR_PATHS],
			    (unsigned long long)ctx->h_ctr[CTR_PATH_WORDS]);
	return 0;
}

// ---- level 2: the global stage on n_paths paths.  hidden_elsewhere: k-mers hidden inside the paths of OTHER ranks
// (this device's CTR_KMERS already holds the ones of its own contraction).
template <int W>
static int graph_stage_global(tagpu_ctx *ctx, const PathStore<W> &ps, uint64_t n_paths, uint64_t hidden_elsewhere);

template <int W>
static int graph_stage_paths(tagpu_ctx *ctx)
{
	PathStore<W> ps;
	if (path_store_own<W>(ctx, ctx->st.n_solid, &ps, false) || contract_local<W>(ctx, ps)) return -1;
	return graph_stage_global<W>(ctx, ps, ctx->h_ctr[CTR_PATHS], 0);
}

template <int W>
static int graph_stage_global(tagpu_ctx *ctx, const PathStore<W> &ps, uint64_t n_paths, uint64_t hidden_elsewhere)
{
	const int k = ctx->k;
	const uint64_t n_solid = ctx->st.n_solid;
	unsigned long long *ctr = ctx->d_ctr;
	// ---- level 2: the global stage on the paths
	const uint64_t slots64 = (n_paths * 5) / 2 + 1024;
	if (slots64 > (1ull << 30)) return fail(ctx, "k-mer table would need %llu slots (> 2^30)", (unsigned long long)slots64);
	const uint32_t n_slots = (uint32_t)slots64;
	ctx->kt_slots = n_slots;
	const size_t mask_bytes = ((size_t)n_slots + 3) / 4 * 4;
	if (ensure_slack(ctx, ctx->kt_keys, (size_t)n_slots * sizeof(Key<W>)) || ensure_slack(ctx, ctx->kt_mask, mask_bytes) ||
	    ensure_slack(ctx, ctx->node_ord, (size_t)n_slots * 4) || ensure_slack(ctx, ctx->vL, (n_paths + 1) * 4) ||
	    ensure_slack(ctx, ctx->vR, (n_paths + 1) * 4) || ensure_slack(ctx, ctx->node_slot, (2 * n_paths + 1) * 4) ||
	    ensure_slack(ctx, ctx->node_ebase, (2 * n_paths + 1) * 4) || ensure_slack(ctx, ctx->chain_slot, (2 * n_paths + 1) * 4))
		return -1;
	KTab<W> t;
	t.keys = (Key<W> *)ctx->kt_keys.p;
	t.mask32 = (uint32_t *)ctx->kt_mask.p;
	t.n_slots = n_slots;
	{ ProfScope ps_(ctx, "memset"); CU(cudaMemsetAsync(t.keys, 0, (size_t)n_slots * sizeof(Key<W>), ctx->stream)); }
	{ ProfScope ps_(ctx, "memset"); CU(cudaMemsetAsync(t.mask32, 0, mask_bytes, ctx->stream)); }
	uint32_t *kind = (uint32_t *)ctx->node_ord.p, *node_slot = (uint32_t *)ctx->node_slot.p,
		 *node_ebase = (uint32_t *)ctx->node_ebase.p, *chain_slot = (uint32_t *)ctx->chain_slot.p,
		 *vL = (uint32_t *)ctx->vL.p, *vR = (uint32_t *)ctx->vR.p;
	if (n_paths) LAUNCH(k_insert_paths<W>, (unsigned)((n_paths + 255) / 256), 256, ps, n_paths, k, t, vL, vR, ctr);
	LAUNCH(k_classify<W>, (n_slots + 1023) / 1024, 1024, t, kind, node_slot, node_ebase, chain_slot, ctr);
	if (read_counters(ctx)) return -1;
	const uint64_t n_nodes = ctx->h_ctr[CTR_NODES], n_e = ctx->h_ctr[CTR_EDGES], n_chain = ctx->h_ctr[CTR_CHAIN];
	ctx->st.n_kmers = ctx->h_ctr[CTR_KMERS] + hidden_elsewhere;
	ctx->st.n_v = 2 * n_nodes;
	ctx->st.n_e = n_e;
	if (n_e > 0xfffffff0ull) return fail(ctx, "too many edges (%llu)", (unsigned long long)n_e);
	const uint32_t n_cv = (uint32_t)(2 * n_chain);
	const uint64_t seq_cap = (n_e * (uint64_t)k + 2 * n_solid) / 16 + n_e + 16;
	if (ensure_slack(ctx, ctx->jump, ((size_t)n_cv + 1) * 8) || ensure_slack(ctx, ctx->vsucc, ((size_t)n_cv + 1) * 4) ||
	    ensure_slack(ctx, ctx->wlast, ((size_t)n_cv + 1) * 4) || ensure_slack(ctx, ctx->vedge, ((size_t)n_cv + 1) * 4) || ensure(ctx, ctx->e_src, (n_e + 1) * 4) ||
	    ensure(ctx, ctx->e_dst, (n_e + 1) * 4) || ensure(ctx, ctx->e_rc, (n_e + 1) * 4) || ensure(ctx, ctx->e_len, (n_e + 1) * 4) ||
	    ensure(ctx, ctx->e_count, (n_e + 1) * 8) || ensure(ctx, ctx->e_off, (n_e + 1) * 8) || ensure(ctx, ctx->e_seq, seq_cap * 4))
		return -1;
	unsigned long long *jump = (unsigned long long *)ctx->jump.p;
	uint32_t *vsucc = (uint32_t *)ctx->vsucc.p, *vedge = (uint32_t *)ctx->vedge.p, *wlast = (uint32_t *)ctx->wlast.p;
	FlatGraph g;
	g.e_src = (uint32_t *)ctx->e_src.p; g.e_dst = (uint32_t *)ctx->e_dst.p; g.e_rc = (uint32_t *)ctx->e_rc.p;
	g.e_len = (uint32_t *)ctx->e_len.p; g.e_count = (unsigned long long *)ctx->e_count.p;
	g.e_off = (unsigned long long *)ctx->e_off.p; g.e_seq = (uint32_t *)ctx->e_seq.p;
	{ ProfScope ps_(ctx, "memset"); CU(cudaMemsetAsync(vedge, 0xff, ((size_t)n_cv + 1) * 4, ctx->stream)); }
	{ ProfScope ps_(ctx, "memset"); CU(cudaMemsetAsync(g.e_seq, 0, seq_cap * 4, ctx->stream)); }
	if (n_cv) {
		LAUNCH(k_succ_paths<W>, (unsigned)((2 * n_paths + 255) / 256), 256, ps, n_paths, vL, vR, kind, jump, vsucc, wlast);
		if (rank_lists(ctx, jump, n_cv)) return -1;
	}
	if (n_paths) {
		LAUNCH(k_heads_paths<W>, (unsigned)((2 * n_paths + TAGPU_HEADS_THREADS - 1) / TAGPU_HEADS_THREADS), TAGPU_HEADS_THREADS, ps, n_paths, k, t, vL, vR, kind, node_ebase, jump, vsucc, wlast, vedge, g, ctr);
		LAUNCH(k_interior_paths<W>, (unsigned)((2 * n_paths + 255) / 256), 256, ps, n_paths, k, vL, vR, kind, jump, wlast, vedge, g);
	}
	if (n_e) LAUNCH(k_rc_links<W>, (unsigned)((n_e + 255) / 256), 256, t, k, (uint32_t)n_e, node_slot, node_ebase, g, ctr);
	if (n_paths && !ctx->skip_counts)
		LAUNCH(k_counts_paths<W>, (unsigned)((n_paths + 255) / 256), 256, ps, n_paths, k, t, vL, kind, node_ebase, vedge, g, ctr);
	CU(cudaEventRecord(ctx->ev[2], ctx->stream));
	if (read_counters(ctx)) return -1;
	ctx->st.jump_rounds = ctx->h_ctr[CTR_JUMP_ROUNDS];
	ctx->st.n_seq_words = ctx->h_ctr[CTR_SEQ_WORDS];
	ctx->st.n_kp1_on_edge = ctx->h_ctr[CTR_KP1_ON_EDGE];
	if (ctx->st.n_seq_words > seq_cap) return fail(ctx, "edge sequence buffer overflow (%llu > %llu words)",
						       (unsigned long long)ctx->st.n_seq_words, (unsigned long long)seq_cap);
	ctx->have_graph = true;
	ctx->contracted = true;
	return 0;
}

static int run(tagpu_ctx *ctx, const uint8_t *d_seq, uint64_t n, int K, bool with_graph)
{
	CU(cudaSetDevice(ctx->device));
	if (K < 18 || K > 64) return fail(ctx, "unsupported k-mer size: k + 1 = %d (supported: 18..64)", K);
	if (ctx->dist) return fail(ctx, "context is in multi-GPU mode (tagpu_dist_plan): use the tagpu_dist_* calls or tagpu_dist_close first");
	ctx->K = K;
	ctx->k = K - 1;
	ctx->W = K <= 32 ? 1 : 2;
	ctx->have_count = ctx->have_graph = false;
	ctx->solid_sharded = false;
	ctx->launches = 0;
	ctx->err[0] = 0;
	memset(&ctx->st, 0, sizeof(ctx->st));
	if (!ctx->local_mode) { ctx->n_garbage = 0; ctx->n_contigs = 0; }
	{ ProfScope ps_(ctx, "memset"); CU(cudaMemsetAsync(ctx->d_ctr, 0, CTR_TOTAL * sizeof(unsigned long long), ctx->stream)); }
	CU(cudaEventRecord(ctx->ev[0], ctx->stream));
	int rc = ctx->W == 1 ? count_stage_partitioned<1>(ctx, d_seq, n) : count_stage_partitioned<2>(ctx, d_seq, n);
	if (rc) return rc;
	CU(cudaEventRecord(ctx->ev[1], ctx->stream));
	if (with_graph && ctx->n_garbage) {
		// (k+1)-mer list of the graph stage = solid (k+1)-mers of the reads, then the contig garbage with count 0
		const size_t key = ctx->W == 1 ? sizeof(Key<1>) : sizeof(Key<2>);
		const uint64_t ns = ctx->st.n_solid, ng = ctx->n_garbage;
		if (ensure(ctx, ctx->comb_key, (ns + ng + 1) * key) || ensure(ctx, ctx->comb_cnt, (ns + ng + 1) * 4)) return -1;
		if (ns) {
			CU(cudaMemcpyAsync(ctx->comb_key.p, ctx->cur_solid_key, ns * key, cudaMemcpyDeviceToDevice, ctx->stream));
			CU(cudaMemcpyAsync(ctx->comb_cnt.p, ctx->cur_solid_cnt, ns * 4, cudaMemcpyDeviceToDevice, ctx->stream));
		}
		CU(cudaMemcpyAsync((char *)ctx->comb_key.p + ns * key, ctx->g_key.p, ng * key, cudaMemcpyDeviceToDevice, ctx->stream));
		CU(cudaMemsetAsync((char *)ctx->comb_cnt.p + ns * 4, 0, ng * 4, ctx->stream));
		ctx->cur_solid_key = ctx->comb_key.p;
		ctx->cur_solid_cnt = ctx->comb_cnt.p;
	}
	ctx->contracted = false;
	if (with_graph && ctx->contract && !ctx->n_garbage && ctx->n_blocks) {
		rc = ctx->W == 1 ? graph_stage_paths<1>(ctx) : graph_stage_paths<2>(ctx);
		if (rc) return rc;
	} else if (with_graph) {
		rc = ctx->W == 1 ? graph_stage<1>(ctx) : graph_stage<2>(ctx);
		if (rc) return rc;
	} else {
		CU(cudaEventRecord(ctx->ev[2], ctx->stream));
		CU(cudaStreamSynchronize(ctx->stream));
	}
	CU(cudaEventElapsedTime(&ctx->st.ms_count, ctx->ev[0], ctx->ev[1]));
	CU(cudaEventElapsedTime(&ctx->st.ms_graph, ctx->ev[1], ctx->ev[2]));
	CU(cudaEventElapsedTime(&ctx->st.ms_total, ctx->ev[0], ctx->ev[2]));
	ctx->st.gpu_launches = ctx->launches;
	prof_finish(ctx);
	return 0;
}

// ------------------------------------------------------------------------------------------------ multi-GPU
// One process per GPU.  Every rank partitions ITS slice of the reads into its own bucket regions (pass 1, local) and
// owns a contiguous range of buckets, whose records it reads from every rank's regions through CUDA-IPC mappings inside
// the counting kernel (pass 2: NVLink peer loads overlapped with shared-memory counting).  All buffers a peer touches
// live in ONE allocation per rank (the "arena") with identical offsets everywhere, so a single 64-byte IPC handle per
// rank is all the host has to exchange.  The host program supplies the barriers between phases (include/tagpu.h).
struct DistState {
	int rank = 0, world = 1, W = 0;
	uint64_t n_total = 0;
	PartCfg cfg;
	uint64_t solid_cap = 0;
	char *arena = nullptr;
	size_t arena_bytes = 0;
	size_t off_cursor = 0, off_extoff = 0, off_ext = 0, off_regions = 0, off_scnt = 0, off_skey = 0;
	char *peer[TAGPU_MAX_RANKS] = { nullptr };      // arena of every rank as mapped here (peer[rank] = arena)
	bool connected = false;
	Buf g_key, g_cnt;                               // solid set gathered from all ranks
	bool have_paths = false;                        // tagpu_dist_contract left this rank's paths at the front of its regions area
	uint64_t paths_cap = 0;                         // ... laid out for this many (= the rank's solid count)
};

static void borrow(Buf &b, void *p, size_t bytes)
{
	b.p = p;
	b.cap = bytes;
}

static void dist_unmap(tagpu_ctx *ctx)
{
	DistState *d = ctx->dist;
	if (!d) return;
	cudaDeviceSynchronize();
	for (int r = 0; r < d->world; ++r)
		if (r != d->rank && d->peer[r]) { cudaIpcCloseMemHandle(d->peer[r]); d->peer[r] = nullptr; }
	d->connected = false;
}

static Buf *const *dist_borrowed(tagpu_ctx *ctx, int *n)
{
	static thread_local Buf *b[6];
	b[0] = &ctx->regions; b[1] = &ctx->cursor; b[2] = &ctx->ext; b[3] = &ctx->ext_off; b[4] = &ctx->solid_key; b[5] = &ctx->solid_cnt;
	*n = 6;
	return b;
}

static void dist_release(tagpu_ctx *ctx)
{
	DistState *d = ctx->dist;
	if (!d) return;
	dist_unmap(ctx);
	int n;
	Buf *const *bb = dist_borrowed(ctx, &n);
	for (int i = 0; i < n; ++i) { bb[i]->p = nullptr; bb[i]->cap = 0; }
	if (d->arena) cudaFree(d->arena);
	if (d->g_key.p) cudaFree(d->g_key.p);
	if (d->g_cnt.p) cudaFree(d->g_cnt.p);
	delete d;
	ctx->dist = nullptr;
}

extern "C" int tagpu_dist_plan(tagpu_ctx *ctx, int rank, int world, uint64_t n_total_bytes, int k, void *handle_out)
{
	CU(cudaSetDevice(ctx->device));
	if (world < 1 || world > TAGPU_MAX_RANKS || rank < 0 || rank >= world) return fail(ctx, "bad rank/world %d/%d (max %d ranks)", rank, world, TAGPU_MAX_RANKS);
	const int K = k + 1;
	if (K < 18 || K > 64) return fail(ctx, "unsupported k-mer size: k + 1 = %d (supported: 18..64)", K);
	dist_release(ctx);
	// the single-GPU buffers with the same roles are dropped: from now on they alias the arena
	int nb;
	Buf *const *bb = dist_borrowed(ctx, &nb);
	for (int i = 0; i < nb; ++i) { if (bb[i]->p) CU(cudaFree(bb[i]->p)); bb[i]->p = nullptr; bb[i]->cap = 0; }
	DistState *d = new DistState();
	ctx->dist = d;
	d->rank = rank; d->world = world; d->n_total = n_total_bytes;
	ctx->count_stream_bytes = n_total_bytes;
	d->W = K <= 32 ? 1 : 2;
	ctx->K = K; ctx->k = k; ctx->W = d->W;
	const size_t rec = d->W == 1 ? sizeof(SkRec<1>) : sizeof(SkRec<2>), key = d->W == 1 ? sizeof(Key<1>) : sizeof(Key<2>);
	d->cfg = plan_cfg(n_total_bytes, K, world, d->W == 1 ? BucketCfg<1>::GROUP_TARGET : BucketCfg<2>::GROUP_TARGET, rec, 0);
	const size_t n_buckets = (size_t)1 << d->cfg.log2_buckets;
	// owned windows ~ N_i / world, bucket ownership is uneven: 1.5x head-room (checked, see count_owned)
	d->solid_cap = n_total_bytes / world / (uint64_t)ctx->ci * 3 / 2 + (1u << 20);
	size_t off = 0;
	auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
	d->off_cursor = take(n_buckets * 8);
	d->off_extoff = take((n_buckets + 1) * 4);
	d->off_ext = take((size_t)d->cfg.overflow_cap * rec);
	d->off_scnt = take(d->solid_cap * 4);
	d->off_skey = take(d->solid_cap * key);
	d->off_regions = take(n_buckets * d->cfg.cap_records * rec);
	d->arena_bytes = off < (8u << 20) ? (8u << 20) : off;       // large enough to be an allocation of its own (IPC maps whole allocations)
	CU(cudaMalloc(&d->arena, d->arena_bytes));
	d->peer[rank] = d->arena;
	borrow(ctx->cursor, d->arena + d->off_cursor, n_buckets * 8);
	borrow(ctx->ext_off, d->arena + d->off_extoff, (n_buckets + 1) * 4);
	borrow(ctx->ext, d->arena + d->off_ext, (size_t)d->cfg.overflow_cap * rec);
	borrow(ctx->solid_cnt, d->arena + d->off_scnt, d->solid_cap * 4);
	borrow(ctx->solid_key, d->arena + d->off_skey, d->solid_cap * key);
	borrow(ctx->regions, d->arena + d->off_regions, n_buckets * d->cfg.cap_records * rec);
	if (ensure(ctx, ctx->overflow, (size_t)d->cfg.overflow_cap * rec) || ensure(ctx, ctx->overflow_bucket, (size_t)d->cfg.overflow_cap * 4) ||
	    ensure(ctx, ctx->ext_count, n_buckets * 4))
		return -1;
	cudaIpcMemHandle_t h;
	memset(&h, 0, sizeof(h));
	if (world > 1) CU(cudaIpcGetMemHandle(&h, d->arena));
	static_assert(sizeof(cudaIpcMemHandle_t) == TAGPU_IPC_HANDLE_BYTES, "IPC handle size");
	memcpy(handle_out, &h, sizeof(h));
	d->connected = world == 1;
	return 0;
}

extern "C" int tagpu_dist_connect(tagpu_ctx *ctx, const void *all_handles)
{
	DistState *d = ctx->dist;
	if (!d) return fail(ctx, "tagpu_dist_connect before tagpu_dist_plan");
	CU(cudaSetDevice(ctx->device));
	for (int r = 0; r < d->world; ++r) {
		if (r == d->rank) continue;
		cudaIpcMemHandle_t h;
		memcpy(&h, (const char *)all_handles + (size_t)r * TAGPU_IPC_HANDLE_BYTES, sizeof(h));
		void *p = nullptr;
		CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
		d->peer[r] = (char *)p;
	}
	d->connected = true;
	return 0;
}

static int dist_partition_impl(tagpu_ctx *ctx, const uint8_t *d_seq, uint64_t n_local_bytes);
static int upload(tagpu_ctx *ctx, const uint8_t *h_seq, uint64_t n);

extern "C" int tagpu_dist_partition(tagpu_ctx *ctx, const uint8_t *d_seq, uint64_t n_local_bytes)
{
	ctx->h_src = nullptr;
	ctx->src_packed = false;
	return dist_partition_impl(ctx, d_seq, n_local_bytes);
}

// same with this rank's slice of the reads still in (pinned) host memory: the upload is overlapped with pass 1
extern "C" int tagpu_dist_partition_host(tagpu_ctx *ctx, const uint8_t *h_seq, uint64_t n_local_bytes)
{
	ctx->src_packed = false;
	if (upload(ctx, h_seq, n_local_bytes)) return -1;
	return dist_partition_impl(ctx, (const uint8_t *)ctx->seq.p, n_local_bytes);
}

extern "C" uint64_t tagpu_packed_bytes(uint64_t n_positions);

// same with this rank's slice as a packed read stream (packed by the rank itself: positions count from the slice start)
extern "C" int tagpu_dist_partition_host_packed(tagpu_ctx *ctx, const uint8_t *h_packed, uint64_t n_local_positions)
{
	ctx->src_packed = true;
	if (upload(ctx, h_packed, tagpu_packed_bytes(n_local_positions))) return -1;
	const int rc = dist_partition_impl(ctx, (const uint8_t *)ctx->seq.p, n_local_positions);
	ctx->src_packed = false;
	return rc;
}

static int dist_partition_impl(tagpu_ctx *ctx, const uint8_t *d_seq, uint64_t n_local_bytes)
{
	DistState *d = ctx->dist;
	if (!d || !d->connected) return fail(ctx, "tagpu_dist_partition before tagpu_dist_plan / tagpu_dist_connect");
	CU(cudaSetDevice(ctx->device));
	if (n_local_bytes > d->n_total / d->world + (1u << 20))
		return fail(ctx, "this rank's slice (%llu bytes) is larger than planned (%llu total over %d ranks)", (unsigned long long)n_local_bytes,
			    (unsigned long long)d->n_total, d->world);
	ctx->have_count = ctx->have_graph = false;
	ctx->solid_sharded = false;
	ctx->launches = 0;
	ctx->err[0] = 0;
	memset(&ctx->st, 0, sizeof(ctx->st));
	{ ProfScope ps_(ctx, "memset"); CU(cudaMemsetAsync(ctx->d_ctr, 0, CTR_TOTAL * sizeof(unsigned long long), ctx->stream)); }
	CU(cudaEventRecord(ctx->ev[0], ctx->stream));
	// partition_local ends with a stream synchronisation: when the host enters the barrier, this rank's regions are complete
	return d->W == 1 ? partition_local<1>(ctx, d_seq, n_local_bytes, d->cfg) : partition_local<2>(ctx, d_seq, n_local_bytes, d->cfg);
}

template <int W>
static int dist_count(tagpu_ctx *ctx)
{
	DistState *d = ctx->dist;
	CountPeers<W> peers;
	memset(&peers, 0, sizeof(peers));
	for (int r = 0; r < d->world; ++r) {
		peers.regions[r] = (const SkRec<W> *)(d->peer[r] + d->off_regions);
		peers.cursor[r] = (const unsigned long long *)(d->peer[r] + d->off_cursor);
		peers.ext[r] = (const SkRec<W> *)(d->peer[r] + d->off_ext);
		peers.ext_off[r] = (const uint32_t *)(d->peer[r] + d->off_extoff);
	}
	return count_owned<W>(ctx, d->cfg, peers, (uint32_t)d->rank * d->cfg.per_rank, d->solid_cap);
}

extern "C" int tagpu_dist_count(tagpu_ctx *ctx, uint64_t stats_out[4])
{
	DistState *d = ctx->dist;
	if (!d || !d->connected) return fail(ctx, "tagpu_dist_count before tagpu_dist_plan / tagpu_dist_connect");
	CU(cudaSetDevice(ctx->device));
	const int rc = d->W == 1 ? dist_count<1>(ctx) : dist_count<2>(ctx);
	if (rc) return rc;
	CU(cudaEventRecord(ctx->ev[1], ctx->stream));
	stats_out[0] = ctx->st.n_instances;   // windows of THIS rank's reads
	stats_out[1] = ctx->st.n_distinct;    // distinct / solid keys of THIS rank's buckets
	stats_out[2] = ctx->st.n_solid;
	stats_out[3] = ctx->st.sum_solid;
	return 0;
}

static int dist_gather_solid(tagpu_ctx *ctx, const uint64_t *all_stats, uint64_t n_total);

// all_stats: world x 4 values, the stats_out of every rank in rank order.  Pulls every rank's solid set over NVLink
// (peer copies out of the mapped arenas), then runs the graph stage on the union.  with_graph = 0 stops after the gather.
extern "C" int tagpu_dist_graph(tagpu_ctx *ctx, const uint64_t *all_stats, int with_graph)
{
	DistState *d = ctx->dist;
	if (!d || !ctx->have_count) return fail(ctx, "tagpu_dist_graph before tagpu_dist_count");
	CU(cudaSetDevice(ctx->device));
	uint64_t tot[4] = { 0, 0, 0, 0 };
	for (int r = 0; r < d->world; ++r)
		for (int j = 0; j < 4; ++j) tot[j] += all_stats[r * 4 + j];
	if (d->world > 1 && dist_gather_solid(ctx, all_stats, tot[2])) return -1;
	ctx->solid_sharded = false;
	ctx->contracted = false;
	ctx->st.n_instances = tot[0];
	ctx->st.n_distinct = tot[1];
	ctx->st.n_solid = tot[2];
	ctx->st.sum_solid = tot[3];
	if (with_graph) {
		const int rc = d->W == 1 ? graph_stage<1>(ctx) : graph_stage<2>(ctx);
		if (rc) return rc;
	} else {
		CU(cudaEventRecord(ctx->ev[2], ctx->stream));
		CU(cudaStreamSynchronize(ctx->stream));
	}
	CU(cudaEventElapsedTime(&ctx->st.ms_count, ctx->ev[0], ctx->ev[1]));
	CU(cudaEventElapsedTime(&ctx->st.ms_graph, ctx->ev[1], ctx->ev[2]));
	CU(cudaEventElapsedTime(&ctx->st.ms_total, ctx->ev[0], ctx->ev[2]));
	ctx->st.gpu_launches = ctx->launches;
	prof_finish(ctx);
	return 0;
}

// ---- two-level graph stage across ranks.  After the stats all-gather (every rank has finished counting, so nobody reads
// anybody's bucket regions any more) each rank contracts ITS OWN solid list into paths, written over the front of its
// regions area in the arena; paths_out = { paths, interior words, hidden k-mers, 1 if it worked this way }.  The host
// all-gathers the 4 values (that is the barrier: every rank's paths are complete), then tagpu_dist_graph_paths pulls all
// paths over NVLink and runs the global stage on them.  If any rank reports 0 in paths_out[3], all fall back to
// tagpu_dist_graph.  The host must put a barrier between tagpu_dist_graph_paths and the next tagpu_dist_partition
// (which overwrites the regions the other ranks pull from).
template <int W>
static int dist_contract(tagpu_ctx *ctx, uint64_t paths_out[4])
{
	DistState *d = ctx->dist;
	paths_out[0] = paths_out[1] = paths_out[2] = paths_out[3] = 0;
	d->have_paths = false;
	const uint64_t n_local = ctx->st.n_solid;
	const size_t region_bytes = ((size_t)1 << d->cfg.log2_buckets) * d->cfg.cap_records * sizeof(SkRec<W>);
	if (!ctx->contract || path_store_bytes<W>(n_local) > region_bytes) return 0;   // (not an error: one-level stage instead)
	if (n_local && !ctx->n_blocks) return 0;
	const PathStore<W> ps = path_store_at<W>(d->arena + d->off_regions, n_local);
	if (contract_local<W>(ctx, ps)) return -1;
	d->have_paths = true;
	d->paths_cap = n_local;
	paths_out[0] = ctx->h_ctr[CTR_PATHS];
	paths_out[1] = ctx->h_ctr[CTR_PATH_WORDS];
	paths_out[2] = ctx->h_ctr[CTR_KMERS];
	paths_out[3] = 1;
	return 0;
}

extern "C" int tagpu_dist_contract(tagpu_ctx *ctx, uint64_t paths_out[4])
{
	DistState *d = ctx->dist;
	if (!d || !ctx->have_count) return fail(ctx, "tagpu_dist_contract before tagpu_dist_count");
	CU(cudaSetDevice(ctx->device));
	return d->W == 1 ? dist_contract<1>(ctx, paths_out) : dist_contract<2>(ctx, paths_out);
}

static int dist_gather_solid(tagpu_ctx *ctx, const uint64_t *all_stats, uint64_t n_total)
{
	DistState *d = ctx->dist;
	const size_t key = d->W == 1 ? sizeof(Key<1>) : sizeof(Key<2>);
	if (ensure(ctx, d->g_key, (n_total + 1) * key) || ensure(ctx, d->g_cnt, (n_total + 1) * 4)) return -1;
	uint64_t o = 0;
	ProfScope ps_(ctx, "peer_gather_solid");
	for (int r = 0; r < d->world; ++r) {
		const uint64_t n = all_stats[r * 4 + 2];
		if (n > d->solid_cap) return fail(ctx, "rank %d reports %llu solid (k+1)-mers, more than the planned %llu", r, (unsigned long long)n, (unsigned long long)d->solid_cap);
		if (n) {
			CU(cudaMemcpyAsync((char *)d->g_key.p + o * key, d->peer[r] + d->off_skey, n * key, cudaMemcpyDeviceToDevice, ctx->stream));
			CU(cudaMemcpyAsync((char *)d->g_cnt.p + o * 4, d->peer[r] + d->off_scnt, n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
		}
		o += n;
	}
	ctx->cur_solid_key = d->g_key.p;
	ctx->cur_solid_cnt = d->g_cnt.p;
	return 0;
}

template <int W>
static int dist_graph_paths(tagpu_ctx *ctx, const uint64_t *all_stats, const uint64_t *all_paths)
{
	DistState *d = ctx->dist;
	PathPeers<W> pp;
	memset(&pp, 0, sizeof(pp));
	pp.world = d->world;
	uint64_t n_paths = 0, n_words = 0, hidden_elsewhere = 0;
	for (int r = 0; r < d->world; ++r) {
		pp.src[r] = path_store_at<W>(d->peer[r] + d->off_regions, all_stats[r * 4 + 2]);
		pp.n_paths[r] = all_paths[r * 4]; pp.n_words[r] = all_paths[r * 4 + 1];
		pp.p0[r] = n_paths; pp.w0[r] = n_words;
		n_paths += pp.n_paths[r]; n_words += pp.n_words[r];
		if (r != d->rank) hidden_elsewhere += all_paths[r * 4 + 2];
	}
	PathStore<W> ps;
	if (path_store_own<W>(ctx, n_paths > n_words ? n_paths : n_words, &ps, true)) return -1;
	if (d->world == 1) ps = pp.src[0];                                  // nothing to pull
	else if (n_paths) LAUNCH(k_gather_paths<W>, 4 * ctx->n_sm, 256, pp, ps);
	return graph_stage_global<W>(ctx, ps, n_paths, hidden_elsewhere);
}

// all_paths: world x 4, the paths_out of every rank in rank order.  with_graph: 1 = graph; 3 = graph, and the solid sets of
// all ranks gathered on every rank as well (for the tagpu_copy_solid / tagpu_copy_kmers / tagpu_write_kmc_db calls).
extern "C" int tagpu_dist_graph_paths(tagpu_ctx *ctx, const uint64_t *all_stats, const uint64_t *all_paths, int with_graph)
{
	DistState *d = ctx->dist;
	if (!d || !ctx->have_count) return fail(ctx, "tagpu_dist_graph_paths before tagpu_dist_count");
	if (!d->have_paths) return fail(ctx, "tagpu_dist_graph_paths before tagpu_dist_contract");
	CU(cudaSetDevice(ctx->device));
	uint64_t tot[4] = { 0, 0, 0, 0 };
	for (int r = 0; r < d->world; ++r) {
		if (!all_paths[r * 4 + 3]) return fail(ctx, "rank %d did not contract its solid set: every rank has to call tagpu_dist_graph instead", r);
		for (int j = 0; j < 4; ++j) tot[j] += all_stats[r * 4 + j];
	}
	ctx->solid_sharded = d->world > 1;
	if ((with_graph & 2) && d->world > 1) {
		if (dist_gather_solid(ctx, all_stats, tot[2])) return -1;
		ctx->solid_sharded = false;
	}
	ctx->st.n_instances = tot[0];
	ctx->st.n_distinct = tot[1];
	ctx->st.n_solid = tot[2];
	ctx->st.sum_solid = tot[3];
	const int rc = d->W == 1 ? dist_graph_paths<1>(ctx, all_stats, all_paths) : dist_graph_paths<2>(ctx, all_stats, all_paths);
	if (rc) return rc;
	CU(cudaEventElapsedTime(&ctx->st.ms_count, ctx->ev[0], ctx->ev[1]));
	CU(cudaEventElapsedTime(&ctx->st.ms_graph, ctx->ev[1], ctx->ev[2]));
	CU(cudaEventElapsedTime(&ctx->st.ms_total, ctx->ev[0], ctx->ev[2]));
	ctx->st.gpu_launches = ctx->launches;
	prof_finish(ctx);
	return 0;
}

// One whole multi-GPU step in a single call: the phases above with the barriers and counter exchanges between them taken
// from a tagpu_shm segment (tagpu_host.c), so that no host-language round trip sits between the kernels of a step.
// src_kind: 0 = device stream, 1 = pinned host ASCII stream, 2 = pinned host packed stream (n = positions).
// flags: bit 0 = build the graph, bit 1 = also gather the solid sets on every rank, bit 2 = the previous step left paths in
// the arenas that other ranks may still be pulling (barrier first).  Returns 0, or -1 (error text in tagpu_last_error);
// *used_paths tells whether the two-level stage ran (then the caller sets bit 2 for the next step).
struct tagpu_shm;
extern "C" void tagpu_shm_barrier(struct tagpu_shm *s);
extern "C" int tagpu_shm_allgather(struct tagpu_shm *s, const uint64_t *mine, int n, uint64_t *all);

extern "C" int tagpu_dist_step(tagpu_ctx *ctx, struct tagpu_shm *shm, const uint8_t *src, uint64_t n, int src_kind, int flags, int *used_paths)
{
	DistState *d = ctx->dist;
	if (!d || !d->connected || !shm) return fail(ctx, "tagpu_dist_step before tagpu_dist_plan / tagpu_dist_connect, or without a rendezvous segment");
	*used_paths = 0;
	if (flags & 4) tagpu_shm_barrier(shm);
	int rc = src_kind == 0 ? tagpu_dist_partition(ctx, src, n) : src_kind == 1 ? tagpu_dist_partition_host(ctx, src, n) : tagpu_dist_partition_host_packed(ctx, src, n);
	// a failing rank still has to meet the others at every rendezvous of the step, or they would wait for ever
	uint64_t mine[4] = { 0, 0, 0, 0 }, all_stats[4 * TAGPU_MAX_RANKS], all_paths[4 * TAGPU_MAX_RANKS];
	tagpu_shm_barrier(shm);
	if (!rc) rc = tagpu_dist_count(ctx, mine);
	const uint64_t bad = ~0ull;
	if (rc) mine[0] = bad;
	tagpu_shm_allgather(shm, mine, 4, all_stats);
	bool any_bad = false;
	for (int r = 0; r < d->world; ++r) any_bad = any_bad || all_stats[4 * r] == bad;
	if (any_bad) return rc ? rc : fail(ctx, "another rank failed in the count stage");
	const bool with_graph = (flags & 1) != 0;
	if (with_graph && ctx->contract) {
		uint64_t paths[4] = { 0, 0, 0, 0 };
		rc = tagpu_dist_contract(ctx, paths);
		if (rc) { paths[0] = bad; paths[3] = 0; }
		tagpu_shm_allgather(shm, paths, 4, all_paths);
		bool all_ok = true;
		for (int r = 0; r < d->world; ++r) {
			if (all_paths[4 * r] == bad) return rc ? rc : fail(ctx, "another rank failed in the contraction");
			all_ok = all_ok && all_paths[4 * r + 3] != 0;
		}
		if (all_ok) {
			*used_paths = 1;
			return tagpu_dist_graph_paths(ctx, all_stats, all_paths, (flags & 2) ? 3 : 1);
		}
	}
	return tagpu_dist_graph(ctx, all_stats, with_graph ? 1 : 0);
}

extern "C" void tagpu_dist_disconnect(tagpu_ctx *ctx)
{
	cudaSetDevice(ctx->device);
	dist_unmap(ctx);
}

extern "C" void tagpu_dist_close(tagpu_ctx *ctx)
{
	cudaSetDevice(ctx->device);
	dist_release(ctx);
}

static int upload(tagpu_ctx *ctx, const uint8_t *h_seq, uint64_t n)
{
	CU(cudaSetDevice(ctx->device));
	if (ensure(ctx, ctx->seq, n + 64)) return -1;
	ctx->h_src = h_seq;                    // uploaded chunk by chunk inside partition_local, overlapped with pass 1
	return 0;
}

extern "C" int tagpu_build_device(tagpu_ctx *ctx, const uint8_t *d_seq, uint64_t n, int k)
{
	ctx->h_src = nullptr;
	ctx->src_packed = false;
	return run(ctx, d_seq, n, k + 1, true);
}
extern "C" int tagpu_count_device(tagpu_ctx *ctx, const uint8_t *d_seq, uint64_t n, int K)
{
	ctx->h_src = nullptr;
	ctx->src_packed = false;
	return run(ctx, d_seq, n, K, false);
}
extern "C" int tagpu_build_host(tagpu_ctx *ctx, const uint8_t *h_seq, uint64_t n, int k)
{
	ctx->src_packed = false;
	const int rc = upload(ctx, h_seq, n) ? -1 : run(ctx, (const uint8_t *)ctx->seq.p, n, k + 1, true);
	ctx->src_ready = nullptr;          // (one use only, also when the build failed before the upload)
	return rc;
}
// Contig-file mode of the stage entry points (n_files < 0, /root/reference/src/kmer_build.c:722-731,779-781): the graph
// from stream A (reads + the contig file), its edge counts from the solid (k+1)-mers of stream B (the reads alone).
// A is built with the one-level graph stage (the count lookup needs every k-mer in the table); B then goes through the
// count stage only — it shares no buffer with the graph — and k_edge_counts_lookup replaces A's edge counts by B's.
template <int W>
static int counts_from_current_solid(tagpu_ctx *ctx, uint64_t n_b, uint64_t n_e)
{
	KTab<W> t;
	t.keys = (Key<W> *)ctx->kt_keys.p;
	t.mask32 = (uint32_t *)ctx->kt_mask.p;
	t.n_slots = ctx->kt_slots;
	FlatGraph g;
	g.e_src = (uint32_t *)ctx->e_src.p; g.e_dst = (uint32_t *)ctx->e_dst.p; g.e_rc = (uint32_t *)ctx->e_rc.p;
	g.e_len = (uint32_t *)ctx->e_len.p; g.e_count = (unsigned long long *)ctx->e_count.p;
	g.e_off = (unsigned long long *)ctx->e_off.p; g.e_seq = (uint32_t *)ctx->e_seq.p;
	// (graph A was built with its own counts, for "Number of (k+1)-mer on edge" — the size of the reference's edge index,
	// kmer_build.c:772 —; they are dropped here and replaced by B's)
	CU(cudaMemsetAsync(g.e_count, 0, (size_t)(n_e + 1) * 8, ctx->stream));
	if (n_b)
		LAUNCH(k_edge_counts_lookup<W>, (unsigned)((n_b + 255) / 256), 256, (const Key<W> *)ctx->cur_solid_key, (const uint32_t *)ctx->cur_solid_cnt, n_b,
		       ctx->k, t, (const uint32_t *)ctx->node_ord.p, (const uint32_t *)ctx->node_ebase.p, (const uint32_t *)ctx->vedge.p, g, ctx->d_ctr);
	return read_counters(ctx);
}

extern "C" int tagpu_build_host_counts_from(tagpu_ctx *ctx, const uint8_t *h_a, uint64_t n_a, const uint8_t *h_b, uint64_t n_b, int k)
{
	const int contract = ctx->contract, skip = ctx->skip_counts;
	ctx->contract = 0;
	ctx->skip_counts = 0;
	ctx->src_packed = false;
	int rc = upload(ctx, h_a, n_a) ? -1 : run(ctx, (const uint8_t *)ctx->seq.p, n_a, k + 1, true);
	ctx->src_ready = nullptr;
	ctx->contract = contract;
	ctx->skip_counts = skip;
	if (rc) return rc;
	const tagpu_stats st_a = ctx->st;
	rc = upload(ctx, h_b, n_b) ? -1 : run(ctx, (const uint8_t *)ctx->seq.p, n_b, k + 1, false);
	if (rc) return rc;
	const uint64_t n_solid_b = ctx->st.n_solid;
	rc = ctx->W == 1 ? counts_from_current_solid<1>(ctx, n_solid_b, st_a.n_e) : counts_from_current_solid<2>(ctx, n_solid_b, st_a.n_e);
	if (rc) return rc;
	// the result is graph A; all figures (instances, distinct, solid, (k+1)-mers on edges) are those of A's (k+1)-mer set
	ctx->st = st_a;
	ctx->have_graph = true;
	ctx->have_count = false;           // (the solid list on the device is B's, not the graph's)
	ctx->contracted = false;
	return 0;
}

extern "C" int tagpu_count_host(tagpu_ctx *ctx, const uint8_t *h_seq, uint64_t n, int K)
{
	ctx->src_packed = false;
	const int rc = upload(ctx, h_seq, n) ? -1 : run(ctx, (const uint8_t *)ctx->seq.p, n, K, false);
	ctx->src_ready = nullptr;
	return rc;
}

// ---- packed read stream (include/tagpu.h): n_positions = bytes of the ASCII stream it was packed from
extern "C" uint64_t tagpu_packed_bytes(uint64_t n_positions)
{
	return (n_positions + TAGPU_TILE_BASES - 1) / TAGPU_TILE_BASES * (uint64_t)TAGPU_PACKED_TILE_BYTES;
}
static int run_packed(tagpu_ctx *ctx, const uint8_t *d_packed, uint64_t n_positions, int K, bool with_graph)
{
	ctx->src_packed = true;
	const int rc = run(ctx, d_packed, n_positions, K, with_graph);
	ctx->src_packed = false;
	return rc;
}
extern "C" int tagpu_build_device_packed(tagpu_ctx *ctx, const uint8_t *d_packed, uint64_t n_positions, int k)
{
	ctx->h_src = nullptr;
	return run_packed(ctx, d_packed, n_positions, k + 1, true);
}
extern "C" int tagpu_build_host_packed(tagpu_ctx *ctx, const uint8_t *h_packed, uint64_t n_positions, int k)
{
	ctx->src_packed = true;            // (upload() stages tagpu_packed_bytes bytes and leaves the chunked copy to pass 1)
	if (upload(ctx, h_packed, tagpu_packed_bytes(n_positions))) return -1;
	return run_packed(ctx, (const uint8_t *)ctx->seq.p, n_positions, k + 1, true);
}
extern "C" int tagpu_count_host_packed(tagpu_ctx *ctx, const uint8_t *h_packed, uint64_t n_positions, int K)
{
	ctx->src_packed = true;
	if (upload(ctx, h_packed, tagpu_packed_bytes(n_positions))) return -1;
	return run_packed(ctx, (const uint8_t *)ctx->seq.p, n_positions, K, false);
}

// ------------------------------------------------------------------------------------------------ raw FASTQ files on the device
// The files entry points read plain FASTQ files into a pinned ring (one kernel copy per byte, no parsing on the host), the
// ring's slots go up as they fill, and the records are parsed here (tagpu_fastq.cuh).
extern "C" void *tagpu_raw_ring(tagpu_ctx *ctx, size_t bytes)
{
	if (cudaSetDevice(ctx->device) != cudaSuccess) return nullptr;
	if (ctx->raw_ring_bytes < bytes) {
		if (ctx->raw_ring) cudaFreeHost(ctx->raw_ring);
		ctx->raw_ring = nullptr;
		ctx->raw_ring_bytes = 0;
		if (cudaMallocHost(&ctx->raw_ring, bytes) != cudaSuccess) { cudaGetLastError(); ctx->raw_ring = nullptr; return nullptr; }
		ctx->raw_ring_bytes = bytes;
	}
	if (!ctx->ev_slot_made) {
		for (int i = 0; i < TAGPU_RAW_SLOTS_MAX; ++i) cudaEventCreateWithFlags(&ctx->ev_slot[i], cudaEventDisableTiming);
		cudaEventCreateWithFlags(&ctx->ev_raw, cudaEventDisableTiming);
		ctx->ev_slot_made = true;
	}
	return ctx->raw_ring;
}

extern "C" int tagpu_raw_begin(tagpu_ctx *ctx, uint64_t total_bytes)
{
	CU(cudaSetDevice(ctx->device));
	return ensure(ctx, ctx->raw, total_bytes + 256);
}

// bytes of a ring slot -> the device file buffer at dev_off (asynchronous; tagpu_raw_slot_wait tells when the slot is free again)
extern "C" int tagpu_raw_put(tagpu_ctx *ctx, uint64_t dev_off, const void *h, uint64_t bytes, int slot)
{
	if (slot < 0 || slot >= TAGPU_RAW_SLOTS_MAX || !ctx->ev_slot_made) return fail(ctx, "tagpu_raw_put: bad slot %d", slot);
	CU(cudaMemcpyAsync((uint8_t *)ctx->raw.p + dev_off, h, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
	CU(cudaEventRecord(ctx->ev_slot[slot], ctx->copy_stream));
	return 0;
}
extern "C" int tagpu_raw_slot_wait(tagpu_ctx *ctx, int slot)
{
	CU(cudaEventSynchronize(ctx->ev_slot[slot]));
	return 0;
}

static int scan_u32(tagpu_ctx *ctx, const uint32_t *in, uint64_t n, unsigned long long *out, unsigned long long *d_total)
{
	const uint64_t per = 1024ull * TAGPU_SCAN_ITEMS, nb = (n + per - 1) / per;
	if (ensure(ctx, ctx->fq_tot, (nb + 1) * 8)) return -1;
	unsigned long long *tot = (unsigned long long *)ctx->fq_tot.p;
	if (nb) LAUNCH(k_scan_a, (unsigned)nb, 1024, in, n, out, tot);
	LAUNCH(k_scan_b, 1, 1024, tot, nb, d_total);
	if (nb) LAUNCH(k_scan_c, (unsigned)nb, 1024, out, n, (const unsigned long long *)tot);
	return 0;
}

// raw file bytes (ctx->raw at off[f], len[f] bytes, ends_nl[f]: the last byte is a newline) -> the read stream in ctx->seq;
// returns its length, -1 on error
static int64_t parse_fastq(tagpu_ctx *ctx, int n_files, const uint64_t *off, const uint64_t *len, const uint8_t *ends_nl)
{
	if (cudaSetDevice(ctx->device) != cudaSuccess) return -1;
	uint64_t bound = 64;
	for (int f = 0; f < n_files; ++f) bound += len[f] / 2 + 2;
	if (ensure(ctx, ctx->seq, bound)) return -1;
	// everything the ring sent up must have landed before the first kernel reads it
	if (ctx->ev_slot_made) {
		if (cudaEventRecord(ctx->ev_raw, ctx->copy_stream) != cudaSuccess || cudaStreamWaitEvent(ctx->stream, ctx->ev_raw, 0) != cudaSuccess) return -1;
	}
	unsigned long long *d_total = ctx->d_ctr + CTR_SPARE3;          // (scratch: the counters are reset by the build that follows)
	uint64_t acc = 0;
	for (int f = 0; f < n_files; ++f) {
		const uint8_t *raw = (const uint8_t *)ctx->raw.p + off[f];
		const uint64_t n = len[f], n_blk = (n + TAGPU_FQ_BLOCK_BYTES - 1) / TAGPU_FQ_BLOCK_BYTES;
		if (!n) continue;
		if (ensure(ctx, ctx->fq_cnt, (n_blk + 1) * 4) || ensure(ctx, ctx->fq_base, (n_blk + 1) * 8)) return -1;
		LAUNCH(k_fq_count, (unsigned)n_blk, TAGPU_FQ_THREADS, raw, n, (uint32_t *)ctx->fq_cnt.p);
		if (scan_u32(ctx, (const uint32_t *)ctx->fq_cnt.p, n_blk, (unsigned long long *)ctx->fq_base.p, d_total)) return -1;
		unsigned long long n_nl = 0;
		if (cudaMemcpyAsync(&n_nl, d_total, 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess) return -1;
		// sequence lines = lines 1, 5, 9, ...; an unterminated last line counts if it is one of them
		const bool tail = !ends_nl[f] && (n_nl & 3ull) == 1ull;
		const uint64_t n_rec = (n_nl + 2) / 4 + (tail ? 1 : 0);
		if (!n_rec) continue;
		if (ensure(ctx, ctx->fq_lo, n_rec * 8) || ensure(ctx, ctx->fq_hi, n_rec * 8) || ensure(ctx, ctx->fq_len, n_rec * 4) || ensure(ctx, ctx->fq_off, n_rec * 8)) return -1;
		unsigned long long *lo = (unsigned long long *)ctx->fq_lo.p, *hi = (unsigned long long *)ctx->fq_hi.p, *o = (unsigned long long *)ctx->fq_off.p;
		LAUNCH(k_fq_mark, (unsigned)n_blk, TAGPU_FQ_THREADS, raw, n, (const unsigned long long *)ctx->fq_base.p, lo, hi, n_rec);
		if (tail && cudaMemcpyAsync(hi + (n_rec - 1), &n, 8, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) return -1;
		LAUNCH(k_fq_len, (unsigned)((n_rec + 255) / 256), 256, raw, (const unsigned long long *)lo, hi, n_rec, (uint32_t *)ctx->fq_len.p);
		if (scan_u32(ctx, (const uint32_t *)ctx->fq_len.p, n_rec, o, d_total)) return -1;
		unsigned long long n_out = 0;
		if (cudaMemcpyAsync(&n_out, d_total, 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess) return -1;
		if (acc + n_out + 64 > ctx->seq.cap) {
			// more sequence than half the bytes of the files (not FASTQ-shaped, e.g. no quality lines): grow, keeping the streams so far
			void *grown = nullptr;
			const size_t cap = (size_t)(acc + n_out + 64) * 2;
			if (cudaMalloc(&grown, cap) != cudaSuccess) { cudaGetLastError(); return -1; }
			if (acc && cudaMemcpyAsync(grown, ctx->seq.p, acc, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess) return -1;
			if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return -1;
			cudaFree(ctx->seq.p);
			ctx->seq.p = grown;
			ctx->seq.cap = cap;
		}
		LAUNCH(k_fq_copy, (unsigned)((n_rec * 32 + 255) / 256), 256, raw, (const unsigned long long *)lo, (const unsigned long long *)hi,
		       (const unsigned long long *)o, n_rec, (uint8_t *)ctx->seq.p + acc);
		acc += n_out;
	}
	return (int64_t)acc;
}

// test / tool entry: parse only, optionally copy the stream back (h_out must hold the returned number of bytes: call twice)
extern "C" int64_t tagpu_parse_fastq_device(tagpu_ctx *ctx, int n_files, const uint64_t *off, const uint64_t *len, const uint8_t *ends_nl, uint8_t *h_out)
{
	ctx->launches = 0;
	const int64_t n = parse_fastq(ctx, n_files, off, len, ends_nl);
	if (n < 0) return n;
	if (h_out && n && (cudaMemcpyAsync(h_out, ctx->seq.p, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)) return -1;
	if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return -1;
	return n;
}

// parse + build (with_graph) or count only; k = node size (the count stage works on K = k + 1)
extern "C" int tagpu_build_fastq_device(tagpu_ctx *ctx, int n_files, const uint64_t *off, const uint64_t *len, const uint8_t *ends_nl, int k, int with_graph)
{
	ctx->h_src = nullptr;
	ctx->src_packed = false;
	ctx->src_ready = nullptr;
	const int64_t n = parse_fastq(ctx, n_files, off, len, ends_nl);
	if (n < 0) return fail(ctx, "FASTQ parsing on the device failed (%s)", cudaGetErrorString(cudaGetLastError()));
	return run(ctx, (const uint8_t *)ctx->seq.p, (uint64_t)n, k + 1, with_graph != 0);
}

// build_local_assembly_graph: reads + the two flanking contigs of the global graph.
// 1. the contigs alone go through the count stage with cutoff 1 -> their distinct canonical (k+1)-mers (the "garbage");
// 2. the reads are counted with the context's cutoff; 3. the graph stage runs on solid ++ garbage; 4. k_garbage_counts.
extern "C" int tagpu_build_local_host(tagpu_ctx *ctx, const uint8_t *h_reads, uint64_t n_bytes, int k, const uint8_t *h_contigs,
				      uint64_t n_contig_bytes, int n_contigs, const uint64_t *contig_off, const uint32_t *contig_len,
				      const double *contig_cov)
{
	if (n_contigs < 0 || n_contigs > 4) return fail(ctx, "at most 4 flanking contigs (got %d)", n_contigs);
	if (ctx->dist) return fail(ctx, "context is in multi-GPU mode");
	CU(cudaSetDevice(ctx->device));
	ctx->src_packed = false;
	ctx->local_mode = false;
	ctx->n_garbage = 0;
	ctx->n_contigs = 0;
	if (n_contigs && n_contig_bytes) {
		if (ensure(ctx, ctx->g_seq, n_contig_bytes + 64)) return -1;
		CU(cudaMemcpyAsync(ctx->g_seq.p, h_contigs, n_contig_bytes, cudaMemcpyHostToDevice, ctx->stream));
		const int ci = ctx->ci;
		ctx->ci = 1;
		ctx->h_src = nullptr;
		const int rc = run(ctx, (const uint8_t *)ctx->g_seq.p, n_contig_bytes, k + 1, false);
		ctx->ci = ci;
		if (rc) return rc;
		const size_t key = ctx->W == 1 ? sizeof(Key<1>) : sizeof(Key<2>);
		const uint64_t ng = ctx->st.n_solid;
		if (ensure(ctx, ctx->g_key, (ng + 1) * key)) return -1;
		if (ng) CU(cudaMemcpyAsync(ctx->g_key.p, ctx->cur_solid_key, ng * key, cudaMemcpyDeviceToDevice, ctx->stream));
		CU(cudaStreamSynchronize(ctx->stream));
		ctx->n_garbage = ng;
		ctx->n_contigs = n_contigs;
		for (int c = 0; c < n_contigs; ++c) {
			ctx->contig_off[c] = contig_off[c];
			ctx->contig_len[c] = contig_len[c];
			ctx->contig_cov[c] = contig_cov[c];
		}
	}
	ctx->local_mode = true;                                     // keeps the garbage across the run() of the reads
	int rc = upload(ctx, h_reads, n_bytes);
	if (!rc) rc = run(ctx, (const uint8_t *)ctx->seq.p, n_bytes, k + 1, true);
	ctx->local_mode = false;
	ctx->n_garbage = 0;
	ctx->n_contigs = 0;
	return rc;
}

extern "C" int tagpu_get_stats(tagpu_ctx *ctx, struct tagpu_stats *out)
{
	*out = ctx->st;
	return 0;
}

// ------------------------------------------------------------------------------------------------ device -> host
extern "C" int tagpu_copy_solid(tagpu_ctx *ctx, uint64_t *hi, uint64_t *lo, uint32_t *count)
{
	if (!ctx->have_count) return fail(ctx, "no count result to copy");
	if (ctx->solid_sharded)
		return fail(ctx, "the solid set of this multi-GPU build stayed with its owner ranks (tagpu_dist_graph_paths with_graph = 3 gathers it)");
	CU(cudaSetDevice(ctx->device));
	const uint64_t n = ctx->st.n_solid;
	if (!n) return 0;
	CU(cudaMemcpyAsync(count, ctx->cur_solid_cnt, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
	if (ctx->W == 1) {
		CU(cudaMemcpyAsync(lo, ctx->cur_solid_key, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
		CU(cudaStreamSynchronize(ctx->stream));
		memset(hi, 0, n * 8);
	} else {
		std::vector<Key<2>> tmp(n);
		CU(cudaMemcpyAsync(tmp.data(), ctx->cur_solid_key, n * 16, cudaMemcpyDeviceToHost, ctx->stream));
		CU(cudaStreamSynchronize(ctx->stream));
		for (uint64_t i = 0; i < n; ++i) { hi[i] = tmp[i].hi; lo[i] = tmp[i].lo; }
	}
	return 0;
}

// k-mers + masks out of a table on the device (~key stored, 0 = empty slot)
static int copy_table_entries(tagpu_ctx *ctx, const void *d_keys, const void *d_mask, uint32_t n_slots, uint64_t *hi, uint64_t *lo, uint8_t *mask)
{
	std::vector<uint8_t> m(n_slots);
	std::vector<uint64_t> keys((size_t)n_slots * ctx->W);
	CU(cudaMemcpyAsync(m.data(), d_mask, n_slots, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaMemcpyAsync(keys.data(), d_keys, (size_t)n_slots * 8 * ctx->W, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	uint64_t o = 0;
	for (uint32_t s = 0; s < n_slots; ++s) {
		uint64_t l = keys[(size_t)s * ctx->W], h = ctx->W == 2 ? keys[(size_t)s * 2 + 1] : 0;
		if ((l | h) == 0) continue;
		if (o >= ctx->st.n_kmers) return fail(ctx, "k-mer table holds more entries than counted");
		lo[o] = ~l;
		hi[o] = ctx->W == 2 ? ~h : 0;
		mask[o] = m[s];
		++o;
	}
	if (o != ctx->st.n_kmers) return fail(ctx, "k-mer table holds %llu entries, counted %llu", (unsigned long long)o, (unsigned long long)ctx->st.n_kmers);
	return 0;
}

// The two-level graph stage keeps only the path-end k-mers in its table; when a caller asks for the full k-mer table it
// is built here, on demand, from the solid list (same kernel as the one-level stage) in scratch memory.
template <int W>
static int copy_kmers_full_table(tagpu_ctx *ctx, uint64_t *hi, uint64_t *lo, uint8_t *mask)
{
	const uint64_t n_solid = ctx->st.n_solid, slots64 = (ctx->st.n_kmers * 5) / 2 + 1024;
	if (slots64 > (1ull << 30)) return fail(ctx, "k-mer table would need %llu slots (> 2^30)", (unsigned long long)slots64);
	const uint32_t n_slots = (uint32_t)slots64;
	const size_t mask_bytes = ((size_t)n_slots + 3) / 4 * 4;
	Buf keys, msk, vl, vr, c;
	int rc = -1;
	if (!(ensure(ctx, keys, (size_t)n_slots * sizeof(Key<W>)) || ensure(ctx, msk, mask_bytes) || ensure(ctx, vl, (n_solid + 1) * 4) ||
	      ensure(ctx, vr, (n_solid + 1) * 4) || ensure(ctx, c, CTR_TOTAL * sizeof(unsigned long long)))) {
		KTab<W> t;
		t.keys = (Key<W> *)keys.p;
		t.mask32 = (uint32_t *)msk.p;
		t.n_slots = n_slots;
		rc = 0;
		if (cudaMemsetAsync(t.keys, 0, (size_t)n_slots * sizeof(Key<W>), ctx->stream) != cudaSuccess ||
		    cudaMemsetAsync(t.mask32, 0, mask_bytes, ctx->stream) != cudaSuccess ||
		    cudaMemsetAsync(c.p, 0, CTR_TOTAL * sizeof(unsigned long long), ctx->stream) != cudaSuccess)
			rc = fail(ctx, "cudaMemsetAsync failed");
		if (!rc && n_solid) {
			k_insert_kmers<W><<<(unsigned)((n_solid + 255) / 256), 256, 0, ctx->stream>>>((const Key<W> *)ctx->cur_solid_key, 0ull, n_solid, 0, ctx->k, t,
												   (uint32_t *)vl.p, (uint32_t *)vr.p, (unsigned long long *)c.p);
			if (cudaGetLastError() != cudaSuccess) rc = fail(ctx, "k_insert_kmers launch failed");
		}
		if (!rc) rc = copy_table_entries(ctx, keys.p, msk.p, n_slots, hi, lo, mask);
	}
	Buf *tmp[] = { &keys, &msk, &vl, &vr, &c };
	for (Buf *b : tmp) if (b->p) cudaFree(b->p);
	return rc;
}

extern "C" int tagpu_copy_kmers(tagpu_ctx *ctx, uint64_t *hi, uint64_t *lo, uint8_t *mask)
{
	if (!ctx->have_graph) return fail(ctx, "no graph result to copy");
	CU(cudaSetDevice(ctx->device));
	if (ctx->contracted && ctx->solid_sharded)
		return fail(ctx, "the solid set of this multi-GPU build stayed with its owner ranks (tagpu_dist_graph_paths with_graph = 3 gathers it)");
	if (ctx->contracted) return ctx->W == 1 ? copy_kmers_full_table<1>(ctx, hi, lo, mask) : copy_kmers_full_table<2>(ctx, hi, lo, mask);
	return copy_table_entries(ctx, ctx->kt_keys.p, ctx->kt_mask.p, ctx->kt_slots, hi, lo, mask);
}

// Order-independent digests of the last build (csrc/tagpu_digest.cuh), computed on the device:
//   out[0], out[1]  sum / xor over the solid (k+1)-mers this context holds; out[2] = how many those are;
//   out[3]          1 if they are the whole solid set, 0 if only this rank's share of a sharded multi-GPU build
//                   (sums and xors of the ranks then add up to the digest of the whole set);
//   out[4], out[5]  sum / xor over all edges; out[6] = sum of edge lengths, out[7] = sum of edge counts, out[8] = n_e.
extern "C" int tagpu_digest(tagpu_ctx *ctx, uint64_t out[9])
{
	if (!ctx->have_count) return fail(ctx, "no result to digest");
	CU(cudaSetDevice(ctx->device));
	unsigned long long *d = nullptr;
	CU(cudaMalloc(&d, 8 * sizeof(unsigned long long)));
	CU(cudaMemsetAsync(d, 0, 8 * sizeof(unsigned long long), ctx->stream));
	const uint64_t n_s = ctx->solid_sharded ? ctx->n_solid_local : ctx->st.n_solid;
	const void *key = ctx->solid_sharded ? ctx->solid_key.p : ctx->cur_solid_key;
	const void *cnt = ctx->solid_sharded ? ctx->solid_cnt.p : ctx->cur_solid_cnt;
	if (n_s) {
		if (ctx->W == 1) k_digest_solid<1><<<4 * ctx->n_sm, 256, 0, ctx->stream>>>((const Key<1> *)key, (const uint32_t *)cnt, n_s, d);
		else k_digest_solid<2><<<4 * ctx->n_sm, 256, 0, ctx->stream>>>((const Key<2> *)key, (const uint32_t *)cnt, n_s, d);
	}
	if (ctx->have_graph && ctx->st.n_e)
		k_digest_edges<<<4 * ctx->n_sm, 256, 0, ctx->stream>>>((const uint32_t *)ctx->e_len.p, (const unsigned long long *)ctx->e_count.p,
									(const unsigned long long *)ctx->e_off.p, (const uint32_t *)ctx->e_seq.p, ctx->st.n_e, d + 4);
	unsigned long long h[8];
	cudaError_t e = cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
	cudaFree(d);
	if (e != cudaSuccess) return fail(ctx, "digest failed: %s", cudaGetErrorString(e));
	out[0] = h[0]; out[1] = h[1]; out[2] = n_s; out[3] = ctx->solid_sharded ? 0 : 1;
	out[4] = h[4]; out[5] = h[5]; out[6] = h[6]; out[7] = h[7]; out[8] = ctx->have_graph ? ctx->st.n_e : 0;
	return 0;
}

extern "C" int tagpu_copy_graph(tagpu_ctx *ctx, struct tagpu_flat_graph *h)
{
	if (!ctx->have_graph) return fail(ctx, "no graph result to copy");
	CU(cudaSetDevice(ctx->device));
	const uint64_t n_nodes = ctx->st.n_v / 2, n_e = ctx->st.n_e, n_w = ctx->st.n_seq_words;
	h->n_nodes = n_nodes; h->n_e = n_e; h->n_seq_words = n_w;
	if (n_nodes) {
		// masks gathered on the device (n_nodes bytes travel instead of the whole table's mask array)
		if (ensure_slack(ctx, ctx->node_mask, n_nodes + 4)) return -1;
		k_gather_node_masks<<<(unsigned)((n_nodes + 255) / 256), 256, 0, ctx->stream>>>((const uint32_t *)ctx->node_slot.p, (const uint32_t *)ctx->kt_mask.p,
												  n_nodes, (uint8_t *)ctx->node_mask.p);
		CU(cudaGetLastError());
		CU(cudaMemcpyAsync(h->node_mask, ctx->node_mask.p, n_nodes, cudaMemcpyDeviceToHost, ctx->stream));
		CU(cudaMemcpyAsync(h->node_ebase, ctx->node_ebase.p, n_nodes * 4, cudaMemcpyDeviceToHost, ctx->stream));
	}
	if (n_e) {
		CU(cudaMemcpyAsync(h->e_src, ctx->e_src.p, n_e * 4, cudaMemcpyDeviceToHost, ctx->stream));
		CU(cudaMemcpyAsync(h->e_dst, ctx->e_dst.p, n_e * 4, cudaMemcpyDeviceToHost, ctx->stream));
		CU(cudaMemcpyAsync(h->e_rc, ctx->e_rc.p, n_e * 4, cudaMemcpyDeviceToHost, ctx->stream));
		CU(cudaMemcpyAsync(h->e_len, ctx->e_len.p, n_e * 4, cudaMemcpyDeviceToHost, ctx->stream));
		CU(cudaMemcpyAsync(h->e_count, ctx->e_count.p, n_e * 8, cudaMemcpyDeviceToHost, ctx->stream));
		CU(cudaMemcpyAsync(h->e_off, ctx->e_off.p, n_e * 8, cudaMemcpyDeviceToHost, ctx->stream));
		CU(cudaMemcpyAsync(h->e_seq, ctx->e_seq.p, n_w * 4, cudaMemcpyDeviceToHost, ctx->stream));
	}
	CU(cudaStreamSynchronize(ctx->stream));
	return 0;
}

// ------------------------------------------------------------------------------------------------ coverage recount (row f4)
// kmer_count_on_edges + add_cnt_to_graph (/root/reference/src/coverage/kmer_count.c:198-240,113-135) for a read stream in
// host memory.  Edges: the flat arrays given (host pointers; e_off in 32-bit words), or — with h_len == NULL — the graph
// of this context's last build, already on the device.  h_count_out[e] = the count the reference leaves in g->edges[e].
extern "C" int tagpu_coverage_recount_host(tagpu_ctx *ctx, const uint8_t *h_seq, uint64_t n_bytes, uint64_t n_e, const uint32_t *h_len,
					   const uint64_t *h_off, const uint32_t *h_words, uint64_t n_words, const uint32_t *h_rc,
					   uint64_t *h_count_out)
{
	CU(cudaSetDevice(ctx->device));
	const uint32_t *d_len, *d_words, *d_rc;
	const unsigned long long *d_off;
	Buf b_len, b_off, b_words, b_rc, b_key, b_cnt, b_raw, b_out, b_err;
	int rc = -1;
	do {
		if (h_len) {
			if (ensure(ctx, b_len, (n_e + 1) * 4) || ensure(ctx, b_off, (n_e + 1) * 8) || ensure(ctx, b_words, (n_words + 4) * 4) || ensure(ctx, b_rc, (n_e + 1) * 4)) break;
			if (cudaMemsetAsync(b_words.p, 0, (n_words + 4) * 4, ctx->stream) != cudaSuccess) break;
			if (n_e && (cudaMemcpyAsync(b_len.p, h_len, n_e * 4, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
				    cudaMemcpyAsync(b_off.p, h_off, n_e * 8, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
				    cudaMemcpyAsync(b_rc.p, h_rc, n_e * 4, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
				    (n_words && cudaMemcpyAsync(b_words.p, h_words, n_words * 4, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)))
				break;
			d_len = (const uint32_t *)b_len.p; d_off = (const unsigned long long *)b_off.p; d_words = (const uint32_t *)b_words.p; d_rc = (const uint32_t *)b_rc.p;
		} else {
			if (!ctx->have_graph) { fail(ctx, "no graph to recount (build one first, or pass the edges)"); break; }
			n_e = ctx->st.n_e;
			n_words = ctx->st.n_seq_words;
			d_len = (const uint32_t *)ctx->e_len.p; d_off = (const unsigned long long *)ctx->e_off.p; d_words = (const uint32_t *)ctx->e_seq.p; d_rc = (const uint32_t *)ctx->e_rc.p;
		}
		// every edge base starts at most one 31-mer: 2x that many slots keeps the load under 0.5
		CovTab t;
		t.n_slots = 2 * (unsigned long long)n_words * 16 + 1024;
		if (ensure(ctx, b_key, t.n_slots * 8) || ensure(ctx, b_cnt, t.n_slots * 8) || ensure(ctx, b_raw, (n_e + 1) * 8) || ensure(ctx, b_out, (n_e + 1) * 8) ||
		    ensure(ctx, b_err, 8) || ensure(ctx, ctx->seq, n_bytes + 64))
			break;
		t.key = (unsigned long long *)b_key.p;
		t.cnt = (unsigned long long *)b_cnt.p;
		if (cudaMemsetAsync(t.key, 0, t.n_slots * 8, ctx->stream) != cudaSuccess || cudaMemsetAsync(t.cnt, 0, t.n_slots * 8, ctx->stream) != cudaSuccess ||
		    cudaMemsetAsync(b_err.p, 0, 8, ctx->stream) != cudaSuccess ||
		    (n_bytes && cudaMemcpyAsync(ctx->seq.p, h_seq, n_bytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess))
			break;
		if (n_e) {
			k_cov_index<<<8 * ctx->n_sm, 256, 0, ctx->stream>>>(d_len, d_off, d_words, n_e, t, (unsigned long long *)b_err.p);
			++ctx->launches;
		}
		if (n_bytes) {
			const unsigned long long n_thr = (n_bytes + TAGPU_COV_SPAN - 1) / TAGPU_COV_SPAN;
			k_cov_count<<<(unsigned)((n_thr + 255) / 256), 256, 0, ctx->stream>>>((const uint8_t *)ctx->seq.p, n_bytes, t);
			++ctx->launches;
		}
		if (n_e) {
			k_cov_sum<<<8 * ctx->n_sm, 256, 0, ctx->stream>>>(d_len, d_off, d_words, n_e, t, (unsigned long long *)b_raw.p);
			k_cov_symmetric<<<(unsigned)((n_e + 255) / 256), 256, 0, ctx->stream>>>(d_rc, (const unsigned long long *)b_raw.p, n_e, (unsigned long long *)b_out.p);
			ctx->launches += 2;
		}
		unsigned long long err = 0;
		if (cudaGetLastError() != cudaSuccess ||
		    (n_e && cudaMemcpyAsync(h_count_out, b_out.p, n_e * 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) ||
		    cudaMemcpyAsync(&err, b_err.p, 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
			fail(ctx, "coverage recount failed: %s", cudaGetErrorString(cudaGetLastError()));
			break;
		}
		if (err) { fail(ctx, "coverage recount: 31-mer table full"); break; }
		rc = 0;
	} while (0);
	Buf *tmp[] = { &b_len, &b_off, &b_words, &b_rc, &b_key, &b_cnt, &b_raw, &b_out, &b_err };
	for (Buf *b : tmp) if (b->p) cudaFree(b->p);
	return rc;
}

// pinned host allocation for the FASTQ loader in tagpu_host.c (keeps cuda_runtime.h out of the C file).  Without a
// usable CUDA device (the CPU-only test box exercises the ingest logic) the buffer is ordinary memory; no compute
// path exists there anyway (tagpu_create fails).
static void *g_unpinned[16];

extern "C" void *tagpu_pinned_alloc(size_t bytes)
{
	void *p = nullptr;
	if (cudaMallocHost(&p, bytes ? bytes : 1) == cudaSuccess) return p;
	cudaGetLastError();
	p = malloc(bytes ? bytes : 1);
	for (int i = 0; p && i < 16; ++i)
		if (!g_unpinned[i]) { g_unpinned[i] = p; return p; }
	free(p);
	return nullptr;
}
extern "C" void tagpu_pinned_free(void *p)
{
	if (!p) return;
	for (int i = 0; i < 16; ++i)
		if (g_unpinned[i] == p) { g_unpinned[i] = nullptr; free(p); return; }
	cudaFreeHost(p);
}
extern "C" int tagpu_ctx_k(tagpu_ctx *ctx) { return ctx->k; }
extern "C" int tagpu_ctx_K(tagpu_ctx *ctx) { return ctx->K; }
extern "C" int tagpu_ctx_cutoff(tagpu_ctx *ctx) { return ctx->ci; }

// developer introspection: per-bucket cursors of the last partition pass (low 32 bits records, high 32 bits instances)
extern "C" uint64_t tagpu_debug_bucket_cursors(tagpu_ctx *ctx, uint64_t *out, uint64_t max_n)
{
	uint64_t n = ctx->cursor.cap / 8;
	if (n > max_n) n = max_n;
	cudaMemcpy(out, ctx->cursor.p, n * 8, cudaMemcpyDeviceToHost);
	return n;
}
