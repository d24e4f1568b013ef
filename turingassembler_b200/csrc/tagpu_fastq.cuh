// FASTQ record parsing on the device: raw file bytes in HBM -> the read stream (sequence lines joined by '\n').
//
// The host's part of the files entry points then is one kernel copy per byte (pread into a pinned ring) — the newline scan,
// the line numbering and the copy of the sequence lines (/root/reference/src/get_buffer.c:339-348: the sequence is line 2
// of every 4; no record-boundary guessing) run here, on whole files, with the rules of the host parser
// (tagpu_host.c:pfq_walk_at; tests/test_ingest.py holds the cases): lines end at '\n'; a '\r' right before the line's end is
// dropped; an unterminated last sequence line gets its '\n'; an empty sequence line still contributes its '\n'.
//
//   k_fq_count   newlines per 4 KB block                      -> scan -> line number at every block start
//   k_fq_mark    line l ends at byte p:  l % 4 == 0 -> the sequence line of record l / 4 starts at p + 1,
//                                        l % 4 == 1 -> it ends at p
//   k_fq_len     stream bytes of every record                 -> scan -> offset of every record in the stream
//   k_fq_copy    one warp per record copies its sequence line
#pragma once
#include "tagpu_key.cuh"

constexpr int TAGPU_FQ_THREADS = 256, TAGPU_FQ_BLOCK_BYTES = TAGPU_FQ_THREADS * 16;

// bit i = byte i of the 16 is '\n' (bytes at or behind `len` never match)
TAGPU_DI uint32_t tagpu_fq_newlines16(const uint8_t *__restrict__ raw, uint64_t len, uint64_t pos)
{
	if (pos >= len) return 0u;
	const uint4 v = __ldg(reinterpret_cast<const uint4 *>(raw + pos));   // (the buffer is allocated in whole 16-byte units)
	const uint32_t w[4] = { v.x, v.y, v.z, v.w };
	uint32_t mask = 0;
#pragma unroll
	for (int i = 0; i < 4; ++i) {
		const uint32_t x = w[i] ^ 0x0a0a0a0au;
		const uint32_t nz = (((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u;     // bit 7 of every byte that is NOT '\n'
		const uint32_t z = (~nz & 0x80808080u) >> 7;                                   // bit 0 of every byte that is
		mask |= (((z * 0x00204081u) >> 21) & 0xfu) << (4 * i);
	}
	const uint64_t left = len - pos;
	return left >= 16 ? mask : mask & ((1u << (uint32_t)left) - 1u);
}

__global__ void __launch_bounds__(TAGPU_FQ_THREADS) k_fq_count(const uint8_t *__restrict__ raw, uint64_t len, uint32_t *__restrict__ blk_cnt)
{
	__shared__ uint32_t s_w[TAGPU_FQ_THREADS / 32];
	const uint64_t pos = ((uint64_t)blockIdx.x * TAGPU_FQ_THREADS + threadIdx.x) * 16;
	uint32_t c = __popc(tagpu_fq_newlines16(raw, len, pos));
	c = __reduce_add_sync(0xffffffffu, c);
	if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
	__syncthreads();
	if (threadIdx.x == 0) {
		uint32_t t = 0;
#pragma unroll
		for (int i = 0; i < TAGPU_FQ_THREADS / 32; ++i) t += s_w[i];
		blk_cnt[blockIdx.x] = t;
	}
}

__global__ void __launch_bounds__(TAGPU_FQ_THREADS) k_fq_mark(const uint8_t *__restrict__ raw, uint64_t len, const unsigned long long *__restrict__ blk_base,
							       unsigned long long *__restrict__ rec_lo, unsigned long long *__restrict__ rec_hi, uint64_t n_rec)
{
	__shared__ uint32_t s_w[TAGPU_FQ_THREADS / 32];
	const uint64_t pos = ((uint64_t)blockIdx.x * TAGPU_FQ_THREADS + threadIdx.x) * 16;
	uint32_t mask = tagpu_fq_newlines16(raw, len, pos);
	const uint32_t c = __popc(mask), lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	uint32_t incl = c;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
		if (lane >= (uint32_t)d) incl += t;
	}
	if (lane == 31) s_w[warp] = incl;
	__syncthreads();
	uint32_t before = 0;
#pragma unroll
	for (int i = 0; i < TAGPU_FQ_THREADS / 32; ++i) before += (uint32_t)i < warp ? s_w[i] : 0u;
	unsigned long long line = blk_base[blockIdx.x] + before + incl - c;     // number of the line that ends at the thread's first newline
	while (mask) {
		const uint64_t p = pos + (uint64_t)(__ffs(mask) - 1);
		mask &= mask - 1;
		const unsigned long long r = line >> 2;
		if (r < n_rec) {
			if ((line & 3ull) == 0ull) rec_lo[r] = p + 1;
			else if ((line & 3ull) == 1ull) rec_hi[r] = p;
		}
		++line;
	}
}

// stream bytes of record r: its sequence line without a trailing '\r', plus the '\n' that ends the read.  n_rec counts an
// unterminated last sequence line too (its end is the end of the file: rec_hi was pre-set to len by the host side).
__global__ void __launch_bounds__(256) k_fq_len(const uint8_t *__restrict__ raw, const unsigned long long *__restrict__ rec_lo,
						 unsigned long long *__restrict__ rec_hi, uint64_t n_rec, uint32_t *__restrict__ out_len)
{
	const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n_rec) return;
	const unsigned long long lo = rec_lo[r];
	unsigned long long hi = rec_hi[r];
	if (hi > lo && raw[hi - 1] == '\r') --hi;
	rec_hi[r] = hi;
	out_len[r] = (uint32_t)(hi - lo) + 1u;
}

__global__ void __launch_bounds__(256) k_fq_copy(const uint8_t *__restrict__ raw, const unsigned long long *__restrict__ rec_lo,
						  const unsigned long long *__restrict__ rec_hi, const unsigned long long *__restrict__ out_off,
						  uint64_t n_rec, uint8_t *__restrict__ stream)
{
	const uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const uint32_t lane = threadIdx.x & 31u;
	if (r >= n_rec) return;
	const unsigned long long lo = rec_lo[r], n = rec_hi[r] - lo;
	uint8_t *dst = stream + out_off[r];
	for (unsigned long long i = lane; i < n; i += 32) dst[i] = raw[lo + i];
	if (lane == 0) dst[n] = '\n';
}

// ---------------------------------------------------------------- exclusive scan, uint32 in -> uint64 out, any length
// k_scan_a: every block scans 4096 items and leaves its total; k_scan_b: one block scans the totals (and leaves the grand
// total in *total); k_scan_c: adds the block offsets.
constexpr int TAGPU_SCAN_ITEMS = 4;
__global__ void __launch_bounds__(1024) k_scan_a(const uint32_t *__restrict__ in, uint64_t n, unsigned long long *__restrict__ out,
						  unsigned long long *__restrict__ tot)
{
	__shared__ unsigned long long s_w[32];
	const uint64_t i0 = ((uint64_t)blockIdx.x * 1024 + threadIdx.x) * TAGPU_SCAN_ITEMS;
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	uint32_t v[TAGPU_SCAN_ITEMS];
	unsigned long long mine = 0;
#pragma unroll
	for (int j = 0; j < TAGPU_SCAN_ITEMS; ++j) { v[j] = i0 + j < n ? in[i0 + j] : 0u; mine += v[j]; }
	unsigned long long incl = mine;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
		if (lane >= (uint32_t)d) incl += t;
	}
	if (lane == 31) s_w[warp] = incl;
	__syncthreads();
	if (warp == 0) {
		unsigned long long x = s_w[lane], y = x;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const unsigned long long t = __shfl_up_sync(0xffffffffu, y, d);
			if (lane >= (uint32_t)d) y += t;
		}
		s_w[lane] = y - x;
		if (lane == 31) tot[blockIdx.x] = y;
	}
	__syncthreads();
	unsigned long long acc = s_w[warp] + incl - mine;
#pragma unroll
	for (int j = 0; j < TAGPU_SCAN_ITEMS; ++j) {
		if (i0 + j < n) out[i0 + j] = acc;
		acc += v[j];
	}
}

__global__ void __launch_bounds__(1024) k_scan_b(unsigned long long *__restrict__ tot, uint64_t nb, unsigned long long *__restrict__ total)
{
	__shared__ unsigned long long s_part[1024];
	const uint64_t per = (nb + 1023) / 1024, lo = min((uint64_t)threadIdx.x * per, nb), hi = min(lo + per, nb);
	unsigned long long sum = 0;
	for (uint64_t i = lo; i < hi; ++i) sum += tot[i];
	s_part[threadIdx.x] = sum;
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned long long acc = 0;
		for (int t = 0; t < 1024; ++t) { const unsigned long long v = s_part[t]; s_part[t] = acc; acc += v; }
		*total = acc;
	}
	__syncthreads();
	unsigned long long acc = s_part[threadIdx.x];
	for (uint64_t i = lo; i < hi; ++i) { const unsigned long long v = tot[i]; tot[i] = acc; acc += v; }
}

__global__ void __launch_bounds__(1024) k_scan_c(unsigned long long *__restrict__ out, uint64_t n, const unsigned long long *__restrict__ tot)
{
	const uint64_t i0 = ((uint64_t)blockIdx.x * 1024 + threadIdx.x) * TAGPU_SCAN_ITEMS;
	const unsigned long long add = tot[blockIdx.x];
#pragma unroll
	for (int j = 0; j < TAGPU_SCAN_ITEMS; ++j)
		if (i0 + j < n) out[i0 + j] += add;
}
