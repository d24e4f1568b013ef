// 2-bit packed k-mer keys for sm_100a: one 64-bit word when the mer has <= 32 bases,
// two words (hi:lo) up to 64 bases.  First base is the most significant 2-bit digit, so
// integer order == the reference's km_cmp order (/root/reference/src/kmer.h:102-112,
// SURVEY.md App. A.2) and canonical = min(fwd, rc) is one unsigned compare.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define TAGPU_DI __device__ __forceinline__
#define TAGPU_HDI __host__ __device__ __forceinline__

template <int W> struct Key;
template <> struct Key<1> { unsigned long long lo; };
template <> struct __align__(16) Key<2> { unsigned long long lo, hi; }; // little-endian u128

TAGPU_HDI uint64_t tagpu_mix64(uint64_t x)
{
	x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
	x ^= x >> 27; x *= 0x94d049bb133111ebull;
	x ^= x >> 31;
	return x;
}

// reverse the order of the 32 2-bit digits of a word and complement them
TAGPU_DI uint64_t tagpu_rc64_full(uint64_t x)
{
	x = ~x;
	x = __brevll(x);
	return ((x & 0xaaaaaaaaaaaaaaaaull) >> 1) | ((x & 0x5555555555555555ull) << 1);
}

// the same for the 16 digits of a 32-bit word
TAGPU_DI uint32_t tagpu_rc32_full(uint32_t x)
{
	x = __brev(~x);
	return ((x & 0xaaaaaaaau) >> 1) | ((x & 0x55555555u) << 1);
}

template <int W> struct KeyOps;

template <> struct KeyOps<1> {
	typedef Key<1> K;
	static TAGPU_HDI K make(uint64_t hi, uint64_t lo) { (void)hi; K r; r.lo = lo; return r; }
	static TAGPU_HDI uint64_t hi(const K &a) { (void)a; return 0; }
	static TAGPU_HDI K mask(int len) { K r; r.lo = len >= 32 ? ~0ull : ((1ull << (2 * len)) - 1); return r; }
	static TAGPU_HDI K band(const K &a, const K &b) { K r; r.lo = a.lo & b.lo; return r; }
	static TAGPU_HDI K bnot(const K &a) { K r; r.lo = ~a.lo; return r; }
	static TAGPU_HDI bool eq(const K &a, const K &b) { return a.lo == b.lo; }
	static TAGPU_HDI bool le(const K &a, const K &b) { return a.lo <= b.lo; }
	static TAGPU_HDI bool is_zero(const K &a) { return a.lo == 0; }
	// (x << 2 | c) & m : append a base on the right
	static TAGPU_HDI K push(const K &x, uint32_t c, const K &m) { K r; r.lo = ((x.lo << 2) | c) & m.lo; return r; }
	// (x >> 2) | c << 2(len-1) : prepend a base on the left (rolling reverse complement)
	static TAGPU_HDI K push_front(const K &x, uint32_t c, int len) { K r; r.lo = (x.lo >> 2) | ((uint64_t)c << (2 * (len - 1))); return r; }
	static TAGPU_HDI K shr2(const K &x) { K r; r.lo = x.lo >> 2; return r; }
	static TAGPU_HDI uint32_t last_base(const K &x) { return (uint32_t)x.lo & 3u; }
	static TAGPU_HDI uint32_t first_base(const K &x, int len) { return (uint32_t)(x.lo >> (2 * (len - 1))) & 3u; }
	static TAGPU_HDI uint32_t base_at(const K &x, int len, int i) { return (uint32_t)(x.lo >> (2 * (len - 1 - i))) & 3u; }
	static TAGPU_DI K rc(const K &x, int len) { K r; r.lo = tagpu_rc64_full(x.lo) >> (64 - 2 * len); return r; }
	static TAGPU_HDI uint64_t hash(const K &x) { return tagpu_mix64(x.lo); }
};

template <> struct KeyOps<2> {
	typedef Key<2> K;
	static TAGPU_HDI K make(uint64_t hi, uint64_t lo) { K r; r.hi = hi; r.lo = lo; return r; }
	static TAGPU_HDI uint64_t hi(const K &a) { return a.hi; }
	static TAGPU_HDI K mask(int len)
	{
		K r;
		if (len >= 64) { r.hi = ~0ull; r.lo = ~0ull; }
		else if (len > 32) { r.hi = (1ull << (2 * (len - 32))) - 1; r.lo = ~0ull; }
		else if (len == 32) { r.hi = 0; r.lo = ~0ull; }
		else { r.hi = 0; r.lo = (1ull << (2 * len)) - 1; }
		return r;
	}
	static TAGPU_HDI K band(const K &a, const K &b) { K r; r.hi = a.hi & b.hi; r.lo = a.lo & b.lo; return r; }
	static TAGPU_HDI K bnot(const K &a) { K r; r.hi = ~a.hi; r.lo = ~a.lo; return r; }
	static TAGPU_HDI bool eq(const K &a, const K &b) { return a.hi == b.hi && a.lo == b.lo; }
	static TAGPU_HDI bool le(const K &a, const K &b) { return a.hi < b.hi || (a.hi == b.hi && a.lo <= b.lo); }
	static TAGPU_HDI bool is_zero(const K &a) { return (a.hi | a.lo) == 0; }
	static TAGPU_HDI K push(const K &x, uint32_t c, const K &m)
	{
		K r;
		r.hi = ((x.hi << 2) | (x.lo >> 62)) & m.hi;
		r.lo = ((x.lo << 2) | c) & m.lo;
		return r;
	}
	static TAGPU_HDI K push_front(const K &x, uint32_t c, int len)
	{
		K r;
		r.lo = (x.lo >> 2) | (x.hi << 62);
		r.hi = x.hi >> 2;
		int sh = 2 * (len - 1);
		if (sh >= 64) r.hi |= (uint64_t)c << (sh - 64); else r.lo |= (uint64_t)c << sh;
		return r;
	}
	static TAGPU_HDI K shr2(const K &x) { K r; r.lo = (x.lo >> 2) | (x.hi << 62); r.hi = x.hi >> 2; return r; }
	static TAGPU_HDI uint32_t last_base(const K &x) { return (uint32_t)x.lo & 3u; }
	static TAGPU_HDI uint32_t first_base(const K &x, int len)
	{
		int sh = 2 * (len - 1);
		return (uint32_t)(sh >= 64 ? x.hi >> (sh - 64) : x.lo >> sh) & 3u;
	}
	static TAGPU_HDI uint32_t base_at(const K &x, int len, int i)
	{
		int sh = 2 * (len - 1 - i);
		return (uint32_t)(sh >= 64 ? x.hi >> (sh - 64) : x.lo >> sh) & 3u;
	}
	static TAGPU_DI K rc(const K &x, int len)
	{
		// full 64-base reverse complement, then drop the (64 - len) leading digits
		uint64_t fhi = tagpu_rc64_full(x.lo), flo = tagpu_rc64_full(x.hi);
		int sh = 128 - 2 * len; // 0..126
		K r;
		if (sh == 0) { r.hi = fhi; r.lo = flo; }
		else if (sh < 64) { r.lo = (flo >> sh) | (fhi << (64 - sh)); r.hi = fhi >> sh; }
		else if (sh == 64) { r.lo = fhi; r.hi = 0; }
		else { r.lo = fhi >> (sh - 64); r.hi = 0; }
		return r;
	}
	static TAGPU_HDI uint64_t hash(const K &x) { return tagpu_mix64(x.lo ^ (x.hi * 0x9e3779b97f4a7c15ull + 0x7f4a7c15ull)); }
};
