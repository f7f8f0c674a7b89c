/*
 * rh_sort.cuh — exact device emulation of klib's radix_sort (reference src/ksort.h:98-151).
 *
 * Why exact: the reference's sort is an UNSTABLE in-place MSD byte sort; the order it leaves
 * equal keys in is visible downstream (which of two equal-score chain ends is backtracked
 * first, which of two equal-position anchors wins a DP tie — SURVEY.md H1).  Bit-identical
 * chains therefore need the same permutation, not just a sorted array.
 *
 * Shape of the algorithm (restated, not copied): every byte level counts a 256-bin
 * histogram, lays the bins out as consecutive regions and then places elements by chasing
 * displacement cycles: take the first unplaced element of the lowest unfinished region, drop
 * it at the front of the region it belongs to, pick up the element that was there, repeat
 * until something that belongs to the starting region turns up.  Regions of > 64 elements
 * recurse on the next byte, smaller ones are finished by a stable insertion sort on the
 * full key.
 *
 * This version is sequential per calling thread (one thread sorts one segment); the
 * pending-segment stack lives in caller-provided global memory.
 */
#ifndef RH_SORT_CUH
#define RH_SORT_CUH

#include "rh_dev.cuh"

struct sort_seg_t { uint32_t beg, len, shift; };

template <class T, class KeyF>
__device__ __forceinline__ void seq_insertion_sort(T *a, uint32_t n, KeyF key)
{
	for (uint32_t i = 1; i < n; ++i) {
		T cur = a[i];
		uint64_t kc = key(cur);
		if (kc < key(a[i - 1])) {
			uint32_t j = i;
			while (j > 0 && kc < key(a[j - 1])) { a[j] = a[j - 1]; --j; }
			a[j] = cur;
		}
	}
}

template <class T, class KeyF>
__device__ void seq_klib_sort(T *a, uint32_t n, KeyF key, sort_seg_t *stack)
{
	if (n <= 64) { seq_insertion_sort(a, n, key); return; }
	uint32_t cnt[256], head[256];
	/* byte positions on which every key agrees are identity passes for every sub-array */
	uint64_t diff = 0;
	{
		const uint64_t k0 = key(a[0]);
#pragma unroll 1
		for (uint32_t i = 1; i < n; ++i) diff |= key(a[i]) ^ k0;
	}
	int sp = 0;
	stack[sp++] = sort_seg_t{0u, n, 56u};
	while (sp > 0) {
		const sort_seg_t s = stack[--sp];
		T *seg = a + s.beg;
		const uint32_t len = s.len, shift = s.shift;
		if (((diff >> shift) & 255ULL) == 0) {
			if (shift > 0) stack[sp++] = sort_seg_t{s.beg, len, shift > 8 ? shift - 8 : 0};
			continue;
		}
#pragma unroll 1
		for (int b = 0; b < 256; ++b) cnt[b] = 0;
#pragma unroll 1
		for (uint32_t i = 0; i < len; ++i) ++cnt[(key(seg[i]) >> shift) & 255];
		const uint32_t next = shift > 8 ? shift - 8 : 0;
		/* all keys share this byte: the placement pass moves nothing */
		const uint32_t b0 = (uint32_t)(key(seg[0]) >> shift) & 255;
		if (cnt[b0] == len) {
			if (shift > 0) stack[sp++] = sort_seg_t{s.beg, len, next};
			continue;
		}
		uint32_t acc = 0;
#pragma unroll 1
		for (int b = 0; b < 256; ++b) { head[b] = acc; acc += cnt[b]; }
		uint32_t region_end = 0;
#pragma unroll 1
		for (uint32_t k = 0; k < 256; ++k) {
			region_end += cnt[k];
			uint32_t hk = head[k];
			while (hk != region_end) {
				T cur = seg[hk];
				uint32_t d = (uint32_t)(key(cur) >> shift) & 255;
				if (d != k) {
					do {
						const uint32_t hd = head[d];
						T displaced = seg[hd];
						seg[hd] = cur;
						head[d] = hd + 1;
						cur = displaced;
						d = (uint32_t)(key(cur) >> shift) & 255;
					} while (d != k);
					seg[hk] = cur;
				}
				++hk;
			}
			head[k] = hk;
		}
		if (shift == 0) continue;
		acc = 0;
#pragma unroll 1
		for (int b = 0; b < 256; ++b) {
			const uint32_t c = cnt[b];
			if (c > 64) stack[sp++] = sort_seg_t{s.beg + acc, c, next};
			else if (c > 1) seq_insertion_sort(seg + acc, c, key);
			acc += c;
		}
	}
}

struct key_of_anchor_x { __device__ __forceinline__ uint64_t operator()(const anchor_t &a) const { return a.x; } };
struct key_of_u64 { __device__ __forceinline__ uint64_t operator()(const uint64_t &a) const { return a; } };

#endif
