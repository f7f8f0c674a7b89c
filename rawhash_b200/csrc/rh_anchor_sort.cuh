/*
 * rh_anchor_sort.cuh — anchor sort (reference src/rmap.cpp:121: radix_sort_128x on anchor.x).
 *
 * The reference's klib radix sort is unstable and its tie order is observable,
 * but any two correct sorts agree on every element whose key is unique.  So:
 *
 *   k_sort_block     one CTA per chunk: stable LSD byte-radix sort of (x, source index) pairs,
 *                    skipping byte positions on which all keys of the chunk agree; then a scan
 *                    for adjacent equal keys.  Chunks without ties (the large majority) are
 *                    finished here: A[j] = B[idx[j]].
 *   k_sort_ties      one warp per chunk WITH ties: replays klib's in-place MSD pass exactly, but
 *                    only along the sub-arrays that contain a tie group — for those it needs the
 *                    true element order the reference would have at that recursion level, which
 *                    it tracks as a permutation of source indices.  Histogram/permute are
 *                    warp-parallel; the displacement-cycle walk itself is order dependent and is
 *                    done by lane 0 on a byte array.  The result patches the index order inside
 *                    the tie-containing terminal buckets, then A[j] = B[idx[j]].
 *
 * Slot region usage (slot_mem): B = anchors as expanded (input), A = sorted output,
 * Z/W = (x, idx) ping-pong, f = ord, p = ord scratch, v = dst, t = tied flags (n bytes) +
 * level bytes (n bytes), U/U2 = segment work lists.
 */
#ifndef RH_ANCHOR_SORT_CUH
#define RH_ANCHOR_SORT_CUH

#include "rh_kernels.cuh"

#define SORT_THREADS 256
#define SORT_WARPS (SORT_THREADS / 32)

struct sort_args_t {
	slot_t *slots; uint32_t n_slots;
	uint8_t *arena;
	uint32_t *tie_list;      /* slots (indices into `slots`) whose chunk has equal keys */
	uint32_t *tie_count;
	unsigned long long *prof;
};

__device__ __forceinline__ uint32_t lanemask_lt() { uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

/* Stable sort of a bucket of c <= 64 (key, payload) items held two per lane (item L and item 32+L):
 * returns each item's rank.  Equivalent to klib's stable insertion sort of a small bucket
 * (ksort.h:105-115), but the 64 x 64 comparisons run out of registers via shuffles instead of a
 * chain of dependent global loads.  All lanes must call. */
__device__ __forceinline__ void warp_rank64(uint64_t k0, uint64_t k1, uint32_t c, uint32_t lane, uint32_t *r0, uint32_t *r1)
{
	const uint32_t FULL = 0xffffffffu;
	uint32_t a0 = 0, a1 = 0;
	const uint32_t c0 = c < 32 ? c : 32;
	for (uint32_t j = 0; j < c0; ++j) { /* items 0..31 */
		const uint64_t kj = __shfl_sync(FULL, k0, j);
		a0 += (kj < k0) || (kj == k0 && j < lane);
		a1 += (kj <= k1); /* every item of the first half precedes item 32+lane */
	}
	for (uint32_t j = 32; j < c; ++j) { /* items 32..c-1 */
		const uint64_t kj = __shfl_sync(FULL, k1, j - 32);
		a0 += (kj < k0);
		a1 += (kj < k1) || (kj == k1 && (j - 32) < lane);
	}
	*r0 = a0; *r1 = a1;
}

/* One stable counting-sort pass of (key, idx) arrays on digit (key >> shift) & 255.
 * s_base[] must hold the exclusive bucket starts on entry. */
template <class KeyT>
__device__ __forceinline__ void sort_scatter_pass(const KeyT *__restrict__ kin, const uint32_t *__restrict__ iin, KeyT *__restrict__ kout, uint32_t *__restrict__ iout,
                                                  uint32_t n, uint32_t shift, uint32_t *s_base, uint32_t (*s_wcnt)[256])
{
	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	for (uint32_t t0 = 0; t0 < n; t0 += SORT_THREADS) {
		for (uint32_t k = tid; k < SORT_WARPS * 256; k += SORT_THREADS) (&s_wcnt[0][0])[k] = 0;
		__syncthreads();
		const uint32_t i = t0 + tid;
		const bool ok = i < n;
		KeyT key = 0; uint32_t ix = 0, d = 0, peers = 0;
		if (ok) { key = kin[i]; ix = iin[i]; d = (uint32_t)(key >> shift) & 255; }
		const uint32_t act = __ballot_sync(0xffffffffu, ok);
		if (ok) {
			peers = __match_any_sync(act, d);
			if ((peers & lanemask_lt()) == 0) s_wcnt[warp][d] = __popc(peers);
		}
		__syncthreads();
		{ /* thread d: prefix over warps for digit d, advance the bucket base */
			uint32_t run = s_base[tid];
#pragma unroll
			for (int w = 0; w < SORT_WARPS; ++w) { const uint32_t c = s_wcnt[w][tid]; s_wcnt[w][tid] = run; run += c; }
			s_base[tid] = run;
		}
		__syncthreads();
		if (ok) { const uint32_t dst = s_wcnt[warp][d] + __popc(peers & lanemask_lt()); kout[dst] = key; iout[dst] = ix; }
		__syncthreads();
	}
}

/* exclusive scan of 256 counters held one per thread -> s_base */
__device__ __forceinline__ void sort_scan256(uint32_t v, uint32_t *s_base, uint32_t *s_tmp)
{
	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	uint32_t incl = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
	if (lane == 31) s_tmp[warp] = incl;
	__syncthreads();
	uint32_t woff = 0;
	for (uint32_t w = 0; w < warp; ++w) woff += s_tmp[w];
	s_base[tid] = woff + incl - v;
	__syncthreads();
}

__global__ void __launch_bounds__(SORT_THREADS) k_sort_block(sort_args_t A)
{
	__shared__ uint32_t s_hist[8][256];
	__shared__ uint32_t s_base[256];
	__shared__ uint32_t s_wcnt[SORT_WARPS][256];
	__shared__ uint32_t s_tmp[SORT_WARPS];
	__shared__ unsigned long long s_diff;
	__shared__ uint32_t s_ties;

	slot_t *S = &A.slots[blockIdx.x];
	const uint32_t n = S->n_anchors;
	if (S->gated || n == 0) return;
	slot_mem_t M = slot_mem(A.arena, S->a_off, n);
	const anchor_t *in = M.B;
	const uint32_t tid = threadIdx.x, lane = tid & 31;
	uint32_t *sidx = (uint32_t *)M.U; /* final sorted order (source indices) */

	if (tid == 0) { s_diff = 0ULL; s_ties = 0; }
	for (uint32_t k = tid; k < 8 * 256; k += SORT_THREADS) (&s_hist[0][0])[k] = 0;
	__syncthreads();
	/* which key bits differ anywhere in the chunk */
	const uint64_t x0 = in[0].x;
	unsigned long long diff = 0;
	for (uint32_t i = tid; i < n; i += SORT_THREADS) diff |= in[i].x ^ x0;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) diff |= __shfl_xor_sync(0xffffffffu, diff, o);
	if (lane == 0 && diff) atomicOr(&s_diff, diff);
	__syncthreads();
	const unsigned long long dmask = s_diff;
	/* anchor.x = strand<<63 | rid<<32 | pos (31 bits): pack the varying low bits of each field into one word.
	 * Order and equality are preserved because the dropped high bits are identical across the chunk. */
	const uint32_t posbits = 32 - __clz((uint32_t)(dmask & 0x7fffffffULL));
	const uint32_t ridbits = 32 - __clz((uint32_t)((dmask >> 32) & 0x7fffffffULL));
	const uint32_t sbit = (uint32_t)(dmask >> 63);
	const uint32_t kbits = posbits + ridbits + sbit;
	uint32_t ties_found = 0;

	if (kbits <= 32) {
		uint32_t *k0 = (uint32_t *)M.Z, *i0 = k0 + n, *k1 = i0 + n, *i1 = k1 + n;
		const uint32_t npass = (kbits + 7) / 8;
		const uint32_t pmask = posbits ? (0xffffffffu >> (32 - posbits)) : 0u, rmask = ridbits ? (0xffffffffu >> (32 - ridbits)) : 0u;
		for (uint32_t i = tid; i < n; i += SORT_THREADS) {
			const uint64_t x = in[i].x;
			const uint32_t key = ((uint32_t)x & pmask) | ((((uint32_t)(x >> 32)) & rmask) << posbits) | (sbit ? ((uint32_t)(x >> 63) << (posbits + ridbits)) : 0u);
			k0[i] = key; i0[i] = i;
			for (uint32_t p = 0; p < npass; ++p) atomicAdd(&s_hist[p][(key >> (8 * p)) & 255], 1u);
		}
		__syncthreads();
		uint32_t *kc = k0, *ic = i0, *kn = k1, *inx = i1;
		for (uint32_t p = 0; p < npass; ++p) {
			sort_scan256(s_hist[p][tid], s_base, s_tmp);
			sort_scatter_pass<uint32_t>(kc, ic, kn, inx, n, 8 * p, s_base, s_wcnt);
			uint32_t *t1 = kc; kc = kn; kn = t1; t1 = ic; ic = inx; inx = t1;
		}
		uint32_t my = 0;
		uint8_t *tied = (uint8_t *)M.t;
		for (uint32_t i = tid; i < n; i += SORT_THREADS) { tied[i] = 0; sidx[i] = ic[i]; }
		__syncthreads();
		for (uint32_t i = tid; i + 1 < n; i += SORT_THREADS)
			if (kc[i] == kc[i + 1]) { tied[ic[i]] = 1; tied[ic[i + 1]] = 1; ++my; }
		if (my) atomicAdd(&s_ties, my);
	} else {
		/* wide keys: same passes on 64-bit keys, skipping bytes on which all keys agree */
		uint64_t *k0 = (uint64_t *)M.Z, *k1 = (uint64_t *)M.W;
		uint32_t *i0 = (uint32_t *)M.f, *i1 = (uint32_t *)M.p;
		for (uint32_t i = tid; i < n; i += SORT_THREADS) {
			const uint64_t x = in[i].x;
			k0[i] = x; i0[i] = i;
#pragma unroll
			for (uint32_t p = 0; p < 8; ++p) atomicAdd(&s_hist[p][(uint32_t)(x >> (8 * p)) & 255], 1u);
		}
		__syncthreads();
		uint64_t *kc = k0, *kn = k1; uint32_t *ic = i0, *inx = i1;
		for (uint32_t p = 0; p < 8; ++p) {
			if (((dmask >> (8 * p)) & 255ULL) == 0) continue;
			sort_scan256(s_hist[p][tid], s_base, s_tmp);
			sort_scatter_pass<uint64_t>(kc, ic, kn, inx, n, 8 * p, s_base, s_wcnt);
			uint64_t *t1 = kc; kc = kn; kn = t1; uint32_t *t2 = ic; ic = inx; inx = t2;
		}
		uint32_t my = 0;
		uint8_t *tied = (uint8_t *)M.t;
		for (uint32_t i = tid; i < n; i += SORT_THREADS) { tied[i] = 0; sidx[i] = ic[i]; }
		__syncthreads();
		for (uint32_t i = tid; i + 1 < n; i += SORT_THREADS)
			if (kc[i] == kc[i + 1]) { tied[ic[i]] = 1; tied[ic[i + 1]] = 1; ++my; }
		if (my) atomicAdd(&s_ties, my);
	}
	__syncthreads();
	ties_found = s_ties;
	if (ties_found == 0 || n <= 64) { /* <=64: klib uses a stable insertion sort -> same as the stable order */
		anchor_t *out = M.A;
		for (uint32_t i = tid; i < n; i += SORT_THREADS) out[i] = in[sidx[i]];
		if (tid == 0) S->n_ties = 0;
	} else if (tid == 0) {
		S->n_ties = ties_found;
		A.tie_list[atomicAdd(A.tie_count, 1u)] = blockIdx.x;
	}
}

/* exact replay along tie-containing sub-arrays; one warp per slot */
/* One warp (= one CTA) per tie-containing chunk.  The byte array the walk chases lives in dynamic
 * shared memory when the chunk fits (smem_cap bytes), else in the slot's global scratch. */
__global__ void __launch_bounds__(32) k_sort_ties(sort_args_t A, uint32_t smem_cap)
{
	extern __shared__ __align__(16) uint8_t s_dyn[];
	uint32_t *cnt = (uint32_t *)s_dyn, *head = cnt + 256, *flag = head + 256;
	uint8_t *s_bytes = s_dyn + 3 * 256 * 4;

	const uint32_t lane = threadIdx.x & 31;
	const uint32_t li = blockIdx.x;
	if (li >= *A.tie_count) return;
	slot_t *S = &A.slots[A.tie_list[li]];
	const uint32_t n = S->n_anchors;
	slot_mem_t M = slot_mem(A.arena, S->a_off, n);
	const anchor_t *in = M.B;
	uint32_t *sidx = (uint32_t *)M.U;
	uint32_t *ord = (uint32_t *)M.f, *ord2 = (uint32_t *)M.p, *dst = (uint32_t *)M.v;
	const uint8_t *tied = (const uint8_t *)M.t;
	uint8_t *bytes = n <= smem_cap ? s_bytes : (uint8_t *)M.t + n;
	uint2 *wl_cur = (uint2 *)M.regs, *wl_nxt = wl_cur + (n / 64 + 2);
	const uint32_t FULL = 0xffffffffu;

	RH_PROF_BEGIN(A.prof);
	for (uint32_t i = lane; i < n; i += 32) ord[i] = i;
	uint32_t n_cur = 1, n_nxt = 0;
	if (lane == 0) wl_cur[0] = make_uint2(0u, n);
	__syncwarp();

	for (int shift = 56; shift >= 0 && n_cur > 0; shift -= 8) {
		n_nxt = 0;
		for (uint32_t s = 0; s < n_cur; ++s) {
			const uint2 seg = wl_cur[s];
			const uint32_t beg = seg.x, len = seg.y;
			for (uint32_t b = lane; b < 256; b += 32) { cnt[b] = 0; flag[b] = 0; }
			__syncwarp();
			for (uint32_t i = lane; i < len; i += 32) {
				const uint32_t o = ord[beg + i];
				const uint32_t b = (uint32_t)(in[o].x >> shift) & 255;
				bytes[beg + i] = (uint8_t)b;
				atomicAdd(&cnt[b], 1u);
				if (tied[o]) flag[b] = 1;
			}
			__syncwarp();
			RH_PROF_MARK(A.prof, 16, lane == 0);
			const uint32_t b0 = bytes[beg];
			if (cnt[b0] == len) { /* nothing moves at this level */
				if (shift > 0) { if (lane == 0) wl_nxt[n_nxt] = seg; ++n_nxt; }
				else { for (uint32_t i = lane; i < len; i += 32) sidx[beg + i] = ord[beg + i]; }
				__syncwarp();
				continue;
			}
			/* bucket starts */
			uint32_t run = 0;
			for (uint32_t b = lane; b < 256; b += 32) {
				const uint32_t v = cnt[b];
				uint32_t incl = v;
#pragma unroll
				for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += u; }
				head[b] = run + incl - v;
				run += __shfl_sync(FULL, incl, 31);
			}
			__syncwarp();
			RH_PROF_MARK(A.prof, 17, lane == 0);
			/* the displacement-cycle walk (ksort.h:126-138), on bytes: dst[i] = slot element i lands in */
			if (lane == 0) {
				uint32_t region_end = 0;
				for (uint32_t k = 0; k < 256; ++k) {
					region_end += cnt[k];
					uint32_t hk = head[k];
					while (hk != region_end) {
						uint32_t e = hk; /* element in hand = original occupant of slot e */
						uint32_t d = bytes[beg + e];
						while (d != k) {
							const uint32_t hd = head[d];
							dst[beg + e] = hd;
							head[d] = hd + 1;
							e = hd;
							d = bytes[beg + e];
						}
						dst[beg + e] = hk;
						++hk;
					}
					head[k] = hk;
				}
			}
			__syncwarp();
			RH_PROF_MARK(A.prof, 18, lane == 0);
			for (uint32_t i = lane; i < len; i += 32) ord2[beg + dst[beg + i]] = ord[beg + i];
			__syncwarp();
			for (uint32_t i = lane; i < len; i += 32) ord[beg + i] = ord2[beg + i];
			__syncwarp();
			RH_PROF_MARK(A.prof, 19, lane == 0);
			/* children that contain a tie group */
			uint32_t acc = 0;
			for (uint32_t bb = 0; bb < 256; bb += 32) {
				const uint32_t b = bb + lane;
				const uint32_t c = cnt[b];
				uint32_t incl = c;
#pragma unroll
				for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += u; }
				const uint32_t start = beg + acc + incl - c;
				acc += __shfl_sync(FULL, incl, 31);
				const bool has = flag[b] && c > 1;
				const bool recurse = has && shift > 0 && c > 64;
				const uint32_t rm = __ballot_sync(FULL, recurse);
				if (recurse) wl_nxt[n_nxt + __popc(rm & lanemask_lt())] = make_uint2(start, c);
				n_nxt += __popc(rm);
				/* terminal buckets, one at a time on the whole warp */
				uint32_t tm = __ballot_sync(FULL, has && !recurse);
				while (tm) {
					const int src = __ffs(tm) - 1; tm &= tm - 1;
					const uint32_t ts = __shfl_sync(FULL, start, src), tc = __shfl_sync(FULL, c, src);
					if (shift > 0) { /* <=64 elements: klib finishes with a stable insertion sort on the full key */
						uint32_t o0 = 0, o1 = 0; uint64_t k0 = 0, k1 = 0;
						if (lane < tc) { o0 = ord[ts + lane]; k0 = in[o0].x; }
						if (32 + lane < tc) { o1 = ord[ts + 32 + lane]; k1 = in[o1].x; }
						uint32_t r0, r1;
						warp_rank64(k0, k1, tc, lane, &r0, &r1);
						if (lane < tc) sidx[ts + r0] = o0;
						if (32 + lane < tc) sidx[ts + r1] = o1;
					} else { /* last byte: the bucket keeps the order the walk left it in */
						for (uint32_t i = lane; i < tc; i += 32) sidx[ts + i] = ord[ts + i];
					}
				}
			}
			__syncwarp();
		}
		RH_PROF_MARK(A.prof, 20, lane == 0);
		uint2 *t = wl_cur; wl_cur = wl_nxt; wl_nxt = t;
		n_cur = n_nxt;
		__syncwarp();
	}
	__syncwarp();
	anchor_t *out = M.A;
	for (uint32_t i = lane; i < n; i += 32) out[i] = in[sidx[i]];
	RH_PROF_MARK(A.prof, 21, lane == 0);
}

#endif
