/*
 * rh_anchor_sort.cuh — anchor sort (reference src/rmap.cpp:121: radix_sort_128x on anchor.x).
 *
 * The reference's klib radix sort is unstable and its tie order is observable,
 * but any two correct sorts agree on every element whose key is unique.  So:
 *
 *   k_sort_smem      one CTA (1024 threads) per chunk: stable LSD byte-radix sort with the packed 32-bit keys
 *                    stationary in shared memory and 16-bit source indices moving; then a scan for adjacent
 *                    equal keys.  Chunks without ties (the large majority) are finished here: A[j] = B[idx[j]].
 *   k_sort_block     the same sort through global memory, for chunks or keys too large for k_sort_smem.
 *   k_sort_ties      one CTA per chunk WITH ties: cta_klib_replay replays klib's in-place MSD pass exactly, but
 *                    only along the sub-arrays that contain a tie group; the result patches the index order
 *                    inside the tie-containing terminal bins, then A[j] = B[idx[j]].
 *   cta_klib_replay  the replay engine (also used by k_chain_finish for full exact sorts): per level, digits of
 *                    the pending sub-arrays, histograms, then either a closed form (two occupied bins) or the
 *                    order-dependent displacement walk, one lane per independent sub-array.
 *
 * Slot region usage (slot_mem): B = anchors as expanded (input), A = sorted output, Z/W = keys ping-pong,
 * f/p = source index ping-pong, v = destinations, t = tied flags (n bytes) + digit bytes (n bytes),
 * U = stable order, U2 = closed-form scratch, regs = pending sub-array lists.
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
	uint32_t *err;
	uint32_t posbits, ridbits; /* bits that hold any target position / target id of this index (fixed key packing) */
};

__device__ __forceinline__ uint32_t lanemask_lt() { uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

/* lanes of `act` holding the same nbits-bit digit as the caller.  Built from one ballot per digit bit: the
 * MATCH.ANY instruction issues far too slowly to sit in the inner loop of a radix pass (measured: the passes
 * ran at ~1 match per ~60 cycles per SM).  Every lane of `act` must call. */
__device__ __forceinline__ uint32_t digit_peers(uint32_t act, uint32_t d, int nbits = 8)
{
	uint32_t peers = act;
#pragma unroll
	for (int b = 0; b < 8; ++b) {
		if (b < nbits) {
			const uint32_t bit = (d >> b) & 1u;
			const uint32_t m = __ballot_sync(act, bit);
			peers &= bit ? m : ~m;
		}
	}
	return peers;
}

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
 * s_base[] must hold the exclusive bucket starts on entry.
 * A round takes SORT_SUB x SORT_THREADS keys: warp w owns SORT_SUB consecutive 32-key sub-tiles of the round, ranks them one
 * after the other against its private counters (keys and ranks stay in registers), then the CTA turns the counters into bases —
 * four barriers per 1024 keys, and four independent loads in flight per thread. */
#define SORT_SUB 4
template <class KeyT>
__device__ __forceinline__ void sort_scatter_pass(const KeyT *__restrict__ kin, const uint32_t *__restrict__ iin, KeyT *__restrict__ kout, uint32_t *__restrict__ iout,
                                                  uint32_t n, uint32_t shift, uint32_t *s_base, uint32_t (*s_wcnt)[256])
{
	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t FULL = 0xffffffffu;
	for (uint32_t t0 = 0; t0 < n; t0 += SORT_THREADS * SORT_SUB) {
		for (uint32_t k = tid; k < SORT_WARPS * 256; k += SORT_THREADS) (&s_wcnt[0][0])[k] = 0;
		__syncthreads();
		KeyT key[SORT_SUB]; uint32_t ix[SORT_SUB], dg[SORT_SUB], rk[SORT_SUB];
		const uint32_t w0 = t0 + warp * (32 * SORT_SUB) + lane;
#pragma unroll
		for (int u = 0; u < SORT_SUB; ++u) { const uint32_t i = w0 + 32 * u; key[u] = 0; ix[u] = 0; if (i < n) { key[u] = kin[i]; ix[u] = iin[i]; } }
#pragma unroll
		for (int u = 0; u < SORT_SUB; ++u) {
			const bool ok = w0 + 32 * u < n;
			const uint32_t act = __ballot_sync(FULL, ok);
			dg[u] = (uint32_t)(key[u] >> shift) & 255; rk[u] = 0;
			if (ok) {
				const uint32_t peers = digit_peers(act, dg[u]);
				const uint32_t off = s_wcnt[warp][dg[u]];
				rk[u] = off + __popc(peers & lanemask_lt());
				__syncwarp(act);
				if ((peers & lanemask_lt()) == 0) s_wcnt[warp][dg[u]] = off + __popc(peers);
			}
			__syncwarp();
		}
		__syncthreads();
		{ /* thread d: prefix over warps for digit d, advance the bucket base */
			uint32_t run = s_base[tid];
#pragma unroll
			for (int w = 0; w < SORT_WARPS; ++w) { const uint32_t c = s_wcnt[w][tid]; s_wcnt[w][tid] = run; run += c; }
			s_base[tid] = run;
		}
		__syncthreads();
#pragma unroll
		for (int u = 0; u < SORT_SUB; ++u)
			if (w0 + 32 * u < n) { const uint32_t dst = s_wcnt[warp][dg[u]] + rk[u]; kout[dst] = key[u]; iout[dst] = ix[u]; }
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

__global__ void __launch_bounds__(SORT_THREADS) k_sort_block(sort_args_t A, uint32_t smem_cap)
{
	__shared__ uint32_t s_hist[8][256];
	__shared__ uint32_t s_base[256];
	__shared__ uint32_t s_wcnt[SORT_WARPS][256];
	__shared__ uint32_t s_tmp[SORT_WARPS];
	__shared__ unsigned long long s_diff;
	__shared__ uint32_t s_ties;

	slot_t *S = &A.slots[blockIdx.x];
	const uint32_t n = S->n_anchors;
	if (S->gated || n == 0 || n <= smem_cap) return; /* n <= smem_cap: k_sort_smem took this chunk */
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

/* =============================================================================================
 * k_sort_smem — the same stable sort for chunks whose packed key fits 32 bits and whose n fits shared memory:
 * keys K[n] stay put in shared memory, only 16-bit source indices move (ping-pong I0/I1), so a chunk costs
 * 8 bytes of shared memory per anchor and no global traffic between the one coalesced read of anchor.x and
 * the final gather.  One CTA of 1024 threads per chunk; warp w owns the w-th contiguous piece of the index
 * array, counts its digits privately (no barrier inside a pass except the three around the prefix), so a
 * pass is: private count -> CTA prefix over (digit, warp) -> private ranked scatter.  Stable because pieces,
 * tiles inside a piece and lanes inside a tile are all taken in order.
 * Key packing uses fixed field widths from the index (bits of the longest target, bits of the target count),
 * so no per-chunk scan for varying bits is needed.
 * ===========================================================================================*/
#define SB_THREADS 1024
#define SB_WARPS (SB_THREADS / 32)
#define SB_FIXED_BYTES (SB_WARPS * 256 * 2 + 256 * 4 + 256 * 4 + 64)
__host__ __device__ inline size_t sort_smem_bytes(uint32_t cap) { return (size_t)cap * 8 + SB_FIXED_BYTES; }

__global__ void __launch_bounds__(SB_THREADS, 1) k_sort_smem(sort_args_t A, uint32_t cap)
{
	extern __shared__ __align__(16) uint8_t s_dyn[];
	uint32_t *K = (uint32_t *)s_dyn;
	uint16_t *I0 = (uint16_t *)(K + cap), *I1 = I0 + cap;
	uint16_t (*wcnt)[256] = (uint16_t (*)[256])(I1 + cap);
	uint32_t *tot = (uint32_t *)(wcnt + SB_WARPS), *dbase = tot + 256;
	uint32_t *s_misc = dbase + 256;

	slot_t *S = &A.slots[blockIdx.x];
	const uint32_t n = S->n_anchors;
	if (S->gated || n == 0 || n > cap) return; /* larger chunks: k_sort_block */
	slot_mem_t M = slot_mem(A.arena, S->a_off, n);
	const anchor_t *__restrict__ in = M.B;
	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t FULL = 0xffffffffu;
	const uint32_t posbits = A.posbits, ridbits = A.ridbits, kbits = posbits + ridbits + 1;
	const uint32_t pmask = posbits >= 32 ? 0xffffffffu : ((1u << posbits) - 1), rmask = ridbits ? ((1u << ridbits) - 1) : 0u;

	if (tid == 0) s_misc[0] = 0;
#pragma unroll 4
	for (uint32_t i = tid; i < n; i += SB_THREADS) {
		const uint64_t x = in[i].x;
		K[i] = ((uint32_t)x & pmask) | ((((uint32_t)(x >> 32)) & rmask) << posbits) | ((uint32_t)(x >> 63) << (posbits + ridbits));
		I0[i] = (uint16_t)i;
	}
	__syncthreads();
	const uint32_t piece = (((n + SB_WARPS - 1) / SB_WARPS) + 31) & ~31u;
	const uint32_t cb = min(warp * piece, n), ce = min(cb + piece, n);
	uint16_t *Ia = I0, *Ib = I1;
	const uint32_t npass = (kbits + 7) / 8;
	for (uint32_t p = 0; p < npass; ++p) {
		const uint32_t shift = 8 * p;
		const int nbits = (int)min(8u, kbits - shift);
#pragma unroll
		for (int q = 0; q < 8; ++q) wcnt[warp][lane * 8 + q] = 0;
		__syncwarp();
		{ /* private digit counts of this warp's piece: shared-memory atomics on the 32-bit word that holds two 16-bit
		   * counters (a piece is far shorter than 65536, so the low half never carries into the high one) */
			uint32_t *cw = (uint32_t *)wcnt[warp];
#pragma unroll 4
			for (uint32_t i = cb + lane; i < ce; i += 32) {
				const uint32_t d = (K[Ia[i]] >> shift) & 255u;
				atomicAdd(&cw[d >> 1], 1u << ((d & 1u) * 16u));
			}
		}
		__syncthreads();
		if (tid < 256) { /* per digit: exclusive prefix over the warps, and the digit total */
			uint32_t run = 0;
#pragma unroll 8
			for (int w = 0; w < SB_WARPS; ++w) { const uint32_t c = wcnt[w][tid]; wcnt[w][tid] = (uint16_t)run; run += c; }
			tot[tid] = run;
		}
		__syncthreads();
		if (warp == 0) { /* exclusive prefix over the digits */
			uint32_t c[8], sum = 0;
#pragma unroll
			for (int q = 0; q < 8; ++q) { c[q] = tot[lane * 8 + q]; sum += c[q]; }
			uint32_t incl = sum;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += u; }
			uint32_t start = incl - sum;
#pragma unroll
			for (int q = 0; q < 8; ++q) { dbase[lane * 8 + q] = start; start += c[q]; }
		}
		__syncthreads();
		for (uint32_t t0 = cb; t0 < ce; t0 += 32) { /* ranked scatter */
			const uint32_t i = t0 + lane;
			const bool ok = i < ce;
			const uint32_t act = __ballot_sync(FULL, ok);
			if (ok) {
				const uint16_t ix = Ia[i];
				const uint32_t d = (K[ix] >> shift) & 255u;
				const uint32_t peers = digit_peers(act, d, nbits);
				const uint32_t off = wcnt[warp][d];
				Ib[dbase[d] + off + __popc(peers & lanemask_lt())] = ix;
				__syncwarp(act);
				if ((peers & lanemask_lt()) == 0) wcnt[warp][d] = (uint16_t)(off + __popc(peers));
			}
			__syncwarp();
		}
		__syncthreads();
		{ uint16_t *t = Ia; Ia = Ib; Ib = t; }
	}
	/* adjacent equal keys */
	uint32_t my = 0;
	for (uint32_t i = tid; i + 1 < n; i += SB_THREADS) my += K[Ia[i]] == K[Ia[i + 1]];
	if (my) atomicAdd(&s_misc[0], my);
	__syncthreads();
	const uint32_t ties = s_misc[0];
	if (ties == 0 || n <= 64) { /* <=64: klib uses a stable insertion sort -> same as the stable order */
		anchor_t *__restrict__ out = M.A;
#pragma unroll 4
		for (uint32_t i = tid; i < n; i += SB_THREADS) out[i] = in[Ia[i]];
		if (tid == 0) S->n_ties = 0;
		return;
	}
	/* chunk with ties: leave the stable order and the tied flags for k_sort_ties */
	uint32_t *sidx = (uint32_t *)M.U;
	uint8_t *tied = (uint8_t *)M.t;
	for (uint32_t i = tid; i < n; i += SB_THREADS) { tied[i] = 0; sidx[i] = Ia[i]; }
	__syncthreads();
	for (uint32_t i = tid; i + 1 < n; i += SB_THREADS)
		if (K[Ia[i]] == K[Ia[i + 1]]) { tied[Ia[i]] = 1; tied[Ia[i + 1]] = 1; }
	if (tid == 0) { S->n_ties = ties; A.tie_list[atomicAdd(A.tie_count, 1u)] = blockIdx.x; }
}

/* =============================================================================================
 * k_sort_ties — exact replay of klib's MSD pass along the sub-arrays that hold a tie group.
 * One CTA (TIE_THREADS) per chunk with ties.
 *
 * Per level (byte of x, from bit 56 down; levels on which the whole chunk agrees are skipped):
 *   - the digit of every element of every pending sub-array goes to shared memory (bytes[]);
 *   - pending sub-arrays are taken TIE_WALKERS at a time; their 256-bin histograms are built by
 *     all threads into one table per sub-array, scanned into packed (end<<16 | head) words;
 *   - a sub-array with ONE occupied bin passes through unchanged; with TWO occupied bins the
 *     displacement walk has a closed form (below) evaluated by all threads; otherwise lane w of
 *     warp 0 runs the order-dependent walk for sub-array w — up to 8 independent walks side by
 *     side, each chasing one shared-memory byte per step;
 *   - (x, source index) pairs are permuted to the order klib would have in memory, and bins that
 *     hold a tie group become the next level's sub-arrays (> 64 elements) or are finished by the
 *     stable rank sort klib's insertion sort is equivalent to (<= 64).
 *
 * Two-bin closed form.  Regions R0=[0,c0) and R1=[c0,len); m_1<..<m_J = positions in R0 holding
 * bin-1 elements, z_1<..<z_J = positions in R1 holding bin-0 elements.  Cycle j drops m_j at the
 * front of what is left of R1 (c0 for j=1, z_{j-1}+1 after), every bin-1 element up to z_j moves
 * one slot right, and z_j falls into the hole m_j.  Elements past z_J and bin-0 elements of R0
 * stay.  (Derived from ksort.h:126-138; checked against the walk in tests.)
 * ===========================================================================================*/
#define TIE_THREADS 256
#define TIE_WARPS (TIE_THREADS / 32)
#define TIE_WALKERS 8
#define TIE_TAB_ROWS (2 + TIE_WARPS > TIE_WALKERS ? 2 + TIE_WARPS : TIE_WALKERS)
#define TIE_CLOSED_MIN 256        /* two-bin sub-arrays at least this long use the closed form */

struct tie_shared_t {
	uint32_t tab[TIE_TAB_ROWS][256];  /* per walker: (end << 16 | head), relative to the sub-array start; rows 2.. double as the per-warp digit counts of fin_radix_pass */
	uint32_t tflag[TIE_WALKERS][8];   /* per walker: bins that hold a tied element (bitmask)              */
	uint32_t wsum[TIE_WARPS];
	uint32_t seg_beg[TIE_WALKERS], seg_len[TIE_WALKERS], seg_kind[TIE_WALKERS], seg_b0[TIE_WALKERS], seg_b1[TIE_WALKERS], seg_c0[TIE_WALKERS];
	uint32_t n_nxt, n_term;
	unsigned long long diff;
	uint32_t big_left;                /* long sub-arrays of the current round still walking */
	uint32_t big_list[TIE_WALKERS];   /* long sub-arrays of the current batch               */
};
/* tables of one long sub-array while its level is replayed (cta_big_*) */
struct big_tab_t {
	uint32_t rs[257];                 /* region starts; rs[256] = len                                      */
	uint32_t qb[257];                 /* first displaced element (index into P/Q) of every region          */
	uint4 st[256];                    /* per region: .x = next displaced element (index into Q), .y = its digit (or BIG_ND_UNKNOWN),
	                                   * .z = end of what the region's ring holds, .w = byte offset of the region's ring */
	uint32_t ja[256];                 /* arrivals a region received before it became the one being closed  */
	uint16_t cid[256];                /* ring slot of every occupied bin                                   */
	uint32_t done, nact, M, RW, stuck;
	uint32_t beg, len;
	const uint8_t *Q;                 /* digits of the displaced elements (global)                         */
	uint32_t *arr;                    /* out: where every displaced element arrives                        */
	uint8_t *ring;                    /* TIE_RING_BYTES of shared memory                                   */
};
enum { TIE_IDENT = 0, TIE_WALK = 1, TIE_TWO = 2, TIE_BIG = 3 };
#define BIG_ND_UNKNOWN 0xffffffffu
#define BIG_SPIN_LIMIT (1u << 24)  /* polls of one ring (tens of milliseconds); a feeder pass takes microseconds */
__device__ __forceinline__ uint4 lds_v4(uint32_t saddr) { uint4 v; asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr)); return v; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr) { uint32_t v; asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr)); return v; }
__device__ __forceinline__ uint32_t lds_u8(uint32_t saddr) { uint32_t v; asm volatile("ld.volatile.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr)); return v; }
__device__ __forceinline__ void sts_v2(uint32_t saddr, uint32_t a, uint32_t b) { asm volatile("st.volatile.shared.v2.u32 [%0], {%1,%2};" :: "r"(saddr), "r"(a), "r"(b) : "memory"); }
__device__ __forceinline__ void sts_u32(uint32_t saddr, uint32_t a) { asm volatile("st.volatile.shared.u32 [%0], %1;" :: "r"(saddr), "r"(a) : "memory"); }
#define TIE_BIG_MIN 4096u         /* sub-arrays at least this long take cta_big_level instead of a lane's walk */
#define TIE_RING_BYTES 8192u      /* shared-memory ring space of cta_big_level: 256 regions x 32 entries       */

/* exclusive rank of `flag` over one tile of TIE_THREADS elements; returns rank within the tile, *total = tile count */
__device__ __forceinline__ uint32_t tie_tile_rank(bool flag, uint32_t *wsum, uint32_t *total)
{
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t m = __ballot_sync(0xffffffffu, flag);
	if (lane == 0) wsum[warp] = __popc(m);
	__syncthreads();
	uint32_t base = 0, tot = 0;
#pragma unroll
	for (uint32_t w = 0; w < TIE_WARPS; ++w) { const uint32_t c = wsum[w]; if (w < warp) base += c; tot += c; }
	__syncthreads();
	*total = tot;
	return base + __popc(m & lanemask_lt());
}

/* modes of cta_klib_replay: a full sort, or only along the sub-arrays that hold an element with an equal key — marked in
 * bit 31 of the key (anchor.x: always 0 there) or in bit 31 of the payload (keys that use all 64 bits) */
enum { KLIB_ALL = 0, KLIB_TIES_KEYFLAG = 1, KLIB_TIES_PAYFLAG = 2 };
#define TIE_FLAG (1ULL << 31)      /* bit 31 of anchor.x is always 0 (31-bit target position): carries "has an equal key" */

/* global work space of one replay, every array sized for the n elements being sorted */
struct klib_ws_t {
	uint64_t *xk, *xk2;      /* keys (| TIE_FLAG in tie mode) in klib's current memory order, ping-pong */
	uint32_t *ord, *ord2;    /* payload, same order, ping-pong                                          */
	uint32_t *dst;           /* destination of every element of the level being replayed                */
	uint32_t *sidx;          /* out: payload of the element at each final position                      */
	uint32_t *zlist, *mlist; /* closed-form scratch, n words each                                       */
	uint2 *term;             /* terminal bins of a level, <= n/2 entries (may alias zlist/mlist)         */
	uint2 *wl0, *wl1;        /* pending sub-arrays (> 64 elements), n/64+2 entries each                 */
	uint8_t *qbytes;         /* cta_big_level: digits of the displaced elements, n + 160 bytes, 16-byte aligned */
	big_tab_t *big;          /* cta_big_*: n_big table sets in shared memory ...                                    */
	uint8_t *ring;           /* ... and n_big x TIE_RING_BYTES of shared-memory rings (free while a long level runs)  */
	uint32_t n_big;          /* long sub-arrays that can walk side by side (<= warps - 1)                          */
	uint32_t *err;           /* device error word: 6 = a walk found its tables inconsistent                          */
};


/* =============================================================================================
 * cta_big_prepare / big_walk / big_feed / cta_big_place — one level of klib's in-place MSD pass (ksort.h:117-138) on ONE long sub-array, by the whole CTA.
 *
 * The displacement walk only ever touches elements that sit outside the region of their own bin ("displaced"
 * elements); everything else keeps its place or moves one slot to the right.  So the level is replayed as
 *   1  ranks: for every position, the number of displaced elements before it; the displaced ones are compacted,
 *      in position order, into P (positions) and Q (their digits) — region b owns the slice [qb[b], qb[b+1])
 *   2  the walk over the displaced elements only: "take the next displaced element of the region I am in; it belongs
 *      to region d; go there" — a rotor walk whose only state is one head counter per region.  Region k, the lowest
 *      unfinished one, is special: an element that returns to k closes the cycle and fills the hole the cycle
 *      started from.  Each step records at which rank the moved element ARRIVES in its own region.
 *      One thread walks; the other three warps keep a 32-entry ring of every region's queue filled from Q, so the
 *      walk's dependent chain is shared-memory loads only.
 *   3  placement, closed form: while region d is not the one being closed, its a-th arrival lands right behind the
 *      (a-1)-th displaced slot of d (the region start for a = 1) and the d-elements between there and the a-th
 *      displaced slot shift one to the right; once d is being closed, an arrival fills the displaced slot itself.
 * Checked against klib for 2..256 bins and skewed histograms (tests: klib tie orders at chunk sizes > 65535).
 * On entry T.tab[w] holds the raw bin counts; on exit dst[beg + i] = new position of element i, relative to beg.
 * Several long sub-arrays of a level walk side by side: one walking thread per warp, one feeder warp for all.
 * ===========================================================================================*/
__device__ void cta_big_prepare(tie_shared_t &T, big_tab_t &B, const uint32_t w, const uint8_t *__restrict__ bytes, const uint32_t beg, const uint32_t len,
                                uint32_t *__restrict__ dst, uint32_t *__restrict__ Pz, uint32_t *__restrict__ arr, uint8_t *__restrict__ Q, uint8_t *ring)
{
	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t FULL = 0xffffffffu;
	const uint8_t *__restrict__ dg = bytes + beg;
	uint32_t *__restrict__ D = dst + beg;
	__syncthreads();
	if (warp == 0) { /* region starts */
		uint32_t c[8], tot = 0;
#pragma unroll
		for (int q = 0; q < 8; ++q) { c[q] = T.tab[w][lane * 8 + q]; tot += c[q]; }
		uint32_t incl = tot;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += u; }
		uint32_t start = incl - tot;
#pragma unroll
		for (int q = 0; q < 8; ++q) { B.rs[lane * 8 + q] = start; start += c[q]; }
		if (lane == 31) B.rs[256] = start;
		/* occupied bins get consecutive ring slots */
		uint32_t nz = 0;
#pragma unroll
		for (int q = 0; q < 8; ++q) nz += c[q] != 0;
		uint32_t nincl = nz;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(FULL, nincl, o); if (lane >= o) nincl += u; }
		uint32_t id = nincl - nz;
#pragma unroll
		for (int q = 0; q < 8; ++q) { B.cid[lane * 8 + q] = (uint16_t)id; id += c[q] != 0; }
		if (lane == 31) B.nact = nincl;
	}
	for (uint32_t b = tid; b < 257; b += TIE_THREADS) B.qb[b] = 0xffffffffu;
	__syncthreads();
	const uint32_t na = B.nact;
	const uint32_t RW = na > 128 ? 32u : na > 64 ? 64u : na > 32 ? 128u : 256u; /* ring entries per region: 32 with all 256 bins occupied */
	/* ---- ranks of the displaced elements; each warp owns a contiguous piece ---- */
	const uint32_t piece = (((len + TIE_WARPS - 1) / TIE_WARPS) + 31) & ~31u;
	const uint32_t cb = min(warp * piece, len), ce = min(cb + piece, len);
	uint32_t hb0 = 0;
	{ /* region of position cb */
		uint32_t lo = 0, hi = 256;
		while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (B.rs[mid + 1] <= cb) lo = mid + 1; else hi = mid; }
		hb0 = lo;
	}
	uint32_t cz = 0;
	{
		uint32_t hb = hb0;
		for (uint32_t i0 = cb; i0 < ce; i0 += 32) {
			const uint32_t i = i0 + lane;
			bool out = false;
			if (i < ce) { while (i >= B.rs[hb + 1]) ++hb; out = dg[i] != hb; }
			cz += __popc(__ballot_sync(FULL, out));
			hb = __shfl_sync(FULL, hb, 0); /* lane 0 holds the smallest position of the next group as well */
		}
	}
	if (lane == 0) T.wsum[warp] = cz;
	__syncthreads();
	uint32_t run = 0, M = 0;
	for (uint32_t w2 = 0; w2 < TIE_WARPS; ++w2) { if (w2 < warp) run += T.wsum[w2]; M += T.wsum[w2]; }
	{
		uint32_t hb = hb0;
		for (uint32_t i0 = cb; i0 < ce; i0 += 32) {
			const uint32_t i = i0 + lane;
			bool out = false; uint32_t d = 0;
			if (i < ce) { while (i >= B.rs[hb + 1]) ++hb; d = dg[i]; out = d != hb; }
			const uint32_t m = __ballot_sync(FULL, out);
			const uint32_t r = run + __popc(m & lanemask_lt());
			if (i < ce) {
				D[i] = r;
				if (out) { Pz[r] = i; Q[r] = (uint8_t)d; }
				if (i == B.rs[hb]) B.qb[hb] = r; /* first position of a non-empty region */
			}
			run += __popc(m);
			hb = __shfl_sync(FULL, hb, 0);
		}
	}
	__syncthreads();
	if (tid == 0) { /* empty regions inherit the next region's start */
		uint32_t nxt = M;
		B.qb[256] = M;
		for (int b = 255; b >= 0; --b) { if (B.qb[b] == 0xffffffffu) B.qb[b] = nxt; else nxt = B.qb[b]; }
		B.done = 0; B.stuck = 0; B.M = M; B.RW = RW; B.beg = beg; B.len = len; B.Q = Q; B.arr = arr; B.ring = ring;
	}
	__syncthreads();
	for (uint32_t b = tid; b < 256; b += TIE_THREADS) {
		B.st[b] = make_uint4(B.qb[b], BIG_ND_UNKNOWN, B.qb[b] & ~15u, (uint32_t)B.cid[b] * RW);
		B.ja[b] = B.qb[b + 1] - B.qb[b];
	}
	__syncthreads();
}

/* the walk of one long sub-array, by ONE thread.  One step = one displaced element taken from the head of the region the
 * walk is in.  The dependent chain of a step is ONE shared-memory load: a region's state word carries the digit of its
 * head element, written there when the element before it was taken. */
__device__ void big_walk(big_tab_t &B, unsigned long long *prof)
{
	const uint32_t M = B.M;
	if (M > 0) {
		uint32_t *__restrict__ arr = B.arr;
		const uint32_t rmask = B.RW - 1; /* a power of two */
		uint32_t st_s = (uint32_t)__cvta_generic_to_shared(B.st), ring_s = (uint32_t)__cvta_generic_to_shared(B.ring);
		asm volatile("" : "+r"(st_s), "+r"(ring_s)); /* opaque: otherwise the shared window base is re-derived (S2UR) on every step */
		uint32_t k = 0;
		while (k < 256 && B.qb[k + 1] == B.qb[k]) ++k;
		B.ja[k] = 0;
		uint32_t spins = 0; /* a ring that never fills means the tables are inconsistent: give up instead of hanging the device */
		uint4 sk = lds_v4(st_s + k * 16);          /* region k's state lives in registers */
		uint32_t ek = B.qb[k + 1];
		for (;;) {
			/* a cycle starts: take k's next displaced element */
			uint32_t d = sk.y;
			if (d == BIG_ND_UNKNOWN) { while (sk.z <= sk.x) { sk.z = lds_u32(st_s + k * 16 + 8); if (++spins > BIG_SPIN_LIMIT) goto stuck; } spins = 0; d = lds_u8(ring_s + sk.w + (sk.x & rmask)); }
			uint32_t ab = sk.x;
			uint4 sc = lds_v4(st_s + d * 16);      /* d != k: a displaced element never belongs to its own region */
			sk.x = ab + 1;
			if (sk.x >= sk.z) sk.z = lds_u32(st_s + k * 16 + 8);
			sk.y = sk.x < sk.z ? lds_u8(ring_s + sk.w + (sk.x & rmask)) : BIG_ND_UNKNOWN;
			sts_u32(st_s + k * 16, sk.x);
			arr[ab] = sc.x | 0x80000000u;          /* lands behind d's previous displaced slot */
			uint32_t cur = d;
			for (;;) { /* follow the cycle until an element of region k turns up */
				d = sc.y;
				if (d == BIG_ND_UNKNOWN) { while (sc.z <= sc.x) { sc.z = lds_u32(st_s + cur * 16 + 8); if (++spins > BIG_SPIN_LIMIT) goto stuck; } spins = 0; d = lds_u8(ring_s + sc.w + (sc.x & rmask)); }
				ab = sc.x;
				const uint32_t nh = ab + 1;
				if (d == k) {
					const uint32_t nd = nh < sc.z ? lds_u8(ring_s + sc.w + (nh & rmask)) : BIG_ND_UNKNOWN;
					sts_v2(st_s + cur * 16, nh, nd);
					arr[ab] = sk.x;                /* fills the hole opened by k's own last pop: displaced slot sk.x - 1 */
					break;
				}
				const uint4 s2 = lds_v4(st_s + d * 16);
				const uint32_t nd = nh < sc.z ? lds_u8(ring_s + sc.w + (nh & rmask)) : BIG_ND_UNKNOWN;
				sts_v2(st_s + cur * 16, nh, nd);
				arr[ab] = s2.x | 0x80000000u;
				cur = d; sc = s2;
			}
			if (sk.x == ek) { /* region k is complete: the next unfinished region takes over */
				do { ++k; } while (k < 256 && lds_u32(st_s + k * 16) == B.qb[k + 1]);
				if (k == 256) break;
				sk = lds_v4(st_s + k * 16);
				ek = B.qb[k + 1];
				B.ja[k] = sk.x - B.qb[k];
			}
		}
		if (prof) { atomicAdd(&prof[27], (unsigned long long)M); atomicAdd(&prof[29], 1ULL); }
	}
	if (false) { stuck: B.stuck = 1; }
	__threadfence_block();
	*(volatile uint32_t *)&B.done = 1;
}

/* ring feeders of up to m concurrent walks: feeder warp `part` of `nparts` serves the region groups q = part, part + nparts, ...
 * (group q = regions 32 q .. 32 q + 31, one per lane) of every walk.  Up to four 16-byte pieces per region and pass, all loads
 * issued before the first store, so that a hot region (the one being closed gives up one element per cycle) is topped up at
 * several times the rate of one piece per pass: the walk is bound by its slowest ring.  Sleeps between idle polls so that its
 * shared-memory traffic stays out of the walks' way. */
__device__ void big_feed(big_tab_t *Bs, const uint32_t m, volatile uint32_t *left, const uint32_t part, const uint32_t nparts)
{
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t FULL = 0xffffffffu;
	for (;;) {
		bool any = false;
		for (uint32_t j = 0; j < m; ++j) {
			big_tab_t &B = Bs[j];
			if (B.M == 0 || *(volatile uint32_t *)&B.done) continue;
			const uint32_t RW = B.RW, rmask = RW - 1;
			const uint32_t st_s = (uint32_t)__cvta_generic_to_shared(B.st);
			const uint8_t *__restrict__ Q = B.Q;
			uint8_t *ring = B.ring;
			for (uint32_t q = part; q < 8; q += nparts) {
				const uint32_t b = lane + 32u * q;
				const uint32_t qe = B.qb[b + 1];
				uint32_t f = 0, cnt = 0;
				if (qe > B.qb[b]) {
					f = lds_u32(st_s + b * 16 + 8);
					const uint32_t lim = lds_u32(st_s + b * 16) + RW; /* entries below the head are free */
					while (cnt < 4 && f + 16 * cnt < qe && f + 16 * cnt + 16 <= lim) ++cnt;
				}
				uint4 v[4];
#pragma unroll
				for (uint32_t u = 0; u < 4; ++u) { v[u] = make_uint4(0, 0, 0, 0); if (u < cnt) v[u] = *(const uint4 *)(Q + f + 16 * u); }
				if (cnt) {
					uint8_t *rb = ring + (uint32_t)B.cid[b] * RW;
#pragma unroll
					for (uint32_t u = 0; u < 4; ++u) if (u < cnt) *(uint4 *)(rb + ((f + 16 * u) & rmask)) = v[u];
					__threadfence_block();
					sts_u32(st_s + b * 16 + 8, f + 16 * cnt);
				}
				any |= cnt != 0;
			}
		}
		if (*left == 0) break;
		if (!__any_sync(FULL, any)) __nanosleep(200);
	}
}

__device__ void cta_big_place(tie_shared_t &T, big_tab_t &B, const uint8_t *__restrict__ bytes, uint32_t *__restrict__ dst, const uint32_t *__restrict__ Pz)
{
	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t FULL = 0xffffffffu;
	const uint32_t beg = B.beg, len = B.len;
	const uint8_t *__restrict__ dg = bytes + beg;
	uint32_t *__restrict__ D = dst + beg;
	const uint32_t *__restrict__ arr = B.arr;
	const uint32_t piece = (((len + TIE_WARPS - 1) / TIE_WARPS) + 31) & ~31u;
	const uint32_t cb = min(warp * piece, len), ce = min(cb + piece, len);
	uint32_t hb = 0;
	{
		uint32_t lo = 0, hi = 256;
		while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (B.rs[mid + 1] <= cb) lo = mid + 1; else hi = mid; }
		hb = lo;
	}
	for (uint32_t i0 = cb; i0 < ce; i0 += 32) {
		const uint32_t i = i0 + lane;
		if (i < ce) {
			while (i >= B.rs[hb + 1]) ++hb;
			const uint32_t d = dg[i], r = D[i];
			uint32_t np;
			if (d == hb) np = i + ((r - B.qb[hb]) < B.ja[hb] ? 1u : 0u);
			else {
				const uint32_t a = arr[r];
				const uint32_t ai = a & 0x7fffffffu; /* index into P/Q */
				if (a & 0x80000000u) np = ai == B.qb[d] ? B.rs[d] : Pz[ai - 1] + 1;
				else np = Pz[ai - 1];
			}
			D[i] = np;
		}
		hb = __shfl_sync(FULL, hb, 0);
	}
	__syncthreads();
}

/* Exact replay of klib's radix_sort on the n > 64 (key, payload) pairs in W.xk/W.ord, by one CTA of
 * TIE_THREADS threads.  ALL = false: only bins that hold a TIE_FLAGged element are followed and only
 * their positions of W.sidx are written (the caller already knows every other position).  ALL = true:
 * a full sort, every position of W.sidx is written. */
template <int MODE>
__device__ void cta_klib_replay(tie_shared_t &T, uint8_t *bytes, klib_ws_t W, const uint32_t n, unsigned long long *prof)
{
	constexpr bool ALL = MODE == KLIB_ALL;
	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t FULL = 0xffffffffu;
	uint64_t *xk = W.xk, *xk2 = W.xk2;
	uint32_t *ord = W.ord, *ord2 = W.ord2;
	uint32_t *__restrict__ dst = W.dst, *__restrict__ sidx = W.sidx, *zlist = W.zlist, *mlist = W.mlist;
	uint2 *wl_cur = W.wl0, *wl_nxt = W.wl1, *term = W.term;
	const uint64_t keymask = MODE == KLIB_TIES_KEYFLAG ? ~TIE_FLAG : ~0ULL;
	const uint32_t plmask = MODE == KLIB_TIES_PAYFLAG ? 0x7fffffffu : 0xffffffffu; /* payload without its flag bit */
	RH_PROF_BEGIN(prof);
	__syncthreads();
	if (tid == 0) { T.diff = 0ULL; T.n_nxt = 0; T.n_term = 0; }
	__syncthreads();
	{
		const uint64_t x0 = xk[0] & keymask;
		unsigned long long diff = 0;
		for (uint32_t i = tid; i < n; i += TIE_THREADS) diff |= (xk[i] & keymask) ^ x0;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) diff |= __shfl_xor_sync(FULL, diff, o);
		if (lane == 0 && diff) atomicOr(&T.diff, diff);
	}
	if (tid == 0) wl_cur[0] = make_uint2(0u, n);
	__syncthreads();
	const unsigned long long dmask = T.diff;
	uint32_t n_cur = 1;
	RH_PROF_MARK(prof, 16, tid == 0);

	for (int shift = 56; shift >= 0 && n_cur > 0; shift -= 8) {
		if (((dmask >> shift) & 255ULL) == 0) continue; /* every sub-array passes through unchanged */
		for (uint32_t s0 = 0; s0 < n_cur; s0 += TIE_WALKERS) {
			const uint32_t nb = min((uint32_t)TIE_WALKERS, n_cur - s0);
			/* ---- digits, histograms, tie flags ---- */
			for (uint32_t k = tid; k < nb * 256; k += TIE_THREADS) (&T.tab[0][0])[k] = 0;
			for (uint32_t k = tid; k < nb * 8; k += TIE_THREADS) (&T.tflag[0][0])[k] = 0;
			if (tid < nb) { const uint2 seg = wl_cur[s0 + tid]; T.seg_beg[tid] = seg.x; T.seg_len[tid] = seg.y; }
			__syncthreads();
			for (uint32_t w = 0; w < nb; ++w) {
				const uint32_t beg = T.seg_beg[w], len = T.seg_len[w];
				const uint64_t *__restrict__ xs = xk + beg;
				const uint32_t *__restrict__ os = ord + beg;
				for (uint32_t t0 = 0; t0 < len; t0 += TIE_THREADS) {
					const uint32_t i = t0 + tid;
					const bool ok = i < len;
					const uint32_t act = __ballot_sync(FULL, ok);
					if (ok) {
						const uint64_t x = xs[i];
						const uint32_t d = (uint32_t)((x & keymask) >> shift) & 255u;
						bytes[beg + i] = (uint8_t)d;
						const uint32_t peers = digit_peers(act, d);
						if ((peers & lanemask_lt()) == 0) atomicAdd(&T.tab[w][d], (uint32_t)__popc(peers));
						if (MODE == KLIB_TIES_KEYFLAG ? (x & TIE_FLAG) != 0 : (MODE == KLIB_TIES_PAYFLAG && (os[i] & 0x80000000u))) atomicOr(&T.tflag[w][d >> 5], 1u << (d & 31));
					}
				}
			}
			__syncthreads();
			/* ---- scan each table into packed (end<<16 | head); classify the sub-array ---- */
			for (uint32_t w = warp; w < nb; w += TIE_WARPS) {
				const uint32_t len = T.seg_len[w];
				uint32_t c[8], tot = 0, nz = 0;
#pragma unroll
				for (int q = 0; q < 8; ++q) { c[q] = T.tab[w][lane * 8 + q]; tot += c[q]; nz += c[q] != 0; }
				uint32_t incl = tot;
#pragma unroll
				for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += u; }
				uint32_t nzt = nz;
#pragma unroll
				for (int o = 16; o > 0; o >>= 1) nzt += __shfl_xor_sync(FULL, nzt, o);
				/* first and second occupied bins */
				uint32_t first = 0xffffffffu, second = 0xffffffffu, first_cnt = 0;
#pragma unroll
				for (int q = 7; q >= 0; --q) if (c[q]) { second = first; first = lane * 8 + q; first_cnt = c[q]; }
				const uint32_t m1 = __ballot_sync(FULL, first != 0xffffffffu);
				const int l1 = __ffs(m1) - 1;
				const uint32_t f_first = __shfl_sync(FULL, first, l1), f_second = __shfl_sync(FULL, second, l1);
				const uint32_t m2 = m1 & ~(1u << l1);
				const uint32_t g_first = __shfl_sync(FULL, first, m2 ? __ffs(m2) - 1 : 0);
				const uint32_t b0 = f_first, b1 = f_second != 0xffffffffu ? f_second : g_first;
				const uint32_t c0 = __shfl_sync(FULL, first_cnt, l1);
				const uint32_t kind = nzt <= 1 ? TIE_IDENT : ((nzt == 2 && len >= TIE_CLOSED_MIN) ? TIE_TWO : (len >= TIE_BIG_MIN ? TIE_BIG : TIE_WALK));
				uint32_t start = incl - tot;
				if (kind != TIE_BIG) {
#pragma unroll
					for (int q = 0; q < 8; ++q) { T.tab[w][lane * 8 + q] = ((start + c[q]) << 16) | start; start += c[q]; }
				}
				if (lane == 0) { T.seg_kind[w] = kind; T.seg_b0[w] = b0; T.seg_b1[w] = b1; T.seg_c0[w] = c0; }
			}
			__syncthreads();
			RH_PROF_MARK(prof, 17, tid == 0);
			/* ---- the order-dependent walks: lane w of warp 0 takes sub-array w.  One iteration = one element
			 *      read (ksort.h:126-138 as a walk over region FIFOs): the element at rp belongs to bin d, goes to
			 *      the head of region d, and the element found there is read next — unless d is the region being
			 *      completed (k), whose hole it fills and whose next slot is read.  Branch free so that walkers with
			 *      different histories stay converged. ---- */
			if (warp == 0) {
				const bool walker = lane < nb && T.seg_kind[lane < nb ? lane : 0] == TIE_WALK;
				uint32_t *tab = T.tab[lane < nb ? lane : 0];
				const uint8_t *bs = bytes + T.seg_beg[lane < nb ? lane : 0];
				/* digits in shared memory are read with ld.shared: through the generic pointer the same load goes
				 * down the global path and costs several times the latency, and it sits on the walk's critical chain */
				const bool bs_shared = __isShared(bytes);
				const uint32_t bs_saddr = bs_shared ? (uint32_t)__cvta_generic_to_shared(bs) : 0u;
				uint32_t *D = dst + T.seg_beg[lane < nb ? lane : 0];
				uint32_t k = 0, rp = 0;
				bool live = walker;
				if (live) { for (;; ++k) { const uint32_t wk = tab[k]; if ((wk & 0xffffu) != (wk >> 16)) { rp = wk & 0xffffu; break; } } }
				/* a single warp issues roughly one dependent instruction every 7-10 cycles, so the walk's speed is its
				 * instruction count per element: no vote, no counters, one load per table */
				if (live) {
					if (bs_shared) {
						for (;;) {
							uint32_t d;
							asm volatile("ld.shared.u8 %0, [%1];" : "=r"(d) : "r"(bs_saddr + rp));
							const uint32_t wd = tab[d];
							const uint32_t h = wd & 0xffffu;
							D[rp] = h;
							tab[d] = wd + 1;
							if (d != k) { rp = h; continue; }
							rp = h + 1;
							if (rp != (wd >> 16)) continue;
							do { ++k; if (k < 256) { const uint32_t wk = tab[k]; rp = wk & 0xffffu; if (rp != (wk >> 16)) break; } } while (k < 256);
							if (k >= 256) break;
						}
					} else {
						const uint32_t beg = T.seg_beg[lane < nb ? lane : 0]; /* uniform bases + one 32-bit index: fewer address instructions */
						const uint8_t *__restrict__ gb = bytes; uint32_t *__restrict__ gd = dst;
						uint32_t ip = beg + rp;
						for (;;) {
							const uint32_t d = gb[ip];
							const uint32_t wd = tab[d];
							const uint32_t h = wd & 0xffffu;
							gd[ip] = h;
							tab[d] = wd + 1;
							if (d != k) { ip = beg + h; continue; }
							ip = beg + h + 1;
							if (h + 1 != (wd >> 16)) continue;
							uint32_t nh = 0;
							do { ++k; if (k < 256) { const uint32_t wk = tab[k]; nh = wk & 0xffffu; if (nh != (wk >> 16)) break; } } while (k < 256);
							if (k >= 256) break;
							ip = beg + nh;
						}
					}
				}
			}
			__syncthreads();
			RH_PROF_MARK(prof, 18, tid == 0);
			/* ---- long sub-arrays: compacted displaced elements, rotor walks from shared-memory rings (up to W.n_big side by
			 *      side: warp 1 feeds the rings, one thread of each other warp walks), closed-form placement ---- */
			{
				if (tid == 0) { uint32_t nbg = 0; for (uint32_t w = 0; w < nb; ++w) if (T.seg_kind[w] == TIE_BIG) T.big_list[nbg++] = w; T.big_left = 0; if (nbg < TIE_WALKERS) T.big_list[nbg] = 0xffffffffu; }
				__syncthreads();
				uint32_t nbg = 0;
				while (nbg < nb && nbg < TIE_WALKERS && T.big_list[nbg] != 0xffffffffu) ++nbg;
				const uint32_t side = min(W.n_big, (uint32_t)TIE_WARPS - 1u);
				for (uint32_t r0 = 0; r0 < nbg; r0 += side) {
					const uint32_t m = min(side, nbg - r0);
					RH_PROF_BEGIN(prof);
					for (uint32_t j = 0; j < m; ++j) {
						const uint32_t w = T.big_list[r0 + j], beg = T.seg_beg[w];
						/* the displaced digits of a sub-array go to a 16-byte aligned slice; sub-arrays that walk side by side get
						 * 16 more bytes per sub-array below them, or a nearly all-displaced one would run into its neighbour's slice */
						uint32_t below = 0;
						for (uint32_t j2 = 0; j2 < m; ++j2) below += T.seg_beg[T.big_list[r0 + j2]] < beg;
						cta_big_prepare(T, W.big[j], w, bytes, beg, T.seg_len[w], dst, zlist + beg, mlist + beg, W.qbytes + ((beg + 15u) & ~15u) + 16u * below, W.ring + j * TIE_RING_BYTES);
					}
					if (tid == 0) T.big_left = m;
					__syncthreads();
					RH_PROF_MARK(prof, 24, tid == 0);
					{ /* walk j runs on lane 0 of warp 0 (j = 0) or warp j + 1; warp 1 and every warp without a walk feed the rings */
						const uint32_t j = warp == 0 ? 0u : warp - 1u;
						const bool walks = warp != 1 && j < m;
						if (!walks) big_feed(W.big, m, (volatile uint32_t *)&T.big_left, warp == 1 ? 0u : warp - m, (uint32_t)TIE_WARPS - m);
						else if (lane == 0) { big_walk(W.big[j], prof); __threadfence_block(); atomicSub(&T.big_left, 1u); }
					}
					__syncthreads();
					if (W.err && tid < m && W.big[tid].stuck) atomicExch(W.err, 6u);
					RH_PROF_MARK(prof, 25, tid == 0);
					for (uint32_t j = 0; j < m; ++j) cta_big_place(T, W.big[j], bytes, dst, zlist + W.big[j].beg);
					RH_PROF_MARK(prof, 26, tid == 0);
				}
			}
			/* ---- two occupied bins: closed form, all threads ---- */
			for (uint32_t w = 0; w < nb; ++w) {
				if (T.seg_kind[w] != TIE_TWO) continue;
				const uint32_t beg = T.seg_beg[w], len = T.seg_len[w], b0 = T.seg_b0[w];
				const uint32_t c0 = T.seg_c0[w]; /* end of region 0 (its start is 0) */
				uint32_t run = 0;
				for (uint32_t t0 = c0; t0 < len; t0 += TIE_THREADS) { /* z_j: bin-0 elements inside region 1 */
					const uint32_t q = t0 + tid;
					const bool f = q < len && bytes[beg + q] == b0;
					uint32_t tot; const uint32_t r = tie_tile_rank(f, T.wsum, &tot);
					if (f) zlist[run + r] = q;
					run += tot;
				}
				const uint32_t J = run;
				__syncthreads();
				run = 0;
				for (uint32_t t0 = 0; t0 < c0; t0 += TIE_THREADS) { /* m_j: bin-1 elements inside region 0 */
					const uint32_t q = t0 + tid;
					const bool in_r = q < c0;
					const bool f = in_r && bytes[beg + q] != b0;
					uint32_t tot; const uint32_t r = tie_tile_rank(f, T.wsum, &tot);
					if (f) { const uint32_t j = run + r; mlist[j] = q; dst[beg + q] = j ? zlist[j - 1] + 1 : c0; }
					else if (in_r) dst[beg + q] = q;
					run += tot;
				}
				__syncthreads();
				const uint32_t zlast = J ? zlist[J - 1] : 0u;
				run = 0;
				for (uint32_t t0 = c0; t0 < len; t0 += TIE_THREADS) {
					const uint32_t q = t0 + tid;
					const bool in_r = q < len;
					const bool f = in_r && bytes[beg + q] == b0;
					uint32_t tot; const uint32_t r = tie_tile_rank(f, T.wsum, &tot);
					if (f) dst[beg + q] = mlist[run + r];
					else if (in_r) dst[beg + q] = (J && q < zlast) ? q + 1 : q;
					run += tot;
				}
				__syncthreads();
			}
			RH_PROF_MARK(prof, 19, tid == 0);
			/* ---- move (x, idx) pairs into klib's memory order, into the other buffer pair ---- */
			for (uint32_t w = 0; w < nb; ++w) {
				const uint32_t beg = T.seg_beg[w], len = T.seg_len[w];
				const bool ident = T.seg_kind[w] == TIE_IDENT;
				const uint64_t *__restrict__ xs = xk + beg; const uint32_t *__restrict__ os = ord + beg; const uint32_t *__restrict__ ds = dst + beg;
				uint64_t *__restrict__ xd = xk2 + beg; uint32_t *__restrict__ od = ord2 + beg;
				uint32_t i = tid;
				for (; i + 3 * TIE_THREADS < len; i += 4 * TIE_THREADS) {
					uint64_t xv[4]; uint32_t ov[4], jv[4];
#pragma unroll
					for (int u = 0; u < 4; ++u) { xv[u] = xs[i + u * TIE_THREADS]; ov[u] = os[i + u * TIE_THREADS]; jv[u] = ident ? i + u * TIE_THREADS : ds[i + u * TIE_THREADS]; }
#pragma unroll
					for (int u = 0; u < 4; ++u) { xd[jv[u]] = xv[u]; od[jv[u]] = ov[u]; }
				}
				for (; i < len; i += TIE_THREADS) { const uint32_t j = ident ? i : ds[i]; xd[j] = xs[i]; od[j] = os[i]; }
			}
			__syncthreads();
			/* ---- children (their data now lives in xk2/ord2) ---- */
			for (uint32_t w = 0; w < nb; ++w) {
				const uint32_t beg = T.seg_beg[w], len = T.seg_len[w], kind = T.seg_kind[w];
				if (kind == TIE_IDENT) {
					if (shift > 0) { if (tid == 0) wl_nxt[atomicAdd(&T.n_nxt, 1u)] = make_uint2(beg, len); }
					else for (uint32_t i = tid; i < len; i += TIE_THREADS) sidx[beg + i] = ord2[beg + i] & plmask;
					continue;
				}
				for (uint32_t b = tid; b < 256; b += TIE_THREADS) {
					if (!ALL && !((T.tflag[w][b >> 5] >> (b & 31)) & 1u)) continue;
					uint32_t start, end;
					if (kind == TIE_TWO) { if (b == T.seg_b0[w]) { start = 0; end = T.seg_c0[w]; } else if (b == T.seg_b1[w]) { start = T.seg_c0[w]; end = len; } else continue; }
					else if (kind == TIE_BIG) { start = 0; for (uint32_t q = 0; q < b; ++q) start += T.tab[w][q]; end = start + T.tab[w][b]; }
					else { start = b ? (T.tab[w][b - 1] >> 16) : 0u; end = T.tab[w][b] >> 16; }
					const uint32_t c = end - start;
					if (ALL && c == 1) sidx[beg + start] = ord2[beg + start];
					if (c < 2) continue;
					if (shift > 0 && c > 64) wl_nxt[atomicAdd(&T.n_nxt, 1u)] = make_uint2(beg + start, c);
					else term[atomicAdd(&T.n_term, 1u)] = make_uint2(beg + start, c);
				}
			}
			__syncthreads();
			/* terminal bins: <= 64 elements finish with klib's stable insertion sort on the full key (a rank
			 * sort here); on the last byte the bin keeps the order the walk left it in */
			{
				const uint32_t nt = T.n_term;
				for (uint32_t q = warp; q < nt; q += TIE_WARPS) {
					const uint2 tb = term[q];
					const uint32_t ts = tb.x, tc = tb.y;
					if (shift > 0) {
						uint32_t o0 = 0, o1 = 0; uint64_t k0 = 0, k1 = 0;
						if (lane < tc) { o0 = ord2[ts + lane]; k0 = xk2[ts + lane] & keymask; }
						if (32 + lane < tc) { o1 = ord2[ts + 32 + lane]; k1 = xk2[ts + 32 + lane] & keymask; }
						uint32_t r0, r1;
						warp_rank64(k0, k1, tc, lane, &r0, &r1);
						if (lane < tc) sidx[ts + r0] = o0 & plmask;
						if (32 + lane < tc) sidx[ts + r1] = o1 & plmask;
					} else {
						for (uint32_t i = lane; i < tc; i += 32) sidx[ts + i] = ord2[ts + i] & plmask;
					}
				}
				__syncthreads();
				if (tid == 0) T.n_term = 0;
				__syncthreads();
			}
			RH_PROF_MARK(prof, 20, tid == 0);
		}
		/* every pending sub-array of the next level was written to the other buffer pair */
		{ uint2 *t = wl_cur; wl_cur = wl_nxt; wl_nxt = t; }
		{ uint64_t *t = xk; xk = xk2; xk2 = t; }
		{ uint32_t *t = ord; ord = ord2; ord2 = t; }
		n_cur = T.n_nxt;
		__syncthreads();
		if (tid == 0) T.n_nxt = 0;
		__syncthreads();
	}
	/* sub-arrays that ran out of varying bytes: fully equal keys keep their current order */
	for (uint32_t s = 0; s < n_cur; ++s) {
		const uint2 seg = wl_cur[s];
		for (uint32_t i = tid; i < seg.y; i += TIE_THREADS) sidx[seg.x + i] = ord[seg.x + i] & plmask;
	}
	__syncthreads();
}

#define TIE_SIDE_WALKS 7           /* long sub-arrays of a tie chunk that walk side by side (warps - 1) */
#define TIE_LARGE_N 131072u        /* chunks from this size on get TIE_SIDE_WALKS table sets (one CTA per SM); smaller ones one set and full occupancy */
__host__ __device__ inline size_t tie_smem_bytes(uint32_t smem_cap, uint32_t n_big) { return ((sizeof(tie_shared_t) + 15) & ~(size_t)15) + n_big * (sizeof(big_tab_t) + TIE_RING_BYTES) + smem_cap; }
__global__ void __launch_bounds__(TIE_THREADS, 4) k_sort_ties(sort_args_t A, uint32_t smem_cap, uint32_t n_lo, uint32_t n_hi, uint32_t n_big)
{
	extern __shared__ __align__(16) uint8_t s_dyn[];
	tie_shared_t &T = *(tie_shared_t *)s_dyn;
	big_tab_t *s_big = (big_tab_t *)(s_dyn + ((sizeof(tie_shared_t) + 15) & ~(size_t)15));
	uint8_t *s_ring = (uint8_t *)(s_big + n_big);
	uint8_t *s_bytes = s_ring + n_big * TIE_RING_BYTES;
	const uint32_t tid = threadIdx.x;
	if (blockIdx.x >= *A.tie_count) return;
	slot_t *S = &A.slots[A.tie_list[blockIdx.x]];
	const uint32_t n = S->n_anchors;
	if (n < n_lo || n >= n_hi) return;   /* the launch for the other size class takes this chunk */
	slot_mem_t M = slot_mem(A.arena, S->a_off, n);
	const anchor_t *__restrict__ in = M.B;
	const uint8_t *__restrict__ tied = (const uint8_t *)M.t;
	klib_ws_t W;
	W.xk = (uint64_t *)M.Z; W.xk2 = (uint64_t *)M.W;
	W.ord = (uint32_t *)M.f; W.ord2 = (uint32_t *)M.p; W.dst = (uint32_t *)M.v;
	W.sidx = (uint32_t *)M.U;                       /* stable order from k_sort_block: right wherever keys are unique */
	W.zlist = (uint32_t *)M.U2; W.mlist = W.zlist + n;
	W.term = (uint2 *)((uint64_t *)M.W + n);        /* M.W is 16n bytes: the upper half */
	W.wl0 = (uint2 *)M.regs; W.wl1 = W.wl0 + (n / 64 + 2);
	uint8_t *bytes = n <= smem_cap ? s_bytes : (uint8_t *)M.t + n;
	W.qbytes = (uint8_t *)(((uintptr_t)((uint8_t *)M.t + 2 * (size_t)n) + 15) & ~(uintptr_t)15); /* M.t is 4n bytes: flags, digits, displaced digits */
	W.big = s_big; W.ring = s_ring; W.n_big = n_big; W.err = A.err;
	for (uint32_t i = tid; i < n; i += TIE_THREADS) { W.xk[i] = in[i].x | (tied[i] ? TIE_FLAG : 0ULL); W.ord[i] = i; }
	cta_klib_replay<KLIB_TIES_KEYFLAG>(T, bytes, W, n, A.prof);
	RH_PROF_BEGIN(A.prof);
	const uint32_t *__restrict__ sidx = W.sidx;
	anchor_t *__restrict__ out = M.A;
	{
		uint32_t i = tid;
		for (; i + 3 * TIE_THREADS < n; i += 4 * TIE_THREADS) {
			uint32_t sv[4]; anchor_t av[4];
#pragma unroll
			for (int u = 0; u < 4; ++u) sv[u] = sidx[i + u * TIE_THREADS];
#pragma unroll
			for (int u = 0; u < 4; ++u) av[u] = in[sv[u]];
#pragma unroll
			for (int u = 0; u < 4; ++u) out[i + u * TIE_THREADS] = av[u];
		}
		for (; i < n; i += TIE_THREADS) out[i] = in[sidx[i]];
	}
	RH_PROF_MARK(A.prof, 21, tid == 0);
}

#endif
