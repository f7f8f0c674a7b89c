/*
 * rh_index_gpu.cu — index construction on the GPU: the reference's
 *     ri_idx_gen / worker_pipeline / worker_post        (src/rindex.c:100-192, 311-363, 900-925)
 *     ri_seq_to_sig                                     (src/rsig.c:13-40)
 *     ri_sketch_reg on both strands                     (src/rsketch.c:143-204)
 *     ri_idx_cal_max_occ / ri_mapopt_update             (src/rindex.c:1018-1053)
 * for an ACGT-only reference and w = 0 (no minimizers); anything else is left to the host builder
 * rh_index_build, which this must equal key for key and position for position.
 *
 * The result STAYS ON THE DEVICE (rh_index_s::dev): a human-size index is ≈40 GB of positions, and the mapper
 * reads it from HBM anyway; the host mirror is downloaded only when a host accessor asks for it.
 *
 * Pipeline (S = one strand of one sequence; events of S are numbered in its own reading direction):
 *   k_idx_qtab      per k-mer: quantised level (dynamic_quantize of the pore level), 4^k bytes
 *   k_idx_filter    thread per 256-event segment of S: rolling k-mer -> pore level -> the diff filter against
 *                   the last KEPT event (rsketch.c:187-189) -> kept bitmap, quantised value per event, count.
 *                   The filter is a sequential recurrence whose only state is the last kept level, so a segment
 *                   first GUESSES its incoming state by running the 64 events before it from "keep the first";
 *   k_idx_verify    compares every guess with the state its predecessor actually left; wrong guesses are
 *                   recomputed from the right state until none is left (a fixed point = the sequential result)
 *   scan            kept counts -> 64-bit prefix (three small kernels)
 *   k_idx_seeds     thread per 64-event word: window of e kept events -> hash64 of the packed quantised values;
 *                   the seed is written at its rank in (sequence, position, strand) order — the order of
 *                   id<<32|pos<<1|strand — obtained from the two strands' kept-prefix counts
 *   sort            ONE stable radix sort by hash (cub::DeviceRadixSort, the library part of this non-hot path):
 *                   positions inside a key stay ascending, which is what worker_post leaves (rindex.c:350)
 *   k_idx_heads*    first element of every key run -> distinct keys + CSR offsets
 *   k_occ_*         mid_occ: exact k-th smallest list length by a two-level 16-bit radix select
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>

#include <cuda_runtime.h>
#include <cub/cub.cuh>

#include "rh_host.h"
#include "rh_dev.cuh"

namespace {

#define IDX_SEG 256u           /* events per filter segment (4 bitmap words) */
#define IDX_WARM 64u           /* events run before a segment to guess its incoming state */
#define IDX_MAX_E 16
#define STATE_NONE 0xffffffffu /* "no event kept yet" (a NaN pattern no pore level has) */

struct strand_desc_t {
	uint64_t base_off;  /* first base of the sequence in the base array                     */
	uint64_t word_off;  /* first bitmap word of this strand (multiple of 4); qv at word_off*64 */
	uint64_t seg_off;   /* first global segment                                            */
	uint64_t seed_base; /* first output slot of this SEQUENCE (both strands interleave)    */
	uint32_t len, n_ev, id, strand;
	uint32_t n_kept, n_valid; /* kept events; seeds = n_kept - e + 1 (or 0)               */
};

__device__ __forceinline__ int strand_base(const uint8_t *__restrict__ s, uint32_t len, uint32_t strand, uint32_t q)
{ /* q-th base in the strand's reading direction, 2-bit code; the reverse strand is the reverse complement */
	return strand ? 3 - (int)__ldg(s + (len - 1 - q)) : (int)__ldg(s + q);
}

__global__ void __launch_bounds__(256) k_idx_encode(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, uint64_t n, uint32_t *__restrict__ bad)
{
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		const uint8_t c = in[i];
		uint8_t v = 4;
		if (c == 'A' || c == 'a') v = 0; else if (c == 'C' || c == 'c') v = 1; else if (c == 'G' || c == 'g') v = 2; else if (c == 'T' || c == 't') v = 3;
		if (v == 4) { *bad = 1; v = 0; }
		out[i] = v;
	}
}

__global__ void __launch_bounds__(256) k_idx_qtab(const float *__restrict__ pore, uint32_t n, int q, float fine_min, float fine_max, float fine_range, uint8_t *__restrict__ qtab)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) qtab[i] = (uint8_t)(quantize_event(pore[i], fine_min, fine_max, fine_range, 1u << q) & ((1u << q) - 1u));
}

__device__ __forceinline__ uint32_t find_desc(const strand_desc_t *__restrict__ D, uint32_t n_desc, uint64_t g)
{ /* last descriptor whose seg_off <= g (descriptors with zero segments never match) */
	uint32_t lo = 0, hi = n_desc;
	while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (D[mid].seg_off <= g) lo = mid; else hi = mid; }
	return lo;
}

/* One segment of the diff filter.  list == nullptr: every segment, incoming state guessed by a warm-up run
 * (exact for the first segment of a strand); otherwise the segments in list[] with the state in in_state[]. */
__global__ void __launch_bounds__(128) k_idx_filter(const uint8_t *__restrict__ bases, const strand_desc_t *__restrict__ D, uint32_t n_desc, uint64_t n_seg_total,
                                                    const float *__restrict__ pore, const uint8_t *__restrict__ qtab, int k, float diff,
                                                    const uint32_t *__restrict__ list, uint32_t n_list,
                                                    uint64_t *__restrict__ bits, uint8_t *__restrict__ qv, uint32_t *__restrict__ cnt,
                                                    uint32_t *__restrict__ in_state, uint32_t *__restrict__ out_state)
{
	const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	uint64_t g;
	if (list) { if (t >= n_list) return; g = list[t]; } else { if (t >= n_seg_total) return; g = t; }
	const uint32_t d = find_desc(D, n_desc, g);
	const strand_desc_t S = D[d];
	const uint32_t sg = (uint32_t)(g - S.seg_off);
	const uint8_t *__restrict__ s = bases + S.base_off;
	const uint32_t j0 = sg * IDX_SEG, j1 = min(j0 + IDX_SEG, S.n_ev);
	const uint32_t kmask = (k >= 16) ? 0xffffffffu : ((1u << (2 * k)) - 1u);
	uint32_t state;
	uint32_t js = j0;
	if (list) state = in_state[g];
	else { js = j0 > IDX_WARM ? j0 - IDX_WARM : 0u; state = STATE_NONE; }
	/* k-mer of event js-1's tail: bases js .. js+k-2 */
	uint32_t kmer = 0;
	for (int q = 0; q < k - 1; ++q) kmer = (kmer << 2) | (uint32_t)strand_base(s, S.len, S.strand, js + (uint32_t)q);
	for (uint32_t j = js; j < j0; ++j) { /* warm-up: state only */
		kmer = ((kmer << 2) | (uint32_t)strand_base(s, S.len, S.strand, j + (uint32_t)k - 1)) & kmask;
		const float v = __ldg(pore + kmer);
		if (state == STATE_NONE || !(fabsf(__fsub_rn(v, __uint_as_float(state))) < diff)) state = __float_as_uint(v);
	}
	if (!list) in_state[g] = (j0 == 0) ? STATE_NONE : state;
	if (j0 == 0) state = STATE_NONE;
	uint32_t kept = 0;
	uint64_t *wout = bits + S.word_off + (uint64_t)sg * (IDX_SEG / 64);
	uint8_t *qout = qv + (S.word_off + (uint64_t)sg * (IDX_SEG / 64)) * 64;
	for (uint32_t w = 0; w < IDX_SEG / 64; ++w) {
		uint64_t m = 0;
		const uint32_t b0 = j0 + w * 64;
		for (uint32_t h = 0; h < 64; h += 16) { /* 16 events -> one 16-byte store of quantised values */
			uint32_t qq[4] = {0, 0, 0, 0};
#pragma unroll
			for (uint32_t u = 0; u < 16; ++u) {
				const uint32_t j = b0 + h + u;
				if (j < j1) {
					kmer = ((kmer << 2) | (uint32_t)strand_base(s, S.len, S.strand, j + (uint32_t)k - 1)) & kmask;
					const float v = __ldg(pore + kmer);
					if (state == STATE_NONE || !(fabsf(__fsub_rn(v, __uint_as_float(state))) < diff)) {
						state = __float_as_uint(v);
						m |= 1ULL << (h + u);
						++kept;
						qq[u >> 2] |= (uint32_t)__ldg(qtab + kmer) << (8 * (u & 3));
					}
				}
			}
			*(uint4 *)(qout + w * 64 + h) = make_uint4(qq[0], qq[1], qq[2], qq[3]);
		}
		wout[w] = m;
	}
	cnt[g] = kept;
	out_state[g] = state;
}

__global__ void __launch_bounds__(256) k_idx_verify(const strand_desc_t *__restrict__ D, uint32_t n_desc, uint64_t n_seg_total,
                                                    uint32_t *__restrict__ in_state, const uint32_t *__restrict__ out_state,
                                                    uint32_t *__restrict__ dirty, uint32_t *__restrict__ n_dirty, uint32_t dirty_cap)
{
	const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= n_seg_total || g == 0) return;
	const uint32_t d = find_desc(D, n_desc, g);
	if (D[d].seg_off == g) return; /* first segment of a strand: exact */
	const uint32_t want = out_state[g - 1];
	if (in_state[g] != want) {
		in_state[g] = want;
		const uint32_t at = atomicAdd(n_dirty, 1u);
		if (at < dirty_cap) dirty[at] = (uint32_t)g;
	}
}

/* ---- 64-bit exclusive scan of 32-bit counts: block sums, scan of block sums (one CTA), add ---- */
#define SCAN_TILE 2048u
__global__ void __launch_bounds__(256) k_scan_block_sums(const uint32_t *__restrict__ v, uint64_t n, uint64_t *__restrict__ bsum)
{
	__shared__ uint64_t sh[8];
	const uint64_t b0 = (uint64_t)blockIdx.x * SCAN_TILE;
	uint64_t acc = 0;
	for (uint32_t i = threadIdx.x; i < SCAN_TILE; i += 256) { const uint64_t j = b0 + i; if (j < n) acc += v[j]; }
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
	if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
	__syncthreads();
	if (threadIdx.x == 0) { uint64_t t = 0; for (int w = 0; w < 8; ++w) t += sh[w]; bsum[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(1024) k_scan_top(uint64_t *__restrict__ bsum, uint64_t nb, uint64_t *__restrict__ total)
{ /* one CTA: exclusive scan of nb block sums in place */
	__shared__ uint64_t sh[32];
	__shared__ uint64_t carry;
	if (threadIdx.x == 0) carry = 0;
	__syncthreads();
	for (uint64_t b0 = 0; b0 < nb; b0 += 1024) {
		const uint64_t i = b0 + threadIdx.x;
		const uint64_t v = i < nb ? bsum[i] : 0;
		uint64_t incl = v;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const uint64_t u = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += u; }
		if ((threadIdx.x & 31) == 31) sh[threadIdx.x >> 5] = incl;
		__syncthreads();
		uint64_t woff = 0;
		for (uint32_t w = 0; w < (threadIdx.x >> 5); ++w) woff += sh[w];
		const uint64_t c = carry;
		if (i < nb) bsum[i] = c + woff + incl - v;
		__syncthreads();
		if (threadIdx.x == 1023) carry = c + woff + incl;
		__syncthreads();
	}
	if (threadIdx.x == 0) *total = carry;
}
__global__ void __launch_bounds__(256) k_scan_apply(const uint32_t *__restrict__ v, uint64_t n, const uint64_t *__restrict__ bsum, uint64_t *__restrict__ out)
{ /* one CTA per tile: thread t owns 8 consecutive items */
	__shared__ uint64_t sh[8];
	const uint64_t b0 = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * 8;
	uint32_t x[8]; uint64_t sum = 0;
#pragma unroll
	for (int u = 0; u < 8; ++u) { x[u] = (b0 + u < n) ? v[b0 + u] : 0u; sum += x[u]; }
	uint64_t incl = sum;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { const uint64_t u = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += u; }
	if ((threadIdx.x & 31) == 31) sh[threadIdx.x >> 5] = incl;
	__syncthreads();
	uint64_t run = bsum[blockIdx.x] + incl - sum;
	for (uint32_t w = 0; w < (threadIdx.x >> 5); ++w) run += sh[w];
#pragma unroll
	for (int u = 0; u < 8; ++u) { if (b0 + u < n) out[b0 + u] = run; run += x[u]; }
}

/* kept events of strand S with position < p (p <= n_ev) */
__device__ __forceinline__ uint64_t kept_before(const strand_desc_t &S, const uint64_t *__restrict__ bits, const uint64_t *__restrict__ seg_pre, uint32_t p)
{
	const uint32_t sg = p / IDX_SEG;
	const uint64_t nseg = ((uint64_t)S.n_ev + IDX_SEG - 1) / IDX_SEG;
	if (sg >= nseg) return S.n_kept;
	uint64_t r = seg_pre[S.seg_off + sg] - seg_pre[S.seg_off];
	const uint64_t *w = bits + S.word_off + (uint64_t)sg * (IDX_SEG / 64);
	const uint32_t wi = (p % IDX_SEG) / 64, bi = p & 63u;
	for (uint32_t q = 0; q < wi; ++q) r += __popcll(w[q]);
	if (bi) r += __popcll(w[wi] & ((1ULL << bi) - 1ULL));
	return r;
}

__global__ void __launch_bounds__(256) k_idx_desc_counts(strand_desc_t *D, uint32_t n_desc, const uint64_t *__restrict__ seg_pre, uint64_t total_kept, int e)
{
	const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
	if (d >= n_desc) return;
	const uint64_t a = seg_pre[D[d].seg_off];
	const uint64_t nseg = ((uint64_t)D[d].n_ev + IDX_SEG - 1) / IDX_SEG;
	const uint64_t b = (d + 1 < n_desc) ? seg_pre[D[d + 1].seg_off] : total_kept;
	(void)nseg;
	const uint32_t nk = (uint32_t)(b - a);
	D[d].n_kept = nk;
	D[d].n_valid = nk >= (uint32_t)e ? nk - (uint32_t)e + 1 : 0u;
}

/* Thread per bitmap word.  A seed = e consecutive kept events; it is completed by its LAST event, so the thread
 * owning that event emits it after looking back over the e-1 kept events before its word. */
__global__ void __launch_bounds__(256) k_idx_seeds(const strand_desc_t *__restrict__ D, uint32_t n_desc, uint64_t n_words_total,
                                                   const uint64_t *__restrict__ bits, const uint8_t *__restrict__ qv, const uint64_t *__restrict__ seg_pre,
                                                   int e, int q, uint32_t *__restrict__ hash_out, uint64_t *__restrict__ y_out)
{
	const uint64_t gw = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (gw >= n_words_total) return;
	const uint64_t mine = bits[gw];
	if (!mine) return;
	/* descriptor by word offset (word_off is monotone like seg_off: both are 4 words per segment) */
	const uint32_t d = find_desc(D, n_desc, gw / (IDX_SEG / 64));
	const strand_desc_t S = D[d];
	if (S.n_valid == 0) return;
	const strand_desc_t O = D[d ^ 1u]; /* the other strand of the same sequence (descriptors come in pairs) */
	const uint32_t lw = (uint32_t)(gw - S.word_off);
	const uint64_t *__restrict__ W = bits + S.word_off;
	const uint8_t *__restrict__ Q = qv + S.word_off * 64;
	const uint64_t mev = (q * e >= 64) ? ~0ULL : ((1ULL << (q * e)) - 1ULL);
	/* rank of this word's first kept event */
	uint64_t r0;
	{
		const uint32_t sg = lw / (IDX_SEG / 64);
		r0 = seg_pre[S.seg_off + sg] - seg_pre[S.seg_off];
		for (uint32_t w = sg * (IDX_SEG / 64); w < lw; ++w) r0 += __popcll(W[w]);
	}
	uint32_t hp[IDX_MAX_E]; /* positions of the last e kept events, hp[IDX_MAX_E-1] newest */
#pragma unroll
	for (int i = 0; i < IDX_MAX_E; ++i) hp[i] = 0;
	uint64_t packed = 0;
	{ /* look back: up to e-1 kept events before this word, oldest first into the window */
		uint32_t prev[IDX_MAX_E]; int np = 0;
		const int need = e - 1;
		for (int64_t w = (int64_t)lw - 1; w >= 0 && np < need; --w) {
			uint64_t m = W[w];
			while (m && np < need) { const int b = 63 - __clzll((long long)m); m &= ~(1ULL << b); prev[np++] = (uint32_t)w * 64u + (uint32_t)b; }
		}
		for (int i = np - 1; i >= 0; --i) {
			packed = ((packed << q) | (uint64_t)Q[prev[i]]) & mev;
#pragma unroll
			for (int u = 0; u + 1 < IDX_MAX_E; ++u) hp[u] = hp[u + 1];
			hp[IDX_MAX_E - 1] = prev[i];
		}
	}
	uint64_t m = mine; uint64_t r = r0;
	while (m) {
		const int b = __ffsll((long long)m) - 1; m &= m - 1;
		const uint32_t pos = lw * 64u + (uint32_t)b;
		packed = ((packed << q) | (uint64_t)Q[pos]) & mev;
#pragma unroll
		for (int u = 0; u + 1 < IDX_MAX_E; ++u) hp[u] = hp[u + 1];
		hp[IDX_MAX_E - 1] = pos;
		++r; /* kept events up to and including this one */
		if (r >= (uint64_t)e) {
			uint32_t first = 0;
#pragma unroll
			for (int u = 0; u < IDX_MAX_E; ++u) if (u == IDX_MAX_E - e) first = hp[u];
			const uint64_t t = r - (uint64_t)e; /* rank of the seed inside its strand */
			/* seeds of the other strand that sort before this one: position < first (strand 0) or <= first (strand 1) */
			uint64_t other = kept_before(O, bits, seg_pre, S.strand ? min(first + 1u, O.n_ev) : first);
			if (other > O.n_valid) other = O.n_valid;
			const uint64_t slot = S.seed_base + t + other;
			hash_out[slot] = (uint32_t)seed_mix(packed);
			y_out[slot] = (uint64_t)S.id << 32 | (uint64_t)first << 1 | (uint64_t)S.strand;
		}
	}
}

/* ---- distinct keys + CSR offsets from the sorted hash array ---- */
#define HEAD_TILE 4096u
__global__ void __launch_bounds__(256) k_idx_head_count(const uint32_t *__restrict__ hs, uint64_t n, uint32_t *__restrict__ bcnt)
{
	__shared__ uint32_t sh;
	if (threadIdx.x == 0) sh = 0;
	__syncthreads();
	const uint64_t b0 = (uint64_t)blockIdx.x * HEAD_TILE;
	uint32_t c = 0;
	for (uint32_t i = threadIdx.x; i < HEAD_TILE; i += 256) { const uint64_t j = b0 + i; if (j < n && (j == 0 || hs[j] != hs[j - 1])) ++c; }
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
	if ((threadIdx.x & 31) == 0 && c) atomicAdd(&sh, c);
	__syncthreads();
	if (threadIdx.x == 0) bcnt[blockIdx.x] = sh;
}
__global__ void __launch_bounds__(256) k_idx_head_write(const uint32_t *__restrict__ hs, uint64_t n, const uint64_t *__restrict__ bpre, uint32_t *__restrict__ keys, uint64_t *__restrict__ off)
{
	__shared__ uint32_t wsum[8];
	__shared__ uint32_t run_s;
	const uint64_t b0 = (uint64_t)blockIdx.x * HEAD_TILE;
	if (threadIdx.x == 0) run_s = 0;
	__syncthreads();
	for (uint32_t t0 = 0; t0 < HEAD_TILE; t0 += 256) {
		const uint64_t j = b0 + t0 + threadIdx.x;
		const bool head = j < n && (j == 0 || hs[j] != hs[j - 1]);
		const uint32_t mk = __ballot_sync(0xffffffffu, head);
		if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = __popc(mk);
		__syncthreads();
		uint32_t base = run_s, tot = 0;
		for (uint32_t w = 0; w < 8; ++w) { if (w < (threadIdx.x >> 5)) base += wsum[w]; tot += wsum[w]; }
		if (head) { const uint64_t r = bpre[blockIdx.x] + base + __popc(mk & ((1u << (threadIdx.x & 31)) - 1u)); keys[r] = hs[j]; off[r] = j; }
		__syncthreads();
		if (threadIdx.x == 0) run_s += tot;
		__syncthreads();
	}
}

/* ---- mid_occ: k-th smallest list length (ks_ksmall over kh_val counts, rindex.c:1018-1036) ---- */
__global__ void __launch_bounds__(256) k_occ_hist(const uint64_t *__restrict__ off, uint64_t n_keys, int level, uint32_t hi_sel, unsigned long long *__restrict__ hist)
{
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_keys; i += (uint64_t)gridDim.x * blockDim.x) {
		const uint64_t c64 = off[i + 1] - off[i];
		const uint32_t c = c64 > 0xffffffffULL ? 0xffffffffu : (uint32_t)c64;
		if (level == 0) atomicAdd(&hist[c >> 16], 1ULL);
		else if ((c >> 16) == hi_sel) atomicAdd(&hist[c & 0xffffu], 1ULL);
	}
}

__global__ void k_bucket_fill(const uint32_t *__restrict__ keys, uint64_t n_keys, int bits, uint32_t *__restrict__ bucket)
{ /* bucket[b] = first key index whose top `bits` bits are >= b, for b in [0, 2^bits] */
	const uint64_t nb = (uint64_t)1 << bits;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n_keys; i += (uint64_t)gridDim.x * blockDim.x) {
		const uint64_t lo = i == 0 ? 0 : ((uint64_t)(keys[i - 1] >> (32 - bits)) + 1);
		const uint64_t hi = i == n_keys ? nb : (uint64_t)(keys[i] >> (32 - bits));
		for (uint64_t b = lo; b <= hi; ++b) bucket[b] = (uint32_t)i;
	}
}

struct dev_free { std::vector<void *> p; ~dev_free() { for (void *q : p) cudaFree(q); } void drop(void *q) { for (auto &x : p) if (x == q) { cudaFree(q); x = nullptr; } } void keep(void *q) { for (auto &x : p) if (x == q) x = nullptr; } };

#define IDX_TRY(call)                                                                                     \
	do {                                                                                                  \
		cudaError_t e_ = (call);                                                                          \
		if (e_ != cudaSuccess) { rh_set_error("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); return bail(); } \
	} while (0)

template <class T> T *dmalloc(dev_free &F, size_t n)
{
	void *p = nullptr;
	if (cudaMalloc(&p, (n ? n : 1) * sizeof(T)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
	F.p.push_back(p);
	return (T *)p;
}

int scan64(const uint32_t *d_v, uint64_t n, uint64_t *d_out, uint64_t *d_bsum, uint64_t *d_total)
{
	const uint64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
	k_scan_block_sums<<<(unsigned)nb, 256>>>(d_v, n, d_bsum);
	k_scan_top<<<1, 1024>>>(d_bsum, nb, d_total);
	k_scan_apply<<<(unsigned)nb, 256>>>(d_v, n, d_bsum, d_out);
	return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

} // namespace

/* the bucket table the mapper's lookup uses; lives with the device index */
int rh_index_dev_make_buckets(rh_index_s *idx)
{
	rh_index_dev_t &V = idx->dev;
	if (V.device < 0) return RH_ERR_ARG;
	if (V.bucket) return RH_OK;
	int bits = 10; while (bits < 26 && ((uint64_t)1 << bits) < V.n_keys) ++bits;
	uint32_t *b = nullptr;
	if (cudaMalloc((void **)&b, (((size_t)1 << bits) + 1) * 4) != cudaSuccess) { rh_set_error("bucket table: out of device memory"); return RH_ERR_NOMEM; }
	k_bucket_fill<<<1184, 256>>>(V.keys, V.n_keys, bits, b);
	if (cudaDeviceSynchronize() != cudaSuccess) { cudaFree(b); rh_set_error("bucket table kernel failed: %s", cudaGetErrorString(cudaGetLastError())); return RH_ERR_CUDA; }
	V.bucket = b; V.bucket_bits = bits;
	return RH_OK;
}

/* ri_idx_cal_max_occ on the device-resident index: value of the k-th smallest list length, k = (uint32)((1-f)*n) */
int rh_index_dev_kth_occ(const rh_index_s *idx, uint64_t kth, uint32_t *out)
{
	const rh_index_dev_t &V = idx->dev;
	if (V.device < 0 || V.n_keys == 0) return RH_ERR_ARG;
	cudaSetDevice(V.device);
	unsigned long long *d_hist = nullptr;
	if (cudaMalloc((void **)&d_hist, 65536 * 8) != cudaSuccess) return RH_ERR_NOMEM;
	std::vector<unsigned long long> h(65536);
	uint32_t hi = 0; uint64_t rem = kth;
	for (int level = 0; level < 2; ++level) {
		cudaMemset(d_hist, 0, 65536 * 8);
		k_occ_hist<<<1184, 256>>>(V.off, V.n_keys, level, hi, d_hist);
		if (cudaMemcpy(h.data(), d_hist, 65536 * 8, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaFree(d_hist); return RH_ERR_CUDA; }
		uint32_t b = 0;
		for (; b < 65536; ++b) { if (rem < h[b]) break; rem -= h[b]; }
		if (b == 65536) b = 65535;
		if (level == 0) hi = b; else *out = hi << 16 | b;
	}
	cudaFree(d_hist);
	return RH_OK;
}

void rh_index_dev_release(rh_index_s *idx)
{
	rh_index_dev_t &V = idx->dev;
	if (V.device < 0) return;
	int cur = 0; cudaGetDevice(&cur);
	cudaSetDevice(V.device);
	if (V.keys) cudaFree(V.keys);
	if (V.off) cudaFree(V.off);
	if (V.pos) cudaFree(V.pos);
	if (V.bucket) cudaFree(V.bucket);
	V = rh_index_dev_t();
	cudaSetDevice(cur);
}

/* host mirror of a device-resident index (accessors, .ind writer, tests) */
int rh_index_sync_host(const rh_index_s *cidx)
{
	rh_index_s *idx = const_cast<rh_index_s *>(cidx);
	if (idx->host_valid) return RH_OK;
	const rh_index_dev_t &V = idx->dev;
	if (V.device < 0) return RH_ERR_ARG;
	int cur = 0; cudaGetDevice(&cur);
	cudaSetDevice(V.device);
	idx->keys.resize(V.n_keys); idx->off.resize(V.n_keys + 1); idx->pos.resize(V.n_pos);
	bool ok = true;
	if (V.n_keys) ok = ok && cudaMemcpy(idx->keys.data(), V.keys, V.n_keys * 4, cudaMemcpyDeviceToHost) == cudaSuccess;
	ok = ok && cudaMemcpy(idx->off.data(), V.off, (V.n_keys + 1) * 8, cudaMemcpyDeviceToHost) == cudaSuccess;
	if (V.n_pos) ok = ok && cudaMemcpy(idx->pos.data(), V.pos, V.n_pos * 8, cudaMemcpyDeviceToHost) == cudaSuccess;
	cudaSetDevice(cur);
	if (!ok) { rh_set_error("index download failed: %s", cudaGetErrorString(cudaGetLastError())); return RH_ERR_CUDA; }
	idx->host_valid = true;
	return RH_OK;
}

/* Builds the index of `n_seq` sequences whose 2-bit base codes (one per byte, values 0..3 = A,C,G,T) already lie
 * back to back in device memory.  Everything stays on `device`. */
extern "C" rh_index_t *rh_index_build_dev(const rh_params_t *p, const float *pore_vals, uint32_t n_pore_vals,
                                           uint32_t n_seq, const char *const *names, const void *d_codes,
                                           const uint32_t *lens, int device)
{
	if (!p || !pore_vals || n_pore_vals < (1u << (2 * p->k)) || (n_seq && (!names || !d_codes || !lens))) { rh_set_error("rh_index_build_dev: bad arguments"); return NULL; }
	if (!(p->w == 0 && p->n == 0 && p->k >= 1 && p->k <= 15 && p->e >= 1 && p->e <= IDX_MAX_E && p->q >= 1 && p->q <= 8 && p->e * p->q <= 64)) {
		rh_set_error("rh_index_build_dev: only w = 0, e <= %d, q <= 8 are built on the device", IDX_MAX_E); return NULL;
	}
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device >= ndev) { rh_set_error("no usable CUDA device (count=%d, asked %d)", ndev, device); return NULL; }
	if (cudaSetDevice(device) != cudaSuccess) { rh_set_error("cudaSetDevice(%d) failed", device); return NULL; }
	const int k = p->k, e = p->e;
	const uint8_t *d_bases = (const uint8_t *)d_codes;

	rh_index_s *idx = new rh_index_s();
	idx->flag = p->idx_flag; idx->w = p->w; idx->e = p->e; idx->n = p->n; idx->q = p->q; idx->k = p->k;
	idx->diff = p->diff; idx->fine_min = p->fine_min; idx->fine_max = p->fine_max; idx->fine_range = p->fine_range;
	for (uint32_t i = 0; i < n_seq; ++i) { idx->names.emplace_back(names[i]); idx->lens.push_back(lens[i]); }
	idx->off.push_back(0);
	auto bail = [&]() -> rh_index_t * { delete idx; return NULL; };

	/* descriptors: always two per sequence (strand 0, strand 1), so that d^1 is the other strand */
	std::vector<strand_desc_t> D;
	uint64_t n_bases = 0, n_words = 0, n_seg = 0;
	for (uint32_t i = 0; i < n_seq; ++i) {
		const uint32_t len = lens[i], ne = len >= (uint32_t)k ? len - (uint32_t)k + 1 : 0;
		for (uint32_t s = 0; s < 2; ++s) {
			strand_desc_t d; memset(&d, 0, sizeof(d));
			d.base_off = n_bases; d.len = len; d.n_ev = ne; d.id = i; d.strand = s;
			d.word_off = n_words; d.seg_off = n_seg;
			D.push_back(d);
			const uint64_t sg = ((uint64_t)ne + IDX_SEG - 1) / IDX_SEG;
			n_seg += sg; n_words += sg * (IDX_SEG / 64);
		}
		n_bases += len;
	}
	if (n_seg == 0) { idx->dev.device = -1; return idx; }
	if (n_seg >= (1ULL << 32)) { rh_set_error("rh_index_build_dev: reference too large"); return bail(); }
	const uint32_t nd = (uint32_t)D.size();

	dev_free F;
	strand_desc_t *d_D = dmalloc<strand_desc_t>(F, nd);
	float *d_pore = dmalloc<float>(F, n_pore_vals);
	uint8_t *d_qtab = dmalloc<uint8_t>(F, n_pore_vals);
	uint64_t *d_bits = dmalloc<uint64_t>(F, n_words);
	uint8_t *d_qv = dmalloc<uint8_t>(F, n_words * 64);
	uint32_t *d_cnt = dmalloc<uint32_t>(F, n_seg), *d_in = dmalloc<uint32_t>(F, n_seg), *d_out = dmalloc<uint32_t>(F, n_seg);
	uint64_t *d_pre = dmalloc<uint64_t>(F, n_seg + 1);
	const uint64_t n_sb = (std::max<uint64_t>(n_seg, 1) + SCAN_TILE - 1) / SCAN_TILE;
	uint64_t *d_bsum = dmalloc<uint64_t>(F, n_sb + 1);
	const uint32_t dirty_cap = 1u << 22;
	uint32_t *d_dirty = dmalloc<uint32_t>(F, dirty_cap), *d_ndirty = dmalloc<uint32_t>(F, 2);
	uint64_t *d_total = dmalloc<uint64_t>(F, 2);
	if (!d_D || !d_pore || !d_qtab || !d_bits || !d_qv || !d_cnt || !d_in || !d_out || !d_pre || !d_bsum || !d_dirty || !d_ndirty || !d_total) { rh_set_error("rh_index_build_dev: out of device memory"); return bail(); }
	IDX_TRY(cudaMemcpy(d_D, D.data(), nd * sizeof(strand_desc_t), cudaMemcpyHostToDevice));
	IDX_TRY(cudaMemcpy(d_pore, pore_vals, (size_t)n_pore_vals * sizeof(float), cudaMemcpyHostToDevice));
	k_idx_qtab<<<(n_pore_vals + 255) / 256, 256>>>(d_pore, n_pore_vals, p->q, p->fine_min, p->fine_max, p->fine_range, d_qtab);
	k_idx_filter<<<(unsigned)((n_seg + 127) / 128), 128>>>(d_bases, d_D, nd, n_seg, d_pore, d_qtab, k, p->diff, nullptr, 0, d_bits, d_qv, d_cnt, d_in, d_out);
	for (int iter = 0;; ++iter) { /* fixed point of the guessed incoming states */
		IDX_TRY(cudaMemset(d_ndirty, 0, 4));
		k_idx_verify<<<(unsigned)((n_seg + 255) / 256), 256>>>(d_D, nd, n_seg, d_in, d_out, d_dirty, d_ndirty, dirty_cap);
		uint32_t ndirty = 0;
		IDX_TRY(cudaMemcpy(&ndirty, d_ndirty, 4, cudaMemcpyDeviceToHost));
		if (ndirty == 0) break;
		if (ndirty > dirty_cap) { /* list overflow: in_state of every wrong guess is already corrected; redo all listed, the rest next turn */
			ndirty = dirty_cap;
		}
		k_idx_filter<<<(ndirty + 127) / 128, 128>>>(d_bases, d_D, nd, n_seg, d_pore, d_qtab, k, p->diff, d_dirty, ndirty, d_bits, d_qv, d_cnt, d_in, d_out);
		if (iter > 1000000) { rh_set_error("rh_index_build_dev: filter fix-up did not converge"); return bail(); }
	}
	if (scan64(d_cnt, n_seg, d_pre, d_bsum, d_total)) { rh_set_error("scan failed"); return bail(); }
	uint64_t total_kept = 0;
	IDX_TRY(cudaMemcpy(&total_kept, d_total, 8, cudaMemcpyDeviceToHost));
	IDX_TRY(cudaMemcpy(d_pre + n_seg, &total_kept, 8, cudaMemcpyHostToDevice));
	k_idx_desc_counts<<<(nd + 255) / 256, 256>>>(d_D, nd, d_pre, total_kept, e);
	IDX_TRY(cudaMemcpy(D.data(), d_D, nd * sizeof(strand_desc_t), cudaMemcpyDeviceToHost));
	uint64_t n_seeds = 0;
	for (uint32_t i = 0; i < nd; i += 2) { D[i].seed_base = D[i + 1].seed_base = n_seeds; n_seeds += (uint64_t)D[i].n_valid + D[i + 1].n_valid; }
	IDX_TRY(cudaMemcpy(d_D, D.data(), nd * sizeof(strand_desc_t), cudaMemcpyHostToDevice));
	if (n_seeds == 0) { idx->dev.device = -1; return idx; }

	uint32_t *d_h0 = dmalloc<uint32_t>(F, n_seeds), *d_h1 = dmalloc<uint32_t>(F, n_seeds);
	uint64_t *d_y0 = dmalloc<uint64_t>(F, n_seeds), *d_y1 = dmalloc<uint64_t>(F, n_seeds);
	if (!d_h0 || !d_h1 || !d_y0 || !d_y1) { rh_set_error("rh_index_build_dev: out of device memory (%.1f GB of sort buffers)", 24.0 * n_seeds / 1e9); return bail(); }
	k_idx_seeds<<<(unsigned)((n_words + 255) / 256), 256>>>(d_D, nd, n_words, d_bits, d_qv, d_pre, e, p->q, d_h0, d_y0);
	IDX_TRY(cudaGetLastError());
	IDX_TRY(cudaDeviceSynchronize());
	F.drop(d_qv); F.drop(d_bits);
	{
		size_t tb = 0;
		cub::DoubleBuffer<uint32_t> kb(d_h0, d_h1);
		cub::DoubleBuffer<uint64_t> vb(d_y0, d_y1);
		IDX_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tb, kb, vb, (int64_t)n_seeds, 0, 32));
		uint8_t *d_tmp = dmalloc<uint8_t>(F, tb);
		if (!d_tmp) { rh_set_error("rh_index_build_dev: out of device memory"); return bail(); }
		IDX_TRY(cub::DeviceRadixSort::SortPairs(d_tmp, tb, kb, vb, (int64_t)n_seeds, 0, 32)); /* stable: list order = (sequence, position, strand) */
		IDX_TRY(cudaDeviceSynchronize());
		F.drop(d_tmp);
		if (kb.Current() != d_h0) std::swap(d_h0, d_h1);
		if (vb.Current() != d_y0) std::swap(d_y0, d_y1);
	}
	F.drop(d_h1); F.drop(d_y1);
	/* d_h0 = sorted hashes, d_y0 = positions in index order */
	const uint64_t n_hb = (n_seeds + HEAD_TILE - 1) / HEAD_TILE;
	uint32_t *d_bc = dmalloc<uint32_t>(F, n_hb);
	uint64_t *d_bp = dmalloc<uint64_t>(F, n_hb + 1);
	const uint64_t n_sb2 = (n_hb + SCAN_TILE - 1) / SCAN_TILE;
	uint64_t *d_bs2 = dmalloc<uint64_t>(F, n_sb2 + 1);
	if (!d_bc || !d_bp || !d_bs2) { rh_set_error("rh_index_build_dev: out of device memory"); return bail(); }
	k_idx_head_count<<<(unsigned)n_hb, 256>>>(d_h0, n_seeds, d_bc);
	if (scan64(d_bc, n_hb, d_bp, d_bs2, d_total)) { rh_set_error("scan failed"); return bail(); }
	uint64_t n_keys = 0;
	IDX_TRY(cudaMemcpy(&n_keys, d_total, 8, cudaMemcpyDeviceToHost));
	uint32_t *d_keys = dmalloc<uint32_t>(F, n_keys);
	uint64_t *d_off = dmalloc<uint64_t>(F, n_keys + 1);
	if (!d_keys || !d_off) { rh_set_error("rh_index_build_dev: out of device memory"); return bail(); }
	k_idx_head_write<<<(unsigned)n_hb, 256>>>(d_h0, n_seeds, d_bp, d_keys, d_off);
	IDX_TRY(cudaMemcpy(d_off + n_keys, &n_seeds, 8, cudaMemcpyHostToDevice));
	IDX_TRY(cudaDeviceSynchronize());
	F.keep(d_keys); F.keep(d_off); F.keep(d_y0);
	idx->dev.device = device; idx->dev.keys = d_keys; idx->dev.off = d_off; idx->dev.pos = d_y0; idx->dev.n_keys = n_keys; idx->dev.n_pos = n_seeds;
	idx->host_valid = false;
	idx->off.clear();
	return idx;
}

/* Same result as rh_index_build (ri_idx_gen semantics), computed on `device` from host sequences.  Falls back to the
 * host builder for inputs the kernels do not cover (non-ACGT bases, minimizers). */
extern "C" rh_index_t *rh_index_build_gpu(const rh_params_t *p, const float *pore_vals, uint32_t n_pore_vals,
                                           uint32_t n_seq, const char *const *names, const char *const *seqs,
                                           const uint32_t *lens, int device)
{
	if (!p || !pore_vals || n_pore_vals < (1u << (2 * p->k)) || (n_seq && (!names || !seqs || !lens))) { rh_set_error("rh_index_build_gpu: bad arguments"); return NULL; }
	const int host_threads = std::max(1, rh_host_threads());
	const bool plain = p->w == 0 && p->n == 0 && p->k >= 1 && p->k <= 15 && p->e >= 1 && p->e <= IDX_MAX_E && p->q >= 1 && p->q <= 8 && p->e * p->q <= 64;
	if (!plain) return rh_index_build(p, pore_vals, n_pore_vals, n_seq, names, seqs, lens, host_threads);
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device >= ndev) { rh_set_error("no usable CUDA device (count=%d, asked %d)", ndev, device); return NULL; }
	if (cudaSetDevice(device) != cudaSuccess) { rh_set_error("cudaSetDevice(%d) failed", device); return NULL; }
	uint64_t total = 0;
	for (uint32_t i = 0; i < n_seq; ++i) total += lens[i];
	uint8_t *d_raw = nullptr, *d_codes = nullptr; uint32_t *d_bad = nullptr;
	auto cleanup = [&]() { if (d_raw) cudaFree(d_raw); if (d_codes) cudaFree(d_codes); if (d_bad) cudaFree(d_bad); };
	const uint64_t slab = std::min<uint64_t>(std::max<uint64_t>(total, 1), (uint64_t)256 << 20);
	if (cudaMalloc((void **)&d_raw, slab) != cudaSuccess || cudaMalloc((void **)&d_codes, total ? total : 1) != cudaSuccess || cudaMalloc((void **)&d_bad, 4) != cudaSuccess) {
		cleanup(); rh_set_error("rh_index_build_gpu: out of device memory"); return NULL;
	}
	cudaMemset(d_bad, 0, 4);
	uint64_t o = 0;
	for (uint32_t i = 0; i < n_seq; ++i) {
		for (uint64_t c0 = 0; c0 < lens[i]; c0 += slab) { /* ASCII -> codes through a bounded staging slab */
			const uint64_t m = std::min<uint64_t>(slab, lens[i] - c0);
			if (cudaMemcpy(d_raw, seqs[i] + c0, m, cudaMemcpyHostToDevice) != cudaSuccess) { cleanup(); rh_set_error("sequence upload failed"); return NULL; }
			k_idx_encode<<<1184, 256>>>(d_raw, d_codes + o + c0, m, d_bad);
		}
		o += lens[i];
	}
	uint32_t bad = 0;
	if (cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost) != cudaSuccess) { cleanup(); rh_set_error("encode kernel failed: %s", cudaGetErrorString(cudaGetLastError())); return NULL; }
	cudaFree(d_raw); d_raw = nullptr;
	if (bad) { /* N and friends keep the previous k-mer in ri_seq_to_sig: a sequential rule, left to the host builder */
		cleanup();
		fprintf(stderr, "[rawhash_b200] reference holds non-ACGT bases: index built on the host (%d threads)\n", host_threads);
		return rh_index_build(p, pore_vals, n_pore_vals, n_seq, names, seqs, lens, host_threads);
	}
	rh_index_t *idx = rh_index_build_dev(p, pore_vals, n_pore_vals, n_seq, names, d_codes, lens, device);
	cleanup();
	return idx;
}
