/*
 * rh_kernels.cuh — the CUDA kernels of the mapping hot path (sm_100a).
 *
 * Stage map (reference function each kernel takes over, paths relative to the RawHash tree):
 *   k_filtered_len      src/rsig.c:496-503        raw int16 -> pA, count samples with 30<pA<200
 *   k_sig_*             src/revent.c:221-316 + src/rsketch.c:143-204   (the event stage, rh_signal.cuh)
 *   k_seed_count        src/rseed.c:60-154 + src/rindex.c:497-514      lookup, occ filter, rep_len
 *   k_seed_expand       src/rmap.cpp:74-116       anchors from position lists + previous chunk's
 *   k_sort_smem/_block/_ties  src/ksort.h:101-151  anchor sort with klib's exact tie order (rh_anchor_sort.cuh)
 *   k_chain_dp          src/lchain.c:385-505      chaining DP
 *   k_chain_finish      src/lchain.c:95-281, src/hit.c:100-150      backtrack, compact_a, regions (rh_chain_finish.cuh)
 *   k_chain_decide      src/hit.c:195-263,338-367,502-539, src/rmap.cpp:423-586   parents, MAPQ, stop rules, records
 *
 * Work decomposition: a "slot" is one (read, chunk) pair of the current chunk round.
 */
#ifndef RH_KERNELS_CUH
#define RH_KERNELS_CUH

#include "rh_dev.cuh"
#include "rh_sort.cuh"

#define RH_WARP 32

/* per-read state carried across chunk rounds (reference: locals of map_worker_for + ri_reg1_t) */
struct read_state_t {
	uint64_t raw_beg, raw_end;     /* sample range of this read in the concatenated raw buffer */
	uint64_t cursor;               /* next raw sample to consume (absolute index)              */
	double cal_offset, cal_scale;  /* slow5 offset; (double)(float)(range/digitisation)        */
	double sum, sum2;              /* running Σx, Σx² of normalize_signal                      */
	uint32_t n_sum;                /* running sample count                                     */
	uint32_t l_sig;                /* total samples surviving the pA filter (qlen)             */
	uint32_t ev_offset;            /* reg->offset: events consumed by earlier chunks           */
	uint32_t prev_n;               /* carried chain anchors                                    */
	uint64_t prev_off;             /* their offset in the carry arena of the previous round    */
	uint32_t name_ub;              /* Rawsamble: #target names <= this read's name             */
	uint32_t done;
};

/* per-slot bookkeeping for one chunk round */
struct slot_t {
	uint32_t read;                 /* read index in the batch                                  */
	uint32_t chunk_len;            /* filtered samples this chunk must consume                 */
	uint32_t c_count;              /* chunk number                                             */
	uint32_t n_sig, n_peaks, n_events, n_seeds;
	uint32_t gated;                /* n_events < min_events                                    */
	uint64_t z_off;                /* offsets into the signal scratch arenas                   */
	uint64_t e_off;
	uint32_t e_cap;
	uint32_t n_new;                /* anchors from seed hits                                   */
	uint32_t n_anchors;            /* n_new + prev_n                                           */
	int32_t  rep_len;
	uint64_t a_off;                /* byte offset of this slot's region in the anchor arena    */
	uint32_t n_u, n_v, n_regs;
	uint32_t raw_used;             /* raw samples streamed by the event kernel for this chunk  */
	uint32_t n_ties;               /* adjacent equal keys found by the anchor sort              */
	uint32_t n_seg;                /* independent DP segments                                   */
	uint32_t ev_done;              /* events already produced by the fast event kernel          */
	uint32_t n_link;               /* anchors with a DP predecessor: chains carry at most 2 n_link anchors to the next chunk */
};

/* layout of a slot's region in the anchor arena (n = n_anchors) */
#ifndef RH_SLOT_BYTES_PER_ANCHOR
#define RH_SLOT_BYTES_PER_ANCHOR 144 /* also in rh_host.h (rh_slot_region_bytes: the host-side planner) */
#endif
__host__ __device__ inline uint64_t slot_region_bytes(uint64_t n) { return ((n * RH_SLOT_BYTES_PER_ANCHOR + 1024 + 255) / 256) * 256; }
/* the region-record area (48n + 1024 bytes at offset 96n): sort scratch of k_chain_finish first, records after */
__host__ __device__ inline uint64_t fin_regs_off(uint64_t n) { return (6 * n + 256 + 63) & ~63ULL; }
__host__ __device__ inline uint64_t fin_regs_cap(uint64_t n) { return (48 * n + 1024 - fin_regs_off(n)) / 64; /* sizeof(dev_reg_t) */ }
struct slot_mem_t {
	anchor_t *A, *B, *Z, *W;
	int32_t *f, *p, *v, *t;
	uint64_t *U, *U2;
	dev_reg_t *regs;
	uint32_t reg_cap;
};
__device__ __forceinline__ slot_mem_t slot_mem(uint8_t *arena, uint64_t a_off, uint64_t n)
{
	slot_mem_t m; uint8_t *b = arena + a_off;
	m.A = (anchor_t *)b; b += 16 * n;
	m.B = (anchor_t *)b; b += 16 * n;
	m.Z = (anchor_t *)b; b += 16 * n;
	m.W = (anchor_t *)b; b += 16 * n;
	m.f = (int32_t *)b; b += 4 * n;
	m.p = (int32_t *)b; b += 4 * n;
	m.v = (int32_t *)b; b += 4 * n;
	m.t = (int32_t *)b; b += 4 * n;
	m.U = (uint64_t *)b; b += 8 * n;
	m.U2 = (uint64_t *)b; b += 8 * n;
	m.regs = (dev_reg_t *)b; /* 48n + 1024 bytes left */
	m.reg_cap = (uint32_t)((48 * n + 1024) / sizeof(dev_reg_t));
	return m;
}

/* =============================================================================================
 * K0  filtered length of every read: one warp per read, 16-byte vector loads.
 * Pure streaming (2 B/sample, ~6 instr/sample): the one HBM-bound kernel of the path.
 * ===========================================================================================*/
__global__ void __launch_bounds__(256) k_filtered_len(const int16_t *__restrict__ raw, read_state_t *__restrict__ rs, uint32_t n_reads)
{
	const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) / RH_WARP, lane = threadIdx.x % RH_WARP;
	if (warp >= n_reads) return;
	const uint64_t beg = rs[warp].raw_beg, end = rs[warp].raw_end;
	const double off = rs[warp].cal_offset, sc = rs[warp].cal_scale;
	uint32_t kept = 0;
	/* head: up to the first 16-byte boundary */
	uint64_t a0 = (beg + 7) & ~7ULL; if (a0 > end) a0 = end;
	for (uint64_t i = beg + lane; i < a0; i += RH_WARP) kept += pa_keep(raw_to_pa(raw[i], off, sc));
	const uint64_t nvec = (end - a0) / 8;
	const int4 *v = (const int4 *)(raw + a0);
	for (uint64_t j = lane; j < nvec; j += RH_WARP) {
		const int4 q = __ldg(v + j);
		const int w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
		for (int t = 0; t < 4; ++t) {
			kept += pa_keep(raw_to_pa((int)(short)(w[t] & 0xffff), off, sc));
			kept += pa_keep(raw_to_pa((int)(short)(w[t] >> 16), off, sc));
		}
	}
	for (uint64_t i = a0 + nvec * 8 + lane; i < end; i += RH_WARP) kept += pa_keep(raw_to_pa(raw[i], off, sc));
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) kept += __shfl_xor_sync(0xffffffffu, kept, o);
	if (lane == 0) rs[warp].l_sig = kept;
}

/* =============================================================================================
 * K2a  seed lookup + occurrence filter + repeat length.  One warp per slot; lanes stride over
 * the slot's seeds.  Index lookup = bucket table on the hash's top bits, then a short scan of
 * the sorted key array.  Per seed we keep (count, position offset); count 0 = absent/filtered.
 * ===========================================================================================*/
struct k2_args_t {
	slot_t *slots; uint32_t n_slots;
	read_state_t *rs;
	const uint32_t *seed_hash; const uint32_t *seed_pos;
	uint32_t *seed_cnt;      /* [e_off..] kept hits of the seed (after filters)          */
	uint64_t *seed_src;      /* [e_off..] offset of the seed's position list in idx.pos   */
	uint32_t *seed_dst;      /* [e_off..] exclusive prefix of seed_cnt inside the slot    */
	uint8_t *arena;          /* anchor arena                                              */
	const anchor_t *carry_in; anchor_t *carry_out;
};

__device__ __forceinline__ bool index_find(const dev_index_t &I, uint32_t h, uint64_t *src, uint32_t *n)
{
	const uint32_t b = I.bucket_bits >= 32 ? h : (h >> (32 - I.bucket_bits));
	uint32_t lo = I.bucket[b], hi = I.bucket[b + 1];
	while (lo < hi) { /* buckets hold ~1 key on average */
		const uint32_t mid = (lo + hi) >> 1;
		const uint32_t kv = I.keys[mid];
		if (kv == h) { const uint64_t o = I.off[mid]; *src = o; *n = (uint32_t)(I.off[mid + 1] - o); return true; }
		if (kv < h) lo = mid + 1; else hi = mid;
	}
	return false;
}

__global__ void __launch_bounds__(256) k_seed_count(k2_args_t A, dev_index_t I, dev_params_t P)
{
	const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) / RH_WARP, lane = threadIdx.x % RH_WARP;
	if (warp >= A.n_slots) return;
	slot_t *S = &A.slots[warp];
	const read_state_t *R = &A.rs[S->read];
	if (S->gated) { if (lane == 0) { S->n_new = 0; S->rep_len = 0; S->n_anchors = R->prev_n; } return; }
	const uint32_t ns = S->n_seeds;
	const uint32_t *sh = A.seed_hash + S->e_off, *sp = A.seed_pos + S->e_off;
	uint32_t *scnt = A.seed_cnt + S->e_off, *sdst = A.seed_dst + S->e_off;
	uint64_t *ssrc = A.seed_src + S->e_off;
	const uint32_t span = (uint32_t)(P.k + P.e - 1);
	uint32_t running = 0;
	int rep_st = 0, rep_en = 0, rep_len = 0;
	for (uint32_t base = 0; base < ns; base += RH_WARP) {
		const uint32_t i = base + lane;
		uint32_t n = 0, flt = 0; uint64_t src = 0;
		if (i < ns) {
			const uint32_t h = sh[i];
			if (index_find(I, h, &src, &n)) {
				if ((int)n > P.mid_occ) flt = 1;
				else if (P.ava) { /* Rawsamble self/duplicate filter, rmap.cpp:86: keep iff qname < target name */
					uint32_t keep = 0;
					for (uint32_t k = 0; k < n; ++k) keep += I.name_rank[(uint32_t)(I.pos[src + k] >> 32)] >= R->name_ub;
					/* the count is what expands; the list is re-filtered at expansion */
					n = keep | 0x80000000u; /* mark: needs per-hit filtering */
				}
			}
		}
		/* repeat length (rseed.c:134-151) is a sequential merge over filtered seeds in read order */
		const uint32_t fmask = __ballot_sync(0xffffffffu, flt);
		if (fmask) {
			uint32_t m = fmask;
			while (m) {
				const int b = __ffs(m) - 1; m &= m - 1;
				const uint32_t qp = __shfl_sync(0xffffffffu, (i < ns) ? sp[i] : 0u, b);
				const int st = (int)qp + 1, en = st + (int)span + 1;
				if (st > rep_en) { rep_len += rep_en - rep_st; rep_st = st; rep_en = en; } else rep_en = en;
			}
		}
		const uint32_t cnt = flt ? 0u : (n & 0x7fffffffu);
		uint32_t incl = cnt;
#pragma unroll
		for (int o = 1; o < RH_WARP; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
		if (i < ns) { scnt[i] = flt ? 0u : n; ssrc[i] = src; sdst[i] = running + incl - cnt; }
		running += __shfl_sync(0xffffffffu, incl, RH_WARP - 1);
	}
	rep_len += rep_en - rep_st;
	if (lane == 0) { S->n_new = running; S->rep_len = rep_len; S->n_anchors = running + R->prev_n; }
}

/* K2b  anchors (rmap.cpp:74-116): for each kept seed, lanes cover its position list (coalesced
 * reads of idx.pos, coalesced 16-byte anchor writes); then the previous chunk's chain anchors. */
__global__ void __launch_bounds__(256) k_seed_expand(k2_args_t A, dev_index_t I, dev_params_t P)
{
	const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) / RH_WARP, lane = threadIdx.x % RH_WARP;
	if (warp >= A.n_slots) return;
	slot_t *S = &A.slots[warp];
	read_state_t *R = &A.rs[S->read];
	const uint32_t n_tot = S->n_anchors;
	if (n_tot == 0) return;
	anchor_t *out = (anchor_t *)(A.arena + S->a_off) + n_tot; /* slot_mem::B: unsorted input of the anchor sort */
	const uint32_t n_new = S->n_new;
	if (!S->gated) {
		const uint32_t ns = S->n_seeds;
		const uint32_t *sh = A.seed_hash + S->e_off, *sp = A.seed_pos + S->e_off;
		const uint32_t *scnt = A.seed_cnt + S->e_off, *sdst = A.seed_dst + S->e_off;
		const uint64_t *ssrc = A.seed_src + S->e_off;
		const uint64_t span = (uint64_t)(P.k + P.e - 1);
		const uint32_t ev_off = R->ev_offset;
		for (uint32_t i = 0; i < ns; ++i) {
			const uint32_t craw = scnt[i];
			if (craw == 0) continue;
			const bool per_hit = craw & 0x80000000u;
			const uint32_t h = sh[i];
			const bool tandem = (i > 0 && sh[i - 1] == h) || (i + 1 < ns && sh[i + 1] == h); /* rseed.c:80-81 */
			const uint64_t y = span << 32 | (uint64_t)(uint32_t)(sp[i] + ev_off) | (tandem ? (1ULL << 38) : 0ULL);
			const uint64_t *src = I.pos + ssrc[i];
			anchor_t *dst = out + sdst[i];
			if (!per_hit) {
				for (uint32_t k = lane; k < craw; k += RH_WARP) {
					const uint64_t hv = src[k];
					anchor_t a;
					a.x = (hv & 0x7fffffff80000000ULL) | (uint64_t)((uint32_t)(hv >> 1) & 0x7fffffffu) | ((hv & 1) << 63);
					a.y = y;
					dst[k] = a;
				}
			} else { /* Rawsamble: list length is the unfiltered count; compact kept hits in order */
				uint32_t full; uint64_t dummy;
				index_find(I, h, &dummy, &full);
				uint32_t w = 0;
				for (uint32_t base = 0; base < full; base += RH_WARP) {
					const uint32_t k = base + lane;
					uint64_t hv = 0; bool keep = false;
					if (k < full) { hv = src[k]; keep = I.name_rank[(uint32_t)(hv >> 32)] >= R->name_ub; }
					const uint32_t km = __ballot_sync(0xffffffffu, keep);
					if (keep) {
						anchor_t a;
						a.x = (hv & 0x7fffffff80000000ULL) | (uint64_t)((uint32_t)(hv >> 1) & 0x7fffffffu) | ((hv & 1) << 63);
						a.y = y;
						dst[w + __popc(km & ((1u << lane) - 1))] = a;
					}
					w += __popc(km);
				}
			}
		}
	}
	/* previous chunk's chain anchors go last (rmap.cpp:111-116) */
	const uint32_t pn = R->prev_n;
	const anchor_t *pv = A.carry_in + R->prev_off;
	for (uint32_t k = lane; k < pn; k += RH_WARP) out[n_new + k] = pv[k];
}

struct k3_args_t {
	slot_t *slots; uint32_t n_slots;
	read_state_t *rs;
	uint8_t *arena;
	anchor_t *carry_out; unsigned long long *carry_top; uint64_t carry_cap;
	const float *logf_tab; uint32_t logf_n;
	rh_map_rec_t *recs; unsigned long long *rec_top; uint64_t rec_cap;
	uint32_t *rec_start, *rec_cnt;   /* per read */
	const uint32_t *seq_len;
	int tap;                         /* tap mode: never stop early */
	uint32_t *err;
	unsigned long long *prof;
	int prof_replay;                 /* RH_PROF: 0 keeps k_chain_finish's score sort out of the replay counters [16..29] (RH_PROF_TIES_ONLY) */
};

/* mg_lchain_dp main loop, reference src/lchain.c:439-505.
 *
 * The recurrence only looks back over anchors of the same strand/target within max_dist_t, and
 * both pieces of loop-carried state (`st`, `max_ii`) re-derive themselves from scratch whenever
 * the previous anchor is on another target or more than max_dist_t behind.  So the sorted anchor
 * array falls apart into independent DP segments at every such gap: one CTA per chunk marks the
 * gaps, then its threads take segments round-robin.  (Random hits are sparse in target space, so
 * a chunk has thousands of short segments and a few long ones around true loci.) */
#define DP_THREADS 256
#define DP_BIG_SEG 64          /* segments at least this long are chained by a whole warp */
__global__ void __launch_bounds__(DP_THREADS) k_chain_dp(k3_args_t A, dev_params_t P)
{
	__shared__ uint32_t s_nseg, s_nbig, s_next, s_nlink;
	slot_t *S = &A.slots[blockIdx.x];
	if (S->gated || S->n_anchors == 0) return;
	const int32_t n = (int32_t)S->n_anchors;
	slot_mem_t M = slot_mem(A.arena, S->a_off, S->n_anchors);
	const anchor_t *a = M.A;
	int32_t *f = M.f, *p = M.p, *v = M.v, *t = M.t;
	uint32_t *starts = (uint32_t *)M.U;
	uint8_t *is_start = (uint8_t *)M.U2;
	int32_t max_t = P.max_t, max_q = P.max_q;
	const int32_t bw = P.bw;
	if (max_t < bw) max_t = bw;
	if (max_q < bw) max_q = bw;
	const uint32_t tid = threadIdx.x;
	uint32_t *big = starts + n; /* segments handed to whole warps (M.U holds 2n words) */
	if (tid == 0) { s_nseg = 0; s_nbig = 0; s_next = 0; s_nlink = 0; }
	__syncthreads();
	uint32_t my_links = 0;
	for (int32_t i = tid; i < n; i += DP_THREADS) {
		t[i] = 0;
		bool st = true;
		if (i > 0) { const uint64_t x = a[i].x, px = a[i - 1].x; st = (x >> 32) != (px >> 32) || x > px + (uint64_t)max_t; }
		is_start[i] = st ? 1 : 0;
		if (st) starts[atomicAdd(&s_nseg, 1u)] = (uint32_t)i;
	}
	__syncthreads();
	const uint32_t n_seg = s_nseg;
	if (tid == 0) S->n_seg = n_seg;
	RH_PROF_BEGIN(A.prof);
	for (uint32_t sg = atomicAdd(&s_next, 1u); sg < n_seg; sg = atomicAdd(&s_next, 1u)) { /* segments differ wildly in cost: hand them out on demand */
		const int32_t i0 = (int32_t)starts[sg];
		{ /* long segments (a true locus: hundreds of anchors, each with a long predecessor window) go to a whole warp */
			int32_t e = i0 + 1;
			while (e < n && e - i0 < DP_BIG_SEG && !is_start[e]) ++e;
			if (e - i0 >= DP_BIG_SEG) { big[atomicAdd(&s_nbig, 1u)] = (uint32_t)i0; continue; }
		}
		int32_t st = i0, band_best = -1;
		for (int32_t i = i0; i < n && (i == i0 || !is_start[i]); ++i) {
			const uint64_t ix = a[i].x, iy = a[i].y;
			int32_t best = (int32_t)((iy >> 32) & 63), best_j = -1, skipped = 0, j;
			while (st < i && ((ix >> 32) != (a[st].x >> 32) || ix > a[st].x + (uint64_t)max_t)) ++st;
			if (i - st > P.max_iter) st = i - P.max_iter;
			for (j = i - 1; j >= st; --j) {
				int32_t sc = pair_score(ix, iy, a[j].x, a[j].y, max_t, max_q, bw, P.pen_gap, P.pen_skip);
				if (sc == INT32_MIN) continue;
				sc += f[j];
				if (sc > best) { best = sc; best_j = j; if (skipped > 0) --skipped; }
				else if (t[j] == i) { if (++skipped > P.max_skip) break; }
				if (p[j] >= 0) t[p[j]] = i;
			}
			const int32_t end_j = j;
			if (band_best < 0 || ix - a[band_best].x > (uint64_t)(int64_t)max_t) {
				int32_t mx = INT32_MIN; band_best = -1;
				for (j = i - 1; j >= st; --j) if (mx < f[j]) { mx = f[j]; band_best = j; }
			}
			if (band_best >= 0 && band_best < end_j) {
				const int32_t sc = pair_score(ix, iy, a[band_best].x, a[band_best].y, max_t, max_q, bw, P.pen_gap, P.pen_skip);
				if (sc != INT32_MIN && best < sc + f[band_best]) { best = sc + f[band_best]; best_j = band_best; }
			}
			f[i] = best; p[i] = best_j;
			my_links += best_j >= 0;
			v[i] = (best_j >= 0 && v[best_j] > best) ? v[best_j] : best;
			if (band_best < 0 || (ix - a[band_best].x <= (uint64_t)(int64_t)max_t && f[band_best] < f[i])) band_best = i;
		}
	}
	RH_PROF_MARK(A.prof, 8, true);
	__syncthreads();
	/* ---- long segments: one warp each, lanes take 32 predecessors at a time.  The order-dependent part of the scan
	 *      (running maximum with "first one wins", the skip counter with its early exit: lchain.c:454-470) is resolved
	 *      from two ballots per tile, identically on every lane. ---- */
	const uint32_t n_big = s_nbig;
	const uint32_t FULL = 0xffffffffu;
	const uint32_t lane = tid & 31, warp = tid >> 5;
	for (uint32_t bg = warp; bg < n_big; bg += DP_THREADS / 32) {
		const int32_t i0 = (int32_t)big[bg];
		int32_t st = i0, band_best = -1;
		for (int32_t i = i0; i < n && (i == i0 || !is_start[i]); ++i) {
			const uint64_t ix = a[i].x, iy = a[i].y;
			int32_t best = (int32_t)((iy >> 32) & 63), best_j = -1, skipped = 0;
			while (st < i && ((ix >> 32) != (a[st].x >> 32) || ix > a[st].x + (uint64_t)max_t)) ++st;
			if (i - st > P.max_iter) st = i - P.max_iter;
			int32_t end_j = st - 1;
			bool broke = false;
			for (int32_t jt = i - 1; jt >= st && !broke; jt -= 32) {
				const int32_t j = jt - (int32_t)lane;
				int32_t sc = INT32_MIN;
				if (j >= st) {
					const int32_t s0 = pair_score(ix, iy, a[j].x, a[j].y, max_t, max_q, bw, P.pen_gap, P.pen_skip);
					if (s0 != INT32_MIN) { sc = s0 + f[j]; const int32_t pj = p[j]; if (pj >= 0) t[pj] = i; } /* marks beyond an early exit are never read */
				}
				__syncwarp();
				const bool marked = sc != INT32_MIN && t[j] == i; /* only an earlier (larger) j can have marked j */
				int32_t incl = sc;
#pragma unroll
				for (int o = 1; o < 32; o <<= 1) { const int32_t u = __shfl_up_sync(FULL, incl, o); if ((int)lane >= o && u > incl) incl = u; }
				int32_t before = __shfl_up_sync(FULL, incl, 1);
				if (lane == 0) before = INT32_MIN;
				if (best > before) before = best;
				const bool improve = sc != INT32_MIN && sc > before;
				const uint32_t im = __ballot_sync(FULL, improve), sm = __ballot_sync(FULL, sc != INT32_MIN && !improve && marked);
				uint32_t ev = im | sm; int brk = -1;
				while (ev) {
					const int l = __ffs(ev) - 1; ev &= ev - 1;
					if ((im >> l) & 1u) { if (skipped > 0) --skipped; }
					else if (++skipped > P.max_skip) { brk = l; break; }
				}
				const uint32_t imv = brk < 0 ? im : (im & ((1u << brk) - 1u));
				if (imv) { const int last = 31 - __clz(imv); best = __shfl_sync(FULL, sc, last); best_j = jt - last; }
				if (brk >= 0) { broke = true; end_j = jt - brk; }
			}
			if (band_best < 0 || ix - a[band_best].x > (uint64_t)(int64_t)max_t) { /* largest f in the window, the largest index among equals */
				int32_t mx = INT32_MIN; band_best = -1;
				for (int32_t jt = i - 1; jt >= st; jt -= 32) {
					const int32_t j = jt - (int32_t)lane;
					const int32_t fj = j >= st ? f[j] : INT32_MIN;
					int32_t m = fj;
#pragma unroll
					for (int o = 16; o > 0; o >>= 1) { const int32_t u = __shfl_xor_sync(FULL, m, o); if (u > m) m = u; }
					const uint32_t who = __ballot_sync(FULL, j >= st && fj == m);
					if (who && mx < m) { mx = m; band_best = jt - (__ffs(who) - 1); }
				}
			}
			if (band_best >= 0 && band_best < end_j) {
				const int32_t sc = pair_score(ix, iy, a[band_best].x, a[band_best].y, max_t, max_q, bw, P.pen_gap, P.pen_skip);
				if (sc != INT32_MIN && best < sc + f[band_best]) { best = sc + f[band_best]; best_j = band_best; }
			}
			const int32_t vi = (best_j >= 0 && v[best_j] > best) ? v[best_j] : best;
			if (band_best < 0 || (ix - a[band_best].x <= (uint64_t)(int64_t)max_t && f[band_best] < best)) band_best = i;
			if (lane == 0) { f[i] = best; p[i] = best_j; v[i] = vi; my_links += best_j >= 0; }
			__syncwarp();
		}
	}
	if (my_links) atomicAdd(&s_nlink, my_links);
	__syncthreads();
	if (tid == 0) S->n_link = s_nlink;
}

#endif
