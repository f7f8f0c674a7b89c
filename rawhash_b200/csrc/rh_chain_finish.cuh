/*
 * rh_chain_finish.cuh — everything after the chaining DP:
 *
 *   k_chain_finish (one CTA per chunk)
 *     mg_chain_backtrack + compact_a          reference src/lchain.c:95-281
 *     mm_gen_regs                             reference src/hit.c:100-150
 *   k_chain_decide (one warp per chunk)
 *     mm_set_parent / mm_select_sub / mm_sync_regs / mm_set_mapq
 *                                             reference src/hit.c:195-263,312-367,502-539
 *     stop rules + final record of map_worker_for   reference src/rmap.cpp:423-586
 *
 * The O(n) passes (candidate collection, gathers, radix passes, scans) run on the whole CTA.  The steps whose
 * result depends on visiting order are cut as small as they can be: the score sort replays klib's displacement
 * walk (cta_klib_replay, rh_anchor_sort.cuh); the backtrack runs one thread per DP segment, because chains never
 * leave their segment and only the relative order of a segment's own candidates matters; the sorts on keys that
 * are unique in practice (chain start, salted region key) use a parallel radix sort and fall back to the exact
 * replay if two equal keys do show up; mm_set_parent is a sequential sweep over regions in score order with the
 * covered query positions kept as a bitset in shared memory.
 */
#ifndef RH_CHAIN_FINISH_CUH
#define RH_CHAIN_FINISH_CUH

#include "rh_kernels.cuh"
#include "rh_anchor_sort.cuh"

#define FIN_THREADS TIE_THREADS
#define FIN_SHORT_RUN 12          /* runs of up to this many candidates are ordered by their own thread */
#define FIN_BYTES_CAP TIE_RING_BYTES /* shared-memory bytes: digit bytes of short sorts, or the rings of cta_big_level */

__device__ __forceinline__ float logf_tab(const k3_args_t &A, int32_t x, uint32_t *flag)
{ /* glibc logf of an integer argument, tabulated on the host so MAPQ is bit-identical */
	if (x >= 0 && (uint32_t)x < A.logf_n) return A.logf_tab[x];
	*flag = 1;
	return logf((float)x);
}


#define FIN_BITS 8192             /* query (event) coordinates covered by the shared-memory bitset */
#define FIN_PCAP 256              /* primaries cached in shared memory                              */
struct prim_cache_t { int qs[FIN_PCAP], qe[FIN_PCAP], idx[FIN_PCAP], subsc[FIN_PCAP], nsub[FIN_PCAP], cnt[FIN_PCAP]; };

/* mm_set_parent (hit.c:195-263), general form: overlaps clipped and sorted, as the reference does. */
__device__ void set_parent_general(dev_reg_t *r, uint32_t n_regs, const dev_params_t &P, const slot_mem_t &M, uint32_t lane)
{
	const uint32_t FULL = 0xffffffffu;
	int *wl = (int *)M.p; uint64_t *cov = M.U2;
	if (lane == 0) { wl[0] = 0; r[0].parent = 0; }
	__syncwarp();
	int kk = 1;
	for (int i = 1; i < (int)n_regs; ++i) {
		const int si = r[i].qs, ei = r[i].qe;
		int n_cov = 0;
		for (int j0 = 0; j0 < kk; j0 += 32) {
			const int j = j0 + (int)lane;
			bool ov = false; int sj = 0, ej = 0;
			if (j < kk) { const dev_reg_t *rp = &r[wl[j]]; sj = rp->qs; ej = rp->qe; ov = !(ej <= si || sj >= ei); }
			const uint32_t m = __ballot_sync(FULL, ov);
			if (ov) { if (sj < si) sj = si; if (ej > ei) ej = ei; cov[n_cov + __popc(m & lanemask_lt())] = (uint64_t)sj << 32 | (uint32_t)ej; }
			n_cov += __popc(m);
		}
		__syncwarp();
		int hit = -1, uncov = 0;
		if (n_cov > 0) {
			if (lane == 0) {
				int x = si;
				seq_klib_sort(cov, (uint32_t)n_cov, key_of_u64(), (sort_seg_t *)M.B);
				for (int c = 0; c < n_cov; ++c) {
					if ((int)(cov[c] >> 32) > x) uncov += (int)(cov[c] >> 32) - x;
					x = (int32_t)cov[c] > x ? (int32_t)cov[c] : x;
				}
				if (ei > x) uncov += ei - x;
			}
			uncov = __shfl_sync(FULL, uncov, 0);
			for (int j0 = 0; j0 < kk && hit < 0; j0 += 32) {
				const int j = j0 + (int)lane;
				bool sec = false;
				if (j < kk) {
					const dev_reg_t *rp = &r[wl[j]];
					const int sj = rp->qs, ej = rp->qe;
					if (!(ej <= si || sj >= ei)) {
						const int mn = ej - sj < ei - si ? ej - sj : ei - si;
						const int mx = ej - sj > ei - si ? ej - sj : ei - si;
						const int ol = si < sj ? (ei < sj ? 0 : ei < ej ? ei - sj : ej - sj) : (ej < si ? 0 : ej < ei ? ej - si : ei - si);
						sec = __fsub_rn(__fdiv_rn((float)ol, (float)mn), __fdiv_rn((float)uncov, (float)mx)) > P.mask_level && uncov <= P.mask_len;
					}
				}
				const uint32_t m = __ballot_sync(FULL, sec);
				if (m) hit = j0 + __ffs(m) - 1;
			}
		}
		if (lane == 0) {
			dev_reg_t *ri = &r[i];
			if (hit >= 0) {
				dev_reg_t *rp = &r[wl[hit]];
				ri->parent = rp->parent;
				rp->subsc = rp->subsc > ri->score ? rp->subsc : ri->score;
				if (ri->cnt >= rp->cnt) ++rp->n_sub;
			} else { wl[kk] = i; ri->parent = i; ri->n_sub = 0; }
		}
		if (hit < 0) ++kk;
		__syncwarp();
	}
}

/* Same function for the common case (query coordinates < FIN_BITS, few primaries): the union of
 * the primaries' query intervals is a bitset in shared memory, so the uncovered length of region
 * i is a popcount, and the primaries' fields live in shared memory.  The uncovered length equals
 * the reference's sorted-interval sweep because all coordinates are integers.
 *
 * Regions are taken 32 at a time, one per lane.  The sweep is sequential only through the set of
 * primaries, and new primaries are rare (a chunk has 10^4..10^5 regions and a handful of primaries):
 * every lane evaluates its region against the current set; the lowest lane that turns out to be a
 * NEW primary ends the batch's valid prefix — the lanes below it saw exactly the set the sequential
 * sweep would have shown them and are committed (subsc is a max, n_sub a count: order free), the new
 * primary is added, and the lanes above it are evaluated again. */
__device__ bool set_parent_bitset(dev_reg_t *r, uint32_t n_regs, const dev_params_t &P, uint32_t *bits, prim_cache_t *pc, uint32_t lane)
{
	const uint32_t FULL = 0xffffffffu;
	for (uint32_t w = lane; w < FIN_BITS / 32; w += 32) bits[w] = 0;
	__syncwarp();
	int kk = 0;
	for (uint32_t i0 = 0; i0 < n_regs; i0 += 32) {
		const uint32_t ii = i0 + lane;
		int si = 0, ei = 0, sci = 0, cni = 0;
		if (ii < n_regs) { const dev_reg_t g = r[ii]; si = g.qs; ei = g.qe; sci = g.score; cni = g.cnt; }
		uint32_t pending = __ballot_sync(FULL, ii < n_regs);
		while (pending) {
			int hit = -1; bool is_new = false;
			if ((pending >> lane) & 1u) {
				int covered = 0;
				if (kk > 0 && ei > si) {
					const int w0 = si >> 5, w1 = (ei - 1) >> 5;
					for (int w = w0; w <= w1; ++w) {
						uint32_t m = bits[w];
						if (w == w0) m &= 0xffffffffu << (si & 31);
						if (w == w1) m &= 0xffffffffu >> (31 - ((ei - 1) & 31));
						covered += __popc(m);
					}
				}
				if (covered > 0) {
					const int uncov = (ei - si) - covered;
					for (int j = 0; j < kk; ++j) {
						const int sj = pc->qs[j], ej = pc->qe[j];
						if (ej <= si || sj >= ei) continue;
						const int mn = ej - sj < ei - si ? ej - sj : ei - si;
						const int mx = ej - sj > ei - si ? ej - sj : ei - si;
						const int ol = si < sj ? (ei < sj ? 0 : ei < ej ? ei - sj : ej - sj) : (ej < si ? 0 : ej < ei ? ej - si : ei - si);
						if (__fsub_rn(__fdiv_rn((float)ol, (float)mn), __fdiv_rn((float)uncov, (float)mx)) > P.mask_level && uncov <= P.mask_len) { hit = j; break; }
					}
				}
				is_new = hit < 0;
			}
			const uint32_t newm = __ballot_sync(FULL, is_new);
			const int first_new = newm ? __ffs(newm) - 1 : 32;
			const uint32_t commit = first_new < 32 ? (pending & ((1u << first_new) - 1u)) : pending;
			if ((commit >> lane) & 1u) { /* secondary of primary `hit` */
				r[ii].parent = pc->idx[hit]; /* a primary's parent is itself */
				atomicMax(&pc->subsc[hit], sci);
				if (cni >= pc->cnt[hit]) atomicAdd(&pc->nsub[hit], 1);
			}
			pending &= ~commit;
			if (first_new < 32) {
				const int ps = __shfl_sync(FULL, si, first_new), pe = __shfl_sync(FULL, ei, first_new);
				if ((int)lane == first_new) { pc->qs[kk] = si; pc->qe[kk] = ei; pc->idx[kk] = (int)ii; pc->subsc[kk] = 0; pc->nsub[kk] = 0; pc->cnt[kk] = cni; r[ii].parent = (int)ii; }
				if (pe > ps) {
					const int w0 = ps >> 5, w1 = (pe - 1) >> 5;
					for (int w = w0 + (int)lane; w <= w1; w += 32) {
						uint32_t m = 0xffffffffu;
						if (w == w0) m &= 0xffffffffu << (ps & 31);
						if (w == w1) m &= 0xffffffffu >> (31 - ((pe - 1) & 31));
						bits[w] |= m;
					}
				}
				++kk;
				pending &= ~(1u << first_new);
			}
			__syncwarp();
			if (kk >= FIN_PCAP) return false; /* primary cache full: caller redoes the chunk with the general form */
		}
	}
	__syncwarp();
	/* write the primaries' accumulated fields back */
	for (int j = (int)lane; j < kk; j += 32) { dev_reg_t *g = &r[pc->idx[j]]; g->subsc = pc->subsc[j]; g->n_sub = pc->nsub[j]; }
	__syncwarp();
	return true;
}

/* inclusive scan of one value per thread over the CTA's current tile; *total = tile sum */
__device__ __forceinline__ uint32_t fin_tile_scan(uint32_t v, uint32_t *wsum, uint32_t *total)
{
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t incl = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
	if (lane == 31) wsum[warp] = incl;
	__syncthreads();
	uint32_t base = 0, tot = 0;
#pragma unroll
	for (uint32_t w = 0; w < TIE_WARPS; ++w) { const uint32_t c = wsum[w]; if (w < warp) base += c; tot += c; }
	__syncthreads();
	*total = tot;
	return base + incl;
}

/* exact klib sort of m (key, payload) pairs already placed in W.xk/W.ord; W.sidx[pos] = payload at sorted position pos */
__device__ __forceinline__ void fin_sort(tie_shared_t &T, uint8_t *s_bytes, uint8_t *g_bytes, const klib_ws_t &W, uint32_t m, unsigned long long *prof)
{
	if (m > 64) { cta_klib_replay<KLIB_ALL>(T, m < TIE_BIG_MIN ? s_bytes : g_bytes, W, m, prof); return; } /* from TIE_BIG_MIN on, s_bytes is cta_big_level's ring */
	__syncthreads();
	if (threadIdx.x < 32 && m > 0) { /* klib: insertion sort (stable) */
		const uint32_t lane = threadIdx.x;
		uint64_t k0 = 0, k1 = 0; uint32_t o0 = 0, o1 = 0;
		if (lane < m) { k0 = W.xk[lane]; o0 = W.ord[lane]; }
		if (32 + lane < m) { k1 = W.xk[32 + lane]; o1 = W.ord[32 + lane]; }
		uint32_t r0, r1;
		warp_rank64(k0, k1, m, lane, &r0, &r1);
		if (lane < m) W.sidx[r0] = o0;
		if (32 + lane < m) W.sidx[r1] = o1;
	}
	__syncthreads();
}

/* One stable counting-sort pass of (key, payload) arrays on digit (key >> shift) & 255 by the whole CTA.
 * Scratch: tie_shared_t::tab rows 0-5 (free whenever no klib replay is running). */
template <class KeyT>
__device__ void fin_radix_pass(tie_shared_t &T, const KeyT *__restrict__ kin, const uint32_t *__restrict__ vin, KeyT *__restrict__ kout, uint32_t *__restrict__ vout,
                               uint32_t n, uint32_t shift)
{
	uint32_t *hist = T.tab[0], *base = T.tab[1];
	uint32_t (*wcnt)[256] = (uint32_t (*)[256])T.tab[2];
	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t FULL = 0xffffffffu;
	for (uint32_t k = tid; k < 256; k += FIN_THREADS) hist[k] = 0;
	__syncthreads();
	for (uint32_t t0 = 0; t0 < n; t0 += FIN_THREADS) {
		const uint32_t i = t0 + tid;
		const bool ok = i < n;
		const uint32_t act = __ballot_sync(FULL, ok);
		if (ok) {
			const uint32_t d = (uint32_t)(kin[i] >> shift) & 255u;
			const uint32_t peers = digit_peers(act, d);
			if ((peers & lanemask_lt()) == 0) atomicAdd(&hist[d], (uint32_t)__popc(peers));
		}
	}
	__syncthreads();
	if (warp == 0) {
		uint32_t c[8], tot = 0;
#pragma unroll
		for (int q = 0; q < 8; ++q) { c[q] = hist[lane * 8 + q]; tot += c[q]; }
		uint32_t incl = tot;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += u; }
		uint32_t start = incl - tot;
#pragma unroll
		for (int q = 0; q < 8; ++q) { base[lane * 8 + q] = start; start += c[q]; }
	}
	__syncthreads();
	for (uint32_t t0 = 0; t0 < n; t0 += FIN_THREADS) {
		for (uint32_t k = tid; k < TIE_WARPS * 256; k += FIN_THREADS) (&wcnt[0][0])[k] = 0;
		__syncthreads();
		const uint32_t i = t0 + tid;
		const bool ok = i < n;
		KeyT key = 0; uint32_t v = 0, d = 0, peers = 0;
		const uint32_t act = __ballot_sync(FULL, ok);
		if (ok) {
			key = kin[i]; v = vin[i]; d = (uint32_t)(key >> shift) & 255u;
			peers = digit_peers(act, d);
			if ((peers & lanemask_lt()) == 0) wcnt[warp][d] = __popc(peers);
		}
		__syncthreads();
		for (uint32_t d2 = tid; d2 < 256; d2 += FIN_THREADS) {
			uint32_t run = base[d2];
#pragma unroll
			for (int w = 0; w < TIE_WARPS; ++w) { const uint32_t c = wcnt[w][d2]; wcnt[w][d2] = run; run += c; }
			base[d2] = run;
		}
		__syncthreads();
		if (ok) { const uint32_t o = wcnt[warp][d] + __popc(peers & lanemask_lt()); kout[o] = key; vout[o] = v; }
		__syncthreads();
	}
}

/* Sort of m (key, payload = position) pairs in W.xk/W.ord whose keys are expected to be unique: every correct
 * sort then equals klib's, so a parallel LSD radix sort is used; if two equal keys do turn up, klib's exact
 * order is replayed instead.  W.sidx[pos] = payload at sorted position pos. */
__device__ void fin_sort_unique(tie_shared_t &T, uint8_t *s_bytes, uint8_t *g_bytes, const klib_ws_t &W, uint32_t m, uint32_t *wsum, unsigned long long *prof = nullptr, int prof_slot = 0)
{
	if (m <= 64) { fin_sort(T, s_bytes, g_bytes, W, m, nullptr); return; }
	const uint32_t tid = threadIdx.x, lane = tid & 31;
	const uint32_t FULL = 0xffffffffu;
	uint64_t *backup = (uint64_t *)W.zlist; /* 8 m <= 4 n bytes */
	__syncthreads();
	if (tid == 0) T.diff = 0ULL;
	__syncthreads();
	{
		const uint64_t x0 = W.xk[0];
		unsigned long long diff = 0;
		for (uint32_t i = tid; i < m; i += FIN_THREADS) { const uint64_t x = W.xk[i]; backup[i] = x; diff |= x ^ x0; }
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) diff |= __shfl_xor_sync(FULL, diff, o);
		if (lane == 0 && diff) atomicOr(&T.diff, diff);
	}
	__syncthreads();
	const unsigned long long dmask = T.diff;
	uint64_t *ka = W.xk, *kb = W.xk2; uint32_t *va = W.ord, *vb = W.ord2;
	for (uint32_t sh = 0; sh < 64; sh += 8) {
		if (((dmask >> sh) & 255ULL) == 0) continue;
		fin_radix_pass<uint64_t>(T, ka, va, kb, vb, m, sh);
		{ uint64_t *t = ka; ka = kb; kb = t; } { uint32_t *t = va; va = vb; vb = t; }
	}
	bool tie = false;
	for (uint32_t i = tid; i + 1 < m; i += FIN_THREADS) tie |= ka[i] == ka[i + 1];
	const int any_tie = __syncthreads_or(tie ? 1 : 0);
	if (!any_tie) {
		for (uint32_t i = tid; i < m; i += FIN_THREADS) W.sidx[i] = va[i];
		__syncthreads();
		return;
	}
	/* equal keys: every position outside a group of equal keys is final in the stable order; klib's order inside the groups is
	 * replayed along the sub-arrays that hold them only (the flag rides in bit 31 of the payload = original position) */
	uint32_t *flag = W.dst; /* free until the replay's first level */
	for (uint32_t i = tid; i < m; i += FIN_THREADS) { W.sidx[i] = va[i]; flag[i] = 0; }
	__syncthreads();
	for (uint32_t i = tid; i + 1 < m; i += FIN_THREADS) if (ka[i] == ka[i + 1]) { flag[va[i]] = 1; flag[va[i + 1]] = 1; }
	__syncthreads();
	for (uint32_t i = tid; i < m; i += FIN_THREADS) { W.xk[i] = backup[i]; W.ord[i] = i | (flag[i] ? 0x80000000u : 0u); }
	__syncthreads();
	if (prof && tid == 0) atomicAdd(&prof[prof_slot], 1ULL);
	cta_klib_replay<KLIB_TIES_PAYFLAG>(T, m < TIE_BIG_MIN ? s_bytes : g_bytes, W, m, nullptr);
}

/* one backtrack attempt from candidate anchor i0 (mg_chain_backtrack body, lchain.c:162-181 + mg_chain_bk_end);
 * returns score<<32 | n_anchors of an accepted chain, 0 otherwise */
__device__ __forceinline__ uint64_t fin_backtrack_one(int32_t i0, const int32_t *__restrict__ f, const int32_t *__restrict__ p, int32_t *t,
                                                      int32_t min_sc, int32_t min_cnt, int32_t max_drop)
{
	if (t[i0] != 0) return 0;
	const int32_t zs = f[i0];
	int32_t end_i = -1, max_i = i0, c = i0, max_s = 0;
	do {
		t[c] = 2;
		end_i = c = p[c];
		const int32_t s = c < 0 ? zs : zs - f[c];
		if (s > max_s) { max_s = s; max_i = c; }
		else if (max_s - s > max_drop) break;
	} while (c >= 0 && t[c] == 0);
	for (c = i0; c >= 0 && c != end_i; c = p[c]) t[c] = 0;
	uint32_t cnt = 0;
	for (c = i0; c != max_i; c = p[c]) { ++cnt; t[c] = 1; }
	const int32_t sc = c < 0 ? zs : zs - f[c];
	if (sc >= min_sc && cnt > 0 && (int32_t)cnt >= min_cnt) return (uint64_t)(uint32_t)sc << 32 | cnt;
	return 0; /* rejected: its anchors stay marked, as in the reference */
}

struct fin_shared_t {
	tie_shared_t T;
	big_tab_t big;
	__align__(16) uint8_t bytes[FIN_BYTES_CAP];
	uint32_t wsum[TIE_WARPS], wsum2[TIE_WARPS];
	unsigned long long carry_off;
};

/* One CTA per chunk.  Phases (all CTA-parallel unless noted):
 *   1  candidates z = {i : f[i] >= min_sc} in index order, and the runs of candidates that share a DP segment
 *   2  exact klib sort of z by score (ties are the norm: SURVEY H1b)
 *   3  backtrack: chains never leave their DP segment, so every run is backtracked by its own thread, visiting
 *      its candidates in the global sorted order; accepted chains are then numbered in that global order
 *   4  compact_a (+ the copy carried to the next chunk), exact sort of chains by target, region keys
 *   5  exact sort of regions, mm_gen_regs
 * The order-dependent remainder runs in k_chain_decide. */
__global__ void __launch_bounds__(FIN_THREADS, 5) k_chain_finish(k3_args_t A, dev_params_t P)
{
	__shared__ fin_shared_t SH;
	const uint32_t FULL = 0xffffffffu;
	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t slot_id = blockIdx.x;
	if (slot_id >= A.n_slots) return;
	slot_t *S = &A.slots[slot_id];
	read_state_t *R = &A.rs[S->read];

	uint32_t n_u = 0, n_v = 0, n_regs = 0;
	dev_reg_t *r = nullptr;
	const int32_t n = (int32_t)S->n_anchors;

	if (S->gated) { /* chunk skipped by the min_events gate: carried anchors stay for the next chunk */
		const uint32_t pn = R->prev_n;
		if (pn) {
			if (tid == 0) SH.carry_off = atomicAdd(A.carry_top, (unsigned long long)pn);
			__syncthreads();
			const unsigned long long o = SH.carry_off;
			if (o + pn > A.carry_cap) { if (tid == 0) { atomicExch(A.err, 2u); R->prev_n = 0; } }
			else {
				const anchor_t *src = (const anchor_t *)(A.arena + S->a_off) + pn; /* slot_mem::B */
				for (uint32_t k = tid; k < pn; k += FIN_THREADS) A.carry_out[o + k] = src[k];
				if (tid == 0) R->prev_off = o;
			}
		}
	} else {
		if (tid == 0) R->prev_n = 0; /* consumed by collect_seed_hits */
		if (n > 0) {
			const uint32_t un = (uint32_t)n;
			slot_mem_t M = slot_mem(A.arena, S->a_off, un);
			anchor_t *a = M.A, *b = M.B;
			int32_t *f = M.f, *p = M.p, *t = M.t;
			uint64_t *u = M.U;
			const uint8_t *is_start = (const uint8_t *)M.U2;
			const int32_t min_sc = P.min_sc, min_cnt = P.min_cnt, max_drop = P.bw;
			klib_ws_t W;
			W.xk = (uint64_t *)M.Z; W.xk2 = W.xk + un;
			W.ord = (uint32_t *)M.W; W.ord2 = W.ord + un; W.dst = W.ord2 + un; W.sidx = W.dst + un;
			W.zlist = (uint32_t *)M.U2; W.mlist = W.zlist + un; W.term = (uint2 *)M.U2;
			W.wl0 = (uint2 *)M.regs; W.wl1 = W.wl0 + (un / 64 + 2);
			uint8_t *g_bytes = (uint8_t *)(W.wl1 + (un / 64 + 2));
			W.qbytes = (uint8_t *)(((uintptr_t)(g_bytes + un) + 15) & ~(uintptr_t)15); /* 2.25 n + 128 bytes of the 6 n + 256 byte sort scratch so far */
			W.big = &SH.big; W.ring = SH.bytes; W.n_big = 1; W.err = A.err;
			uint32_t *z_idx = (uint32_t *)M.v;
			RH_PROF_BEGIN(A.prof);

			/* ---- 1: candidates in index order; DP segment id of each candidate ---- */
			uint32_t n_z = 0, n_runs = 0;
			{
				uint32_t *zseg = W.dst;
				/* each warp owns a contiguous quarter of the anchors: count, exchange four totals, then place — two
				 * barriers for the whole phase instead of four per 128 anchors */
				const uint32_t piece = (((un + TIE_WARPS - 1) / TIE_WARPS) + 31) & ~31u;
				const uint32_t cb = min(warp * piece, un), ce = min(cb + piece, un);
				uint32_t cz = 0, cs = 0;
				for (uint32_t i0 = cb; i0 < ce; i0 += 32) {
					const uint32_t i = i0 + lane;
					bool ok = false, st = false;
					if (i < ce) { ok = f[i] >= min_sc; st = is_start[i] != 0; t[i] = 0; }
					cz += __popc(__ballot_sync(FULL, ok)); cs += __popc(__ballot_sync(FULL, st));
				}
				if (lane == 0) { SH.wsum[warp] = cz; SH.wsum2[warp] = cs; }
				__syncthreads();
				uint32_t bz = 0, bs = 0;
				for (uint32_t w2 = 0; w2 < TIE_WARPS; ++w2) { if (w2 < warp) { bz += SH.wsum[w2]; bs += SH.wsum2[w2]; } n_z += SH.wsum[w2]; }
				for (uint32_t i0 = cb; i0 < ce; i0 += 32) {
					const uint32_t i = i0 + lane;
					int32_t fi = 0; bool ok = false, st = false;
					if (i < ce) { fi = f[i]; ok = fi >= min_sc; st = is_start[i] != 0; }
					const uint32_t zm = __ballot_sync(FULL, ok), sm = __ballot_sync(FULL, st);
					if (ok) {
						const uint32_t j = bz + __popc(zm & lanemask_lt());
						z_idx[j] = i; zseg[j] = bs + __popc(sm & (lanemask_lt() | (1u << lane))); /* starts up to and including i */
						W.xk[j] = (uint64_t)(int64_t)fi; W.ord[j] = j;
					}
					bz += __popc(zm); bs += __popc(sm);
				}
				__syncthreads();
				uint32_t *runs = (uint32_t *)M.U, *runidx = runs + un; /* M.U is 8n bytes */
				for (uint32_t j0 = 0; j0 < n_z; j0 += FIN_THREADS) {
					const uint32_t j = j0 + tid;
					const bool first = j < n_z && (j == 0 || zseg[j] != zseg[j - 1]);
					uint32_t tot; const uint32_t incl = fin_tile_scan(first ? 1u : 0u, SH.wsum, &tot);
					if (j < n_z) { const uint32_t rr = n_runs + incl - 1; runidx[j] = rr; if (first) runs[rr] = j; }
					n_runs += tot;
				}
				__syncthreads();
			}
			RH_PROF_MARK(A.prof, 32, tid == 0);
			if (A.prof && tid == 0) { atomicAdd(&A.prof[48], (unsigned long long)n_z); atomicAdd(&A.prof[49], (unsigned long long)n_runs); atomicAdd(&A.prof[53], (unsigned long long)un); }
			if (n_z > 0) {
				/* ---- 2: z sorted by score exactly as klib leaves it ---- */
				fin_sort(SH.T, SH.bytes, g_bytes, W, n_z, A.prof_replay ? A.prof : nullptr);
				RH_PROF_MARK(A.prof, 33, tid == 0);
				/* ---- 3: backtrack, best score first inside every run.  A run's candidates are contiguous in index order
				 *      (a DP segment is a contiguous anchor range); what is needed is their order by position in the sorted z.
				 *      inv[] = position of every candidate; short runs (nearly all: 2-3 candidates) pick their next-best candidate
				 *      by a scan of their own few positions, long runs are rank-sorted by a warp first ---- */
				const uint32_t *__restrict__ sidx = W.sidx;
				uint64_t *acc = W.xk2;
				const uint32_t *runs = (const uint32_t *)M.U, *runidx = runs + un;
				uint32_t *inv = (uint32_t *)W.xk, *longs = inv + un, *grouped = W.ord;
				(void)runidx;
				if (tid == 0) SH.T.n_nxt = 0;
				for (uint32_t pos = tid; pos < n_z; pos += FIN_THREADS) { inv[sidx[pos]] = pos; acc[pos] = 0ULL; }
				__syncthreads();
				for (uint32_t rr = tid; rr < n_runs; rr += FIN_THREADS) {
					const uint32_t lo = runs[rr], hi = rr + 1 < n_runs ? runs[rr + 1] : n_z;
					const uint32_t L = hi - lo;
					if (L > FIN_SHORT_RUN) { longs[atomicAdd(&SH.T.n_nxt, 1u)] = rr; continue; }
					/* a lone candidate without a predecessor (most runs against a large index: a single random hit) is a chain of
					 * one anchor, rejected by min_cnt >= 2, and nothing can reach its mark: skip the dependent loads of the attempt */
					if (L == 1 && min_cnt > 1 && p[z_idx[lo]] < 0) continue;
					uint32_t bound = 0xffffffffu;
					for (uint32_t s2 = 0; s2 < L; ++s2) {
						uint32_t best = 0, bj = lo; bool any = false;
						for (uint32_t j = lo; j < hi; ++j) { const uint32_t v = inv[j]; if (v < bound && (!any || v > best)) { best = v; bj = j; any = true; } }
						bound = best;
						const uint64_t res = fin_backtrack_one((int32_t)z_idx[bj], f, p, t, min_sc, min_cnt, max_drop);
						if (res) acc[best] = res;
					}
				}
				__syncthreads();
				{
					const uint32_t n_long = SH.T.n_nxt;
					for (uint32_t q = warp; q < n_long; q += TIE_WARPS) {
						const uint32_t rr = longs[q];
						const uint32_t lo = runs[rr], hi = rr + 1 < n_runs ? runs[rr + 1] : n_z;
						for (uint32_t j = lo + lane; j < hi; j += 32) { /* rank = candidates of the run at a lower position */
							const uint32_t v = inv[j];
							uint32_t rank = 0;
							for (uint32_t j2 = lo; j2 < hi; ++j2) rank += inv[j2] < v;
							grouped[lo + rank] = j;
						}
						__syncwarp();
						if (lane == 0) {
							for (uint32_t k = hi; k-- > lo;) {
								const uint32_t j = grouped[k];
								const uint64_t res = fin_backtrack_one((int32_t)z_idx[j], f, p, t, min_sc, min_cnt, max_drop);
								if (res) acc[inv[j]] = res;
							}
						}
					}
				}
				__syncthreads();
				RH_PROF_MARK(A.prof, 34, tid == 0);
				/* chains numbered in discovery order = descending position in sorted z */
				uint32_t *chain_i0 = (uint32_t *)M.t, *koff = (uint32_t *)M.f;
				for (uint32_t q0 = 0; q0 < n_z; q0 += FIN_THREADS) {
					const uint32_t q = q0 + tid;
					uint64_t av = 0; uint32_t pos = 0;
					if (q < n_z) { pos = n_z - 1 - q; av = acc[pos]; }
					uint32_t ct, vt;
					const uint32_t cr = tie_tile_rank(av != 0, SH.wsum, &ct);
					const uint32_t vincl = fin_tile_scan((uint32_t)av, SH.wsum, &vt);
					if (av) { const uint32_t ci = n_u + cr; u[ci] = av; chain_i0[ci] = z_idx[sidx[pos]]; koff[ci] = n_v + vincl - (uint32_t)av; }
					n_u += ct; n_v += vt;
				}
				__syncthreads();
			}
			if (n_u > fin_regs_cap(un)) { if (tid == 0) atomicExch(A.err, 3u); n_u = 0; n_v = 0; }
			if (n_u > 0) {
				/* ---- 4: compact_a: forward-order gather (= next chunk's prev_anchors), then chains by target ---- */
				if (tid == 0) SH.carry_off = atomicAdd(A.carry_top, (unsigned long long)n_v);
				__syncthreads();
				const unsigned long long co = SH.carry_off;
				const bool carry_ok = co + n_v <= A.carry_cap;
				if (!carry_ok && tid == 0) atomicExch(A.err, 2u);
				const uint32_t *chain_i0 = (const uint32_t *)M.t; const uint32_t *koff = (const uint32_t *)M.f;
				for (uint32_t ci = tid; ci < n_u; ci += FIN_THREADS) { /* one thread per chain: chains are a few anchors long */
					const uint32_t ni = (uint32_t)u[ci], k0 = koff[ci];
					int32_t c = (int32_t)chain_i0[ci];
					anchor_t x; x.x = x.y = 0;
					for (uint32_t q = 0; q < ni; ++q) {
						x = a[c];
						const uint32_t j = k0 + (ni - 1 - q);
						b[j] = x;
						if (carry_ok) A.carry_out[co + j] = x;
						c = p[c];
					}
					W.xk[ci] = x.x; W.ord[ci] = ci; /* the chain's first anchor */
				}
				if (tid == 0 && carry_ok) { R->prev_off = co; R->prev_n = n_v; }
				__syncthreads();
				RH_PROF_MARK(A.prof, 35, tid == 0);
				fin_sort_unique(SH.T, SH.bytes, g_bytes, W, n_u, SH.wsum, A.prof, 50);
				RH_PROF_MARK(A.prof, 36, tid == 0);
				/* output offsets in target order, then copy chains and pre-compute the region keys (mm_gen_regs) */
				uint32_t *kout = (uint32_t *)M.t;      /* chain_i0 is no longer needed */
				uint64_t *u2 = (uint64_t *)W.dst;      /* 8 n_u <= 4 n bytes           */
				uint64_t *keys2 = (uint64_t *)M.v;     /* z_idx is no longer needed; offset 72n is 8-byte aligned for every n */
				const uint32_t rhash = wang32(wang32(R->ev_offset + S->n_events) + wang32(11u)); /* rmap.cpp:346-348 */
				{
					uint32_t run = 0;
					for (uint32_t q0 = 0; q0 < n_u; q0 += FIN_THREADS) {
						const uint32_t pos = q0 + tid;
						uint32_t ci = 0, c = 0; uint64_t uv = 0;
						if (pos < n_u) { ci = W.sidx[pos]; uv = u[ci]; c = (uint32_t)uv; }
						uint32_t tot; const uint32_t incl = fin_tile_scan(c, SH.wsum, &tot);
						if (pos < n_u) {
							const uint32_t k0 = run + incl - c;
							const anchor_t *src = b + koff[ci];
							for (uint32_t q = 0; q < c; ++q) a[k0 + q] = src[q];
							const anchor_t fa = src[0];
							const uint32_t h = (uint32_t)mix64((mix64(fa.x) + mix64(fa.y)) ^ rhash); /* hit.c:120 */
							const uint64_t key = uv ^ h;
							kout[pos] = k0; u2[pos] = uv; keys2[pos] = key;
						}
						run += tot;
					}
				}
				__syncthreads();
				for (uint32_t pos = tid; pos < n_u; pos += FIN_THREADS) { u[pos] = u2[pos]; W.xk[pos] = keys2[pos]; W.ord[pos] = pos; }
				__syncthreads();
				RH_PROF_MARK(A.prof, 37, tid == 0);
				/* ---- 5: mm_gen_regs ---- */
				fin_sort_unique(SH.T, SH.bytes, g_bytes, W, n_u, SH.wsum, A.prof, 51);
				RH_PROF_MARK(A.prof, 38, tid == 0);
				r = (dev_reg_t *)((uint8_t *)M.regs + fin_regs_off(un)); /* after the sort scratch */
				for (uint32_t i = tid; i < n_u; i += FIN_THREADS) { /* descending key */
					const uint32_t pos = W.sidx[n_u - 1 - i];
					const uint64_t key = keys2[pos];
					dev_reg_t g;
					g.id = (int32_t)i; g.parent = -1; g.subsc = 0; g.n_sub = 0; g.mapq = 0;
					g.score = g.score0 = (int32_t)(key >> 32); g.hash = (uint32_t)key;
					g.cnt = (int32_t)(uint32_t)u[pos]; g.as = (int32_t)kout[pos];
					const anchor_t fa = a[g.as], la = a[g.as + g.cnt - 1];
					g.rev = (uint32_t)(fa.x >> 63); g.rid = (int32_t)(fa.x << 1 >> 33);
					g.rs = (int32_t)fa.x; g.re = (int32_t)la.x + 1; g.qs = (int32_t)fa.y; g.qe = (int32_t)la.y + 1;
					r[i] = g;
				}
				n_regs = n_u;
				__syncthreads();
				RH_PROF_MARK(A.prof, 39, tid == 0);
			}
		}
	}
	if (A.prof && tid == 0) { atomicAdd(&A.prof[52], (unsigned long long)n_u); atomicAdd(&A.prof[54], (unsigned long long)n_v); }
	if (tid == 0) { S->n_u = n_u; S->n_v = n_v; S->n_regs = n_regs; }
}

/* The order-dependent remainder, one WARP per chunk at full occupancy: mm_set_parent, mm_select_sub,
 * mm_sync_regs, mm_set_mapq, then the stop rules and the final record of map_worker_for. */
#define DEC_WARPS 4
__global__ void __launch_bounds__(DEC_WARPS * 32, 8) k_chain_decide(k3_args_t A, dev_params_t P)
{
	__shared__ uint32_t s_bits[DEC_WARPS][FIN_BITS / 32];
	__shared__ prim_cache_t s_pc[DEC_WARPS];
	const uint32_t FULL = 0xffffffffu;
	const uint32_t wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t slot_id = blockIdx.x * DEC_WARPS + wib;
	if (slot_id >= A.n_slots) return;
	slot_t *S = &A.slots[slot_id];
	read_state_t *R = &A.rs[S->read];
	uint32_t *bits = s_bits[wib];
	prim_cache_t *pc = &s_pc[wib];
	const uint32_t n_u = S->n_u, n_v = S->n_v;
	uint32_t n_regs = S->n_regs;
	uint32_t inexact = 0;
	dev_reg_t *r = nullptr;
	if (n_regs > 0) {
		slot_mem_t M = slot_mem(A.arena, S->a_off, S->n_anchors);
		r = (dev_reg_t *)((uint8_t *)M.regs + fin_regs_off(S->n_anchors));

		RH_PROF_BEGIN(A.prof);
		/* ---- mm_set_parent: regions in score order ---- */
		int maxq = 0;
		for (uint32_t i = lane; i < n_regs; i += 32) maxq = max(maxq, r[i].qe);
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) maxq = max(maxq, __shfl_xor_sync(FULL, maxq, o));
		bool done_sp = false;
		if (maxq <= FIN_BITS) done_sp = set_parent_bitset(r, n_regs, P, bits, pc, lane);
		if (!done_sp) set_parent_general(r, n_regs, P, M, lane);
		__syncwarp();
		RH_PROF_MARK(A.prof, 40, lane == 0);
		/* ---- mm_select_sub + mm_sync_regs, mm_set_mapq (small; lane 0) ---- */
		if (!P.ava && P.pri_ratio > 0.0f && P.best_n == 0) {
			/* mm_select_sub with best_n = 0 keeps exactly the primaries (hit.c:350-364); mm_sync_regs then
			 * renumbers them, and a primary's parent is itself: an order-preserving compaction. */
			uint32_t kept = 0;
			for (uint32_t i0 = 0; i0 < n_regs; i0 += 32) {
				const uint32_t i = i0 + lane;
				dev_reg_t g; bool keep = false;
				if (i < n_regs) { g = r[i]; keep = g.parent == (int32_t)i; }
				const uint32_t m = __ballot_sync(FULL, keep);
				if (keep) { const uint32_t d = kept + __popc(m & lanemask_lt()); g.id = (int32_t)d; g.parent = (int32_t)d; r[d] = g; }
				kept += __popc(m);
				__syncwarp();
			}
			n_regs = kept;
		}
		if (lane == 0) {
			if (!P.ava && P.pri_ratio > 0.0f && P.best_n != 0) {
				int kept = 0, n2 = 0;
				const int nn = (int)n_regs;
				for (int i = 0; i < nn; ++i) {
					const int pp = r[i].parent;
					bool keep = false;
					if (pp == i) keep = true;
					else if ((float)r[i].score >= __fmul_rn((float)r[pp].score, P.pri_ratio) && n2 < P.best_n) {
						if (!(r[i].qs == r[pp].qs && r[i].qe == r[pp].qe && r[i].rid == r[pp].rid && r[i].rs == r[pp].rs && r[i].re == r[pp].re)) { keep = true; ++n2; }
					} else if (n2 < P.best_n && r[i].score > P.min_strand_sc && r[i].rev != r[pp].rev) { keep = true; ++n2; }
					if (keep) { if (kept != i) r[kept] = r[i]; ++kept; }
				}
				if (kept != nn) {
					int *tmp = (int *)M.v;
					int max_id = -1;
					for (int i = 0; i < kept; ++i) max_id = max_id > r[i].id ? max_id : r[i].id;
					for (int i = 0; i <= max_id; ++i) tmp[i] = -1;
					for (int i = 0; i < kept; ++i) if (r[i].id >= 0) tmp[r[i].id] = i;
					for (int i = 0; i < kept; ++i) {
						dev_reg_t *g = &r[i];
						g->id = i;
						if (g->parent == -2) g->parent = i;
						else if (g->parent >= 0 && g->parent <= max_id && tmp[g->parent] >= 0) g->parent = tmp[g->parent];
						else g->parent = -1;
					}
				}
				n_regs = (uint32_t)kept;
			}
			long long sum_sc = 0;
			for (uint32_t i = 0; i < n_regs; ++i) if (r[i].parent == r[i].id) sum_sc += r[i].score;
			const float uniq = __fdiv_rn((float)sum_sc, (float)(sum_sc + (long long)S->rep_len));
			for (uint32_t i = 0; i < n_regs; ++i) { /* hit.c:519-538 as compiled */
				dev_reg_t *g = &r[i];
				const double s1 = g->score > 100 ? 1.0 : __dmul_rn(0.01, (double)g->score);
				const float pen_s1 = __double2float_rn(__dmul_rn(s1, (double)uniq));
				float pen_cm = g->cnt > 10 ? 1.0f : __fmul_rn(0.1f, (float)g->cnt);
				pen_cm = pen_s1 < pen_cm ? pen_s1 : pen_cm;
				const int subsc = g->subsc > P.min_sc ? g->subsc : P.min_sc;
				const float x = __fdiv_rn((float)subsc, (float)g->score0);
				const float lead = __fmul_rn(__fmul_rn(__fmul_rn(pen_cm, 40.0f), __fsub_rn(1.0f, x)), logf_tab(A, g->score, &inexact));
				int mapq = (int)lead;
				mapq -= (int)__fmaf_rn(logf_tab(A, g->n_sub + 1, &inexact), 4.343f, .499f);
				mapq = mapq > 0 ? mapq : 0;
				g->mapq = mapq < 60 ? mapq : 60;
			}
		}
		n_regs = __shfl_sync(FULL, n_regs, 0);
		RH_PROF_MARK(A.prof, 41, lane == 0);
	}
	if (lane != 0) return;
	if (inexact) atomicExch(A.err, 4u);
	S->n_u = n_u; S->n_v = n_v; S->n_regs = n_regs;
	if (!S->gated) R->ev_offset += S->n_events;
	if (A.tap) return;

	/* ---- stop rules after this chunk (rmap.cpp:423-500) ---- */
	const uint32_t qlen = R->l_sig;
	const uint32_t l_chunk = (P.chunk_size > qlen || P.noadapt) ? qlen : P.chunk_size;
	const uint32_t max_chunk = P.noadapt ? 1u : P.max_num_chunk;
	uint32_t c_count = S->c_count;
	unsigned long long rec_base = 0; uint32_t n_maps = 0;
	auto push_map = [&](uint32_t cid) {
		if (n_maps == 0) rec_base = atomicAdd(A.rec_top, (unsigned long long)(P.ava ? (n_regs ? n_regs : 1u) : 1u));
		if (rec_base + n_maps < A.rec_cap) A.recs[rec_base + n_maps].c_id = cid;
		++n_maps;
	};
	bool stop = false;
	if (n_regs == 1 && (int)r[0].mapq >= P.min_mapq) { push_map(0); stop = true; }
	if (!stop) {
		float meanC = 0.0f, meanQ = 0.0f;
		for (uint32_t i = 0; i < n_regs; ++i) { meanC = __fadd_rn(meanC, (float)r[i].score); meanQ = __fadd_rn(meanQ, (float)r[i].mapq); }
		if (n_regs > 0) { meanC = __fdiv_rn(meanC, (float)n_regs); meanQ = __fdiv_rn(meanQ, (float)n_regs); }
		const uint32_t n_chains = (P.ava || n_regs < 1) ? n_regs : 1u;
		for (uint32_t ic = 0; ic < n_chains; ++ic) {
			float weighted = 0.0f;
			const float bestQ = (float)r[ic].mapq, bestC = (float)r[ic].score;
			if (!P.ava) {
				float r_q = bestQ > 0 ? __fdiv_rn(bestQ, 30.0f) : 0.0f; if (r_q > 1) r_q = 1.0f;
				float r_mq = bestQ > 0 ? __fsub_rn(1.0f, __fdiv_rn(meanQ, bestQ)) : 0.0f; if (r_mq < 0) r_mq = 0.0f;
				float r_mc = bestC > 0 ? __fsub_rn(1.0f, __fdiv_rn(meanC, bestC)) : 0.0f; if (r_mc < 0) r_mc = 0.0f;
				weighted = __fmaf_rn(r_mc, P.w_bestmc, __fmaf_rn(r_q, P.w_bestq, __fmul_rn(P.w_bestmq, r_mq)));
			}
			if (weighted >= P.w_threshold || (P.ava && r[ic].score >= P.min_sc2)) push_map(ic);
		}
		if (n_maps > 0) stop = true;
	}
	bool exhausted = false;
	if (!stop) {
		const uint64_t s_qs_next = (uint64_t)(c_count + 1) * l_chunk;
		++c_count;
		if (!(s_qs_next < qlen && c_count < max_chunk)) { exhausted = true; if (c_count > 0) --c_count; /* rmap.cpp:507 */ }
	}
	if (!stop && !exhausted) return;

	/* ---- final record(s) (rmap.cpp:507-586) ---- */
	R->done = 1;
	const uint32_t offset = R->ev_offset;
	const float scale = offset == 0 ? 0.0f : (P.sample_per_base == 0.0f ? 0.0f :
		__fdiv_rn(__fdiv_rn(__fmul_rn((float)(c_count + 1), (float)l_chunk), (float)offset), P.sample_per_base));
	if (n_maps == 0 && n_regs > 0 && (int)r[0].mapq > P.min_mapq) push_map(0);
	rh_map_rec_t base; memset(&base, 0, sizeof(base));
	base.read_idx = S->read; base.ci = c_count + 1; base.sl = qlen;
	if (n_maps == 0) {
		rec_base = atomicAdd(A.rec_top, 1ULL);
		base.read_length = P.sig_target ? offset : (uint32_t)__fmul_rn(scale, (float)offset);
		if (n_regs >= 1) { base.cm = r[0].cnt; base.nc = (int32_t)n_regs; base.s1 = r[0].score; }
		if (rec_base < A.rec_cap) A.recs[rec_base] = base; else atomicExch(A.err, 5u);
		A.rec_start[S->read] = (uint32_t)rec_base; A.rec_cnt[S->read] = 1;
		return;
	}
	if (rec_base + n_maps > A.rec_cap) { atomicExch(A.err, 5u); A.rec_cnt[S->read] = 0; return; }
	for (uint32_t m = 0; m < n_maps; ++m) {
		const uint32_t cid = A.recs[rec_base + m].c_id;
		const dev_reg_t g = r[cid];
		rh_map_rec_t o = base;
		o.c_id = cid; o.cm = g.cnt; o.nc = (int32_t)n_regs; o.s1 = g.score;
		o.read_length = P.sig_target ? offset : (uint32_t)__fmul_rn(scale, (float)g.qe);
		o.ref_id = (uint32_t)g.rid;
		o.read_start_position = P.sig_target ? (uint32_t)g.qs : (uint32_t)__fmul_rn(scale, (float)g.qs);
		o.read_end_position = P.sig_target ? (uint32_t)g.qe : (uint32_t)__fmul_rn(scale, (float)g.qe);
		o.fragment_start_position = g.rev ? (uint32_t)(A.seq_len[g.rid] + 1 - g.re) : (uint32_t)g.rs;
		o.fragment_length = (uint32_t)(g.re - g.rs + 1);
		o.mapq = (uint8_t)g.mapq; o.rev = g.rev ? 1 : 0; o.mapped = 1;
		A.recs[rec_base + m] = o;
	}
	A.rec_start[S->read] = (uint32_t)rec_base; A.rec_cnt[S->read] = n_maps;
}

#endif
