/*
 * rh_chain_finish.cuh — everything after the chaining DP, one WARP per chunk:
 *
 *   mg_chain_backtrack + compact_a          reference src/lchain.c:95-281
 *   mm_gen_regs / mm_set_parent / mm_select_sub / mm_sync_regs / mm_set_mapq
 *                                           reference src/hit.c:100-150,195-263,312-367,502-539
 *   stop rules + final record of map_worker_for   reference src/rmap.cpp:423-586
 *
 * The O(n) passes (candidate collection, mark reset, gathers, histograms, permutes) and the
 * O(n_u * n_primary) overlap tests of mm_set_parent run on all 32 lanes; only the steps whose
 * result depends on visiting order (the klib sort's displacement walk, the backtrack itself,
 * the parent assignment order) are executed by lane 0.
 */
#ifndef RH_CHAIN_FINISH_CUH
#define RH_CHAIN_FINISH_CUH

#include "rh_kernels.cuh"
#include "rh_anchor_sort.cuh"

#define FIN_WARPS 4

struct fin_scratch_t { uint8_t *bytes; uint32_t *dst; uint2 *wl0, *wl1; };
__device__ __forceinline__ fin_scratch_t fin_scratch(const slot_mem_t &M, uint64_t n)
{
	fin_scratch_t s; uint8_t *b = (uint8_t *)M.regs;
	s.bytes = b; b += (n + 15) & ~15ULL;
	s.dst = (uint32_t *)b; b += 4 * n;
	s.wl0 = (uint2 *)(((uintptr_t)b + 7) & ~(uintptr_t)7); b = (uint8_t *)s.wl0 + 8 * (n / 64 + 2);
	s.wl1 = (uint2 *)b;
	return s;
}

/* Exact klib radix sort (see rh_sort.cuh for the algorithm) of (x = key, y = payload) pairs,
 * cooperative over one warp.  Every lane must call it. */
__device__ void warp_klib_sort_pairs(anchor_t *a, uint32_t n, anchor_t *tmp, const fin_scratch_t &X, uint32_t *cnt, uint32_t *head, uint32_t lane)
{
	const uint32_t FULL = 0xffffffffu;
	if (n <= 64) {
		anchor_t e0, e1; e0.x = e0.y = e1.x = e1.y = 0;
		if (lane < n) e0 = a[lane];
		if (32 + lane < n) e1 = a[32 + lane];
		uint32_t r0, r1;
		warp_rank64(e0.x, e1.x, n, lane, &r0, &r1);
		__syncwarp();
		if (lane < n) a[r0] = e0;
		if (32 + lane < n) a[32 * 0 + r1] = e1;
		__syncwarp();
		return;
	}
	unsigned long long diff = 0;
	const uint64_t k0 = a[0].x;
	for (uint32_t i = lane; i < n; i += 32) diff |= a[i].x ^ k0;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) diff |= __shfl_xor_sync(FULL, diff, o);
	uint2 *wl_cur = X.wl0, *wl_nxt = X.wl1;
	uint32_t n_cur = 1, n_nxt = 0;
	if (lane == 0) wl_cur[0] = make_uint2(0u, n);
	__syncwarp();
	for (int shift = 56; shift >= 0 && n_cur > 0; shift -= 8) {
		if (((diff >> shift) & 255ULL) == 0) continue; /* identity level for every sub-array */
		n_nxt = 0;
		for (uint32_t s = 0; s < n_cur; ++s) {
			const uint2 seg = wl_cur[s];
			const uint32_t beg = seg.x, len = seg.y;
			for (uint32_t b = lane; b < 256; b += 32) cnt[b] = 0;
			__syncwarp();
			for (uint32_t i = lane; i < len; i += 32) {
				const uint32_t b = (uint32_t)(a[beg + i].x >> shift) & 255;
				X.bytes[beg + i] = (uint8_t)b;
				atomicAdd(&cnt[b], 1u);
			}
			__syncwarp();
			if (cnt[X.bytes[beg]] == len) {
				if (shift > 0) { if (lane == 0) wl_nxt[n_nxt] = seg; ++n_nxt; }
				__syncwarp();
				continue;
			}
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
			if (lane == 0) { /* displacement-cycle walk on the byte array */
				uint32_t region_end = 0;
				for (uint32_t k = 0; k < 256; ++k) {
					region_end += cnt[k];
					uint32_t hk = head[k];
					while (hk != region_end) {
						uint32_t e = hk, d = X.bytes[beg + e];
						while (d != k) {
							const uint32_t hd = head[d];
							X.dst[beg + e] = hd; head[d] = hd + 1;
							e = hd; d = X.bytes[beg + e];
						}
						X.dst[beg + e] = hk;
						++hk;
					}
					head[k] = hk;
				}
			}
			__syncwarp();
			for (uint32_t i = lane; i < len; i += 32) tmp[beg + X.dst[beg + i]] = a[beg + i];
			__syncwarp();
			for (uint32_t i = lane; i < len; i += 32) a[beg + i] = tmp[beg + i];
			__syncwarp();
			if (shift > 0) {
				uint32_t acc = 0;
				for (uint32_t bb = 0; bb < 256; bb += 32) {
					const uint32_t c = cnt[bb + lane];
					uint32_t incl = c;
#pragma unroll
					for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += u; }
					const uint32_t start = beg + acc + incl - c;
					acc += __shfl_sync(FULL, incl, 31);
					const bool recurse = c > 64;
					const uint32_t rm = __ballot_sync(FULL, recurse);
					if (recurse) wl_nxt[n_nxt + __popc(rm & lanemask_lt())] = make_uint2(start, c);
					n_nxt += __popc(rm);
					uint32_t tm = __ballot_sync(FULL, !recurse && c > 1); /* small buckets: stable sort, one bucket per pass */
					while (tm) {
						const int src = __ffs(tm) - 1; tm &= tm - 1;
						const uint32_t ts = __shfl_sync(FULL, start, src), tc = __shfl_sync(FULL, c, src);
						anchor_t e0, e1; e0.x = e0.y = e1.x = e1.y = 0;
						if (lane < tc) e0 = a[ts + lane];
						if (32 + lane < tc) e1 = a[ts + 32 + lane];
						uint32_t r0, r1;
						warp_rank64(e0.x, e1.x, tc, lane, &r0, &r1);
						__syncwarp();
						if (lane < tc) a[ts + r0] = e0;
						if (32 + lane < tc) a[ts + r1] = e1;
					}
				}
				__syncwarp();
			}
		}
		uint2 *t = wl_cur; wl_cur = wl_nxt; wl_nxt = t;
		n_cur = n_nxt;
		__syncwarp();
	}
	__syncwarp();
}

__device__ __forceinline__ float logf_tab(const k3_args_t &A, int32_t x, uint32_t *flag)
{ /* glibc logf of an integer argument, tabulated on the host so MAPQ is bit-identical */
	if (x >= 0 && (uint32_t)x < A.logf_n) return A.logf_tab[x];
	*flag = 1;
	return logf((float)x);
}


#define FIN_BITS 8192             /* query (event) coordinates covered by the shared-memory bitset */
#define FIN_PCAP 64               /* primaries cached in shared memory                              */
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
 * i is a popcount, and the primaries' fields live in shared memory instead of behind two
 * dependent global loads.  The uncovered length equals the reference's sorted-interval sweep
 * because all coordinates are integers. */
__device__ bool set_parent_bitset(dev_reg_t *r, uint32_t n_regs, const dev_params_t &P, uint32_t *bits, prim_cache_t *pc, uint32_t lane)
{
	const uint32_t FULL = 0xffffffffu;
	for (uint32_t w = lane; w < FIN_BITS / 32; w += 32) bits[w] = 0;
	__syncwarp();
	int kk = 0;
	for (uint32_t i0 = 0; i0 < n_regs; i0 += 32) { /* 32 regions' fields are fetched at once */
		const uint32_t ii = i0 + lane;
		int my_qs = 0, my_qe = 0, my_score = 0, my_cnt = 0;
		if (ii < n_regs) { const dev_reg_t g = r[ii]; my_qs = g.qs; my_qe = g.qe; my_score = g.score; my_cnt = g.cnt; }
		const uint32_t nb = min(32u, n_regs - i0);
		for (uint32_t b = 0; b < nb; ++b) {
			const int i = (int)(i0 + b);
			const int si = __shfl_sync(FULL, my_qs, b), ei = __shfl_sync(FULL, my_qe, b);
			const int sci = __shfl_sync(FULL, my_score, b), cni = __shfl_sync(FULL, my_cnt, b);
			/* covered positions of [si, ei) */
			int covered = 0;
			if (kk > 0 && ei > si) {
				const int w0 = si >> 5, w1 = (ei - 1) >> 5;
				for (int w = w0 + (int)lane; w <= w1; w += 32) {
					uint32_t m = bits[w];
					if (w == w0) m &= 0xffffffffu << (si & 31);
					if (w == w1) m &= 0xffffffffu >> (31 - ((ei - 1) & 31));
					covered += __popc(m);
				}
#pragma unroll
				for (int o = 16; o > 0; o >>= 1) covered += __shfl_xor_sync(FULL, covered, o);
			}
			int hit = -1;
			if (covered > 0) {
				const int uncov = (ei - si) - covered;
				for (int j0 = 0; j0 < kk && hit < 0; j0 += 32) {
					const int j = j0 + (int)lane;
					bool sec = false;
					if (j < kk) {
						int sj, ej;
						sj = pc->qs[j]; ej = pc->qe[j];
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
			if (hit >= 0) {
				if (lane == 0) {
					r[i].parent = pc->idx[hit]; /* a primary's parent is itself */
					pc->subsc[hit] = pc->subsc[hit] > sci ? pc->subsc[hit] : sci;
					if (cni >= pc->cnt[hit]) ++pc->nsub[hit];
				}
			} else {
				if (lane == 0) { pc->qs[kk] = si; pc->qe[kk] = ei; pc->idx[kk] = i; pc->subsc[kk] = 0; pc->nsub[kk] = 0; pc->cnt[kk] = cni; r[i].parent = i; }
				if (ei > si) {
					const int w0 = si >> 5, w1 = (ei - 1) >> 5;
					for (int w = w0 + (int)lane; w <= w1; w += 32) {
						uint32_t m = 0xffffffffu;
						if (w == w0) m &= 0xffffffffu << (si & 31);
						if (w == w1) m &= 0xffffffffu >> (31 - ((ei - 1) & 31));
						bits[w] |= m;
					}
				}
				++kk;
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

__global__ void __launch_bounds__(FIN_WARPS * 32, 8) k_chain_finish(k3_args_t A, dev_params_t P)
{
	__shared__ uint32_t s_cnt[FIN_WARPS][256];
	__shared__ uint32_t s_head[FIN_WARPS][256];
	__shared__ uint32_t s_bits[FIN_WARPS][FIN_BITS / 32];
	__shared__ prim_cache_t s_pc[FIN_WARPS];
	const uint32_t FULL = 0xffffffffu;
	const uint32_t wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t slot_id = blockIdx.x * FIN_WARPS + wib;
	if (slot_id >= A.n_slots) return;
	slot_t *S = &A.slots[slot_id];
	read_state_t *R = &A.rs[S->read];
	uint32_t *cnt = s_cnt[wib], *head = s_head[wib], *bits = s_bits[wib];
	prim_cache_t *pc = &s_pc[wib];

	uint32_t n_u = 0, n_v = 0, n_regs = 0;
	dev_reg_t *r = nullptr;
	uint32_t inexact = 0;
	const int32_t n = (int32_t)S->n_anchors;

	if (S->gated) { /* chunk skipped by the min_events gate: carried anchors stay for the next chunk */
		const uint32_t pn = R->prev_n;
		if (pn) {
			unsigned long long o = 0;
			if (lane == 0) o = atomicAdd(A.carry_top, (unsigned long long)pn);
			o = __shfl_sync(FULL, o, 0);
			if (o + pn > A.carry_cap) { if (lane == 0) { atomicExch(A.err, 2u); R->prev_n = 0; } }
			else {
				const anchor_t *src = (const anchor_t *)(A.arena + S->a_off) + pn; /* slot_mem::B */
				for (uint32_t k = lane; k < pn; k += 32) A.carry_out[o + k] = src[k];
				if (lane == 0) R->prev_off = o;
			}
		}
	} else {
		if (lane == 0) R->prev_n = 0; /* consumed by collect_seed_hits */
		if (n > 0) {
			slot_mem_t M = slot_mem(A.arena, S->a_off, S->n_anchors);
			const fin_scratch_t X = fin_scratch(M, S->n_anchors);
			anchor_t *a = M.A, *b = M.B, *z = M.Z, *w = M.W;
			int32_t *f = M.f, *p = M.p, *v = M.v, *t = M.t;
			uint64_t *u = M.U, *u2 = M.U2;
			const int32_t min_sc = P.min_sc, min_cnt = P.min_cnt, max_drop = P.bw;
			RH_PROF_BEGIN(A.prof);
			/* ---- candidates z = {(f[i], i) : f[i] >= min_sc}, in index order ---- */
			uint32_t n_z = 0;
			for (int32_t i0 = 0; i0 < n; i0 += 32) {
				const int32_t i = i0 + (int32_t)lane;
				int32_t fi = 0; bool ok = false;
				if (i < n) { fi = f[i]; ok = fi >= min_sc; t[i] = 0; }
				const uint32_t m = __ballot_sync(FULL, ok);
				if (ok) { anchor_t e; e.x = (uint64_t)(int64_t)fi; e.y = (uint64_t)i; z[n_z + __popc(m & lanemask_lt())] = e; }
				n_z += __popc(m);
			}
			__syncwarp();
			if (n_z > 0) {
				RH_PROF_MARK(A.prof, 32, lane == 0);
				warp_klib_sort_pairs(z, n_z, w, X, cnt, head, lane);
				RH_PROF_MARK(A.prof, 33, lane == 0);
				/* ---- backtrack, best score first.  Visiting order matters, so lane 0 walks the chains;
				 *      the test that skips candidates already swallowed by an earlier chain (the vast
				 *      majority) is prefetched 32 candidates at a time.  t[] only ever goes 0 -> nonzero
				 *      for good, so a nonzero prefetch is final and a zero one is re-read by lane 0. ---- */
				for (int64_t kb = (int64_t)n_z - 1; kb >= 0; kb -= 32) {
					const int64_t kq = kb - (int64_t)lane;
					int32_t ci0 = 0, czs = 0; bool open = false;
					if (kq >= 0) { const anchor_t e = z[kq]; ci0 = (int32_t)e.y; czs = (int32_t)e.x; open = t[ci0] == 0; }
					uint32_t cand = __ballot_sync(FULL, open);
					while (cand) {
						const int bsel = __ffs(cand) - 1; cand &= cand - 1;
						const int32_t i0 = __shfl_sync(FULL, ci0, bsel), zs = __shfl_sync(FULL, czs, bsel);
						if (lane == 0 && t[i0] == 0) {
							int32_t end_i = -1, max_i = i0, c = i0, max_s = 0; /* mg_chain_bk_end */
							do {
								t[c] = 2;
								end_i = c = p[c];
								const int32_t s = c < 0 ? zs : zs - f[c];
								if (s > max_s) { max_s = s; max_i = c; }
								else if (max_s - s > max_drop) break;
							} while (c >= 0 && t[c] == 0);
							for (c = i0; c >= 0 && c != end_i; c = p[c]) t[c] = 0;
							const uint32_t n_v0 = n_v;
							for (c = i0; c != max_i; c = p[c]) { v[n_v++] = c; t[c] = 1; }
							const int32_t sc = c < 0 ? zs : zs - f[c];
							if (sc >= min_sc && n_v > n_v0 && (int32_t)(n_v - n_v0) >= min_cnt) u[n_u++] = (uint64_t)sc << 32 | (n_v - n_v0);
							else n_v = n_v0;
						}
					}
					__syncwarp();
				}
				RH_PROF_MARK(A.prof, 34, lane == 0);
				n_u = __shfl_sync(FULL, n_u, 0); n_v = __shfl_sync(FULL, n_v, 0);
				__syncwarp();
			}
			if (n_u > 0) {
				/* ---- compact_a: forward-order gather (= next chunk's prev_anchors), then chains by target ---- */
				unsigned long long co = 0;
				if (lane == 0) co = atomicAdd(A.carry_top, (unsigned long long)n_v);
				co = __shfl_sync(FULL, co, 0);
				const bool carry_ok = co + n_v <= A.carry_cap;
				if (!carry_ok && lane == 0) atomicExch(A.err, 2u);
				/* chain start offsets in backtrack order: exclusive scan of the chain sizes */
				uint32_t *koff = (uint32_t *)t; /* t[] is free once the backtrack is over */
				{
					uint32_t run = 0;
					for (uint32_t c0 = 0; c0 < n_u; c0 += 32) {
						const uint32_t ci = c0 + lane;
						const uint32_t ni = ci < n_u ? (uint32_t)u[ci] : 0u;
						uint32_t incl = ni;
#pragma unroll
						for (int o = 1; o < 32; o <<= 1) { const uint32_t q = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += q; }
						if (ci < n_u) koff[ci] = run + incl - ni;
						run += __shfl_sync(FULL, incl, 31);
					}
				}
				__syncwarp();
				for (uint32_t ci = lane; ci < n_u; ci += 32) { /* one lane per chain: chains are a few anchors long */
					const uint32_t ni = (uint32_t)u[ci], k0 = koff[ci];
					for (uint32_t j = 0; j < ni; ++j) {
						const anchor_t x = a[v[k0 + (ni - j - 1)]];
						b[k0 + j] = x;
						if (carry_ok) A.carry_out[co + k0 + j] = x;
						if (j == 0) { anchor_t e; e.x = x.x; e.y = (uint64_t)k0 << 32 | ci; w[ci] = e; }
					}
				}
				if (lane == 0 && carry_ok) { R->prev_off = co; R->prev_n = n_v; }
				__syncwarp();
				RH_PROF_MARK(A.prof, 35, lane == 0);
				warp_klib_sort_pairs(w, n_u, z, X, cnt, head, lane);
				RH_PROF_MARK(A.prof, 36, lane == 0);
				/* output offsets in target order, then copy chains and pre-compute the region keys (mm_gen_regs) */
				uint32_t *kout = (uint32_t *)v; /* the backtrack order list is no longer needed */
				{
					uint32_t run = 0;
					for (uint32_t c0 = 0; c0 < n_u; c0 += 32) {
						const uint32_t ci = c0 + lane;
						const uint32_t ni = ci < n_u ? (uint32_t)u[(uint32_t)w[ci].y] : 0u;
						uint32_t incl = ni;
#pragma unroll
						for (int o = 1; o < 32; o <<= 1) { const uint32_t q = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += q; }
						if (ci < n_u) kout[ci] = run + incl - ni;
						run += __shfl_sync(FULL, incl, 31);
					}
				}
				__syncwarp();
				const uint32_t rhash = wang32(wang32(R->ev_offset + S->n_events) + wang32(11u)); /* rmap.cpp:346-348 */
				for (uint32_t ci = lane; ci < n_u; ci += 32) {
					const uint32_t j = (uint32_t)w[ci].y, c = (uint32_t)u[j], k0 = kout[ci];
					const anchor_t *src = b + (w[ci].y >> 32);
					for (uint32_t q = 0; q < c; ++q) a[k0 + q] = src[q];
					u2[ci] = u[j];
					const anchor_t fa = src[0];
					const uint32_t h = (uint32_t)mix64((mix64(fa.x) + mix64(fa.y)) ^ rhash); /* hit.c:120 */
					anchor_t e; e.x = u[j] ^ h; e.y = (uint64_t)k0 << 32 | c;
					z[ci] = e;
				}
				__syncwarp();
				for (uint32_t ci = lane; ci < n_u; ci += 32) u[ci] = u2[ci];
				__syncwarp();

				/* ---- mm_gen_regs ---- */
				if (n_u > fin_regs_cap(S->n_anchors)) { if (lane == 0) atomicExch(A.err, 3u); n_u = 0; }
			}
			if (n_u > 0) {
				r = (dev_reg_t *)((uint8_t *)M.regs + fin_regs_off(S->n_anchors)); /* after the sort scratch (fin_scratch) */
				RH_PROF_MARK(A.prof, 37, lane == 0);
				warp_klib_sort_pairs(z, n_u, w, X, cnt, head, lane);
				RH_PROF_MARK(A.prof, 38, lane == 0);
				for (uint32_t i = lane; i < n_u; i += 32) { /* descending score */
					const anchor_t zz = z[n_u - 1 - i];
					dev_reg_t g;
					g.id = (int32_t)i; g.parent = -1; g.subsc = 0; g.n_sub = 0; g.mapq = 0;
					g.score = g.score0 = (int32_t)(zz.x >> 32); g.hash = (uint32_t)zz.x;
					g.cnt = (int32_t)zz.y; g.as = (int32_t)(zz.y >> 32);
					const anchor_t fa = a[g.as], la = a[g.as + g.cnt - 1];
					g.rev = (uint32_t)(fa.x >> 63); g.rid = (int32_t)(fa.x << 1 >> 33);
					g.rs = (int32_t)fa.x; g.re = (int32_t)la.x + 1; g.qs = (int32_t)fa.y; g.qe = (int32_t)la.y + 1;
					r[i] = g;
				}
				n_regs = n_u;
				__syncwarp();
				/* ---- mm_set_parent: regions in score order ---- */
				int maxq = 0;
				for (uint32_t i = lane; i < n_regs; i += 32) maxq = max(maxq, r[i].qe);
#pragma unroll
				for (int o = 16; o > 0; o >>= 1) maxq = max(maxq, __shfl_xor_sync(FULL, maxq, o));
				RH_PROF_MARK(A.prof, 39, lane == 0);
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
		}
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
