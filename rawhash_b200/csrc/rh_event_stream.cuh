/*
 * rh_event_stream.cuh — the event stage for ordinary chunks (<= EVS_MAXN samples) without intermediates in HBM.
 *
 * Same reference functions as rh_signal.cuh (src/rsig.c:496-503, src/revent.c:23-74,91-255, src/rsketch.c:143-204),
 * same bits; what changes is where the intermediates live.  The round-1 stage wrote z, both prefix sums and both
 * t-statistic arrays (5 x 16 KB per chunk) to HBM between its seven launches: 17.5 x the algorithmic bytes in DRAM
 * traffic.  Here nothing but peaks (one word per event) crosses HBM between three launches; the raw samples are read
 * three times instead (2 B each), because every pass is cheaper to recompute from them than to store:
 *
 *   k_evt_sums     warp / chunk   raw -> pA -> (30,200) filter -> exact Σx, Σx² -> mean, sd of the read so far
 *   k_evt_stream   LANE / chunk   raw -> pA -> z -> |z|<3 filter -> the two sequential float prefix sums, both
 *                                 t-statistics and both coupled peak detectors in ONE streaming loop.  The order-
 *                                 dependent steps (prefix sums, detector state machines: SURVEY H3) cost one lane per
 *                                 chunk, 32 chunks share every instruction; a t-statistic needs the prefix sums w
 *                                 positions either side, so the detectors run w2 positions behind the sample being
 *                                 read and the last 32 prefix values per chunk sit in a shared-memory ring.
 *   k_evt_finish   CTA  / chunk   z again into shared memory, per-segment sort (every thread ranks one sample inside
 *                                 its segment), IQR-filtered means = events, then the sketch: quantise all events in
 *                                 parallel, the diff filter's short dependent chain on one thread, pack + hash64 of
 *                                 every window of e kept events in parallel.
 *
 * Exact arithmetic without the IEEE division sequences where a cheaper exact form exists:
 *   - x / w for the window length w (five per position and window): q = x * (1/w); r = fma(-q, w, x); q' = fma(r, 1/w, q)
 *     is the correctly rounded quotient for the w it is used with — checked EXHAUSTIVELY over all 2^32 float patterns
 *     at rh_gpu_init (k_selftest_divw); a w that fails the check keeps __fdiv_rn.
 *   - z = (float)((x - mean) / sd) in double: Markstein's two-FMA refinement of (x - mean) * (1/sd) is within one double
 *     ulp of the quotient, and the float it rounds to can only differ when the quotient lies within a few double ulps
 *     of a float rounding boundary; those cases (2^-27 of all) take the IEEE division.
 */
#ifndef RH_EVENT_STREAM_CUH
#define RH_EVENT_STREAM_CUH

#include "rh_signal.cuh"

#define EVS_MAXN 4096          /* chunk lengths the streaming stage takes (the default chunk is 4000 samples) */
#define EVS_RING 32            /* prefix values kept per chunk: >= 2 * w2 + 1, w2 <= 15 */
#define EVS_THREADS 128

struct chunk_norm_t { double mean, sd; uint64_t cur0; uint32_t got, pad; };

struct evt_args_t {
	const int16_t *raw;
	read_state_t *rs;
	slot_t *slots;
	uint32_t n_slots;
	chunk_norm_t *norm;       /* per slot */
	const uint2 *groups;      /* k_evt_sums: (first slot, count) — consecutive chunks of ONE read, summed in order */
	uint32_t n_groups;
	uint32_t *peaks;          /* [e_off .. +e_cap) */
	float *events;
	uint32_t *seed_hash, *seed_pos, *min_hash, *min_pos;
	int fast_w1, fast_w2;     /* the exact multiply-by-reciprocal division holds for this window length */
};

/* ---- exhaustive check of the reciprocal division for one divisor ---- */
__device__ __forceinline__ float div_by_const(float x, float w, float rw)
{
	const float q = __fmul_rn(x, rw);
	const float r = __fmaf_rn(-q, w, x);
	return __fmaf_rn(r, rw, q);
}
__global__ void __launch_bounds__(256) k_selftest_divw(float w, unsigned long long *bad)
{
	const float rw = __frcp_rn(w);
	unsigned long long my = 0;
	for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < (1ULL << 32); b += (uint64_t)gridDim.x * blockDim.x) {
		const float x = __uint_as_float((uint32_t)b);
		if (!(fabsf(x) <= FLT_MAX)) continue; /* NaN, inf: never a prefix-sum difference */
		if ((uint32_t)b == 0x80000000u) continue; /* -0 / w: the reciprocal form gives +0.  The sign of a zero quotient never reaches
		                                           * the t-statistic: the variance is clamped to FLT_MIN, the numerator goes through fabsf */
		const bool diff = __float_as_uint(div_by_const(x, w, rw)) != __float_as_uint(__fdiv_rn(x, w));
		my += diff;
		if (diff && (x == 0.0f || (fabsf(x) >= 0x1p-100f && fabsf(x) <= 0x1p100f))) my += 1ULL << 32; /* high word: inside the range sums of |z|<3 live in */
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) my += __shfl_xor_sync(0xffffffffu, my, o);
	if ((threadIdx.x & 31) == 0 && my) atomicAdd(bad, my);
}

/* z of one pA sample, bit-identical to (float)(((double)pa - mean) / sd) */
__device__ __forceinline__ float z_of(float pa, double mean, double sd, double rsd)
{
	const double a = __dsub_rn((double)pa, mean);
	const double q0 = __dmul_rn(a, rsd);
	const double e = __fma_rn(-q0, sd, a);
	const double q1 = __fma_rn(e, rsd, q0);
	const uint32_t lo = (uint32_t)__double2loint(q1) & 0x1fffffffu; /* the 29 mantissa bits a float drops */
	if (lo - 0x0ffffff8u <= 16u) return __double2float_rn(__ddiv_rn(a, sd)); /* within 8 ulps of the rounding boundary */
	return __double2float_rn(q1);
}

/* =============================================================================================
 * k_evt_sums: warp per chunk.  Consumes the chunk's raw samples (the cursor moves), adds to the read's running sums.
 * ===========================================================================================*/
__global__ void __launch_bounds__(128) k_evt_sums(evt_args_t A)
{
	const uint32_t FULL = 0xffffffffu;
	const uint32_t gid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (gid >= A.n_groups) return;
	const uint2 G = A.groups[gid];
	const int16_t *__restrict__ raw = A.raw;
	const uint32_t lt = (1u << lane) - 1;
	for (uint32_t slot_id = G.x; slot_id < G.x + G.y; ++slot_id) { /* normalize_signal carries its sums from chunk to chunk (revent.c:221-255) */
		slot_t *S = &A.slots[slot_id];
		read_state_t *R = &A.rs[S->read];
		const double off = R->cal_offset, scale = R->cal_scale;
		const uint64_t cur0 = R->cursor, rend = R->raw_end;
		const uint32_t want = S->chunk_len;
		double sum = 0.0, sum2 = 0.0;
		uint32_t got = 0; uint64_t c = cur0;
		while (c < rend && got < want) {
			const uint64_t idx = c + lane;
			float pa = 0.0f; bool keep = false;
			if (idx < rend) { pa = raw_to_pa(raw[idx], off, scale); keep = pa_keep(pa); }
			uint32_t m = __ballot_sync(FULL, keep);
			const uint32_t nk = __popc(m);
			if (got + nk >= want) { /* the chunk ends right after its want-th kept sample */
				const uint32_t need = want - got;
				keep = keep && (uint32_t)__popc(m & lt) < need;
				m = __ballot_sync(FULL, keep);
				if (keep) { sum += (double)pa; sum2 += (double)__fmul_rn(pa, pa); }
				c += 32 - __clz(m);
				got = want;
				break;
			}
			if (keep) { sum += (double)pa; sum2 += (double)__fmul_rn(pa, pa); }
			got += nk;
			c = (c + 32 < rend) ? c + 32 : rend;
		}
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) { sum += __shfl_xor_sync(FULL, sum, o); sum2 += __shfl_xor_sync(FULL, sum2, o); }
		sum = __dadd_rn(R->sum, sum); sum2 = __dadd_rn(R->sum2, sum2);
		const uint32_t n_tot = R->n_sum + got;
		__syncwarp();
		if (lane == 0) {
			R->sum = sum; R->sum2 = sum2; R->n_sum = n_tot; R->cursor = c; S->raw_used = (uint32_t)(c - cur0);
			chunk_norm_t N;
			N.mean = __ddiv_rn(sum, (double)n_tot);
			N.sd = __dsqrt_rn(__fma_rn(-N.mean, N.mean, __ddiv_rn(sum2, (double)n_tot)));
			N.cur0 = cur0; N.got = got; N.pad = 0;
			A.norm[slot_id] = N;
			S->n_peaks = 0; S->n_sig = 0;
		}
		__syncwarp();
		__threadfence_block(); /* the next chunk of this read starts from the cursor and sums just written */
	}
}

/* t-statistic at position c from the ring of prefix values (comp_tstat body, revent.c:49-69) */
template <bool FAST>
__device__ __forceinline__ float tstat_ring(const float *__restrict__ rps, const float *__restrict__ rpq, uint32_t c, uint32_t w, float fw, float rw, uint32_t stride)
{
	const uint32_t i0 = ((c - w) & (EVS_RING - 1)) * stride, i1 = (c & (EVS_RING - 1)) * stride, i2 = ((c + w) & (EVS_RING - 1)) * stride;
	const float a = rps[i1], aq = rpq[i1];
	const float s1 = __fsub_rn(a, rps[i0]), q1 = __fsub_rn(aq, rpq[i0]);   /* c == w reads ps[0] = 0: a - 0 = a */
	const float s2 = __fsub_rn(rps[i2], a), q2 = __fsub_rn(rpq[i2], aq);
	float m1, m2, d1, d2;
	if (FAST) { m1 = div_by_const(s1, fw, rw); m2 = div_by_const(s2, fw, rw); d1 = div_by_const(q1, fw, rw); d2 = div_by_const(q2, fw, rw); }
	else { m1 = __fdiv_rn(s1, fw); m2 = __fdiv_rn(s2, fw); d1 = __fdiv_rn(q1, fw); d2 = __fdiv_rn(q2, fw); }
	float acc = __fmaf_rn(-m1, m1, d1);
	acc = __fadd_rn(acc, d2);
	acc = __fmaf_rn(-m2, m2, acc);
	const float var = fmaxf(FAST ? div_by_const(acc, fw, rw) : __fdiv_rn(acc, fw), FLT_MIN);
	return __fdiv_rn(fabsf(__fsub_rn(m2, m1)), __fsqrt_rn(var));
}

/* =============================================================================================
 * k_evt_stream: lane per chunk.
 * ===========================================================================================*/
template <bool F1, bool F2>
__global__ void __launch_bounds__(EVS_THREADS) k_evt_stream(evt_args_t A, dev_params_t P)
{
	__shared__ float s_ps[EVS_RING * EVS_THREADS], s_pq[EVS_RING * EVS_THREADS];
	const uint32_t tid = threadIdx.x;
	const uint32_t slot_id = blockIdx.x * EVS_THREADS + tid;
	const bool live = slot_id < A.n_slots;
	slot_t *S = live ? &A.slots[slot_id] : nullptr;
	float *__restrict__ rps = s_ps + tid, *__restrict__ rpq = s_pq + tid; /* column `tid` of the rings: conflict free */
	uint32_t raw_used = 0; double mean = 0, sd = 1, rsd = 1, off = 0, scale = 0; const int16_t *__restrict__ rp = A.raw;
	uint32_t *__restrict__ peaks = A.peaks;
	if (live) {
		const chunk_norm_t N = A.norm[slot_id];
		const read_state_t *R = &A.rs[S->read];
		mean = N.mean; sd = N.sd; rsd = __ddiv_rn(1.0, sd); off = R->cal_offset; scale = R->cal_scale;
		rp = A.raw + N.cur0; raw_used = S->raw_used;
		peaks = A.peaks + S->e_off;
	}
	const uint32_t w1 = P.w1, w2 = P.w2;
	const float fw1 = (float)w1, fw2 = (float)w2, rw1 = __frcp_rn(fw1), rw2 = __frcp_rn(fw2);
	const float thr1 = P.thr1, thr2 = P.thr2, height = P.height;
	const bool use1 = w1 >= 2, use2 = w2 >= 2;
	const uint32_t lag = w1 > w2 ? w1 : w2; /* the detectors run this far behind the sample being read */
	peak_det_t d1 = {0u, -1, FLT_MAX, 0}, d2 = {0u, -1, FLT_MAX, 0};
	uint32_t n_peaks = 0, n = 0; /* n = samples that passed both filters = index of the next prefix value */
	float ps = 0.0f, pq = 0.0f;
	rps[0] = 0.0f; rpq[0] = 0.0f;
	auto step = [&](int sample) {
		const float pa = raw_to_pa(sample, off, scale);
		if (!pa_keep(pa)) return;
		const float z = z_of(pa, mean, sd, rsd);
		if (!(z < 3.0f && z > -3.0f)) return;
		ps = __fadd_rn(ps, z); pq = __fmaf_rn(z, z, pq);
		++n;
		rps[(n & (EVS_RING - 1)) * EVS_THREADS] = ps; rpq[(n & (EVS_RING - 1)) * EVS_THREADS] = pq;
		if (n >= lag) { /* position c: everything up to ps[c + lag] is known */
			const uint32_t c = n - lag;
			const float t1 = (use1 && c >= w1) ? tstat_ring<F1>(rps, rpq, c, w1, fw1, rw1, EVS_THREADS) : 0.0f;
			const float t2 = (use2 && c >= w2) ? tstat_ring<F2>(rps, rpq, c, w2, fw2, rw2, EVS_THREADS) : 0.0f;
			detector_step(d1, &d2, c, t1, thr1, w1, w1, height, peaks, n_peaks);
			detector_step(d2, nullptr, c, t2, thr2, w2, w1, height, peaks, n_peaks);
		}
	};
	{ /* the lane's samples: scalar loads up to the first 16-byte boundary, then eight samples per load */
		uint32_t j = 0;
		const uint32_t head = min(raw_used, (uint32_t)(((16u - ((uint32_t)(uintptr_t)rp & 15u)) & 15u) >> 1));
		for (; j < head; ++j) step((int)__ldg(rp + j));
		for (; j + 8 <= raw_used; j += 8) {
			const uint4 v = __ldg((const uint4 *)(rp + j));
			const uint64_t lo = (uint64_t)v.y << 32 | v.x, hi = (uint64_t)v.w << 32 | v.z;
#pragma unroll 1
			for (uint32_t u = 0; u < 8; ++u) step((int)(short)(((u < 4 ? lo : hi) >> (16u * (u & 3u))) & 0xffffu));
		}
		for (; j < raw_used; ++j) step((int)__ldg(rp + j));
	}
	/* the last lag - 1 positions: a window that no longer fits gives 0 */
	for (uint32_t c = n >= lag ? n - lag + 1 : 0u; c < n; ++c) {
		const float t1 = (use1 && c >= w1 && c + w1 <= n) ? tstat_ring<F1>(rps, rpq, c, w1, fw1, rw1, EVS_THREADS) : 0.0f;
		const float t2 = (use2 && c >= w2 && c + w2 <= n) ? tstat_ring<F2>(rps, rpq, c, w2, fw2, rw2, EVS_THREADS) : 0.0f;
		detector_step(d1, &d2, c, t1, thr1, w1, w1, height, peaks, n_peaks);
		detector_step(d2, nullptr, c, t2, thr2, w2, w1, height, peaks, n_peaks);
	}
	if (live) { S->n_sig = n; S->n_peaks = n_peaks; }
}

/* =============================================================================================
 * k_evt_finish: CTA per chunk.
 * ===========================================================================================*/
#define EVF2_THREADS 128
__global__ void __launch_bounds__(EVF2_THREADS) k_evt_finish(evt_args_t A, dev_params_t P)
{
	__shared__ float zs[EVS_MAXN], so[EVS_MAXN];
	__shared__ uint16_t seg_of[EVS_MAXN];
	__shared__ uint32_t s_w[EVF2_THREADS / 32], s_nlong, s_bad, s_kept;
	const uint32_t FULL = 0xffffffffu;
	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	slot_t *S = &A.slots[blockIdx.x];
	const uint32_t n_peaks = S->n_peaks, n = S->n_sig;
	const bool gated = n_peaks < P.min_events;
	uint32_t n_seeds = 0;
	if (n_peaks > 0 && n <= EVS_MAXN && n_peaks < 0xffffu) {
		const uint32_t *__restrict__ pk = A.peaks + S->e_off;
		float *__restrict__ ev = A.events + S->e_off;
		/* ---- z of the chunk, again from the raw samples: every warp compacts a contiguous quarter of the raw range into its
		 *      own stretch of `so`, then the stretches are laid end to end in zs (two barriers for the whole chunk) ---- */
		{
			const chunk_norm_t N = A.norm[blockIdx.x];
			const read_state_t *R = &A.rs[S->read];
			const double mean = N.mean, sd = N.sd, rsd = __ddiv_rn(1.0, sd), off = R->cal_offset, scale = R->cal_scale;
			const int16_t *__restrict__ rp = A.raw + N.cur0;
			const uint32_t raw_used = S->raw_used;
			if (raw_used <= EVS_MAXN) {
				const uint32_t piece = (((raw_used + EVF2_THREADS / 32 - 1) / (EVF2_THREADS / 32)) + 31) & ~31u;
				const uint32_t cb = min(warp * piece, raw_used), ce = min(cb + piece, raw_used);
				uint32_t cnt = 0;
				for (uint32_t j0 = cb; j0 < ce; j0 += 32) {
					const uint32_t j = j0 + lane;
					float z = 0.0f; bool keep = false;
					if (j < ce) {
						const float pa = raw_to_pa((int)__ldg(rp + j), off, scale);
						if (pa_keep(pa)) { z = z_of(pa, mean, sd, rsd); keep = z < 3.0f && z > -3.0f; }
					}
					const uint32_t m = __ballot_sync(FULL, keep);
					if (keep) so[cb + cnt + __popc(m & ((1u << lane) - 1u))] = z; /* kept <= raw: stays inside the warp's stretch */
					cnt += __popc(m);
				}
				if (lane == 0) s_w[warp] = cnt;
				__syncthreads();
				uint32_t base = 0;
				for (uint32_t w2 = 0; w2 < warp; ++w2) base += s_w[w2];
				for (uint32_t i = lane; i < cnt; i += 32) { zs[base + i] = so[cb + i]; seg_of[base + i] = 0xffffu; }
				__syncthreads();
			} else { /* many outliers dropped: more raw samples than the shared buffers hold; ordered compaction tile by tile */
				uint32_t base = 0;
				for (uint32_t j0 = 0; j0 < raw_used; j0 += EVF2_THREADS) {
					const uint32_t j = j0 + tid;
					float z = 0.0f; bool keep = false;
					if (j < raw_used) {
						const float pa = raw_to_pa((int)__ldg(rp + j), off, scale);
						if (pa_keep(pa)) { z = z_of(pa, mean, sd, rsd); keep = z < 3.0f && z > -3.0f; }
					}
					const uint32_t m = __ballot_sync(FULL, keep);
					if (lane == 0) s_w[warp] = __popc(m);
					__syncthreads();
					uint32_t woff = 0, tot = 0;
#pragma unroll
					for (uint32_t w2 = 0; w2 < EVF2_THREADS / 32; ++w2) { const uint32_t c = s_w[w2]; if (w2 < warp) woff += c; tot += c; }
					if (keep) { const uint32_t i = base + woff + __popc(m & ((1u << lane) - 1u)); if (i < EVS_MAXN) { zs[i] = z; seg_of[i] = 0xffffu; } }
					base += tot;
					__syncthreads();
				}
			}
		}
		if (tid == 0) { s_nlong = 0; s_bad = 0; }
		__syncthreads();
		/* ---- gen_events (revent.c:193-219): segment j = z[peaks[j-1] .. peaks[j]) ---- */
		for (uint32_t j = tid; j < n_peaks; j += EVF2_THREADS) {
			const uint32_t p = pk[j], start = j ? pk[j - 1] : 0u;
			if (p > n || start > p) { s_bad = 1; continue; } /* the reference assumes increasing peaks inside the chunk */
			const uint32_t len = p - start;
			if (len > EVF_LONG) { atomicAdd(&s_nlong, 1u); continue; }
			for (uint32_t i = start; i < p; ++i) seg_of[i] = (uint16_t)j;
		}
		__syncthreads();
		const bool bad = s_bad != 0;
		if (!bad) {
			/* per-segment sort: every thread ranks one sample inside its segment (count of smaller values, index as tie-break), so
			 * the lanes of a warp work on neighbouring samples of the same few segments.  (A warp-per-segment version that ranks
			 * with shuffles was measured and dropped: 1.5x the instructions, a segment of 11 samples fills a third of a warp.) */
			for (uint32_t i = tid; i < n; i += EVF2_THREADS) {
				const uint32_t j = seg_of[i];
				if (j == 0xffffu) continue; /* after the last peak, or inside a long segment */
				const uint32_t start = j ? pk[j - 1] : 0u, end = pk[j];
				const float v = zs[i];
				uint32_t r = 0;
				for (uint32_t k = start; k < end; ++k) { const float u = zs[k]; r += (u < v) || (u == v && k < i); }
				so[start + r] = v;
			}
			__syncthreads();
			if (s_nlong) { /* long segments (rare): bitonic sort by the whole CTA, one at a time */
				for (uint32_t j = 0; j < n_peaks; ++j) {
					const uint32_t p = pk[j], start = j ? pk[j - 1] : 0u, len = p - start;
					if (len <= EVF_LONG) continue;
					/* sort zs[start..p) in place, padded virtually with +inf */
					uint32_t N2 = 1; while (N2 < len) N2 <<= 1;
					for (uint32_t k = 2; k <= N2; k <<= 1)
						for (uint32_t jj = k >> 1; jj > 0; jj >>= 1) {
							for (uint32_t i = tid; i < N2; i += EVF2_THREADS) {
								const uint32_t l = i ^ jj;
								if (l > i) {
									const float x = i < len ? zs[start + i] : FLT_MAX, y = l < len ? zs[start + l] : FLT_MAX;
									const bool up = (i & k) == 0;
									if ((x > y) == up) { if (i < len) zs[start + i] = y; if (l < len) zs[start + l] = x; }
								}
							}
							__syncthreads();
						}
					for (uint32_t i = tid; i < len; i += EVF2_THREADS) so[start + i] = zs[start + i];
					__syncthreads();
				}
			}
			for (uint32_t j = tid; j < n_peaks; j += EVF2_THREADS) {
				const uint32_t p = pk[j], start = j ? pk[j - 1] : 0u, len = p - start;
				ev[j] = len ? filtered_mean_sorted(so + start, len, 1) : 0.0f;
			}
		} else {
			for (uint32_t j = tid; j < n_peaks; j += EVF2_THREADS) ev[j] = 0.0f; /* cannot happen: peaks are emitted in increasing order */
		}
		__syncthreads();
		/* ---- ri_sketch_reg (rsketch.c:143-204) ---- */
		if (!gated) {
			uint8_t *qv = (uint8_t *)zs;                    /* zs is dead: quantised events, n_peaks bytes        */
			uint16_t *kept = (uint16_t *)(zs + EVS_MAXN / 2); /* kept event positions, n_peaks entries (8 KB half) */
			float *evs = so;                                /* events in shared memory for the dependent chain     */
			uint32_t *__restrict__ sh = A.seed_hash + S->e_off, *__restrict__ sp = A.seed_pos + S->e_off;
			const int e = P.e, q = P.q;
			const uint64_t mev = (q * e >= 64) ? ~0ULL : ((1ULL << (q * e)) - 1), mq = (1ULL << q) - 1;
			__syncthreads();
			for (uint32_t j = tid; j < n_peaks; j += EVF2_THREADS) { const float v = ev[j]; evs[j] = v; qv[j] = (uint8_t)(quantize_event(v, P.fine_min, P.fine_max, P.fine_range, 1u << q) & mq); }
			__syncthreads();
			if (tid == 0) { /* keep iff first or |v - last kept| >= diff */
				float last = 0.0f; uint32_t nk = 0;
				for (uint32_t i = 0; i < n_peaks; ++i) {
					const float v = evs[i];
					if (i && fabsf(__fsub_rn(v, last)) < P.diff) continue;
					last = v; kept[nk++] = (uint16_t)i;
				}
				s_kept = nk;
			}
			__syncthreads();
			const uint32_t nk = s_kept;
			n_seeds = nk >= (uint32_t)e ? nk - (uint32_t)e + 1 : 0u;
			for (uint32_t t = tid; t < nk; t += EVF2_THREADS) sp[t] = kept[t];
			for (uint32_t t = tid; t < n_seeds; t += EVF2_THREADS) { /* seed t = kept events t .. t+e-1 */
				uint64_t packed = 0;
				for (int m2 = 0; m2 < e; ++m2) packed = ((packed << q) | (uint64_t)qv[kept[t + m2]]) & mev;
				sh[t] = (uint32_t)seed_mix(packed);
			}
			if (P.w > 0 && n_seeds > 0) {
				__syncthreads();
				if (tid == 0) s_kept = minimizer_select(sh, sp, n_seeds, P.w, e, A.min_hash + S->e_off, A.min_pos + S->e_off, S->e_cap);
				__syncthreads();
				n_seeds = s_kept;
			}
		}
	}
	if (tid == 0) { S->n_events = n_peaks; S->n_seeds = n_seeds; S->gated = gated ? 1u : 0u; S->ev_done = 1; }
}

#endif
