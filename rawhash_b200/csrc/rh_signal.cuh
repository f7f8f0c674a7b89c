/*
 * rh_signal.cuh — the event stage: raw int16 current samples -> events -> seeds.
 *
 * Reference functions taken over (paths relative to the RawHash tree):
 *   src/rsig.c:496-503      int16 -> pA, drop samples outside (30,200) pA
 *   src/revent.c:221-255    normalize_signal (running Σx, Σx² over the read so far, z, |z|<3 filter)
 *   src/revent.c:23-36      comp_prefix_prefixsq (sequential float prefix sums)
 *   src/revent.c:38-74      comp_tstat (two window lengths)
 *   src/revent.c:91-150     gen_peaks (two coupled detectors)
 *   src/revent.c:158-219    gen_events / calculate_mean_of_filtered_segment
 *   src/rsketch.c:143-204   ri_sketch_reg (diff filter, dynamic_quantize, pack, hash64)
 *
 * What is parallel and what is not.  Inside one chunk three steps are order dependent and cannot
 * be re-associated without changing bits: the float prefix sums, the peak state machine and the
 * diff filter.  Everything else — pA conversion, both filters, Σx/Σx² (exact in
 * double, hence order free), the t-statistics, the per-segment sort/mean, quantise/hash — is
 * data parallel.  The stage is therefore a handful of back-to-back launches that alternate between
 * "all lanes on one chunk" and "one lane per chunk" work shapes, so the serial steps cost one
 * lane each instead of a whole warp:
 *
 *   k_sig_norm    warp / chunk     raw -> pA -> filter -> sums -> z -> filter
 *   k_sig_prefix  lane / chunk     the two sequential float prefix sums
 *   k_sig_tstat   CTA  / chunk     t-statistics for both windows, one thread per position
 *   k_sig_peaks   lane / chunk     the two coupled peak detectors
 *   k_sig_events  CTA  / chunk     one thread per segment: sort, IQR filter, mean
 *   k_sig_sketch  lane / chunk     diff filter, quantise, pack e events, hash
 *
 * Scratch per chunk (float, chunk-major, every chunk 16-byte aligned): z[len], ps/pq/t1/t2[len+1]; peaks/events/seeds[e_cap].
 */
#ifndef RH_SIGNAL_CUH
#define RH_SIGNAL_CUH

#include "rh_kernels.cuh"

struct sig_args_t {
	const int16_t *raw;
	read_state_t *rs;
	slot_t *slots;
	uint32_t n_slots;
	float *z;              /* [z_off .. +chunk_len)                      */
	float *ps, *pq;        /* [z_off + slot .. +chunk_len+1)             */
	float *t1, *t2;        /* [z_off + slot .. +chunk_len+1)             */
	uint32_t *peaks;       /* [e_off .. +e_cap)                          */
	float *events;         /* [e_off .. +e_cap)                          */
	uint32_t *seed_hash;   /* [e_off .. +e_cap)                          */
	uint32_t *seed_pos;    /* [e_off .. +e_cap)                          */
	uint32_t *min_hash, *min_pos; /* [e_off .. +e_cap) minimizer output before it replaces the seed stream (w > 0) */
	unsigned long long *prof;
};

/* ------------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(128) k_sig_norm(sig_args_t A)
{
	const uint32_t FULL = 0xffffffffu;
	const uint32_t slot_id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (slot_id >= A.n_slots) return;
	slot_t *S = &A.slots[slot_id];
	read_state_t *R = &A.rs[S->read];
	const int16_t *raw = A.raw;
	const double off = R->cal_offset, scale = R->cal_scale;
	const uint64_t cur0 = R->cursor, rend = R->raw_end;
	const uint32_t want = S->chunk_len;
	float *xz = A.z + S->z_off;
	const uint32_t lt = (1u << lane) - 1;

	/* pass 1: pA, outlier drop, compaction; Σx and Σx² are exact in double -> any order */
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
			if (keep) { xz[got + __popc(m & lt)] = pa; sum += (double)pa; sum2 += (double)__fmul_rn(pa, pa); }
			c += 32 - __clz(m); /* one past the last consumed sample */
			got = want;
			break;
		}
		if (keep) { xz[got + __popc(m & lt)] = pa; sum += (double)pa; sum2 += (double)__fmul_rn(pa, pa); }
		got += nk;
		c = (c + 32 < rend) ? c + 32 : rend;
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) { sum += __shfl_xor_sync(FULL, sum, o); sum2 += __shfl_xor_sync(FULL, sum2, o); }
	sum = __dadd_rn(R->sum, sum); sum2 = __dadd_rn(R->sum2, sum2);
	const uint32_t n_tot = R->n_sum + got;
	__syncwarp();
	if (lane == 0) { R->sum = sum; R->sum2 = sum2; R->n_sum = n_tot; R->cursor = c; S->raw_used = (uint32_t)(c - cur0); }
	const double mean = __ddiv_rn(sum, (double)n_tot);
	const double sd = __dsqrt_rn(__fma_rn(-mean, mean, __ddiv_rn(sum2, (double)n_tot)));

	/* pass 2: z-normalise, |z| < 3 filter, in-place compaction (writes never pass the reads) */
	uint32_t n = 0;
	for (uint32_t t0 = 0; t0 < got; t0 += 32) {
		const uint32_t i = t0 + lane;
		float zv = 0.0f; bool keep = false;
		if (i < got) { zv = __double2float_rn(__ddiv_rn(__dsub_rn((double)xz[i], mean), sd)); keep = zv < 3.0f && zv > -3.0f; }
		const uint32_t m = __ballot_sync(FULL, keep);
		if (keep) xz[n + __popc(m & lt)] = zv;
		n += __popc(m);
	}
	__syncwarp();

	if (lane == 0) { S->n_sig = n; S->n_peaks = 0; }
}

/* float prefix sums, comp_prefix_prefixsq (revent.c:23-36): two strictly sequential recurrences (the second one an
 * FMA in the reference as compiled), so they cannot be re-associated.  One LANE per chunk runs them — a warp works
 * on 32 chunks at once instead of 32 lanes repeating one chunk's chain — with 16-byte loads one group ahead and
 * 16-byte stores. */
__global__ void __launch_bounds__(128) k_sig_prefix(sig_args_t A)
{
	const uint32_t slot_id = blockIdx.x * blockDim.x + threadIdx.x;
	if (slot_id >= A.n_slots) return;
	const slot_t *S = &A.slots[slot_id];
	const uint32_t n = S->n_sig;
	const float4 *__restrict__ z4 = (const float4 *)(A.z + S->z_off);
	const uint64_t o = S->z_off + 4ull * slot_id;
	float4 *__restrict__ ps4 = (float4 *)(A.ps + o), *__restrict__ pq4 = (float4 *)(A.pq + o);
	/* ps[0] = 0, ps[i+1] = ps[i] + z[i]; group g stores ps[4g .. 4g+3] and needs z[4g-1 .. 4g+2] */
	float run_s = 0.0f, run_q = 0.0f;
	const uint32_t ng = n / 4 + 1; /* groups covering ps[0..n] */
	float4 cur = n ? z4[0] : make_float4(0, 0, 0, 0);
	for (uint32_t g = 0; g < ng; ++g) {
		const float4 zc = cur;
		if (4 * (g + 1) < n) cur = z4[g + 1]; /* next group's samples are in flight while this one is consumed */
		float4 s, q;
		s.x = run_s; q.x = run_q;
		const float v0 = 4 * g + 0 < n ? zc.x : 0.0f, v1 = 4 * g + 1 < n ? zc.y : 0.0f, v2 = 4 * g + 2 < n ? zc.z : 0.0f, v3 = 4 * g + 3 < n ? zc.w : 0.0f;
		s.y = __fadd_rn(s.x, v0); q.y = __fmaf_rn(v0, v0, q.x);
		s.z = __fadd_rn(s.y, v1); q.z = __fmaf_rn(v1, v1, q.y);
		s.w = __fadd_rn(s.z, v2); q.w = __fmaf_rn(v2, v2, q.z);
		run_s = __fadd_rn(s.w, v3); run_q = __fmaf_rn(v3, v3, q.w);
		ps4[g] = s; pq4[g] = q; /* entries past ps[n] fall in the chunk's padding and are never read as data */
	}
}

/* ------------------------------------------------------------------------------------------- */
__device__ __forceinline__ float tstat_at(const float *__restrict__ ps, const float *__restrict__ pq, uint32_t i, uint32_t w, float fw)
{ /* comp_tstat body, revent.c:49-69, contractions as in the compiled reference */
	const float a = ps[i], aq = pq[i];
	float s1 = a, q1 = aq;
	if (i > w) { s1 = __fsub_rn(a, ps[i - w]); q1 = __fsub_rn(aq, pq[i - w]); }
	const float s2 = __fsub_rn(ps[i + w], a), q2 = __fsub_rn(pq[i + w], aq);
	const float m1 = __fdiv_rn(s1, fw), m2 = __fdiv_rn(s2, fw);
	float acc = __fmaf_rn(-m1, m1, __fdiv_rn(q1, fw));
	acc = __fadd_rn(acc, __fdiv_rn(q2, fw));
	acc = __fmaf_rn(-m2, m2, acc);
	const float var = fmaxf(__fdiv_rn(acc, fw), FLT_MIN);
	return __fdiv_rn(fabsf(__fsub_rn(m2, m1)), __fsqrt_rn(var));
}

__global__ void __launch_bounds__(256) k_sig_tstat(sig_args_t A, dev_params_t P)
{
	const slot_t *S = &A.slots[blockIdx.x];
	const uint32_t n = S->n_sig;
	if (n == 0) return;
	const uint64_t o = S->z_off + 4ull * blockIdx.x;
	const float *ps = A.ps + o, *pq = A.pq + o;
	float *t1 = A.t1 + o, *t2 = A.t2 + o;
	const uint32_t w1 = P.w1, w2 = P.w2;
	const float fw1 = (float)w1, fw2 = (float)w2;
	const bool ok1 = w1 >= 2 && n >= 2 * w1, ok2 = w2 >= 2 && n >= 2 * w2;
	const uint32_t n_pad = (n + 4) & ~3u; /* n+1 values, padded to the 16-byte groups k_sig_peaks loads */
	for (uint32_t i = threadIdx.x; i < n_pad; i += blockDim.x) {
		t1[i] = (ok1 && i >= w1 && i + w1 <= n) ? tstat_at(ps, pq, i, w1, fw1) : 0.0f;
		t2[i] = (ok2 && i >= w2 && i + w2 <= n) ? tstat_at(ps, pq, i, w2, fw2) : 0.0f;
	}
}

/* ------------------------------------------------------------------------------------------- */
struct peak_det_t { uint32_t masked_to; int peak_pos; float peak_val; int valid; };

/* One detector at one position, revent.c:100-146, written as selects so that the 32 lanes of a warp — each
 * running the state machine of its own chunk — execute one instruction stream instead of diverging over the
 * detector states.  `other` (the long detector, reset and masked by the short one) may be null. */
__device__ __forceinline__ void detector_step(peak_det_t &D, peak_det_t *other, uint32_t i, float cur,
                                              float thr, uint32_t win, uint32_t win0, float height, uint32_t *__restrict__ peaks, uint32_t &n_peaks)
{
	const bool active = !(D.masked_to >= i);
	const bool searching = active && D.peak_pos == -1, tracking = active && D.peak_pos != -1;
	/* searching for a rise: follow the signal down, latch when it climbs more than `height` above the low */
	const bool lower = cur < D.peak_val;
	const bool rise = !lower && __fsub_rn(cur, D.peak_val) > height;
	float pv = (searching && (lower || rise)) ? cur : D.peak_val;
	int pp = (searching && rise) ? (int)i : D.peak_pos;
	/* tracking a peak */
	const bool higher = tracking && cur > D.peak_val;
	pv = higher ? cur : pv;
	pp = higher ? (int)i : pp;
	if (other) {
		const bool mask = tracking && pv > thr;
		other->masked_to = mask ? (uint32_t)(pp + (int)win0) : other->masked_to;
		other->peak_pos = mask ? -1 : other->peak_pos;
		other->peak_val = mask ? FLT_MAX : other->peak_val;
		other->valid = mask ? 0 : other->valid;
	}
	const int valid = (D.valid || (tracking && __fsub_rn(pv, cur) > height && pv > thr)) ? 1 : 0;
	const bool emit = tracking && valid && (i - (uint32_t)pp) > win / 2;
	if (emit) peaks[n_peaks] = (uint32_t)pp;
	n_peaks += emit ? 1u : 0u;
	D.peak_pos = emit ? -1 : pp;
	D.peak_val = emit ? cur : pv;
	D.valid = emit ? 0 : valid;
}

__global__ void __launch_bounds__(128) k_sig_peaks(sig_args_t A, dev_params_t P)
{
	const uint32_t slot_id = blockIdx.x * blockDim.x + threadIdx.x;
	if (slot_id >= A.n_slots) return;
	slot_t *S = &A.slots[slot_id];
	const uint32_t n = S->n_sig;
	const uint64_t o = S->z_off + 4ull * slot_id; /* 16-byte aligned: z_off advances in multiples of 4 floats */
	const float4 *__restrict__ t1 = (const float4 *)(A.t1 + o), *__restrict__ t2 = (const float4 *)(A.t2 + o);
	uint32_t *__restrict__ peaks = A.peaks + S->e_off;
	peak_det_t d1 = {0u, -1, FLT_MAX, 0}, d2 = {0u, -1, FLT_MAX, 0};
	uint32_t n_peaks = 0;
	const uint32_t w1 = P.w1, w2 = P.w2;
	const float thr1 = P.thr1, thr2 = P.thr2, height = P.height;
	const uint32_t nq = (n + 3) / 4; /* the arrays hold n+1 values padded to a multiple of 4: reading the pad is harmless */
	float4 a = nq ? t1[0] : make_float4(0, 0, 0, 0), b = nq ? t2[0] : make_float4(0, 0, 0, 0);
	for (uint32_t q = 0; q < nq; ++q) {
		const float4 ca = a, cb = b;
		if (q + 1 < nq) { a = t1[q + 1]; b = t2[q + 1]; } /* next four positions are in flight while these are consumed */
		const float va[4] = {ca.x, ca.y, ca.z, ca.w}, vb[4] = {cb.x, cb.y, cb.z, cb.w};
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			const uint32_t i = 4 * q + k;
			if (i < n) {
				detector_step(d1, &d2, i, va[k], thr1, w1, w1, height, peaks, n_peaks);
				detector_step(d2, nullptr, i, vb[k], thr2, w2, w1, height, peaks, n_peaks);
			}
		}
	}
	S->n_peaks = n_peaks;
}

/* ------------------------------------------------------------------------------------------- */
#define EV_THREADS 128
#define EV_SMALL 64                /* segment lengths sorted by one thread in shared memory */
#define EV_BIG (EV_THREADS * EV_SMALL) /* longest segment the CTA sorts in shared memory */

__device__ __forceinline__ float filtered_mean_sorted(const float *seg, uint32_t len, uint32_t stride)
{ /* calculate_mean_of_filtered_segment after the sort, revent.c:163-179 (ascending accumulation order) */
	const float q1 = seg[(len / 4) * stride], q3 = seg[(3 * len / 4) * stride];
	const float iqr = __fsub_rn(q3, q1), lob = __fsub_rn(q1, iqr), hib = __fadd_rn(q3, iqr);
	float sum = 0.0f; uint32_t cnt = 0;
	for (uint32_t a = 0; a < len; ++a) { const float v = seg[a * stride]; if (v >= lob && v <= hib) { sum = __fadd_rn(sum, v); ++cnt; } }
	return cnt ? __fdiv_rn(sum, (float)cnt) : 0.0f;
}

__device__ void heapsort_global(float *a, uint32_t n)
{ /* segments too long for shared memory (whole-read Rawsamble chunks); any correct sort will do */
	for (uint32_t start = n / 2; start-- > 0;) {
		uint32_t r = start; const float v = a[r];
		for (;;) { uint32_t ch = 2 * r + 1; if (ch >= n) break; if (ch + 1 < n && a[ch + 1] > a[ch]) ++ch; if (a[ch] <= v) break; a[r] = a[ch]; r = ch; }
		a[r] = v;
	}
	for (uint32_t end = n; end-- > 1;) {
		const float v = a[end]; a[end] = a[0];
		uint32_t r = 0;
		for (;;) { uint32_t ch = 2 * r + 1; if (ch >= end) break; if (ch + 1 < end && a[ch + 1] > a[ch]) ++ch; if (a[ch] <= v) break; a[r] = a[ch]; r = ch; }
		a[r] = v;
	}
}

/* Fast path of gen_events for ordinary chunks (<= EVF_MAXN samples, no segment longer than EVF_LONG): the chunk's
 * samples sit in shared memory and every THREAD ranks one SAMPLE inside its segment (count of smaller values, index
 * as tie-break), so the lanes of a warp work on neighbouring samples of the same few segments and stay converged —
 * the thread-per-segment insertion sort ran with 5 of 32 lanes active (ncu).  The order-dependent float sum in
 * ascending order (revent.c:169-176) is then one short loop per segment over the sorted copy.  Chunks it cannot
 * take are left to k_sig_events. */
#define EVF_MAXN 4096
#define EVF_LONG 512
__global__ void __launch_bounds__(EV_THREADS) k_sig_events_fast(sig_args_t A)
{
	__shared__ float zs[EVF_MAXN], so[EVF_MAXN];
	__shared__ uint16_t seg_of[EVF_MAXN];
	__shared__ uint32_t s_maxlen, s_bad;
	slot_t *S = &A.slots[blockIdx.x];
	const uint32_t n_peaks = S->n_peaks, n = S->n_sig, tid = threadIdx.x;
	if (n_peaks == 0) { if (tid == 0) S->ev_done = 1; return; }
	if (n > EVF_MAXN || n_peaks >= 0xffffu) return;
	const uint32_t *__restrict__ pk = A.peaks + S->e_off;
	const float *__restrict__ zz = A.z + S->z_off;
	float *__restrict__ ev = A.events + S->e_off;
	if (tid == 0) { s_maxlen = 0; s_bad = 0; }
	for (uint32_t i = tid; i < n; i += EV_THREADS) { zs[i] = zz[i]; seg_of[i] = 0xffffu; }
	__syncthreads();
	for (uint32_t j = tid; j < n_peaks; j += EV_THREADS) { /* gen_events (revent.c:193-219): segment j = z[peaks[j-1] .. peaks[j]) */
		const uint32_t p = pk[j], start = j ? pk[j - 1] : 0u;
		if (p > n || start > p) { s_bad = 1; continue; } /* the reference assumes increasing peaks inside the chunk */
		const uint32_t len = p - start;
		atomicMax(&s_maxlen, len);
		if (len <= EVF_LONG) for (uint32_t i = start; i < p; ++i) seg_of[i] = (uint16_t)j;
	}
	__syncthreads();
	if (s_bad || s_maxlen > EVF_LONG) return; /* CTA-uniform: the general kernel takes this chunk */
	for (uint32_t i = tid; i < n; i += EV_THREADS) {
		const uint32_t j = seg_of[i];
		if (j == 0xffffu) continue; /* samples after the last peak produce no event */
		const uint32_t start = j ? pk[j - 1] : 0u, end = pk[j];
		const float v = zs[i];
		uint32_t r = 0;
		for (uint32_t k = start; k < end; ++k) { const float u = zs[k]; r += (u < v) || (u == v && k < i); }
		so[start + r] = v;
	}
	__syncthreads();
	for (uint32_t j = tid; j < n_peaks; j += EV_THREADS) {
		const uint32_t p = pk[j], start = j ? pk[j - 1] : 0u, len = p - start;
		ev[j] = len ? filtered_mean_sorted(so + start, len, 1) : 0.0f;
	}
	if (tid == 0) S->ev_done = 1;
}

__global__ void __launch_bounds__(EV_THREADS) k_sig_events(sig_args_t A)
{
	__shared__ float s_buf[EV_BIG];
	__shared__ uint32_t s_long[256];
	__shared__ uint32_t s_nlong;
	slot_t *S = &A.slots[blockIdx.x];
	const uint32_t n_peaks = S->n_peaks, tid = threadIdx.x;
	if (n_peaks == 0 || S->ev_done) return; /* ev_done: k_sig_events_fast already produced this chunk's events */
	const uint32_t *pk = A.peaks + S->e_off;
	float *zz = A.z + S->z_off;
	float *ev = A.events + S->e_off;
	if (tid == 0) s_nlong = 0;
	__syncthreads();
	/* gen_events (revent.c:193-219): segment j = z[peaks[j-1] .. peaks[j]) */
	for (uint32_t j = tid; j < n_peaks; j += EV_THREADS) {
		const uint32_t p = pk[j], start = j ? pk[j - 1] : 0u;
		const uint32_t len = p > start ? p - start : 0u; /* the reference assumes increasing peaks */
		if (len == 0) { ev[j] = 0.0f; continue; }
		if (len > EV_SMALL) { const uint32_t q = atomicAdd(&s_nlong, 1u); if (q < 256) s_long[q] = j; else { heapsort_global(zz + start, len); ev[j] = filtered_mean_sorted(zz + start, len, 1); } continue; }
		float *seg = s_buf + tid;
		const float *src = zz + start;
		for (uint32_t a = 0; a < len; ++a) { /* insertion sort while loading; order of equal values is immaterial */
			const float cur = src[a];
			uint32_t b = a;
			while (b > 0 && seg[(b - 1) * EV_THREADS] > cur) { seg[b * EV_THREADS] = seg[(b - 1) * EV_THREADS]; --b; }
			seg[b * EV_THREADS] = cur;
		}
		ev[j] = filtered_mean_sorted(seg, len, EV_THREADS);
	}
	__syncthreads();
	const uint32_t nl = min(s_nlong, 256u);
	for (uint32_t q = 0; q < nl; ++q) { /* long segments: the whole CTA sorts one at a time (bitonic, padded with +inf) */
		const uint32_t j = s_long[q];
		const uint32_t p = pk[j], start = j ? pk[j - 1] : 0u, len = p - start;
		if (len > EV_BIG) { if (tid == 0) { heapsort_global(zz + start, len); ev[j] = filtered_mean_sorted(zz + start, len, 1); } __syncthreads(); continue; }
		uint32_t N = 1; while (N < len) N <<= 1;
		for (uint32_t i = tid; i < N; i += EV_THREADS) s_buf[i] = i < len ? zz[start + i] : FLT_MAX;
		__syncthreads();
		for (uint32_t k = 2; k <= N; k <<= 1)
			for (uint32_t jj = k >> 1; jj > 0; jj >>= 1) {
				for (uint32_t i = tid; i < N; i += EV_THREADS) {
					const uint32_t l = i ^ jj;
					if (l > i) {
						const float x = s_buf[i], y = s_buf[l];
						const bool up = (i & k) == 0;
						if ((x > y) == up) { s_buf[i] = y; s_buf[l] = x; }
					}
				}
				__syncthreads();
			}
		if (tid == 0) ev[j] = filtered_mean_sorted(s_buf, len, 1);
		__syncthreads();
	}
}

/* ------------------------------------------------------------------------------------------- */
#define RH_MAX_W 32
/* ri_sketch_min, rsketch.c:94-140: minimizer selection over the seed stream (one (hash, position) per window of e
 * kept events), minimap2's window logic including the equal-minimum cases.  Seeds compare by hash (the low 6 bits
 * of x hold the same span for all), an empty window slot is "larger than everything".  Output goes to oh/op and
 * replaces the stream at the end.  Every stream entry is emitted at most once, so n_out <= n. */
__device__ uint32_t minimizer_select(uint32_t *sh, uint32_t *sp, uint32_t n, int w, int e, uint32_t *oh, uint32_t *op, uint32_t cap)
{
	const uint64_t EMPTY = ~0ULL;
	uint64_t wx[RH_MAX_W]; uint32_t wy[RH_MAX_W];
	for (int j = 0; j < w; ++j) { wx[j] = EMPTY; wy[j] = 0xffffffffu; }
	uint64_t mx = EMPTY; uint32_t my = 0xffffffffu;
	int wp = 0, mp = 0;
	uint32_t n_out = 0;
#define RH_PUSH(hx, py) do { if (n_out < cap) { oh[n_out] = (uint32_t)(hx); op[n_out] = (py); } ++n_out; } while (0)
	for (uint32_t t = 0; t < n; ++t) {
		const uint32_t l = t + (uint32_t)e; /* kept events seen so far */
		const uint64_t cx = sh[t]; const uint32_t cy = sp[t];
		wx[wp] = cx; wy[wp] = cy;
		if (l == (uint32_t)(w + e - 1) && mx != EMPTY) { /* first full window: equal minima not stored yet */
			for (int j = wp + 1; j < w; ++j) if (mx == wx[j] && wy[j] != my) RH_PUSH(wx[j], wy[j]);
			for (int j = 0; j < wp; ++j) if (mx == wx[j] && wy[j] != my) RH_PUSH(wx[j], wy[j]);
		}
		if (cx <= mx) {
			if (l >= (uint32_t)(w + e) && mx != EMPTY) RH_PUSH(mx, my);
			mx = cx; my = cy; mp = wp;
		} else if (wp == mp) { /* the minimum slid out of the window */
			if (l >= (uint32_t)(w + e - 1) && mx != EMPTY) RH_PUSH(mx, my);
			mx = EMPTY;
			for (int j = wp + 1; j < w; ++j) if (mx >= wx[j]) { mx = wx[j]; my = wy[j]; mp = j; }
			for (int j = 0; j <= wp; ++j) if (mx >= wx[j]) { mx = wx[j]; my = wy[j]; mp = j; }
			if (l >= (uint32_t)(w + e - 1) && mx != EMPTY) {
				for (int j = wp + 1; j < w; ++j) if (mx == wx[j] && my != wy[j]) RH_PUSH(wx[j], wy[j]);
				for (int j = 0; j <= wp; ++j) if (mx == wx[j] && my != wy[j]) RH_PUSH(wx[j], wy[j]);
			}
		}
		if (++wp == w) wp = 0;
	}
	if (mx != EMPTY) RH_PUSH(mx, my);
#undef RH_PUSH
	if (n_out > cap) n_out = cap; /* cannot happen (n_out <= n <= cap); keeps the copy below in bounds */
	for (uint32_t t = 0; t < n_out; ++t) { sh[t] = oh[t]; sp[t] = op[t]; }
	return n_out;
}

/* ------------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(128) k_sig_sketch(sig_args_t A, dev_params_t P)
{ /* ri_sketch_reg, rsketch.c:143-204: diff filter against the last KEPT event, quantise, pack, hash */
	const uint32_t slot_id = blockIdx.x * blockDim.x + threadIdx.x;
	if (slot_id >= A.n_slots) return;
	slot_t *S = &A.slots[slot_id];
	const uint32_t n_events = S->n_peaks; /* every emitted peak lies in (0, n_sig) */
	uint32_t n_seeds = 0;
	const bool gated = n_events < P.min_events;
	if (!gated) {
		const float *__restrict__ ev = A.events + S->e_off;
		uint32_t *__restrict__ sh = A.seed_hash + S->e_off, *__restrict__ sp = A.seed_pos + S->e_off;
		const int e = P.e, q = P.q;
		const uint64_t mev = (q * e >= 64) ? ~0ULL : ((1ULL << (q * e)) - 1), mq = (1ULL << q) - 1;
		uint64_t packed = 0; float last = 0.0f; uint32_t kept = 0;
		for (uint32_t i = 0; i < n_events; ++i) {
			const float v = ev[i];
			if (i && fabsf(__fsub_rn(v, last)) < P.diff) continue;
			last = v;
			packed = ((packed << q) | (quantize_event(v, P.fine_min, P.fine_max, P.fine_range, 1u << q) & mq)) & mev;
			sp[kept] = i; /* position of the kept event; seed j starts at kept event j */
			++kept;
			if (kept >= (uint32_t)e) sh[n_seeds++] = (uint32_t)seed_mix(packed);
		}
		if (P.w > 0 && n_seeds > 0) n_seeds = minimizer_select(sh, sp, n_seeds, P.w, e, A.min_hash + S->e_off, A.min_pos + S->e_off, S->e_cap);
	}
	S->n_events = n_events; S->n_seeds = n_seeds; S->gated = gated ? 1u : 0u;
}

#endif
