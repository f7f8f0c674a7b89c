/* rh_dev.cuh — device-side types and small helpers shared by the kernels in rh_gpu.cu. */
#ifndef RH_DEV_CUH
#define RH_DEV_CUH

#include <stdint.h>
#include <float.h>
#include <cuda_runtime.h>

struct __align__(16) anchor_t { uint64_t x, y; };

/* parameters every kernel reads (copied from rh_params_t once per context) */
struct dev_params_t {
	int w, e, q, k;
	float diff, fine_min, fine_max, fine_range;
	uint32_t w1, w2;
	float thr1, thr2, height;
	uint32_t min_events;
	int mid_occ;
	int bw, max_t, max_q, max_iter, max_skip, min_cnt, min_sc, min_sc2;
	float pen_gap, pen_skip;
	float mask_level; int mask_len; float pri_ratio; int best_n; int min_strand_sc;
	float w_bestq, w_bestmq, w_bestmc, w_threshold;
	int min_mapq;
	uint32_t max_num_chunk, chunk_size;
	float sample_per_base;
	int ava, noadapt, sig_target;
};

/* flattened index in HBM */
struct dev_index_t {
	const uint32_t *keys;     /* ascending distinct hashes                    */
	const uint64_t *off;      /* n_keys+1                                      */
	const uint64_t *pos;      /* id<<32|pos<<1|strand, ascending within a key  */
	const uint32_t *bucket;   /* (1<<bucket_bits)+1 : first key index whose top bits >= b */
	int bucket_bits;
	uint64_t n_keys;
	const uint32_t *seq_len;
	const uint32_t *name_rank; /* Rawsamble: position of each target name in sorted order */
	uint32_t n_seq;
};

/* region record kept per chain (subset of mm_reg1_t, reference src/chain.h:27-45) */
struct dev_reg_t {
	int32_t id, cnt, rid, score, qs, qe, rs, re, parent, subsc, as, n_sub, score0;
	uint32_t mapq, rev, hash;
};
static_assert(sizeof(dev_reg_t) == 64, "fin_regs_cap() assumes 64-byte region records");

/* ---- exact arithmetic helpers: the float/double operations of the reference as compiled
 *      (x86-64 + FMA contraction).  The file is built with --fmad=false so nothing else fuses. */
__device__ __forceinline__ float raw_to_pa(int raw, double off, double scale)
{ /* reference src/rsig.c:496-499: (int16 + double offset) * (double)(float scale) -> float */
	return __double2float_rn(__dmul_rn(__dadd_rn((double)raw, off), scale));
}
__device__ __forceinline__ bool pa_keep(float pa) { return pa > 30.0f && pa < 200.0f; }

__device__ __forceinline__ uint64_t seed_mix(uint64_t key)
{ /* hash64 masked to 32 bits, reference src/rsketch.c:7-16 */
	const uint64_t m = 0xffffffffULL;
	key = (~key + (key << 21)) & m; key ^= key >> 24;
	key = (key + (key << 3) + (key << 8)) & m; key ^= key >> 14;
	key = (key + (key << 2) + (key << 4)) & m; key ^= key >> 28;
	key = (key + (key << 31)) & m;
	return key;
}

__device__ __forceinline__ uint64_t mix64(uint64_t key)
{ /* unmasked variant, reference src/hit.c:73-83 */
	key = ~key + (key << 21); key ^= key >> 24;
	key = key + (key << 3) + (key << 8); key ^= key >> 14;
	key = key + (key << 2) + (key << 4); key ^= key >> 28;
	key = key + (key << 31);
	return key;
}

__device__ __forceinline__ uint32_t wang32(uint32_t key)
{ /* __ac_Wang_hash, reference src/khash.h:400-409 */
	key += ~(key << 15); key ^= (key >> 10); key += (key << 3);
	key ^= (key >> 6); key += ~(key << 11); key ^= (key >> 16);
	return key;
}

__device__ __forceinline__ uint32_t quantize_event(float v, float fine_min, float fine_max, float fine_range, uint32_t nb)
{ /* dynamic_quantize, reference src/rsketch.c:18-53; every product/sum rounds separately */
	const float lo = -3.0f, width = 6.0f;
	float c1 = __fdiv_rn(__fsub_rn(1.0f, fine_range), 2.0f);
	float c2 = __fadd_rn(fine_range, c1);
	float u = __fdiv_rn(__fsub_rn(v, lo), width);
	float r;
	if (v >= fine_min && v <= fine_max) {
		float a = __fdiv_rn(__fsub_rn(fine_min, lo), width), b = __fdiv_rn(__fsub_rn(fine_max, lo), width);
		r = __fmul_rn(fine_range, __fdiv_rn(__fsub_rn(u, a), __fsub_rn(b, a)));
	} else {
		float t = __fmul_rn(c1, u);
		r = (u < 0.5f) ? __fadd_rn(fine_range, t) : __fadd_rn(c2, t);
	}
	return (uint32_t)__fmul_rn(r, (float)(nb - 1));
}

__device__ __forceinline__ float approx_log2(float x)
{ /* mg_log2, reference src/lchain.c:23-31 (two FMAs as compiled) */
	uint32_t i = __float_as_uint(x);
	float r = (float)((int)((i >> 23) & 255) - 128);
	i &= ~(255u << 23); i += 127u << 23;
	float z = __uint_as_float(i);
	return __fadd_rn(r, __fmaf_rn(__fmaf_rn(-0.34484843f, z, 2.02466578f), z, -0.67487759f));
}

__device__ __forceinline__ int32_t pair_score(uint64_t ix, uint64_t iy, uint64_t jx, uint64_t jy,
                                              int32_t max_t, int32_t max_q, int32_t bw, float pen_gap, float pen_skip)
{ /* compute_score, reference src/lchain.c:297-356 */
	int32_t dq = (int32_t)iy - (int32_t)jy;
	if (dq <= 0 || dq > max_q) return INT32_MIN;
	int32_t dr = (int32_t)(ix - jx);
	if (dr == 0 || dr > max_t) return INT32_MIN;
	int32_t dd = dr > dq ? dr - dq : dq - dr;
	if (dd > bw || dr > max_q) return INT32_MIN;
	int32_t dg = dr < dq ? dr : dq;
	int32_t qs = (int32_t)((jy >> 32) & 63);
	int32_t sc = qs < dg ? qs : dg;
	if (dd || dg > qs) {
		float lin = __fmaf_rn(pen_gap, (float)dd, __fmul_rn(pen_skip, (float)dg));
		float lg = dd >= 1 ? approx_log2((float)(dd + 1)) : 0.0f;
		sc -= (int)__fadd_rn(lin, __fmul_rn(.5f, lg));
	}
	return sc;
}


/* optional phase timing (RH_PROF=1 at run time): cycles summed over calling warps/threads */
#define RH_PROF_BEGIN(P) long long prof_t_ = (P) ? clock64() : 0
#define RH_PROF_MARK(P, idx, leader) do { if ((P) && (leader)) { const long long n_ = clock64(); atomicAdd(&(P)[idx], (unsigned long long)(n_ - prof_t_)); prof_t_ = n_; } } while (0)

#endif
