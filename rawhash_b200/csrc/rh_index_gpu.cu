/*
 * rh_index_gpu.cu — index construction on the GPU: the reference's
 *     ri_idx_gen / worker_pipeline / worker_post        (src/rindex.c:100-192, 311-363, 900-925)
 *     ri_seq_to_sig                                     (src/rsig.c:13-40)
 *     ri_sketch_reg on both strands                     (src/rsketch.c:143-204)
 * for an ACGT-only reference and w = 0 (no minimizers); anything else is left to the host builder
 * rh_index_build, which this must equal key for key and position for position.
 *
 *   k_idx_events   thread per (sequence, strand, k-mer): expected event value = pore level of the k-mer
 *   k_idx_keep     thread per (sequence, strand): the diff filter against the last KEPT event is a
 *                  sequential recurrence (rsketch.c:187-189); it only writes the list of kept positions
 *   k_idx_seeds    thread per window of e kept events: quantise, pack, hash64 -> (hash, y)
 *   sort           two stable radix sorts (by y, then by hash) = ascending positions inside each key
 *                  (what worker_post's sort leaves, rindex.c:350); CUB device primitives — this is the
 *                  library part of a non-hot path
 *   k_idx_heads    first element of every key run -> distinct keys + CSR offsets
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include <cuda_runtime.h>
#include <cub/cub.cuh>

#include "rh_host.h"
#include "rh_dev.cuh"

namespace {

struct seq_desc_t {
	uint64_t seq_off;   /* first base of the sequence in the concatenated base array       */
	uint64_t ev_off;    /* first event of this (sequence, strand) in the event arrays     */
	uint32_t len;       /* bases                                                          */
	uint32_t n_ev;      /* len - k + 1 (0 if the sequence is shorter than k)              */
	uint32_t id, strand;
	uint32_t n_kept;    /* out: events that survive the diff filter                       */
	uint64_t seed_off;  /* first seed of this (sequence, strand) in the seed arrays       */
};

__device__ __forceinline__ int base2(uint8_t c)
{
	return (c == 'A' || c == 'a') ? 0 : (c == 'C' || c == 'c') ? 1 : (c == 'G' || c == 'g') ? 2 : 3;
}

__global__ void __launch_bounds__(256) k_idx_events(const uint8_t *__restrict__ bases, const seq_desc_t *__restrict__ D, uint32_t n_desc,
                                                    const float *__restrict__ pore, int k, float *__restrict__ ev)
{
	const uint32_t d = blockIdx.y;
	if (d >= n_desc) return;
	const seq_desc_t S = D[d];
	const uint8_t *s = bases + S.seq_off;
	for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < S.n_ev; j += gridDim.x * blockDim.x) {
		/* event j covers bases j .. j+k-1 of the strand being read (ri_seq_to_sig, rsig.c:13-40); the reverse strand
		 * reads the sequence backwards and complemented */
		uint32_t kmer = 0;
		for (int t = 0; t < k; ++t) {
			const uint32_t q = j + (uint32_t)t;
			const int c = S.strand ? 3 - base2(s[S.len - 1 - q]) : base2(s[q]);
			kmer = (kmer << 2) | (uint32_t)c;
		}
		ev[S.ev_off + j] = pore[kmer];
	}
}

__global__ void k_idx_keep(seq_desc_t *D, uint32_t n_desc, const float *__restrict__ ev, uint32_t *__restrict__ kept_pos, float diff)
{
	const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
	if (d >= n_desc) return;
	const uint32_t n = D[d].n_ev;
	const float *__restrict__ v = ev + D[d].ev_off;
	uint32_t *__restrict__ out = kept_pos + D[d].ev_off;
	uint32_t kept = 0; float last = 0.0f;
	for (uint32_t i = 0; i < n; ++i) { /* ri_sketch_reg, rsketch.c:176-189: the first event is always kept */
		const float x = v[i];
		if (i && fabsf(__fsub_rn(x, last)) < diff) continue;
		last = x;
		out[kept++] = i;
	}
	D[d].n_kept = kept;
}

__global__ void __launch_bounds__(256) k_idx_seeds(const seq_desc_t *__restrict__ D, uint32_t n_desc, const float *__restrict__ ev, const uint32_t *__restrict__ kept_pos,
                                                   int e, int q, float fine_min, float fine_max, float fine_range,
                                                   uint32_t *__restrict__ hash_out, uint64_t *__restrict__ y_out)
{
	const uint32_t d = blockIdx.y;
	if (d >= n_desc) return;
	const seq_desc_t S = D[d];
	if (S.n_kept < (uint32_t)e) return;
	const uint32_t n_seeds = S.n_kept - (uint32_t)e + 1;
	const float *__restrict__ v = ev + S.ev_off;
	const uint32_t *__restrict__ kp = kept_pos + S.ev_off;
	const uint64_t mev = (q * e >= 64) ? ~0ULL : ((1ULL << (q * e)) - 1), mq = (1ULL << q) - 1;
	for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n_seeds; t += gridDim.x * blockDim.x) {
		uint64_t packed = 0; /* window of e kept events starting at kept event t (rsketch.c:191-202) */
		for (int m = 0; m < e; ++m) packed = ((packed << q) | (quantize_event(v[kp[t + m]], fine_min, fine_max, fine_range, 1u << q) & mq)) & mev;
		hash_out[S.seed_off + t] = (uint32_t)seed_mix(packed);
		y_out[S.seed_off + t] = (uint64_t)S.id << 32 | (uint64_t)kp[t] << 1 | (uint64_t)S.strand;
	}
}

__global__ void __launch_bounds__(256) k_idx_heads(const uint32_t *__restrict__ hash_sorted, uint64_t n, uint32_t *__restrict__ is_head)
{
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
		is_head[i] = (i == 0 || hash_sorted[i] != hash_sorted[i - 1]) ? 1u : 0u;
}

__global__ void __launch_bounds__(256) k_idx_csr(const uint32_t *__restrict__ hash_sorted, const uint32_t *__restrict__ is_head, const uint32_t *__restrict__ head_rank /* exclusive scan of is_head */,
                                                 uint64_t n, uint32_t *__restrict__ keys, uint64_t *__restrict__ off)
{
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
		if (is_head[i]) { keys[head_rank[i]] = hash_sorted[i]; off[head_rank[i]] = i; }
}

struct dev_free { std::vector<void *> p; ~dev_free() { for (void *q : p) cudaFree(q); } };

#define IDX_TRY(call)                                                                                     \
	do {                                                                                                  \
		cudaError_t e_ = (call);                                                                          \
		if (e_ != cudaSuccess) { rh_set_error("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); return NULL; } \
	} while (0)

template <class T> T *dmalloc(dev_free &F, size_t n)
{
	void *p = nullptr;
	if (cudaMalloc(&p, (n ? n : 1) * sizeof(T)) != cudaSuccess) return nullptr;
	F.p.push_back(p);
	return (T *)p;
}

} // namespace

/* Same result as rh_index_build (ri_idx_gen semantics), computed on `device`.  Falls back to the host builder for
 * inputs the kernels do not cover (non-ACGT bases, minimizers). */
static rh_index_t *index_build_gpu_one(const rh_params_t *p, const float *pore_vals, uint32_t n_pore_vals,
                                       uint32_t n_seq, const char *const *names, const char *const *seqs,
                                       const uint32_t *lens, int device);

extern "C" rh_index_t *rh_index_build_gpu(const rh_params_t *p, const float *pore_vals, uint32_t n_pore_vals,
                                           uint32_t n_seq, const char *const *names, const char *const *seqs,
                                           const uint32_t *lens, int device)
{
	if (!p || !pore_vals || n_pore_vals < (1u << (2 * p->k)) || (n_seq && (!names || !seqs || !lens))) { rh_set_error("rh_index_build_gpu: bad arguments"); return NULL; }
	bool plain = p->w == 0 && p->n == 0 && p->k >= 1 && p->k <= 15 && p->e >= 1 && p->e * p->q <= 64;
	uint64_t total = 0;
	for (uint32_t i = 0; i < n_seq && plain; ++i) {
		total += lens[i];
		for (uint32_t j = 0; j < lens[i]; ++j) { const char c = seqs[i][j]; if (!(c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == 'a' || c == 'c' || c == 'g' || c == 't')) { plain = false; break; } }
	}
	if (!plain) return rh_index_build(p, pore_vals, n_pore_vals, n_seq, names, seqs, lens, 8);
	/* one pass sorts every seed of its sequences on the device (≈50 B of buffers per base); references beyond the group
	 * size are built contig group by contig group and merged on the host (rh_index_build_grouped) */
	const uint64_t group = rh_index_group_bases((uint64_t)512 << 20);
	if (n_seq > 1 && total > group) return rh_index_build_grouped(index_build_gpu_one, group, p, pore_vals, n_pore_vals, n_seq, names, seqs, lens, device);
	return index_build_gpu_one(p, pore_vals, n_pore_vals, n_seq, names, seqs, lens, device);
}

static rh_index_t *index_build_gpu_one(const rh_params_t *p, const float *pore_vals, uint32_t n_pore_vals,
                                       uint32_t n_seq, const char *const *names, const char *const *seqs,
                                       const uint32_t *lens, int device)
{
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device >= ndev) { rh_set_error("no usable CUDA device (count=%d, asked %d)", ndev, device); return NULL; }
	IDX_TRY(cudaSetDevice(device));
	const int k = p->k, e = p->e;

	/* descriptors: two strands per sequence, in the order the host builder emits seeds (sequence, then strand) */
	std::vector<seq_desc_t> D;
	uint64_t n_bases = 0, n_ev = 0;
	for (uint32_t i = 0; i < n_seq; ++i) {
		const uint32_t len = lens[i], ne = len >= (uint32_t)k ? len - (uint32_t)k + 1 : 0;
		for (uint32_t s = 0; s < 2; ++s) {
			seq_desc_t d; memset(&d, 0, sizeof(d));
			d.seq_off = n_bases; d.ev_off = n_ev; d.len = len; d.n_ev = ne; d.id = i; d.strand = s;
			if (len > 0) D.push_back(d);
			n_ev += ne;
		}
		n_bases += len;
	}
	rh_index_s *idx = new rh_index_s();
	idx->flag = p->idx_flag; idx->w = p->w; idx->e = p->e; idx->n = p->n; idx->q = p->q; idx->k = p->k;
	idx->diff = p->diff; idx->fine_min = p->fine_min; idx->fine_max = p->fine_max; idx->fine_range = p->fine_range;
	for (uint32_t i = 0; i < n_seq; ++i) { idx->names.emplace_back(names[i]); idx->lens.push_back(lens[i]); }
	idx->off.push_back(0);
	if (D.empty() || n_ev == 0) return idx;
	auto bail = [&]() -> rh_index_t * { delete idx; return NULL; };

	dev_free F;
	uint8_t *d_bases = dmalloc<uint8_t>(F, n_bases);
	seq_desc_t *d_D = dmalloc<seq_desc_t>(F, D.size());
	float *d_pore = dmalloc<float>(F, n_pore_vals), *d_ev = dmalloc<float>(F, n_ev);
	uint32_t *d_kept = dmalloc<uint32_t>(F, n_ev);
	if (!d_bases || !d_D || !d_pore || !d_ev || !d_kept) { rh_set_error("rh_index_build_gpu: out of device memory"); return bail(); }
	{
		uint64_t o = 0;
		for (uint32_t i = 0; i < n_seq; ++i) { if (lens[i] && cudaMemcpy(d_bases + o, seqs[i], lens[i], cudaMemcpyHostToDevice) != cudaSuccess) { rh_set_error("sequence upload failed"); return bail(); } o += lens[i]; }
	}
	if (cudaMemcpy(d_D, D.data(), D.size() * sizeof(seq_desc_t), cudaMemcpyHostToDevice) != cudaSuccess ||
	    cudaMemcpy(d_pore, pore_vals, (size_t)n_pore_vals * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) { rh_set_error("upload failed"); return bail(); }
	const uint32_t nd = (uint32_t)D.size();
	if (nd > 65535) { rh_set_error("rh_index_build_gpu: more than 32767 sequences"); return bail(); }
	k_idx_events<<<dim3(296, nd), 256>>>(d_bases, d_D, nd, d_pore, k, d_ev);
	k_idx_keep<<<(nd + 31) / 32, 32>>>(d_D, nd, d_ev, d_kept, p->diff);
	if (cudaMemcpy(D.data(), d_D, D.size() * sizeof(seq_desc_t), cudaMemcpyDeviceToHost) != cudaSuccess) { rh_set_error("event kernels failed: %s", cudaGetErrorString(cudaGetLastError())); return bail(); }
	uint64_t n_seeds = 0;
	for (seq_desc_t &d : D) { d.seed_off = n_seeds; if (d.n_kept >= (uint32_t)e) n_seeds += d.n_kept - (uint32_t)e + 1; }
	if (n_seeds == 0) return idx;
	if (n_seeds >= (1ULL << 31)) { rh_set_error("rh_index_build_gpu: more than 2^31 seeds"); return bail(); }
	if (cudaMemcpy(d_D, D.data(), D.size() * sizeof(seq_desc_t), cudaMemcpyHostToDevice) != cudaSuccess) { rh_set_error("upload failed"); return bail(); }
	uint32_t *d_h0 = dmalloc<uint32_t>(F, n_seeds), *d_h1 = dmalloc<uint32_t>(F, n_seeds);
	uint64_t *d_y0 = dmalloc<uint64_t>(F, n_seeds), *d_y1 = dmalloc<uint64_t>(F, n_seeds);
	if (!d_h0 || !d_h1 || !d_y0 || !d_y1) { rh_set_error("rh_index_build_gpu: out of device memory"); return bail(); }
	k_idx_seeds<<<dim3(296, nd), 256>>>(d_D, nd, d_ev, d_kept, e, p->q, p->fine_min, p->fine_max, p->fine_range, d_h0, d_y0);
	/* order: hash, then y — stable sort by y first, then by hash */
	const int n_items = (int)n_seeds;
	size_t tb1 = 0, tb2 = 0, tb3 = 0;
	cub::DeviceRadixSort::SortPairs(nullptr, tb1, d_y0, d_y1, d_h0, d_h1, n_items);
	cub::DeviceRadixSort::SortPairs(nullptr, tb2, d_h1, d_h0, d_y1, d_y0, n_items);
	cub::DeviceScan::ExclusiveSum(nullptr, tb3, d_h1, d_h1, n_items);
	const size_t tb = std::max(tb1, std::max(tb2, tb3));
	uint8_t *d_tmp = dmalloc<uint8_t>(F, tb);
	if (!d_tmp) { rh_set_error("rh_index_build_gpu: out of device memory"); return bail(); }
	size_t t = tb;
	IDX_TRY(cub::DeviceRadixSort::SortPairs(d_tmp, t, d_y0, d_y1, d_h0, d_h1, n_items));   /* keys y: (y0,h0) -> (y1,h1) */
	t = tb;
	IDX_TRY(cub::DeviceRadixSort::SortPairs(d_tmp, t, d_h1, d_h0, d_y1, d_y0, n_items));   /* keys hash: (h1,y1) -> (h0,y0) */
	/* d_h0 = sorted hashes, d_y0 = positions in index order */
	uint32_t *d_head = d_h1, *d_rank = dmalloc<uint32_t>(F, n_seeds);
	if (!d_rank) { rh_set_error("rh_index_build_gpu: out of device memory"); return bail(); }
	k_idx_heads<<<1184, 256>>>(d_h0, n_seeds, d_head);
	t = tb;
	IDX_TRY(cub::DeviceScan::ExclusiveSum(d_tmp, t, d_head, d_rank, n_items));
	uint32_t last_rank = 0, last_head = 0;
	IDX_TRY(cudaMemcpy(&last_rank, d_rank + n_seeds - 1, 4, cudaMemcpyDeviceToHost));
	IDX_TRY(cudaMemcpy(&last_head, d_head + n_seeds - 1, 4, cudaMemcpyDeviceToHost));
	const uint64_t n_keys = (uint64_t)last_rank + last_head;
	uint32_t *d_keys = dmalloc<uint32_t>(F, n_keys);
	uint64_t *d_off = dmalloc<uint64_t>(F, n_keys);
	if (!d_keys || !d_off) { rh_set_error("rh_index_build_gpu: out of device memory"); return bail(); }
	k_idx_csr<<<1184, 256>>>(d_h0, d_head, d_rank, n_seeds, d_keys, d_off);
	idx->keys.resize(n_keys); idx->off.resize(n_keys + 1); idx->pos.resize(n_seeds);
	if (cudaMemcpy(idx->keys.data(), d_keys, n_keys * 4, cudaMemcpyDeviceToHost) != cudaSuccess ||
	    cudaMemcpy(idx->off.data(), d_off, n_keys * 8, cudaMemcpyDeviceToHost) != cudaSuccess ||
	    cudaMemcpy(idx->pos.data(), d_y0, n_seeds * 8, cudaMemcpyDeviceToHost) != cudaSuccess) { rh_set_error("index download failed: %s", cudaGetErrorString(cudaGetLastError())); return bail(); }
	idx->off[n_keys] = n_seeds;
	return idx;
}
