/*
 * rh_gpu.cu — host driver of the GPU mapping path behind the C-ABI (include/rawhash_b200.h).
 *
 * rh_gpu_map_batch_raw() is the drop-in for the reference's
 *     kt_for(p->n_threads, map_worker_for, in, s->n_sig)          (src/rmap.cpp:700)
 * The per-read chunk loop of map_worker_for (src/rmap.cpp:415-501) becomes "chunk rounds":
 * round c processes chunk c of every read that has not stopped yet, all reads in parallel on
 * the device; reads that satisfy a stop rule (or run out of signal/chunks) emit their final
 * record in that round and drop out of the active set.
 *
 * There is no CPU fallback anywhere in this file: without a CUDA device every entry point
 * fails with RH_ERR_CUDA.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <algorithm>
#include <chrono>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "rh_host.h"
#include "rh_kernels.cuh"
#include "rh_signal.cuh"
#include "rh_event_stream.cuh"
#include "rh_anchor_sort.cuh"
#include "rh_chain_finish.cuh"

#define CUDA_TRY(call)                                                                                   \
	do {                                                                                                 \
		cudaError_t e_ = (call);                                                                         \
		if (e_ != cudaSuccess) {                                                                         \
			rh_set_error("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));            \
			return RH_ERR_CUDA;                                                                          \
		}                                                                                                \
	} while (0)

namespace {

template <class T> struct dbuf { /* grow-only device buffer */
	T *p = nullptr; size_t cap = 0;
	int reserve(size_t n)
	{
		if (n <= cap) return RH_OK;
		if (p) cudaFree(p);
		p = nullptr; cap = 0;
		size_t want = n + n / 4 + 64;
		cudaError_t e = cudaMalloc((void **)&p, want * sizeof(T));
		if (e != cudaSuccess) { rh_set_error("cudaMalloc(%zu bytes): %s", want * sizeof(T), cudaGetErrorString(e)); return RH_ERR_NOMEM; }
		cap = want;
		return RH_OK;
	}
	int reserve_exact(size_t n) /* no head room: for buffers whose size the caller has already padded */
	{
		if (n <= cap) return RH_OK;
		if (p) cudaFree(p);
		p = nullptr; cap = 0;
		cudaError_t e = cudaMalloc((void **)&p, n * sizeof(T));
		if (e != cudaSuccess) { rh_set_error("cudaMalloc(%zu bytes): %s", n * sizeof(T), cudaGetErrorString(e)); return RH_ERR_NOMEM; }
		cap = n;
		return RH_OK;
	}
	void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct timed_span { cudaEvent_t a, b; int kind; };

} // namespace

enum { T_EVENT = 0, T_SEED, T_SORT, T_CHAIN, T_POST, T_LEN, T_TIES, T_NKIND };

/* One worker = one CUDA stream with its own scratch, arenas and per-read state.  A batch is cut into
 * contiguous read ranges, one per worker; the workers run their chunk rounds concurrently so that the
 * latency-bound tail of one range (late rounds hold few, slow chunks) overlaps the bulk of the others
 * and the host->device copy of the raw samples overlaps compute.  Index, log table and raw buffer are shared. */
struct rh_worker {
	int device = 0;
	rh_params_t P;
	dev_params_t D;
	const rh_index_s *idx = nullptr;
	dev_index_t I;
	cudaStream_t stream = nullptr, own_stream = nullptr, stream2 = nullptr; /* stream2: the heavy lane of run_round */
	/* shared, owned by the context */
	dbuf<float> d_logf; uint32_t logf_n = 0;
	dbuf<uint32_t> d_seqlen;
	const std::vector<uint32_t> *name_order = nullptr;
	const int16_t *raw_ptr = nullptr; /* raw buffer of the current call (the context's or the caller's) */
	/* per-batch */
	dbuf<read_state_t> d_rs;
	dbuf<slot_t> d_slots;
	dbuf<float> d_z, d_events, d_ps, d_pq, d_t1, d_t2;
	dbuf<chunk_norm_t> d_norm;
	dbuf<uint2> d_groups;
	dbuf<slot_t> d_pre;            /* slots of the chunks whose event stage ran ahead of their round */
	uint32_t n_splits = 0;        /* times a read range had to be halved because it did not fit the device */
	int fast_w1 = 0, fast_w2 = 0; /* exhaustive self-test at init: x / w by reciprocal + two FMAs is exact for this window length */
	dbuf<uint32_t> d_peaks, d_seed_hash, d_seed_pos, d_seed_cnt, d_seed_dst;
	dbuf<uint64_t> d_seed_src;
	dbuf<uint8_t> d_arena;
	dbuf<anchor_t> d_carry[2];
	dbuf<unsigned long long> d_counters; /* [0] carry_top, [1] rec_top */
	dbuf<uint32_t> d_err, d_rec_start, d_rec_cnt, d_tie_list, d_tie_count, d_tie_list2, d_tie_count2;
	dbuf<unsigned long long> d_prof; bool prof_on = false; bool trace = false;
	dbuf<rh_map_rec_t> d_recs;
	size_t arena_bytes = 0, sig_budget = 0;
	uint32_t sort_posbits = 0, sort_ridbits = 0, sort_smem_cap = 0;
	double hits_per_key = 0.0;    /* index positions per distinct key: first estimate of a chunk's anchors (scheduler choice) */
	unsigned long long carry_known = 0; /* carry_top of the round's output arena at the last sync */
	uint64_t rec_hint = 0, rec_cap_now = 0; bool retry_same = false; /* record arena overflow (all-vs-all: one record per chain): size it from the count and map the range again */
	std::vector<timed_span> spans;
	std::vector<cudaEvent_t> ev_pool; size_t ev_used = 0;
	rh_gpu_stats_t st;
	char err[512];
};

struct rh_gpu_ctx_s {
	int device = 0;
	rh_params_t P;
	dev_params_t D;
	const rh_index_s *idx = nullptr;
	dev_index_t I;
	/* index storage: the flattened index lives with the index handle (rh_index_s::dev) unless that one is resident on
	 * another GPU, in which case this context owns a private copy */
	rh_index_dev_t own_dev; bool own_dev_valid = false;
	dbuf<uint32_t> d_seqlen, d_namerank;
	dbuf<float> d_logf;
	uint32_t logf_n = 0;
	std::vector<uint32_t> name_order; /* sorted target names (indices) for Rawsamble */
	dbuf<int16_t> d_raw;              /* raw samples of the current host-buffer call, shared by the workers */
	uint32_t sort_posbits = 0, sort_ridbits = 0, sort_smem_cap = 0;
	double hits_per_key = 0.0;
	std::vector<rh_worker *> workers;
	uint32_t n_active = 1;            /* workers used per call */
	void *user_stream = nullptr;
	rh_gpu_stats_t st;
};

namespace {

cudaEvent_t get_event(rh_worker *c)
{
	if (c->ev_used == c->ev_pool.size()) { cudaEvent_t e; cudaEventCreate(&e); c->ev_pool.push_back(e); }
	return c->ev_pool[c->ev_used++];
}
#define RH_TRACE(c, what) do { if ((c)->trace) { cudaError_t e_ = cudaStreamSynchronize((c)->stream); fprintf(stderr, "[trace] %s: %s\n", what, cudaGetErrorString(e_)); } } while (0)
#define RH_TRACE_ON(c, st, what) do { if ((c)->trace) { cudaError_t e_ = cudaStreamSynchronize(st); fprintf(stderr, "[trace] %s: %s\n", what, cudaGetErrorString(e_)); } } while (0)
struct span_guard {
	rh_worker *c; timed_span s; int n_launch;
	cudaStream_t st;
	span_guard(rh_worker *c_, int kind, int n_launch_ = 1, cudaStream_t st_ = nullptr) : c(c_), n_launch(n_launch_), st(st_ ? st_ : c_->stream) { s.kind = kind; s.a = get_event(c); s.b = get_event(c); cudaEventRecord(s.a, st); }
	~span_guard() { cudaEventRecord(s.b, st); c->spans.push_back(s); c->st.kernel_launches += n_launch; if (s.kind == T_EVENT) c->st.event_kernel_launches++; }
};

void fill_dev_params(const rh_params_t &P, dev_params_t &D)
{
	memset(&D, 0, sizeof(D));
	D.w = P.w; D.e = P.e; D.q = P.q; D.k = P.k;
	D.diff = P.diff; D.fine_min = P.fine_min; D.fine_max = P.fine_max; D.fine_range = P.fine_range;
	D.w1 = P.window_length1; D.w2 = P.window_length2; D.thr1 = P.threshold1; D.thr2 = P.threshold2; D.height = P.peak_height;
	D.min_events = P.min_events; D.mid_occ = P.mid_occ;
	D.bw = P.bw; D.max_t = P.max_target_gap_length; D.max_q = P.max_query_gap_length; D.max_iter = P.max_chain_iter;
	D.max_skip = P.max_num_skips; D.min_cnt = P.min_num_anchors; D.min_sc = P.min_chaining_score; D.min_sc2 = P.min_chaining_score2;
	/* rmap.cpp:318: evaluated in double, narrowed to float */
	D.pen_gap = (float)(P.chain_gap_scale * 0.01 * (P.e + P.k - 1));
	D.pen_skip = (float)(P.chain_skip_scale * 0.01 * (P.e + P.k - 1));
	D.mask_level = P.mask_level; D.mask_len = P.mask_len; D.pri_ratio = P.pri_ratio; D.best_n = P.best_n;
	D.min_strand_sc = (int)(P.max_target_gap_length * 0.8); /* rmap.cpp:354 */
	D.w_bestq = P.w_bestq; D.w_bestmq = P.w_bestmq; D.w_bestmc = P.w_bestmc; D.w_threshold = P.w_threshold;
	D.min_mapq = P.min_mapq; D.max_num_chunk = P.max_num_chunk; D.chunk_size = P.chunk_size; D.sample_per_base = P.sample_per_base;
	D.ava = (P.map_flag & RH_M_ALL_CHAINS) ? 1 : 0; D.noadapt = (P.map_flag & RH_M_NO_ADAPTIVE) ? 1 : 0;
	D.sig_target = (P.idx_flag & RH_I_SIG_TARGET) ? 1 : 0;
}

template <class T> int upload(dbuf<T> &d, const std::vector<T> &h, cudaStream_t s)
{
	int rc = d.reserve(h.size() ? h.size() : 1);
	if (rc) return rc;
	if (h.size()) CUDA_TRY(cudaMemcpyAsync(d.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s));
	return RH_OK;
}

struct batch_in {
	uint32_t n;
	const int16_t *const *raw; const uint64_t *raw_len;   /* host pointers, or null when d_raw given */
	const void *d_raw; const uint64_t *raw_off;            /* device-resident variant                  */
	const double *offset, *range, *digitisation;
	const char *const *names;
};

int check_params(const rh_params_t &P)
{
	if (P.window_length1 > 15 || P.window_length2 > 15) { rh_set_error("segmentation windows > 15 are not supported"); return RH_ERR_ARG; }
	if (P.e < 1 || P.q < 1 || P.e * P.q > 64) { rh_set_error("bad e/q"); return RH_ERR_ARG; }
	if (P.w < 0 || P.w > RH_MAX_W) { rh_set_error("minimizer window w must be in [0, %d]", RH_MAX_W); return RH_ERR_ARG; }
	if (P.n != 0) { rh_set_error("BLEND seeding (n>0) is disabled in the reference and unsupported here"); return RH_ERR_ARG; }
	if (P.min_num_anchors < 2) { rh_set_error("min_num_anchors < 2 is not supported (chains are at least two anchors; scratch sizing relies on it)"); return RH_ERR_ARG; }
	return RH_OK;
}

/* ---- one chunk round over a set of slots ---------------------------------------------------- */
struct round_io {
	std::vector<slot_t> slots;         /* host copy */
	int tap = 0;
	rh_tap_t *tap_out = nullptr; uint64_t *tap_off = nullptr; /* running offsets: ev, seed, anc, u, ca, reg */
	uint32_t tap_chunk = 0;
	bool events_done = false;          /* slots already carry the event stage's results (computed ahead, event_stage()) */
	/* streaming scheduler: slots [0, n_mandatory) must run in this round (their reads carry chain anchors in the round's
	 * input arena); the slots after them are first chunks of reads waiting for admission and run only as far as they fill
	 * the last anchor-arena group (at most max_optional of them).  n_run = slots that did run. */
	uint32_t n_mandatory = 0xffffffffu, max_optional = 0, n_run = 0;
};

/* The event stage over `hs.size()` chunks whose slots live at d_slots: assigns the scratch offsets, uploads the slots,
 * launches.  groups = runs of consecutive slots that are successive chunks of ONE read (their sums chain). */
int event_stage(rh_worker *c, std::vector<slot_t> &hs, dbuf<slot_t> &d_slots, const std::vector<uint2> &groups, bool preset_off = false)
{
	const uint32_t ns = (uint32_t)hs.size();
	if (ns == 0) return RH_OK;
	cudaStream_t s = c->stream;
	uint64_t zt = 0, et = 0;
	uint32_t max_len = 0;
	for (slot_t &sl : hs) {
		sl.z_off = zt; zt += ((uint64_t)sl.chunk_len + 3) & ~3ULL; /* chunks start 16-byte aligned */
		if (preset_off) et = std::max<uint64_t>(et, sl.e_off + sl.e_cap); /* the caller placed the chunk's scratch (a pool that outlives this call) */
		else {
			sl.e_cap = (uint32_t)((uint64_t)sl.chunk_len * 2 / 3 + 8);
			sl.e_off = et; et += sl.e_cap;
		}
		max_len = std::max(max_len, sl.chunk_len);
	}
	int rc;
	const bool streaming = max_len <= EVS_MAXN && !getenv("RH_EVENT_OLD"); /* ordinary chunks: no intermediates in HBM (rh_event_stream.cuh) */
	if (preset_off && (!streaming || et > c->d_events.cap || et > c->d_seed_src.cap)) { rh_set_error("internal: event scratch pool too small"); return RH_ERR_ARG; }
	if ((rc = c->d_events.reserve(et)) || (rc = c->d_peaks.reserve(et)) || (rc = c->d_seed_hash.reserve(et)) || (rc = c->d_seed_pos.reserve(et)) ||
	    (rc = c->d_seed_cnt.reserve(et)) || (rc = c->d_seed_dst.reserve(et)) || (rc = c->d_seed_src.reserve(et))) return rc;
	if ((rc = upload(d_slots, hs, s))) return rc;
	c->st.h2d_bytes += ns * sizeof(slot_t);
	if (streaming) {
		if ((rc = c->d_norm.reserve(ns)) || (rc = upload(c->d_groups, groups, s))) return rc;
		c->st.h2d_bytes += groups.size() * sizeof(uint2);
		evt_args_t e1;
		e1.raw = c->raw_ptr; e1.rs = c->d_rs.p; e1.slots = d_slots.p; e1.n_slots = ns; e1.norm = c->d_norm.p;
		e1.groups = c->d_groups.p; e1.n_groups = (uint32_t)groups.size();
		e1.peaks = c->d_peaks.p; e1.events = c->d_events.p; e1.seed_hash = c->d_seed_hash.p; e1.seed_pos = c->d_seed_pos.p;
		e1.min_hash = c->d_seed_cnt.p; e1.min_pos = c->d_seed_dst.p; /* free until k_seed_count */
		e1.fast_w1 = c->fast_w1; e1.fast_w2 = c->fast_w2;
		span_guard g(c, T_EVENT, 3);
		k_evt_sums<<<(e1.n_groups * 32 + 127) / 128, 128, 0, s>>>(e1);
		const uint32_t gb = (ns + EVS_THREADS - 1) / EVS_THREADS;
		if (c->fast_w1 && c->fast_w2) k_evt_stream<true, true><<<gb, EVS_THREADS, 0, s>>>(e1, c->D);
		else if (c->fast_w1) k_evt_stream<true, false><<<gb, EVS_THREADS, 0, s>>>(e1, c->D);
		else if (c->fast_w2) k_evt_stream<false, true><<<gb, EVS_THREADS, 0, s>>>(e1, c->D);
		else k_evt_stream<false, false><<<gb, EVS_THREADS, 0, s>>>(e1, c->D);
		k_evt_finish<<<ns, EVF2_THREADS, 0, s>>>(e1, c->D);
	} else {
		for (const uint2 &g : groups) if (g.y != 1) { rh_set_error("internal: long chunks are detected one per read and launch"); return RH_ERR_ARG; }
		if ((rc = c->d_z.reserve(zt))) return rc;
		if ((rc = c->d_ps.reserve(zt + 4 * (uint64_t)ns + 4)) || (rc = c->d_pq.reserve(zt + 4 * (uint64_t)ns + 4)) || (rc = c->d_t1.reserve(zt + 4 * (uint64_t)ns + 4)) || (rc = c->d_t2.reserve(zt + 4 * (uint64_t)ns + 4))) return rc;
		sig_args_t a1;
		a1.raw = c->raw_ptr; a1.rs = c->d_rs.p; a1.slots = d_slots.p; a1.n_slots = ns;
		a1.z = c->d_z.p; a1.ps = c->d_ps.p; a1.pq = c->d_pq.p; a1.t1 = c->d_t1.p; a1.t2 = c->d_t2.p;
		a1.peaks = c->d_peaks.p; a1.events = c->d_events.p; a1.seed_hash = c->d_seed_hash.p; a1.seed_pos = c->d_seed_pos.p;
		a1.min_hash = c->d_seed_cnt.p; a1.min_pos = c->d_seed_dst.p; /* free until k_seed_count */
		a1.prof = c->prof_on ? c->d_prof.p : nullptr;
		/* whole-read chunks (Rawsamble) and other long chunks: seven launches back to back, timed as one span */
		span_guard g(c, T_EVENT, 7);
		k_sig_norm<<<(ns * 32 + 127) / 128, 128, 0, s>>>(a1);
		k_sig_prefix<<<(ns + 127) / 128, 128, 0, s>>>(a1);
		k_sig_tstat<<<ns, 256, 0, s>>>(a1, c->D);
		k_sig_peaks<<<(ns + 127) / 128, 128, 0, s>>>(a1, c->D);
		k_sig_events_fast<<<ns, EV_THREADS, 0, s>>>(a1);
		k_sig_events<<<ns, EV_THREADS, 0, s>>>(a1);
		k_sig_sketch<<<(ns + 127) / 128, 128, 0, s>>>(a1, c->D);
	}
	return RH_OK;
}

int run_round(rh_worker *c, round_io &io, int carry_in_idx)
{
	const uint32_t ns = (uint32_t)io.slots.size();
	if (ns == 0) return RH_OK;
	cudaStream_t s = c->stream;
	int rc;
	if (io.events_done) { /* computed ahead: the slots carry their counts and scratch offsets already */
		if ((rc = upload(c->d_slots, io.slots, s))) return rc;
		c->st.h2d_bytes += ns * sizeof(slot_t);
	} else {
		std::vector<uint2> groups(ns);
		for (uint32_t i = 0; i < ns; ++i) groups[i] = make_uint2(i, 1u);
		if ((rc = event_stage(c, io.slots, c->d_slots, groups))) return rc;
	}
	k2_args_t a2;
	a2.slots = c->d_slots.p; a2.n_slots = ns; a2.rs = c->d_rs.p;
	a2.seed_hash = c->d_seed_hash.p; a2.seed_pos = c->d_seed_pos.p; a2.seed_cnt = c->d_seed_cnt.p; a2.seed_src = c->d_seed_src.p; a2.seed_dst = c->d_seed_dst.p;
	a2.arena = c->d_arena.p; a2.carry_in = c->d_carry[carry_in_idx].p; a2.carry_out = c->d_carry[carry_in_idx ^ 1].p;
	const uint32_t wblocks = (ns * RH_WARP + 255) / 256;
	{
		span_guard g(c, T_SEED);
		k_seed_count<<<wblocks, 256, 0, s>>>(a2, c->I, c->D);
	}
	RH_TRACE(c, "k_seed_count");
	CUDA_TRY(cudaMemcpyAsync(io.slots.data(), c->d_slots.p, ns * sizeof(slot_t), cudaMemcpyDeviceToHost, s));
	CUDA_TRY(cudaMemcpyAsync(&c->carry_known, c->d_counters.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
	CUDA_TRY(cudaStreamSynchronize(s));
	c->st.d2h_bytes += ns * sizeof(slot_t) + 8;
	dbuf<anchor_t> &carry_out = c->d_carry[carry_in_idx ^ 1];

	/* anchor-arena groups */
	k3_args_t a3;
	a3.slots = nullptr; a3.n_slots = 0; a3.rs = c->d_rs.p; a3.arena = c->d_arena.p;
	a3.carry_out = carry_out.p; a3.carry_top = c->d_counters.p; a3.carry_cap = carry_out.cap;
	a3.logf_tab = c->d_logf.p; a3.logf_n = c->logf_n;
	a3.recs = c->d_recs.p; a3.rec_top = c->d_counters.p + 1; a3.rec_cap = c->rec_cap_now;
	a3.rec_start = c->d_rec_start.p; a3.rec_cnt = c->d_rec_cnt.p; a3.seq_len = c->d_seqlen.p; a3.tap = io.tap; a3.err = c->d_err.p; a3.prof = c->prof_on ? c->d_prof.p : nullptr; a3.prof_replay = getenv("RH_PROF_TIES_ONLY") ? 0 : 1;

	/* The layout of the round: order of the chunks (heaviest first), arena groups, heavy lane — rh_plan_round_impl (rh_host.cpp). */
	const uint32_t n_mand = std::min(io.n_mandatory, ns);
	rh_round_plan_t plan;
	{
		std::vector<uint32_t> sizes(ns);
		for (uint32_t q = 0; q < ns; ++q) sizes[q] = io.slots[q].n_anchors;
		const bool heavy_ok = !io.tap && io.n_mandatory != 0xffffffffu && c->stream2 && !getenv("RH_NO_HEAVY_LANE");
		if ((rc = rh_plan_round_impl(sizes.data(), ns, n_mand, io.max_optional, c->arena_bytes, !io.tap, heavy_ok, &plan))) return rc;
		std::vector<slot_t> ordered(n_mand);
		for (uint32_t q = 0; q < n_mand; ++q) ordered[q] = io.slots[plan.order[q]]; /* waiting reads keep their order */
		std::copy(ordered.begin(), ordered.end(), io.slots.begin());
		for (uint32_t q = 0; q < plan.n_run; ++q) io.slots[q].a_off = plan.a_off[q];
	}

	unsigned long long committed = c->carry_known; /* upper bound of carry_top once everything launched so far has run */
	/* front half of a group on stream `st`: slots up, anchors, sort (+ exact tie order), chaining DP */
	auto front = [&](cudaStream_t st, uint32_t g0, uint32_t g1, dbuf<uint32_t> &tie_list, dbuf<uint32_t> &tie_count) -> int {
		const uint32_t gn = g1 - g0;
		int rc2;
		CUDA_TRY(cudaMemcpyAsync(c->d_slots.p + g0, io.slots.data() + g0, gn * sizeof(slot_t), cudaMemcpyHostToDevice, st));
		c->st.h2d_bytes += gn * sizeof(slot_t);
		k2_args_t b2 = a2; b2.slots = c->d_slots.p + g0; b2.n_slots = gn;
		k3_args_t b3 = a3; b3.slots = c->d_slots.p + g0; b3.n_slots = gn;
		const uint32_t gw = (gn * RH_WARP + 255) / 256;
		if (c->trace) fprintf(stderr, "[trace] group of %u slots at %u%s\n", gn, g0, st == s ? "" : " (heavy lane)");
		{ span_guard g(c, T_SEED, 1, st); k_seed_expand<<<gw, 256, 0, st>>>(b2, c->I, c->D); }
		RH_TRACE_ON(c, st, "k_seed_expand");
		if ((rc2 = tie_list.reserve(gn)) || (rc2 = tie_count.reserve(1))) return rc2;
		CUDA_TRY(cudaMemsetAsync(tie_count.p, 0, 4, st));
		sort_args_t as; as.slots = b3.slots; as.n_slots = gn; as.arena = c->d_arena.p; as.tie_list = tie_list.p; as.tie_count = tie_count.p; as.prof = c->prof_on ? c->d_prof.p : nullptr; as.err = c->d_err.p;
		as.posbits = c->sort_posbits; as.ridbits = c->sort_ridbits;
		uint32_t maxn = 0;
		for (uint32_t q = g0; q < g1; ++q) if (!io.slots[q].gated) maxn = std::max(maxn, io.slots[q].n_anchors);
		{ /* chunks that fit shared memory (nearly all against a small index) sort there; the rest take the global-memory kernel */
			const uint32_t cap = std::min(c->sort_smem_cap, (maxn + 31) & ~31u); /* small groups: less shared memory, more CTAs per SM */
			span_guard g(c, T_SORT, (cap ? 1 : 0) + (maxn > cap ? 1 : 0), st);
			if (cap) k_sort_smem<<<gn, SB_THREADS, sort_smem_bytes(cap), st>>>(as, cap);
			if (maxn > cap) k_sort_block<<<gn, SORT_THREADS, 0, st>>>(as, cap);
		}
		RH_TRACE_ON(c, st, "k_sort");
		{
			/* digit bytes live in the chunk's global scratch, shared memory holds only the walk tables: measured faster
			 * than keeping the bytes in shared memory (149 vs 184 ms per 100 k reads) because twice as many chunks are
			 * resident per SM and the walk is latency bound either way */
			span_guard g(c, T_TIES, 1, st);
			k_sort_ties<<<gn, TIE_THREADS, tie_smem_bytes(0, 1), st>>>(as, 0u, 0u, TIE_LARGE_N, 1u);
			if (maxn >= TIE_LARGE_N) k_sort_ties<<<gn, TIE_THREADS, tie_smem_bytes(0, TIE_SIDE_WALKS), st>>>(as, 0u, TIE_LARGE_N, 0xffffffffu, (uint32_t)TIE_SIDE_WALKS);
		}
		RH_TRACE_ON(c, st, "k_sort_ties");
		if (io.tap) { /* sorted anchor list of the single tapped slot */
			rh_tap_t *T = io.tap_out; const slot_t &sl = io.slots[0];
			if (!sl.gated) {
				if (io.tap_off[2] + sl.n_anchors > T->cap_anchors) return RH_ERR_NOMEM;
				CUDA_TRY(cudaMemcpyAsync(T->anchors + 2 * io.tap_off[2], c->d_arena.p + sl.a_off, (size_t)sl.n_anchors * 16, cudaMemcpyDeviceToHost, st));
				CUDA_TRY(cudaStreamSynchronize(st));
				io.tap_off[2] += sl.n_anchors;
			}
		}
		if (c->prof_on) { /* RH_PROF=1: distribution of chunk sizes and tie chunks in this group (debug aid, adds a sync) */
			std::vector<slot_t> dbg(gn);
			cudaMemcpyAsync(dbg.data(), c->d_slots.p + g0, gn * sizeof(slot_t), cudaMemcpyDeviceToHost, st);
			cudaStreamSynchronize(st);
			std::vector<uint32_t> na, nt;
			for (const slot_t &q : dbg) { if (q.gated || q.n_anchors == 0) continue; na.push_back(q.n_anchors); if (q.n_ties) nt.push_back(q.n_anchors); }
			std::sort(na.begin(), na.end()); std::sort(nt.begin(), nt.end());
			auto pc = [](const std::vector<uint32_t> &v, double f) { return v.empty() ? 0u : v[std::min(v.size() - 1, (size_t)(f * v.size()))]; };
			fprintf(stderr, "[RH_PROF] group: %u slots, %zu chained (n p50=%u p90=%u p99=%u max=%u), %zu with ties (n p50=%u p90=%u max=%u)\n",
			        gn, na.size(), pc(na, .5), pc(na, .9), pc(na, .99), pc(na, 1.0), nt.size(), pc(nt, .5), pc(nt, .9), pc(nt, 1.0));
		}
		{ span_guard g(c, T_CHAIN, 1, st); k_chain_dp<<<gn, DP_THREADS, 0, st>>>(b3, c->D); }
		RH_TRACE_ON(c, st, "k_chain_dp");
		return RH_OK;
	};
	/* back half: room in the output carry arena for what the group's chains carry into the next round, then backtrack,
	 * regions, decisions.  A chain of m anchors uses m-1 distinct anchors that have a DP predecessor, so a chunk carries at
	 * most min(n_anchors, 2 n_link) anchors (a gated chunk: its prev_n). */
	auto back = [&](cudaStream_t st, uint32_t g0, uint32_t g1) -> int {
		const uint32_t gn = g1 - g0;
		int rc2;
		std::vector<slot_t> after(gn);
		CUDA_TRY(cudaMemcpyAsync(after.data(), c->d_slots.p + g0, gn * sizeof(slot_t), cudaMemcpyDeviceToHost, st));
		CUDA_TRY(cudaStreamSynchronize(st));
		c->st.d2h_bytes += gn * sizeof(slot_t);
		unsigned long long need = 0;
		for (uint32_t q = 0; q < gn; ++q) {
			const slot_t &sl = after[q];
			need += (sl.gated || sl.n_anchors == 0) ? sl.n_anchors : std::min<unsigned long long>(sl.n_anchors, 2ULL * sl.n_link);
		}
		if (committed + need > carry_out.cap) { /* grow: nothing may be writing the arena while it moves */
			CUDA_TRY(cudaDeviceSynchronize());
			unsigned long long top = 0;
			CUDA_TRY(cudaMemcpy(&top, c->d_counters.p, sizeof(top), cudaMemcpyDeviceToHost));
			dbuf<anchor_t> bigger;
			if ((rc2 = bigger.reserve_exact((size_t)((committed + need) * 9 / 8 + 4096)))) return rc2;
			if (top) CUDA_TRY(cudaMemcpy(bigger.p, carry_out.p, top * sizeof(anchor_t), cudaMemcpyDeviceToDevice));
			carry_out.release();
			carry_out = bigger;
			a2.carry_out = carry_out.p; a3.carry_out = carry_out.p; a3.carry_cap = carry_out.cap;
		}
		committed += need;
		k3_args_t b3 = a3; b3.slots = c->d_slots.p + g0; b3.n_slots = gn;
		{ span_guard g(c, T_POST, 2, st); k_chain_finish<<<gn, FIN_THREADS, 0, st>>>(b3, c->D); RH_TRACE_ON(c, st, "k_chain_finish"); k_chain_decide<<<(gn + DEC_WARPS - 1) / DEC_WARPS, DEC_WARPS * 32, 0, st>>>(b3, c->D); }
		RH_TRACE_ON(c, st, "k_chain_decide");
		if (io.tap) {
			CUDA_TRY(cudaMemcpyAsync(io.slots.data(), c->d_slots.p, sizeof(slot_t), cudaMemcpyDeviceToHost, st));
			CUDA_TRY(cudaStreamSynchronize(st));
		}
		return RH_OK;
	};

	/* Heavy lane (streaming scheduler): the largest chunks of the iteration (late chunks of reads that do not map: each
	 * drags all chain anchors of its predecessors along) take several times as long as an average chunk in every kernel,
	 * and a kernel lasts as long as its slowest chunk.  They run as a group of their own on a second stream, in a slice at
	 * the top of the arena, next to the ordinary groups: front half first, back half once the first ordinary group's back
	 * half has been launched. */
	cudaStream_t sh = c->stream2;
	const rh_round_group_t *heavy = (!plan.groups.empty() && plan.groups[0].heavy) ? &plan.groups[0] : nullptr;
	bool heavy_back_due = false;
	if (heavy) {
		if ((rc = front(sh, heavy->first, heavy->first + heavy->count, c->d_tie_list2, c->d_tie_count2))) return rc;
		heavy_back_due = true;
	}
	for (const rh_round_group_t &g : plan.groups) {
		if (g.heavy) continue;
		if ((rc = front(s, g.first, g.first + g.count, c->d_tie_list, c->d_tie_count))) return rc;
		if ((rc = back(s, g.first, g.first + g.count))) return rc;
		if (heavy_back_due) { heavy_back_due = false; if ((rc = back(sh, heavy->first, heavy->first + heavy->count))) return rc; } /* its front half ran beside this group's */
	}
	if (heavy_back_due) { if ((rc = back(sh, heavy->first, heavy->first + heavy->count))) return rc; }
	if (heavy) { /* the round ends when both lanes have */
		cudaEvent_t ev = get_event(c);
		CUDA_TRY(cudaEventRecord(ev, sh));
		CUDA_TRY(cudaStreamWaitEvent(s, ev, 0));
	}
	const uint32_t n_total = plan.n_run;
	io.n_run = n_total;
	for (uint32_t q = 0; q < n_total; ++q) {
		const slot_t &sl = io.slots[q];
		c->st.raw_samples_consumed += sl.raw_used; c->st.n_chains += sl.n_u;
		c->st.n_chunks++; c->st.n_events += sl.n_events; c->st.n_seeds += sl.n_seeds; c->st.n_anchors += sl.n_anchors;
		if (!io.events_done) { c->st.event_stage_samples += sl.raw_used; c->st.event_stage_seeds += sl.n_seeds; }
	}
	return RH_OK;
}

int collect_spans(rh_worker *c)
{
	if (c->prof_on) {
		unsigned long long h[64];
		cudaMemcpy(h, c->d_prof.p, sizeof(h), cudaMemcpyDeviceToHost);
		cudaMemset(c->d_prof.p, 0, sizeof(h));
		fprintf(stderr, "[RH_PROF] Mcycles:");
		for (int i = 0; i < 64; ++i) if (h[i]) fprintf(stderr, " [%d]=%.1f", i, h[i] / 1e6);
		fprintf(stderr, "  (chunks=%llu)\n", (unsigned long long)c->st.n_chunks);
	}
	double ms[T_NKIND] = {0};
	for (const timed_span &sp : c->spans) { float t = 0; if (cudaEventElapsedTime(&t, sp.a, sp.b) == cudaSuccess) ms[sp.kind] += t; }
	c->st.ms_event_kernel += ms[T_EVENT]; c->st.ms_seed += ms[T_SEED]; c->st.ms_sort += ms[T_SORT] + ms[T_TIES]; c->st.ms_sort_ties += ms[T_TIES];
	c->st.ms_chain += ms[T_CHAIN]; c->st.ms_post += ms[T_POST] + ms[T_LEN];
	c->spans.clear(); c->ev_used = 0;
	return RH_OK;
}

int check_dev_err(rh_worker *c)
{
	uint32_t e = 0;
	CUDA_TRY(cudaMemcpyAsync(&e, c->d_err.p, 4, cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	if (e == 0) return RH_OK;
	if (e == 5) { /* rec_top kept counting past the capacity: it is the number of records the range needs */
		unsigned long long tops[2] = {0, 0};
		cudaMemcpy(tops, c->d_counters.p, sizeof(tops), cudaMemcpyDeviceToHost);
		if (tops[1] + tops[1] / 4 + 1024 > c->rec_hint) { c->rec_hint = tops[1] + tops[1] / 4 + 1024; c->retry_same = true; }
		rh_set_error("device reported: record arena exhausted (%llu records)", tops[1]);
		return RH_ERR_NOMEM;
	}
	if (e == 2) {
		unsigned long long tops[2] = {0, 0};
		cudaMemcpy(tops, c->d_counters.p, sizeof(tops), cudaMemcpyDeviceToHost);
		rh_set_error("device reported: carry arena exhausted (top %llu, capacities %zu / %zu)", tops[0], c->d_carry[0].cap, c->d_carry[1].cap);
		return RH_ERR_NOMEM;
	}
	const char *what = e == 2 ? "carry arena exhausted" : e == 3 ? "region scratch exhausted" : e == 4 ? "logf argument outside the exact table" : e == 5 ? "record arena exhausted" : e == 6 ? "sort replay found inconsistent tables" : "device error";
	rh_set_error("device reported: %s (code %u)", what, e);
	return e == 4 ? RH_ERR_ARG : e == 6 ? RH_ERR_CUDA : RH_ERR_NOMEM;
}

/* set up per-read state on the device; raw already resident in c->d_raw */
int setup_reads(rh_worker *c, const batch_in &in, const std::vector<uint64_t> &beg, const std::vector<uint64_t> &len, std::vector<uint32_t> &l_sig)
{
	const uint32_t n = in.n;
	std::vector<read_state_t> rs(n);
	/* Rawsamble: #target names <= read name, against the sorted target-name list */
	const bool ava = c->D.ava != 0;
	for (uint32_t i = 0; i < n; ++i) {
		read_state_t &r = rs[i];
		memset(&r, 0, sizeof(r));
		r.raw_beg = beg[i]; r.raw_end = beg[i] + len[i]; r.cursor = beg[i];
		r.cal_offset = in.offset[i];
		r.cal_scale = (double)(float)(in.range[i] / in.digitisation[i]); /* rsig.c:493: float scale */
		if (ava && in.names) {
			const char *q = in.names[i];
			uint32_t lo = 0, hi = (uint32_t)c->name_order->size();
			while (lo < hi) { uint32_t mid = (lo + hi) / 2; if (strcmp(c->idx->names[(*c->name_order)[mid]].c_str(), q) <= 0) lo = mid + 1; else hi = mid; }
			r.name_ub = lo;
		}
	}
	int rc;
	if ((rc = upload(c->d_rs, rs, c->stream))) return rc;
	c->st.h2d_bytes += n * sizeof(read_state_t);
	if ((rc = c->d_rec_start.reserve(n)) || (rc = c->d_rec_cnt.reserve(n))) return rc;
	CUDA_TRY(cudaMemsetAsync(c->d_rec_cnt.p, 0xff, n * 4, c->stream));
	CUDA_TRY(cudaMemsetAsync(c->d_rec_start.p, 0, n * 4, c->stream));
	CUDA_TRY(cudaMemsetAsync(c->d_counters.p, 0, 2 * sizeof(unsigned long long), c->stream));
	CUDA_TRY(cudaMemsetAsync(c->d_err.p, 0, 4, c->stream));
	{
		span_guard g(c, T_LEN);
		k_filtered_len<<<(n * RH_WARP + 255) / 256, 256, 0, c->stream>>>(c->raw_ptr, c->d_rs.p, n);
	}
	CUDA_TRY(cudaMemcpyAsync(rs.data(), c->d_rs.p, n * sizeof(read_state_t), cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	c->st.d2h_bytes += n * sizeof(read_state_t);
	l_sig.resize(n);
	for (uint32_t i = 0; i < n; ++i) l_sig[i] = rs[i].l_sig;
	return RH_OK;
}

int map_resident(rh_worker *c, const batch_in &in, const std::vector<uint64_t> &beg, const std::vector<uint64_t> &len, rh_map_rec_t **recs_out, uint64_t *n_recs_out)
{
	const uint32_t n = in.n;
	std::vector<uint32_t> l_sig;
	if (c->stream2) cudaStreamSynchronize(c->stream2); /* a call that failed half way (and is being retried in halves) may have left its heavy lane running */
	int rc = setup_reads(c, in, beg, len, l_sig);
	if (rc) return rc;
	const rh_params_t &P = c->P;
	const bool noadapt = c->D.noadapt != 0;
	const uint32_t max_chunk = noadapt ? 1u : P.max_num_chunk;
	/* first guess: one record per read, 64 in all-vs-all mode (every chain is a record); an overflow re-sizes the arena from
	 * the device's count and the range is mapped again (RH_REC_CAP_TEST: a deliberately small first guess, for the test) */
	uint64_t rec_cap = (uint64_t)n * (c->D.ava ? 64 : 1) + 1024;
	if (const char *e = getenv("RH_REC_CAP_TEST")) rec_cap = (uint64_t)std::max(1, atoi(e));
	rec_cap = std::max<uint64_t>(rec_cap, c->rec_hint);
	if ((rc = c->d_recs.reserve(rec_cap))) return rc;
	c->rec_cap_now = rec_cap; /* not the buffer's (padded) capacity: the test hook needs the overflow */

	std::vector<uint32_t> active;
	for (uint32_t i = 0; i < n; ++i) if (l_sig[i] > 0) active.push_back(i);
	std::vector<uint32_t> rec_cnt(n);
	auto chunk_of = [&](uint32_t r, uint32_t round, uint32_t *len) { /* chunk geometry of map_worker_for, rmap.cpp:402-421 */
		const uint32_t qlen = l_sig[r];
		const uint32_t l_chunk = (P.chunk_size > qlen || noadapt) ? qlen : P.chunk_size;
		const uint64_t s_qs = (uint64_t)round * l_chunk;
		if (s_qs >= qlen) return false;
		*len = (uint32_t)std::min<uint64_t>(l_chunk, qlen - s_qs);
		return true;
	};
	/* Event detection never looks at mapping results (only at the signal and the read's running sums), so the event stage
	 * of the next `ahead` chunks of every active read runs in ONE launch set ("wave") ahead of their rounds: launches large
	 * enough to fill the GPU instead of one small launch set per round, at the price of events for chunks a read may never
	 * map.  Long chunks (whole-read Rawsamble) are detected one per read and launch. */
	bool long_chunks = false;
	for (uint32_t r : active) { uint32_t len0; if (chunk_of(r, 0, &len0) && len0 > EVS_MAXN) { long_chunks = true; break; } }
	uint32_t ahead = 1;
	if (!long_chunks && !noadapt && !getenv("RH_EVENT_OLD")) {
		const char *e = getenv("RH_EVENT_AHEAD");
		ahead = e ? (uint32_t)std::max(1, atoi(e)) : (uint32_t)std::max<size_t>(1, (size_t)131072 / std::max<size_t>(active.size(), 1));
		ahead = std::min(ahead, max_chunk);
	}
	/* Which loop: the streaming scheduler pays when the anchor arena holds far fewer chunks than there are reads (large
	 * indexes: 5·10⁵ anchors per chunk against 3 Gb), so that plain chunk rounds would end in many nearly empty launches.
	 * When a round of the whole batch is a handful of arena groups (small indexes), rounds launch less and are faster.
	 * First estimate of a chunk's anchors: seeds per chunk (one per ~15 samples) x index positions per key x 1.7 (measured:
	 * 1.85 at 12 Mb, 1.7 at 3.09 Gb — frequent keys are hit more often).  RH_SCHED_ROUNDS / RH_SCHED_STREAM force either. */
	bool stream_sched = ahead > 1;
	if (stream_sched && !getenv("RH_SCHED_STREAM")) {
		const double est_anchors = 1.7 * ((double)P.chunk_size / 15.0) * std::max(c->hits_per_key, 1.0);
		const double capacity = (double)c->arena_bytes / (double)slot_region_bytes((uint64_t)est_anchors);
		if ((double)active.size() <= 4.0 * capacity) stream_sched = false;
	}
	if (getenv("RH_SCHED_ROUNDS")) stream_sched = false;
	if (stream_sched) {
		/* ---- streaming scheduler -------------------------------------------------------------------------------------
		 * Reads are independent and a read's chunks are sequential (chunk c+1 chains onto the anchors chunk c carried over),
		 * so a late chunk round holds few, large, slow chunks: run as rounds of the whole batch, most launches would be
		 * nearly empty and still last as long as their slowest chunk.  Instead every iteration runs the next chunk of every
		 * read in flight (mandatory: their carried anchors live in this iteration's input arena) and tops the last
		 * anchor-arena group up with first chunks of reads that have not started yet.  The event stage of a cohort of
		 * waiting reads (all their chunks: it never looks at mapping results) runs ahead in one launch set; its seeds live
		 * in a pool block owned by the read until its final record is out. */
		const uint32_t chunk_cap = max_chunk;
		const uint32_t e_cap_max = (uint32_t)((uint64_t)P.chunk_size * 2 / 3 + 8);
		uint32_t max_inflight = 3072, cohort = 4096; /* cohort: event-stage launches of ~40 000 chunks */
		if (const char *e = getenv("RH_MAX_INFLIGHT")) max_inflight = (uint32_t)std::max(1, atoi(e));
		if (const char *e = getenv("RH_COHORT")) cohort = (uint32_t)std::max(1, atoi(e));
		const uint32_t n_blocks = (uint32_t)std::min<size_t>(active.size(), (size_t)max_inflight + cohort + cohort / 2);
		const uint64_t pool_entries = (uint64_t)n_blocks * chunk_cap * e_cap_max;
		if ((rc = c->d_events.reserve_exact(pool_entries)) || (rc = c->d_peaks.reserve_exact(pool_entries)) || (rc = c->d_seed_hash.reserve_exact(pool_entries)) || (rc = c->d_seed_pos.reserve_exact(pool_entries)) ||
		    (rc = c->d_seed_cnt.reserve_exact(pool_entries)) || (rc = c->d_seed_dst.reserve_exact(pool_entries)) || (rc = c->d_seed_src.reserve_exact(pool_entries))) return rc;
		std::vector<slot_t> table((size_t)n_blocks * chunk_cap);   /* host copy of every pool block's slots after their event stage */
		std::vector<uint32_t> free_blocks(n_blocks);
		for (uint32_t b = 0; b < n_blocks; ++b) free_blocks[b] = n_blocks - 1 - b;
		std::vector<uint32_t> blk(n, 0xffffffffu), next_cc(n, 0), n_cc(n, 0);
		std::vector<uint32_t> inflight, ready;  /* reads with a chunk behind them / reads whose events are staged */
		size_t unseen = 0;                      /* next entry of `active` (reads with signal, input order) not yet staged */
		const std::vector<uint32_t> order = active;
		std::vector<uint32_t> rec_now(n);
		uint32_t iter = 0;
		const bool sched_verbose = getenv("RH_SCHED_VERBOSE") != NULL;
		const auto t_sched0 = std::chrono::steady_clock::now();
		while (!inflight.empty() || !ready.empty() || unseen < order.size()) {
			/* stage the next cohort while few reads wait */
			if (ready.size() < cohort / 2 && unseen < order.size() && !free_blocks.empty()) {
				const size_t take = std::min<size_t>(std::min<size_t>(cohort, order.size() - unseen), free_blocks.size());
				std::vector<slot_t> hs; std::vector<uint2> groups; std::vector<uint32_t> who;
				hs.reserve(take * chunk_cap);
				for (size_t q = 0; q < take; ++q) {
					const uint32_t r = order[unseen + q];
					const uint32_t b = free_blocks.back(); free_blocks.pop_back();
					blk[r] = b; next_cc[r] = 0;
					const uint32_t first = (uint32_t)hs.size();
					for (uint32_t cc = 0; cc < chunk_cap; ++cc) {
						uint32_t len2;
						if (!chunk_of(r, cc, &len2)) break;
						slot_t sl; memset(&sl, 0, sizeof(sl));
						sl.read = r; sl.chunk_len = len2; sl.c_count = cc;
						sl.e_cap = (uint32_t)((uint64_t)len2 * 2 / 3 + 8);
						sl.e_off = ((uint64_t)b * chunk_cap + cc) * e_cap_max;
						hs.push_back(sl);
					}
					groups.push_back(make_uint2(first, (uint32_t)hs.size() - first));
					n_cc[r] = (uint32_t)hs.size() - first;
					who.push_back(r);
				}
				if (c->trace) fprintf(stderr, "[trace] staging %zu reads, %zu chunks, %zu free blocks left\n", take, hs.size(), free_blocks.size());
				if ((rc = event_stage(c, hs, c->d_pre, groups, true))) return rc;
				RH_TRACE(c, "event stage");
				CUDA_TRY(cudaMemcpyAsync(hs.data(), c->d_pre.p, hs.size() * sizeof(slot_t), cudaMemcpyDeviceToHost, c->stream));
				CUDA_TRY(cudaStreamSynchronize(c->stream));
				c->st.d2h_bytes += hs.size() * sizeof(slot_t);
				for (size_t q = 0; q < who.size(); ++q) {
					const uint2 g = groups[q];
					for (uint32_t k = 0; k < g.y; ++k) table[(size_t)blk[who[q]] * chunk_cap + k] = hs[g.x + k];
					ready.push_back(who[q]);
				}
				for (const slot_t &sl : hs) { c->st.event_stage_samples += sl.raw_used; c->st.event_stage_seeds += sl.n_seeds; }
				unseen += take;
			}
			CUDA_TRY(cudaMemsetAsync(c->d_counters.p, 0, sizeof(unsigned long long), c->stream)); /* carry_top of the iteration's output arena */
			round_io io;
			io.events_done = true;
			io.slots.reserve(inflight.size() + ready.size());
			for (uint32_t r : inflight) io.slots.push_back(table[(size_t)blk[r] * chunk_cap + next_cc[r]]);
			for (uint32_t r : ready) io.slots.push_back(table[(size_t)blk[r] * chunk_cap]);
			io.n_mandatory = (uint32_t)inflight.size();
			io.max_optional = inflight.size() >= max_inflight ? 0u : (uint32_t)std::min<size_t>(ready.size(), max_inflight - inflight.size());
			if ((rc = run_round(c, io, (int)(iter & 1)))) return rc;
			c->st.n_rounds++;
			CUDA_TRY(cudaMemcpyAsync(rec_now.data(), c->d_rec_cnt.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
			CUDA_TRY(cudaStreamSynchronize(c->stream));
			c->st.d2h_bytes += n * 4;
			if (c->prof_on && getenv("RH_PROF_ROUNDS")) {
				unsigned long long h[64];
				cudaMemcpy(h, c->d_prof.p, sizeof(h), cudaMemcpyDeviceToHost);
				cudaMemset(c->d_prof.p, 0, sizeof(h));
				fprintf(stderr, "[RH_PROF] iteration %u (%zu in flight, %u admitted of %zu waiting) raw counters:", iter, inflight.size(), io.n_run - io.n_mandatory, ready.size());
				for (int i = 0; i < 64; ++i) if (h[i]) fprintf(stderr, " [%d]=%llu", i, h[i]);
				fprintf(stderr, "\n");
			}
			const uint32_t n_adm = io.n_run - io.n_mandatory;
			if (sched_verbose) {
				uint64_t na = 0; for (uint32_t q = 0; q < io.n_run; ++q) na += io.slots[q].n_anchors;
				fprintf(stderr, "[rawhash_b200] iteration %u: %zu in flight + %u admitted (%zu waiting, %zu not staged), %.1f M anchors, carry arenas %zu/%zu MB, %.2f s since the call began\n", iter, inflight.size(), n_adm,
				        ready.size() - n_adm, order.size() - unseen, na / 1e6, c->d_carry[0].cap * sizeof(anchor_t) >> 20, c->d_carry[1].cap * sizeof(anchor_t) >> 20,
				        std::chrono::duration<double>(std::chrono::steady_clock::now() - t_sched0).count());
			}
			std::vector<uint32_t> next;
			next.reserve(inflight.size() + n_adm);
			auto settle = [&](uint32_t r) {
				if (rec_now[r] == 0xffffffffu) { ++next_cc[r]; next.push_back(r); }
				else { rec_cnt[r] = rec_now[r]; free_blocks.push_back(blk[r]); blk[r] = 0xffffffffu; }
			};
			for (uint32_t r : inflight) settle(r);
			for (uint32_t q = 0; q < n_adm; ++q) settle(ready[q]);
			ready.erase(ready.begin(), ready.begin() + n_adm);
			if (inflight.empty() && n_adm == 0 && !ready.empty()) { rh_set_error("internal: the scheduler admitted no read"); return RH_ERR_CUDA; }
			for (uint32_t r : next) if (next_cc[r] >= n_cc[r]) { rh_set_error("internal: read %u still active after its last chunk", r); return RH_ERR_CUDA; }
			inflight.swap(next);
			++iter;
		}
		active.clear();
	}
	std::vector<slot_t> pre;            /* host copy of the wave's slots after their event stage */
	std::vector<uint32_t> pre_first(n, 0xffffffffu);
	uint32_t wave_base = 0, wave_end = 0;
	uint32_t round = 0;
	while (!stream_sched && !active.empty() && round < max_chunk) {
		CUDA_TRY(cudaMemsetAsync(c->d_counters.p, 0, sizeof(unsigned long long), c->stream)); /* carry_top of the round's output arena */
		if (ahead > 1 && round == wave_end) { /* next wave: chunks [round, round + ahead) of every active read */
			pre.clear();
			std::vector<uint2> groups;
			for (uint32_t r : active) {
				const uint32_t first = (uint32_t)pre.size();
				for (uint32_t cc = round; cc < round + ahead && cc < max_chunk; ++cc) {
					uint32_t len2;
					if (!chunk_of(r, cc, &len2)) break;
					slot_t sl; memset(&sl, 0, sizeof(sl));
					sl.read = r; sl.chunk_len = len2; sl.c_count = cc;
					pre.push_back(sl);
				}
				pre_first[r] = first;
				if (pre.size() > first) groups.push_back(make_uint2(first, (uint32_t)pre.size() - first));
			}
			if ((rc = event_stage(c, pre, c->d_pre, groups))) return rc;
			CUDA_TRY(cudaMemcpyAsync(pre.data(), c->d_pre.p, pre.size() * sizeof(slot_t), cudaMemcpyDeviceToHost, c->stream));
			CUDA_TRY(cudaStreamSynchronize(c->stream));
			c->st.d2h_bytes += pre.size() * sizeof(slot_t);
			for (const slot_t &sl : pre) { c->st.event_stage_samples += sl.raw_used; c->st.event_stage_seeds += sl.n_seeds; }
			wave_base = round; wave_end = round + ahead;
		}
		if (ahead > 1) {
			round_io io;
			io.events_done = true;
			io.slots.reserve(active.size());
			for (uint32_t r : active) io.slots.push_back(pre[pre_first[r] + (round - wave_base)]);
			if ((rc = run_round(c, io, round & 1))) return rc;
		} else {
			/* split the active set into groups whose signal scratch fits the budget */
			size_t a0 = 0;
			while (a0 < active.size()) {
				round_io io;
				uint64_t samples = 0; size_t a1 = a0;
				while (a1 < active.size()) {
					const uint32_t r = active[a1];
					uint32_t len2 = 0;
					chunk_of(r, round, &len2);
					if (samples + len2 > c->sig_budget && a1 > a0) break;
					slot_t sl; memset(&sl, 0, sizeof(sl));
					sl.read = r; sl.chunk_len = len2; sl.c_count = round;
					io.slots.push_back(sl);
					samples += len2; ++a1;
				}
				if ((rc = run_round(c, io, round & 1))) return rc;
				a0 = a1;
			}
		}
		c->st.n_rounds++;
		CUDA_TRY(cudaMemcpyAsync(rec_cnt.data(), c->d_rec_cnt.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
		CUDA_TRY(cudaStreamSynchronize(c->stream));
		if (c->prof_on && getenv("RH_PROF_ROUNDS")) { /* per-round phase cycles (debug aid) */
			unsigned long long h[64];
			cudaMemcpy(h, c->d_prof.p, sizeof(h), cudaMemcpyDeviceToHost);
			cudaMemset(c->d_prof.p, 0, sizeof(h));
			fprintf(stderr, "[RH_PROF] round %u (%zu reads active) units (cycles 1e6; [48] n_z [49] n_runs [50] start-ties [51] key-ties [52] n_u [53] n [54] n_v):", round, active.size());
			for (int i = 0; i < 64; ++i) if (h[i]) fprintf(stderr, " [%d]=%.1f", i, h[i] / 1e6);
			fprintf(stderr, "\n");
		}
		c->st.d2h_bytes += n * 4;
		std::vector<uint32_t> next;
		for (uint32_t r : active) if (rec_cnt[r] == 0xffffffffu) next.push_back(r);
		active.swap(next);
		++round;
	}
	if ((rc = check_dev_err(c))) return rc;
	if (!active.empty()) { rh_set_error("internal: %zu reads still active after the last round", active.size()); return RH_ERR_CUDA; }
	/* gather records in read order */
	unsigned long long tops[2];
	CUDA_TRY(cudaMemcpyAsync(tops, c->d_counters.p, sizeof(tops), cudaMemcpyDeviceToHost, c->stream));
	std::vector<uint32_t> rec_start(n);
	CUDA_TRY(cudaMemcpyAsync(rec_start.data(), c->d_rec_start.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	std::vector<rh_map_rec_t> dev_recs(tops[1]);
	if (tops[1]) CUDA_TRY(cudaMemcpyAsync(dev_recs.data(), c->d_recs.p, tops[1] * sizeof(rh_map_rec_t), cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	c->st.d2h_bytes += tops[1] * sizeof(rh_map_rec_t) + n * 4;
	uint64_t total = 0;
	for (uint32_t i = 0; i < n; ++i) total += (l_sig[i] == 0) ? 1 : rec_cnt[i];
	rh_map_rec_t *out = (rh_map_rec_t *)malloc((total ? total : 1) * sizeof(rh_map_rec_t));
	uint64_t k = 0;
	for (uint32_t i = 0; i < n; ++i) {
		if (l_sig[i] == 0) { /* loop of map_worker_for never runs: one unmapped record, ci = 1 */
			rh_map_rec_t r; memset(&r, 0, sizeof(r)); r.read_idx = i; r.ci = 1; r.sl = 0;
			out[k++] = r; continue;
		}
		for (uint32_t m = 0; m < rec_cnt[i]; ++m) out[k++] = dev_recs[rec_start[i] + m];
	}
	*recs_out = out; *n_recs_out = k;
	c->st.n_reads += n;
	return RH_OK;
}

void begin_call(rh_worker *c) { memset(&c->st, 0, sizeof(c->st)); c->spans.clear(); c->ev_used = 0; }

} // namespace

/* ============================================================================================ */
namespace {

void destroy_worker(rh_worker *w)
{
	if (!w) return;
	w->d_ps.release(); w->d_pq.release(); w->d_t1.release(); w->d_t2.release(); w->d_norm.release(); w->d_groups.release(); w->d_pre.release();
	w->d_rs.release(); w->d_slots.release(); w->d_z.release(); w->d_events.release(); w->d_peaks.release();
	w->d_seed_hash.release(); w->d_seed_pos.release(); w->d_seed_cnt.release(); w->d_seed_dst.release(); w->d_seed_src.release();
	w->d_arena.release(); w->d_carry[0].release(); w->d_carry[1].release(); w->d_counters.release(); w->d_err.release();
	w->d_rec_start.release(); w->d_rec_cnt.release(); w->d_recs.release(); w->d_tie_list.release(); w->d_tie_count.release(); w->d_prof.release();
	for (cudaEvent_t e : w->ev_pool) cudaEventDestroy(e);
	if (w->own_stream) cudaStreamDestroy(w->own_stream);
	if (w->stream2) cudaStreamDestroy(w->stream2);
	w->d_tie_list2.release(); w->d_tie_count2.release();
	delete w;
}

rh_worker *make_worker(rh_gpu_ctx *c, size_t arena_bytes)
{
	rh_worker *w = new rh_worker();
	w->device = c->device; w->P = c->P; w->D = c->D; w->idx = c->idx; w->I = c->I;
	w->sort_posbits = c->sort_posbits; w->sort_ridbits = c->sort_ridbits; w->sort_smem_cap = c->sort_smem_cap; w->hits_per_key = c->hits_per_key;
	w->d_logf = c->d_logf; w->logf_n = c->logf_n; w->d_seqlen = c->d_seqlen; w->name_order = &c->name_order; /* aliases: the context owns them */
	w->err[0] = 0;
	if (cudaStreamCreateWithFlags(&w->own_stream, cudaStreamNonBlocking) != cudaSuccess) { rh_set_error("cudaStreamCreate failed"); destroy_worker(w); return nullptr; }
	w->stream = w->own_stream;
	if (cudaStreamCreateWithFlags(&w->stream2, cudaStreamNonBlocking) != cudaSuccess) w->stream2 = nullptr;
	w->prof_on = getenv("RH_PROF") != NULL; w->trace = getenv("RH_TRACE") != NULL;
	if (w->d_counters.reserve(2) || w->d_err.reserve(1) || w->d_prof.reserve(64)) { destroy_worker(w); return nullptr; }
	cudaMemset(w->d_prof.p, 0, 64 * 8);
	w->arena_bytes = arena_bytes;
	const size_t carry_elems = arena_bytes / 8 / sizeof(anchor_t); /* initial size: grows on demand (run_round) */
	w->sig_budget = std::max<size_t>(arena_bytes / 8 / 22, (size_t)1 << 20); /* ~85 KB of peak/event/seed scratch per 4000-sample chunk */
	if (w->d_arena.reserve(arena_bytes) || w->d_carry[0].reserve(carry_elems) || w->d_carry[1].reserve(carry_elems)) { destroy_worker(w); return nullptr; }
	w->arena_bytes = w->d_arena.cap;
	return w;
}

void add_stats(rh_gpu_stats_t &a, const rh_gpu_stats_t &b)
{
	a.n_reads += b.n_reads; a.n_chunks += b.n_chunks; a.n_rounds = std::max(a.n_rounds, b.n_rounds);
	a.raw_samples_consumed += b.raw_samples_consumed; a.n_events += b.n_events; a.n_seeds += b.n_seeds; a.n_anchors += b.n_anchors; a.n_chains += b.n_chains;
	a.kernel_launches += b.kernel_launches; a.event_kernel_launches += b.event_kernel_launches;
	a.ms_total = std::max(a.ms_total, b.ms_total);
	a.ms_event_kernel += b.ms_event_kernel; a.ms_seed += b.ms_seed; a.ms_sort += b.ms_sort; a.ms_sort_ties += b.ms_sort_ties; a.ms_chain += b.ms_chain; a.ms_post += b.ms_post;
	a.h2d_bytes += b.h2d_bytes; a.d2h_bytes += b.d2h_bytes;
	a.event_stage_samples += b.event_stage_samples; a.event_stage_seeds += b.event_stage_seeds;
}

/* contiguous read ranges with (nearly) equal raw sample counts: b[0..k] */
std::vector<uint32_t> split_ranges(uint32_t n, const std::vector<uint64_t> &len, uint32_t k)
{
	std::vector<uint32_t> b(k + 1, n);
	rh_split_by_samples(n, len.data(), k, b.data());
	return b;
}

/* one worker's share of a batch: optional upload of its raw samples, then the chunk rounds */
struct job_t {
	rh_worker *w; uint32_t lo, hi;
	const int16_t *const *raw; int16_t *d_raw;   /* host pointers + shared device buffer (null when already resident) */
	const batch_in *in; const std::vector<uint64_t> *beg, *len;
	rh_map_rec_t *recs = nullptr; uint64_t n_recs = 0; int rc = RH_OK; uint64_t h2d = 0;
};

/* H2D copies of one worker's reads on its stream; called from the submitting thread for worker 0, 1, ... in turn so
 * that the copy engine serves the ranges in that order and worker r computes while the ranges after it still load */
int issue_uploads(job_t *j)
{
	rh_worker *w = j->w;
	int rc = RH_OK;
	if (!j->raw) return rc;
	const std::vector<uint64_t> &B = *j->beg, &L = *j->len;
	for (uint32_t i = j->lo; i < j->hi && rc == RH_OK;) { /* one copy per run of reads contiguous on both sides */
		uint32_t e = i; uint64_t bytes = L[i] * 2;
		while (e + 1 < j->hi && L[e] > 0 && j->raw[e + 1] == j->raw[e] + L[e] && B[e + 1] == B[e] + L[e] && bytes < ((uint64_t)256 << 20)) { ++e; bytes += L[e] * 2; }
		if (bytes && cudaMemcpyAsync(j->d_raw + B[i], j->raw[i], bytes, cudaMemcpyHostToDevice, w->stream) != cudaSuccess) { rh_set_error("H2D copy of reads %u..%u failed: %s", i, e, cudaGetErrorString(cudaGetLastError())); rc = RH_ERR_CUDA; }
		j->h2d += bytes;
		i = e + 1;
	}
	return rc;
}

void run_job(job_t *j)
{
	rh_worker *w = j->w;
	if (cudaSetDevice(w->device) != cudaSuccess) { j->rc = RH_ERR_CUDA; snprintf(w->err, sizeof(w->err), "cudaSetDevice(%d) failed", w->device); return; }
	begin_call(w);
	cudaEvent_t t0 = get_event(w), t1 = get_event(w);
	cudaEventRecord(t0, w->stream);
	const uint32_t n = j->hi - j->lo;
	const int rc0 = j->rc; /* outcome of issue_uploads (its error text is already in w->err) */
	int rc = rc0;
	w->st.h2d_bytes += j->h2d;
	if (rc == RH_OK && n) {
		batch_in sub = *j->in;
		sub.n = n; sub.offset += j->lo; sub.range += j->lo; sub.digitisation += j->lo;
		if (sub.names) sub.names += j->lo;
		/* a range that does not fit the device (the chains carried between chunk rounds grow with the number of reads in
		 * flight: ~3 MB per read against a human-size index) is mapped in halves, one after the other */
		struct piece_t { uint32_t lo, hi; };
		std::vector<piece_t> todo{{0u, n}};
		std::vector<rh_map_rec_t> acc;
		rh_gpu_stats_t st_sum; memset(&st_sum, 0, sizeof(st_sum));
		while (!todo.empty() && rc == RH_OK) {
			const piece_t pc = todo.back(); todo.pop_back();
			batch_in part = sub;
			part.n = pc.hi - pc.lo; part.offset += pc.lo; part.range += pc.lo; part.digitisation += pc.lo;
			if (part.names) part.names += pc.lo;
			const std::vector<uint64_t> b(j->beg->begin() + j->lo + pc.lo, j->beg->begin() + j->lo + pc.hi), l(j->len->begin() + j->lo + pc.lo, j->len->begin() + j->lo + pc.hi);
			rh_map_rec_t *recs = nullptr; uint64_t n_recs = 0;
			const rh_gpu_stats_t keep = w->st;
			const int r1 = map_resident(w, part, b, l, &recs, &n_recs);
			if (r1 == RH_ERR_NOMEM && w->retry_same) { /* the record arena was too small and has been re-sized from the count */
				w->retry_same = false; w->st = keep; cudaGetLastError();
				todo.push_back(pc);
				continue;
			}
			if (r1 == RH_ERR_NOMEM && part.n >= 128) { /* nothing of this piece was kept: retry as two */
				w->st = keep;
				cudaGetLastError();
				const uint32_t mid = pc.lo + part.n / 2;
				todo.push_back({mid, pc.hi}); todo.push_back({pc.lo, mid}); /* the lower half first: records stay in read order */
				w->n_splits++;
				fprintf(stderr, "[rawhash_b200] %u reads did not fit the device (%s): mapping them as two halves\n", part.n, rh_gpu_last_error());
				continue;
			}
			rc = r1;
			if (rc == RH_OK) { for (uint64_t q = 0; q < n_recs; ++q) { recs[q].read_idx += pc.lo; acc.push_back(recs[q]); } free(recs); }
		}
		if (rc == RH_OK) {
			j->n_recs = acc.size();
			j->recs = (rh_map_rec_t *)malloc((acc.size() ? acc.size() : 1) * sizeof(rh_map_rec_t));
			memcpy(j->recs, acc.data(), acc.size() * sizeof(rh_map_rec_t));
		}
	}
	cudaEventRecord(t1, w->stream);
	cudaEventSynchronize(t1);
	float ms = 0; cudaEventElapsedTime(&ms, t0, t1); w->st.ms_total = ms;
	collect_spans(w);
	j->rc = rc;
	if (rc != RH_OK && rc0 == RH_OK) snprintf(w->err, sizeof(w->err), "%s", rh_gpu_last_error()); /* the error text is thread local: hand it to the caller */
}

int map_batch(rh_gpu_ctx *c, const batch_in &in, const std::vector<uint64_t> &beg, const std::vector<uint64_t> &len, const int16_t *raw_dev,
              rh_map_rec_t **recs, uint64_t *n_recs)
{
	const uint32_t n = in.n;
	memset(&c->st, 0, sizeof(c->st));
	uint32_t k = std::min<uint32_t>(c->n_active, (uint32_t)c->workers.size());
	if (n < 64 * k) k = 1; /* tiny batches: one range */
	if (c->user_stream) k = 1; /* the caller's stream carries the whole batch */
	const std::vector<uint32_t> b = split_ranges(n, len, k);
	std::vector<job_t> jobs(k);
	for (uint32_t r = 0; r < k; ++r) {
		job_t &j = jobs[r];
		j.w = c->workers[r]; j.lo = b[r]; j.hi = b[r + 1];
		j.raw = in.raw; j.d_raw = c->d_raw.p; j.in = &in; j.beg = &beg; j.len = &len;
		j.w->raw_ptr = raw_dev;
		j.w->stream = (c->user_stream && r == 0) ? (cudaStream_t)c->user_stream : j.w->own_stream;
	}
	for (uint32_t r = 0; r < k; ++r) { jobs[r].rc = issue_uploads(&jobs[r]); if (jobs[r].rc != RH_OK) snprintf(jobs[r].w->err, sizeof(jobs[r].w->err), "%s", rh_gpu_last_error()); }
	if (k == 1) run_job(&jobs[0]);
	else {
		std::vector<std::thread> th;
		for (uint32_t r = 0; r < k; ++r) th.emplace_back(run_job, &jobs[r]);
		for (std::thread &t : th) t.join();
	}
	int rc = RH_OK;
	uint64_t total = 0;
	for (job_t &j : jobs) { if (j.rc != RH_OK && rc == RH_OK) { rc = j.rc; rh_set_error("%s", j.w->err); } total += j.n_recs; add_stats(c->st, j.w->st); }
	if (rc == RH_OK) {
		rh_map_rec_t *out = (rh_map_rec_t *)malloc((total ? total : 1) * sizeof(rh_map_rec_t));
		uint64_t o = 0;
		for (job_t &j : jobs) {
			for (uint64_t q = 0; q < j.n_recs; ++q) { out[o] = j.recs[q]; out[o].read_idx += j.lo; ++o; } /* rank order = input order */
		}
		{ /* mt:f: — every read is charged its share (by chunks consumed) of the batch's stream time */
			unsigned long long chunks = 0; uint32_t prev = 0xffffffffu;
			for (uint64_t q = 0; q < total; ++q) if (out[q].read_idx != prev) { prev = out[q].read_idx; chunks += out[q].ci; }
			const double per_chunk = chunks ? c->st.ms_total / (double)chunks : 0.0;
			for (uint64_t q = 0; q < total; ++q) out[q].mt_ms = (float)(per_chunk * out[q].ci);
		}
		*recs = out; *n_recs = total;
	}
	for (job_t &j : jobs) free(j.recs);
	return rc;
}

} // namespace

extern "C" rh_gpu_ctx *rh_gpu_init(const rh_index_t *idx, const rh_params_t *p, int device, size_t arena_bytes)
{
	if (!idx || !p) { rh_set_error("rh_gpu_init: null argument"); return NULL; }
	if (check_params(*p) != RH_OK) return NULL;
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device >= ndev) { rh_set_error("no usable CUDA device (count=%d, asked %d)", ndev, device); return NULL; }
	if (cudaSetDevice(device) != cudaSuccess) { rh_set_error("cudaSetDevice(%d) failed", device); return NULL; }
	rh_gpu_ctx *c = new rh_gpu_ctx();
	c->device = device; c->P = *p; c->idx = idx;
	fill_dev_params(*p, c->D);
	auto fail = [&](const char *what) -> rh_gpu_ctx * { if (what) rh_set_error("%s", what); rh_gpu_destroy(c); return NULL; };
	/* ---- index: a device-resident index on this device is used in place; a host index is uploaded once and becomes
	 *      device-resident (several contexts on one GPU share it) ---- */
	rh_index_s *midx = const_cast<rh_index_s *>(idx);
	if (midx->dev.device >= 0 && midx->dev.device != device) {
		if (rh_index_sync_host(idx) != RH_OK) return fail(NULL);
	}
	if (midx->dev.device != device) { /* upload the host arrays */
		rh_index_dev_t V; V.device = device; V.n_keys = idx->keys.size(); V.n_pos = idx->pos.size();
		if (midx->dev.device >= 0) { /* lives on another GPU: this context keeps a private copy (c->own_dev) */ }
		if (cudaMalloc((void **)&V.keys, std::max<size_t>(V.n_keys, 1) * 4) != cudaSuccess || cudaMalloc((void **)&V.off, (V.n_keys + 1) * 8) != cudaSuccess ||
		    cudaMalloc((void **)&V.pos, std::max<size_t>(V.n_pos, 1) * 8) != cudaSuccess) {
			if (V.keys) cudaFree(V.keys); if (V.off) cudaFree(V.off); if (V.pos) cudaFree(V.pos);
			return fail("index upload: out of device memory");
		}
		const uint64_t zero = 0;
		cudaMemcpy(V.keys, idx->keys.data(), V.n_keys * 4, cudaMemcpyHostToDevice);
		if (idx->off.size() == V.n_keys + 1) cudaMemcpy(V.off, idx->off.data(), (V.n_keys + 1) * 8, cudaMemcpyHostToDevice);
		else cudaMemcpy(V.off, &zero, 8, cudaMemcpyHostToDevice);
		cudaMemcpy(V.pos, idx->pos.data(), V.n_pos * 8, cudaMemcpyHostToDevice);
		if (midx->dev.device < 0) midx->dev = V; /* the index handle owns it from here on */
		else { c->own_dev = V; c->own_dev_valid = true; }
	}
	rh_index_dev_t *V = c->own_dev_valid ? &c->own_dev : &midx->dev;
	if (!V->bucket) {
		rh_index_s tmp_holder; /* rh_index_dev_make_buckets works on an index handle */
		if (c->own_dev_valid) { tmp_holder.dev = c->own_dev; if (rh_index_dev_make_buckets(&tmp_holder) != RH_OK) { tmp_holder.dev = rh_index_dev_t(); return fail(NULL); } c->own_dev = tmp_holder.dev; tmp_holder.dev = rh_index_dev_t(); }
		else if (rh_index_dev_make_buckets(midx) != RH_OK) return fail(NULL);
	}
	const size_t nk = (size_t)V->n_keys;
	c->hits_per_key = nk ? (double)V->n_pos / (double)nk : 0.0;
	const int bits = V->bucket_bits;
	std::vector<uint32_t> rank(idx->names.size());
	c->name_order.resize(idx->names.size());
	std::iota(c->name_order.begin(), c->name_order.end(), 0u);
	std::sort(c->name_order.begin(), c->name_order.end(), [&](uint32_t a, uint32_t b) { int s = strcmp(idx->names[a].c_str(), idx->names[b].c_str()); return s ? s < 0 : a < b; });
	for (uint32_t i = 0; i < c->name_order.size(); ++i) rank[c->name_order[i]] = i;
	std::vector<float> lt;
	c->logf_n = 1u << 20;
	lt.resize(c->logf_n);
	for (uint32_t i = 0; i < c->logf_n; ++i) lt[i] = logf((float)i); /* host libm: identical to the reference's calls */
	cudaStream_t s0 = nullptr;
	if (upload(c->d_seqlen, idx->lens, s0) || upload(c->d_namerank, rank, s0) || upload(c->d_logf, lt, s0)) return fail(NULL);
	c->I.keys = V->keys; c->I.off = V->off; c->I.pos = V->pos; c->I.bucket = V->bucket; c->I.bucket_bits = bits;
	c->I.n_keys = nk; c->I.seq_len = c->d_seqlen.p; c->I.name_rank = c->d_namerank.p; c->I.n_seq = (uint32_t)idx->names.size();
	if (p->mid_occ <= 0) { rh_params_t q = *p; rh_index_update_mapopt(idx, &q); c->P.mid_occ = q.mid_occ; c->D.mid_occ = q.mid_occ; }
	{ /* fixed key packing for the shared-memory anchor sort: strand | target id | target position */
		uint32_t maxlen = 1;
		for (uint32_t l : idx->lens) maxlen = std::max(maxlen, l);
		uint32_t pb = 1; while (pb < 31 && (1u << pb) < maxlen) ++pb;
		uint32_t rb = 0; while (rb < 31 && (1u << rb) < (uint32_t)idx->lens.size()) ++rb;
		c->sort_posbits = pb; c->sort_ridbits = rb;
		int max_smem = 0;
		cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
		c->sort_smem_cap = 0;
		if (pb + rb + 1 <= 32 && (size_t)max_smem > SB_FIXED_BYTES + 8 * 1024) {
			c->sort_smem_cap = std::min<uint32_t>(((uint32_t)max_smem - SB_FIXED_BYTES) / 8, 65535u) & ~31u;
			if (cudaFuncSetAttribute(k_sort_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_smem_bytes(c->sort_smem_cap)) != cudaSuccess) return fail("cudaFuncSetAttribute(k_sort_smem) failed");
		}
	}
	if (cudaFuncSetAttribute(k_sort_ties, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tie_smem_bytes(0, TIE_SIDE_WALKS)) != cudaSuccess) return fail("cudaFuncSetAttribute(k_sort_ties) failed");
	if (cudaDeviceSynchronize() != cudaSuccess) return fail("index upload failed");
	/* ---- workers: the work arenas are split evenly ---- */
	size_t free_b = 0, total_b = 0;
	cudaMemGetInfo(&free_b, &total_b);
	if (arena_bytes == 0) arena_bytes = free_b / 2;
	if (arena_bytes > free_b * 7 / 10) arena_bytes = free_b * 7 / 10;
	uint32_t nw = 1;
	if (const char *e = getenv("RH_WORKERS")) nw = (uint32_t)std::max(1, std::min(16, atoi(e)));
	while (nw > 1 && arena_bytes / nw < ((size_t)64 << 20)) --nw; /* keep every worker's arena useful */
	for (uint32_t r = 0; r < nw; ++r) {
		rh_worker *w = make_worker(c, (size_t)((double)arena_bytes / nw / 1.25) /* dbuf over-allocates by 25 % */);
		if (!w) return fail(NULL);
		c->workers.push_back(w);
	}
	{ /* exhaustive check of the reciprocal division for the two window lengths of this context */
		unsigned long long *d_bad = nullptr, h_bad[2] = {0, 0};
		if (cudaMalloc((void **)&d_bad, 16) != cudaSuccess) return fail("out of device memory");
		cudaMemset(d_bad, 0, 16);
		k_selftest_divw<<<1184, 256>>>((float)p->window_length1, d_bad);
		k_selftest_divw<<<1184, 256>>>((float)p->window_length2, d_bad + 1);
		if (cudaMemcpy(h_bad, d_bad, 16, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaFree(d_bad); return fail("division self-test failed to run"); }
		cudaFree(d_bad);
		for (rh_worker *w : c->workers) { w->fast_w1 = h_bad[0] == 0; w->fast_w2 = h_bad[1] == 0; }
		if (getenv("RH_CLI_VERBOSE") || getenv("RH_PROF"))
			fprintf(stderr, "[rawhash_b200] x/w by reciprocal: w=%u mismatches %llu (in range %llu), w=%u mismatches %llu (in range %llu)\n", p->window_length1,
			        h_bad[0] & 0xffffffffULL, h_bad[0] >> 32, p->window_length2, h_bad[1] & 0xffffffffULL, h_bad[1] >> 32);
	}
	c->n_active = nw;
	return c;
}

extern "C" void rh_gpu_destroy(rh_gpu_ctx *c)
{
	if (!c) return;
	cudaSetDevice(c->device);
	for (rh_worker *w : c->workers) destroy_worker(w);
	if (c->own_dev_valid) { rh_index_s h; h.dev = c->own_dev; /* released by ~rh_index_s */ }
	c->d_seqlen.release(); c->d_namerank.release(); c->d_logf.release();
	c->d_raw.release();
	delete c;
}

extern "C" void rh_gpu_set_stream(rh_gpu_ctx *c, void *cuda_stream) { if (c) c->user_stream = cuda_stream; }

extern "C" int rh_gpu_set_workers(rh_gpu_ctx *c, int n_workers)
{
	if (!c || n_workers < 1) { rh_set_error("rh_gpu_set_workers: bad argument"); return RH_ERR_ARG; }
	c->n_active = std::min<uint32_t>((uint32_t)n_workers, (uint32_t)c->workers.size());
	return (int)c->n_active;
}

extern "C" void rh_gpu_get_stats(const rh_gpu_ctx *c, rh_gpu_stats_t *st) { *st = c->st; }

extern "C" int rh_gpu_map_batch_raw(rh_gpu_ctx *c, uint32_t n, const int16_t *const *raw, const uint64_t *raw_len,
                                    const double *offset, const double *range, const double *digitisation,
                                    const char *const *names, rh_map_rec_t **recs, uint64_t *n_recs)
{
	if (!c || !recs || !n_recs || (n && (!raw || !raw_len || !offset || !range || !digitisation))) { rh_set_error("rh_gpu_map_batch_raw: null argument"); return RH_ERR_ARG; }
	if (c->D.ava && n && !names) { rh_set_error("read names are required in all-vs-all mode"); return RH_ERR_ARG; }
	CUDA_TRY(cudaSetDevice(c->device));
	std::vector<uint64_t> beg(n), len(n);
	uint64_t tot = 0;
	for (uint32_t i = 0; i < n; ++i) { /* reads that follow each other in host memory keep doing so on the device: one copy per run */
		len[i] = raw_len[i];
		if (i > 0 && raw_len[i - 1] > 0 && raw[i] == raw[i - 1] + raw_len[i - 1]) beg[i] = beg[i - 1] + raw_len[i - 1];
		else { tot = (tot + 7) & ~7ULL; beg[i] = tot; } /* a new run starts 16-byte aligned */
		tot = beg[i] + raw_len[i];
	}
	int rc = c->d_raw.reserve(tot + 8);
	if (rc) return rc;
	batch_in in{n, raw, raw_len, nullptr, nullptr, offset, range, digitisation, names};
	return map_batch(c, in, beg, len, c->d_raw.p, recs, n_recs);
}

extern "C" int rh_gpu_map_batch_dev(rh_gpu_ctx *c, uint32_t n, const void *d_raw, const uint64_t *raw_off,
                                    const double *offset, const double *range, const double *digitisation,
                                    const char *const *names, rh_map_rec_t **recs, uint64_t *n_recs)
{
	if (!c || !recs || !n_recs || (n && (!d_raw || !raw_off || !offset || !range || !digitisation))) { rh_set_error("rh_gpu_map_batch_dev: null argument"); return RH_ERR_ARG; }
	if (((uintptr_t)d_raw & 15) != 0) { rh_set_error("device raw buffer must be 16-byte aligned"); return RH_ERR_ARG; }
	if (c->D.ava && n && !names) { rh_set_error("read names are required in all-vs-all mode"); return RH_ERR_ARG; }
	CUDA_TRY(cudaSetDevice(c->device));
	std::vector<uint64_t> beg(n), len(n);
	for (uint32_t i = 0; i < n; ++i) { beg[i] = raw_off[i]; len[i] = raw_off[i + 1] - raw_off[i]; }
	batch_in in{n, nullptr, nullptr, d_raw, raw_off, offset, range, digitisation, names};
	return map_batch(c, in, beg, len, (const int16_t *)d_raw, recs, n_recs);
}

/* raw samples of a few reads into the context's shared buffer, on one worker's stream (tap / index paths) */
static int upload_raw(rh_gpu_ctx *ctx, rh_worker *c, uint32_t n, const int16_t *const *raw, const uint64_t *raw_len, std::vector<uint64_t> &beg, std::vector<uint64_t> &len)
{
	beg.assign(n, 0); len.assign(n, 0);
	uint64_t tot = 0;
	for (uint32_t i = 0; i < n; ++i) { beg[i] = tot; len[i] = raw_len[i]; tot += (raw_len[i] + 7) & ~7ULL; } /* 16-byte aligned starts */
	int rc = ctx->d_raw.reserve(tot + 8);
	if (rc) return rc;
	for (uint32_t i = 0; i < n; ++i)
		if (raw_len[i]) { CUDA_TRY(cudaMemcpyAsync(ctx->d_raw.p + beg[i], raw[i], raw_len[i] * 2, cudaMemcpyHostToDevice, c->stream)); c->st.h2d_bytes += raw_len[i] * 2; }
	c->raw_ptr = ctx->d_raw.p;
	return RH_OK;
}

/* Stage tap: one read, every chunk, no stop rules (parity tests). */
extern "C" int rh_gpu_tap_read(rh_gpu_ctx *ctx, const int16_t *raw, uint64_t raw_len, double offset, double range, double digitisation,
                               const char *name, rh_tap_t *tap)
{
	if (!ctx || !raw || !tap) { rh_set_error("rh_gpu_tap_read: null argument"); return RH_ERR_ARG; }
	CUDA_TRY(cudaSetDevice(ctx->device));
	rh_worker *c = ctx->workers[0];
	c->stream = c->own_stream;
	begin_call(c);
	std::vector<uint64_t> beg, len;
	const int16_t *rp[1] = {raw}; const uint64_t rl[1] = {raw_len};
	const char *nm[1] = {name ? name : ""};
	int rc = upload_raw(ctx, c, 1, rp, rl, beg, len);
	if (rc) return rc;
	batch_in in{1, rp, rl, nullptr, nullptr, &offset, &range, &digitisation, nm};
	std::vector<uint32_t> l_sig;
	if ((rc = setup_reads(c, in, beg, len, l_sig))) return rc;
	if ((rc = c->d_recs.reserve(1024))) return rc;
	const rh_params_t &P = c->P;
	const bool noadapt = c->D.noadapt != 0;
	const uint32_t qlen = l_sig[0];
	const uint32_t l_chunk = (P.chunk_size > qlen || noadapt) ? qlen : P.chunk_size;
	const uint32_t max_chunk = noadapt ? 1u : P.max_num_chunk;
	uint64_t toff[6] = {0, 0, 0, 0, 0, 0}; /* ev, seed, anc, u, chain, reg */
	uint32_t cc = 0;
	cudaStream_t s = c->stream;
	for (uint64_t s_qs = 0; s_qs < qlen && cc < max_chunk; s_qs += l_chunk, ++cc) {
		if (cc >= tap->cap_chunks) return RH_ERR_NOMEM;
		CUDA_TRY(cudaMemsetAsync(c->d_counters.p, 0, sizeof(unsigned long long), s));
		round_io io; io.tap = 1; io.tap_out = tap; io.tap_off = toff; io.tap_chunk = cc;
		slot_t sl; memset(&sl, 0, sizeof(sl));
		sl.read = 0; sl.chunk_len = (uint32_t)std::min<uint64_t>(l_chunk, qlen - s_qs); sl.c_count = cc;
		io.slots.push_back(sl);
		if ((rc = run_round(c, io, cc & 1))) return rc;
		const slot_t &r = io.slots[0];
		int32_t *cnt = tap->cnt + (size_t)cc * RH_TAP_NCNT;
		memset(cnt, 0, sizeof(int32_t) * RH_TAP_NCNT);
		cnt[RH_TAP_NSIG] = r.n_sig; cnt[RH_TAP_NEVENTS] = r.n_events;
		if (toff[0] + r.n_events > tap->cap_events) return RH_ERR_NOMEM;
		if (r.n_events) CUDA_TRY(cudaMemcpyAsync(tap->events + toff[0], c->d_events.p + r.e_off, r.n_events * 4, cudaMemcpyDeviceToHost, s));
		toff[0] += r.n_events;
		if (r.gated) { CUDA_TRY(cudaStreamSynchronize(s)); continue; }
		cnt[RH_TAP_NSEEDS] = r.n_seeds; cnt[RH_TAP_NANCHORS] = r.n_anchors; cnt[RH_TAP_NU] = r.n_u; cnt[RH_TAP_NV] = r.n_v;
		cnt[RH_TAP_NREGS] = r.n_regs; cnt[RH_TAP_REPLEN] = r.rep_len;
		if (toff[1] + r.n_seeds > tap->cap_seeds || toff[3] + r.n_u > tap->cap_u || toff[4] + r.n_v > tap->cap_chain_a || toff[4] + r.n_v > tap->cap_prev_a ||
		    toff[5] + r.n_regs > tap->cap_regs) return RH_ERR_NOMEM;
		std::vector<uint32_t> sh(r.n_seeds), sp(r.n_seeds);
		if (r.n_seeds) {
			CUDA_TRY(cudaMemcpyAsync(sh.data(), c->d_seed_hash.p + r.e_off, r.n_seeds * 4, cudaMemcpyDeviceToHost, s));
			CUDA_TRY(cudaMemcpyAsync(sp.data(), c->d_seed_pos.p + r.e_off, r.n_seeds * 4, cudaMemcpyDeviceToHost, s));
		}
		std::vector<dev_reg_t> regs(r.n_regs);
		read_state_t rs;
		CUDA_TRY(cudaMemcpyAsync(&rs, c->d_rs.p, sizeof(rs), cudaMemcpyDeviceToHost, s));
		if (r.n_anchors) {
			const uint64_t n = r.n_anchors;
			const uint8_t *base = c->d_arena.p + r.a_off;
			if (r.n_u) CUDA_TRY(cudaMemcpyAsync(tap->u + toff[3], base + 80 * n, r.n_u * 8, cudaMemcpyDeviceToHost, s));   /* slot_mem: U  */
			if (r.n_v) CUDA_TRY(cudaMemcpyAsync(tap->chain_a + 2 * toff[4], base, (size_t)r.n_v * 16, cudaMemcpyDeviceToHost, s)); /* A */
			if (r.n_regs) CUDA_TRY(cudaMemcpyAsync(regs.data(), base + 96 * n + fin_regs_off(n), r.n_regs * sizeof(dev_reg_t), cudaMemcpyDeviceToHost, s));
		}
		CUDA_TRY(cudaStreamSynchronize(s));
		if (r.n_v) CUDA_TRY(cudaMemcpy(tap->prev_a + 2 * toff[4], c->d_carry[(cc & 1) ^ 1].p + rs.prev_off, (size_t)r.n_v * 16, cudaMemcpyDeviceToHost));
		const uint64_t span = (uint64_t)(P.k + P.e - 1);
		for (uint32_t i = 0; i < r.n_seeds; ++i) { tap->seeds[2 * (toff[1] + i)] = (uint64_t)sh[i] << 6 | span; tap->seeds[2 * (toff[1] + i) + 1] = (uint64_t)sp[i] << 1; }
		for (uint32_t i = 0; i < r.n_regs; ++i) {
			const dev_reg_t &g = regs[i];
			int32_t *f = tap->regs + (toff[5] + i) * RH_TAP_REG_NF;
			f[0] = g.score; f[1] = g.cnt; f[2] = g.rid; f[3] = (int32_t)g.rev; f[4] = g.qs; f[5] = g.qe; f[6] = g.rs; f[7] = g.re;
			f[8] = g.parent; f[9] = g.subsc; f[10] = g.n_sub; f[11] = (int32_t)g.mapq; f[12] = g.as; f[13] = g.score0;
		}
		toff[1] += r.n_seeds; toff[3] += r.n_u; toff[4] += r.n_v; toff[5] += r.n_regs;
	}
	tap->n_chunks = cc;
	rc = check_dev_err(c);
	collect_spans(c);
	return rc;
}

/* Rawsamble index (ri_idx_siggen / worker_sig_pipeline, src/rindex.c:239-309,927-969): whole-read
 * event detection with fresh sums on the GPU (the same k_signal_to_seeds phases A-C), sketching of
 * the target side on the host like the sequence index. */
extern "C" rh_index_t *rh_index_build_sig(const rh_params_t *p, uint32_t n_reads, const char *const *names,
                                           const int16_t *const *raw, const uint64_t *raw_len,
                                           const double *offset, const double *range, const double *digitisation)
{
	if (!p || (n_reads && (!names || !raw || !raw_len))) { rh_set_error("rh_index_build_sig: null argument"); return NULL; }
	rh_index_s empty;
	rh_params_t q = *p; q.w = 0; q.mid_occ = 1; /* event detection only; w/mid_occ are irrelevant to it */
	q.map_flag |= RH_M_NO_ADAPTIVE;
	rh_gpu_ctx *ctx = rh_gpu_init(&empty, &q, 0, (size_t)256 << 20);
	if (!ctx) return NULL;
	rh_worker *c = ctx->workers[0];
	rh_index_s *idx = new rh_index_s();
	idx->flag = p->idx_flag; idx->w = p->w; idx->e = p->e; idx->n = p->n; idx->q = p->q; idx->k = p->k;
	idx->diff = p->diff; idx->fine_min = p->fine_min; idx->fine_max = p->fine_max; idx->fine_range = p->fine_range;
	std::vector<rh_seed_t> all;
	int rc = RH_OK;
	const uint32_t B = 2048;
	for (uint32_t b0 = 0; b0 < n_reads && rc == RH_OK; b0 += B) {
		const uint32_t bn = std::min(B, n_reads - b0);
		begin_call(c);
		std::vector<uint64_t> beg, len;
		rc = upload_raw(ctx, c, bn, raw + b0, raw_len + b0, beg, len);
		if (rc) break;
		batch_in in{bn, raw + b0, raw_len + b0, nullptr, nullptr, offset + b0, range + b0, digitisation + b0, names + b0};
		std::vector<uint32_t> l_sig;
		if ((rc = setup_reads(c, in, beg, len, l_sig))) break;
		round_io io;
		for (uint32_t i = 0; i < bn; ++i) {
			if (l_sig[i] == 0) continue;
			slot_t sl; memset(&sl, 0, sizeof(sl)); sl.read = i; sl.chunk_len = l_sig[i];
			io.slots.push_back(sl);
		}
		/* events only: run phases A-D, then read the events back */
		const uint32_t ns = (uint32_t)io.slots.size();
		std::vector<float> ev; std::vector<uint32_t> ev_off(ns), ev_n(ns);
		if (ns) {
			uint64_t zt = 0, et = 0;
			for (slot_t &sl : io.slots) { sl.z_off = zt; zt += ((uint64_t)sl.chunk_len + 3) & ~3ULL; /* chunks start 16-byte aligned */ sl.e_cap = (uint32_t)((uint64_t)sl.chunk_len * 2 / 3 + 8); sl.e_off = et; et += sl.e_cap; }
			if ((rc = c->d_z.reserve(zt)) || (rc = c->d_events.reserve(et)) || (rc = c->d_peaks.reserve(et)) || (rc = c->d_seed_hash.reserve(et)) ||
			    (rc = c->d_seed_pos.reserve(et)) || (rc = upload(c->d_slots, io.slots, c->stream))) break;
			if ((rc = c->d_ps.reserve(zt + 4 * (uint64_t)ns + 4)) || (rc = c->d_pq.reserve(zt + 4 * (uint64_t)ns + 4)) || (rc = c->d_t1.reserve(zt + 4 * (uint64_t)ns + 4)) || (rc = c->d_t2.reserve(zt + 4 * (uint64_t)ns + 4))) break;
			sig_args_t a1; a1.prof = nullptr; a1.min_hash = nullptr; a1.min_pos = nullptr;
			a1.raw = c->raw_ptr; a1.rs = c->d_rs.p; a1.slots = c->d_slots.p; a1.n_slots = ns;
			a1.z = c->d_z.p; a1.ps = c->d_ps.p; a1.pq = c->d_pq.p; a1.t1 = c->d_t1.p; a1.t2 = c->d_t2.p;
			a1.peaks = c->d_peaks.p; a1.events = c->d_events.p; a1.seed_hash = c->d_seed_hash.p; a1.seed_pos = c->d_seed_pos.p;
			k_sig_norm<<<(ns * 32 + 127) / 128, 128, 0, c->stream>>>(a1);
			k_sig_prefix<<<(ns + 127) / 128, 128, 0, c->stream>>>(a1);
			k_sig_tstat<<<ns, 256, 0, c->stream>>>(a1, c->D);
			k_sig_peaks<<<(ns + 127) / 128, 128, 0, c->stream>>>(a1, c->D);
			k_sig_events_fast<<<ns, EV_THREADS, 0, c->stream>>>(a1);
			k_sig_events<<<ns, EV_THREADS, 0, c->stream>>>(a1);
			if (cudaMemcpyAsync(io.slots.data(), c->d_slots.p, ns * sizeof(slot_t), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) { rc = RH_ERR_CUDA; break; }
			ev.resize(et);
			if (cudaMemcpyAsync(ev.data(), c->d_events.p, et * 4, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) { rc = RH_ERR_CUDA; break; }
			if (cudaStreamSynchronize(c->stream) != cudaSuccess) { rh_set_error("event kernel failed: %s", cudaGetErrorString(cudaGetLastError())); rc = RH_ERR_CUDA; break; }
		}
		size_t si = 0;
		for (uint32_t i = 0; i < bn; ++i) {
			idx->names.emplace_back(names[b0 + i]); idx->lens.push_back(l_sig[i]);
			if (l_sig[i] == 0) continue;
			const slot_t &sl = io.slots[si++];
			if (sl.n_peaks) rh_host_sketch(*p, ev.data() + sl.e_off, sl.n_peaks, b0 + i, 0, all);
		}
	}
	rh_gpu_destroy(ctx);
	if (rc != RH_OK) { delete idx; return NULL; }
	rh_index_from_seeds(idx, all, 8);
	return idx;
}
