/*
 * rh_host.cpp — host side of the C-ABI (include/rawhash_b200.h): option presets, pore-model
 * parsing, index build / `.ind` loading / flattening, PAF formatting.
 *
 * None of this is on the per-read hot path (that is rh_gpu.cu); it is the reference's
 * surrounding host surface kept so the library drops in where rawhash2's own host code sits.
 * Reference citations are relative to the RawHash tree.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include <math.h>
#include <limits.h>
#include <algorithm>
#include <atomic>
#include <thread>
#include <mutex>

#include "rh_host.h"

static thread_local char g_err[512] = "";
void rh_set_error(const char *fmt, ...)
{
	va_list ap; va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
}
extern "C" const char *rh_gpu_last_error(void) { return g_err; }
extern "C" void rh_free(void *p) { free(p); }

/* ---- options -------------------------------------------------------------------------------- */
extern "C" void rh_params_init(rh_params_t *p)
{ /* ri_idxopt_init + ri_mapopt_init, src/roptions.c:4-138 */
	memset(p, 0, sizeof(*p));
	p->e = 8; p->q = 4; p->k = 6; p->lev_col = 1;
	p->diff = 0.35f; p->fine_min = -2.0f; p->fine_max = 2.0f; p->fine_range = 0.4;
	p->window_length1 = 3; p->window_length2 = 9;
	p->threshold1 = 4.0f; p->threshold2 = 3.5f; p->peak_height = 0.4f;
	p->bp_per_sec = 450; p->sample_rate = 4000; p->chunk_size = 4000;
	p->sample_per_base = (float)p->sample_rate / p->bp_per_sec;
	p->mid_occ_frac = 1e-2f; p->min_mid_occ = 50; p->max_mid_occ = 500000;
	p->min_events = 50;
	p->bw = 500; p->max_target_gap_length = 2500; p->max_query_gap_length = 2500;
	p->max_chain_iter = 200; p->max_num_skips = 5; p->min_num_anchors = 2;
	p->min_chaining_score = 15; p->min_chaining_score2 = 0;
	p->chain_gap_scale = 0.8f; p->chain_skip_scale = 0.0f;
	p->mask_level = 0.5f; p->mask_len = INT_MAX; p->pri_ratio = 0.3f; p->best_n = 0; p->alt_drop = 0.15f;
	p->w_bestq = 0.35f; p->w_bestmq = 0.05f; p->w_bestmc = 0.6f; p->w_threshold = 0.45f;
	p->max_num_chunk = 10; p->min_mapq = 2;
}

static void overlap_mode(rh_params_t *p)
{ /* shared tail of the four Rawsamble presets, src/main.cpp:128-201 */
	p->diff = 0.45f;
	p->max_target_gap_length = 2500; p->max_query_gap_length = 2500;
	p->idx_flag |= RH_I_SIG_TARGET;
	p->map_flag |= RH_M_ALL_CHAINS | RH_M_NO_ADAPTIVE;
	p->pri_ratio = 0.0f;
}

extern "C" int rh_params_preset(rh_params_t *p, const char *preset)
{ /* ri_set_opt, src/main.cpp:111-210 */
	if (!preset || !*preset || !strcmp(preset, "sensitive") || !strcmp(preset, "sequence-until")) return RH_OK;
	if (!strcmp(preset, "viral")) {
		p->e = 6; p->bw = 100; p->max_target_gap_length = p->max_query_gap_length = 500;
		p->max_num_chunk = 5; p->min_chaining_score = 10; p->chain_gap_scale = 1.2f; p->chain_skip_scale = 0.3f;
	} else if (!strcmp(preset, "fast")) {
		p->fine_range = 0.6; p->min_mapq = 5; p->min_chaining_score = 10; p->chain_gap_scale = 0.6f;
	} else if (!strcmp(preset, "faster")) {
		p->e = 11; p->w = 3; p->fine_range = 0.6;
		p->max_num_chunk = 5; p->min_mapq = 5; p->min_chaining_score = 10; p->chain_gap_scale = 0.6f;
	} else if (!strcmp(preset, "ava-viral")) {
		overlap_mode(p);
		p->e = 6; p->w = 0; p->chain_gap_scale = 1.2f; p->chain_skip_scale = 0.3f;
		p->min_chaining_score = 20; p->min_chaining_score2 = 30; p->min_num_anchors = 5; p->min_mapq = 5; p->bw = 1000;
	} else if (!strcmp(preset, "ava")) {
		overlap_mode(p);
		p->w = 3; p->min_chaining_score = 40; p->min_chaining_score2 = 75; p->min_num_anchors = 5; p->min_mapq = 5; p->bw = 5000;
	} else if (!strcmp(preset, "ava-sensitive")) {
		overlap_mode(p);
		p->w = 0; p->min_chaining_score = 75; p->min_chaining_score2 = 100; p->min_num_anchors = 5; p->min_mapq = 5; p->bw = 1000;
	} else if (!strcmp(preset, "ava-large")) {
		overlap_mode(p);
		p->fine_range = 0.6; p->chain_gap_scale = 0.6f; p->w = 5;
		p->min_chaining_score = 20; p->min_chaining_score2 = 50; p->min_num_anchors = 2; p->min_mapq = 2; p->bw = 5000;
	} else return RH_ERR_ARG;
	return RH_OK;
}

extern "C" void rh_params_r10(rh_params_t *p)
{ /* --r10, src/main.cpp:363-376 */
	p->k = 9;
	p->window_length1 = 3; p->window_length2 = 6;
	p->threshold1 = 6.5f; p->threshold2 = 4.0f; p->peak_height = 0.2f;
	p->chain_gap_scale = 1.2f;
}

/* ---- pore model ------------------------------------------------------------------------------ */
extern "C" int rh_pore_load(const char *path, int k, int lev_col, float **vals, uint32_t *n_vals)
{ /* load_pore, src/rutils.c:133-178: levels z-normalised with the population mean / std
     (variance evaluated as fma(-mean, mean, E[x^2]), the contraction the reference is built with) */
	FILE *fp = fopen(path, "r");
	if (!fp) { rh_set_error("cannot open pore model %s", path); return RH_ERR_IO; }
	size_t cap = (size_t)1 << (2 * k), n = 0;
	float *v = (float *)malloc(cap * sizeof(float));
	double acc = 0, acc2 = 0;
	char line[1024];
	while (fgets(line, sizeof(line), fp)) {
		if (strncmp(line, "kmer", 4) == 0) continue;
		char *cursor = line, *field; int col = 0;
		while ((field = strsep(&cursor, "\t")) != NULL) {
			if (col++ != lev_col) continue;
			float level;
			if (n >= cap || sscanf(field, "%f", &level) != 1) { fclose(fp); free(v); rh_set_error("bad pore model line %zu", n); return RH_ERR_FORMAT; }
			v[n] = level; acc += level; acc2 += level * level;
			break;
		}
		++n;
	}
	fclose(fp);
	if (n == 0) { free(v); return RH_ERR_FORMAT; }
	double mean = acc / (int)n, sd = sqrt(fma(-mean, mean, acc2 / (int)n));
	for (size_t i = 0; i < n && i < cap; ++i) v[i] = (v[i] - mean) / sd;
	*vals = v; *n_vals = (uint32_t)cap;
	return RH_OK;
}

/* ---- host sketch (index side) ----------------------------------------------------------------- */
static inline uint64_t seed_mix(uint64_t key, uint64_t mask)
{ /* hash64, src/rsketch.c:7-16 */
	key = (~key + (key << 21)) & mask; key ^= key >> 24;
	key = (key + (key << 3) + (key << 8)) & mask; key ^= key >> 14;
	key = (key + (key << 2) + (key << 4)) & mask; key ^= key >> 28;
	key = (key + (key << 31)) & mask;
	return key;
}

static inline uint32_t bucket_of(float v, const rh_params_t &P)
{ /* dynamic_quantize, src/rsketch.c:18-53 (no FMA: the reference object has none here) */
	const float lo = -3.0f, width = 6.0f;
	float c1 = (1 - P.fine_range) / 2, c2 = P.fine_range + c1;
	float u = (v - lo) / width, a = (P.fine_min - lo) / width, b = (P.fine_max - lo) / width, r;
	if (v >= P.fine_min && v <= P.fine_max) r = P.fine_range * ((u - a) / (b - a));
	else { float t = c1 * u; r = (u < 0.5) ? P.fine_range + t : c2 + t; }
	return (uint32_t)(r * ((1u << P.q) - 1));
}

void rh_host_sketch(const rh_params_t &P, const float *ev, uint32_t len, uint32_t id, int strand, std::vector<rh_seed_t> &out)
{ /* ri_sketch, src/rsketch.c:271-290.  Step 1: kept events (diff filter) -> (hash, first-event y).
     Step 2 (w>0): minimizer selection over that stream. */
	if (len == 0) return;
	const int e = P.e, q = P.q, w = P.w;
	const uint64_t span = P.k + e - 1, idb = (uint64_t)id << 32, m32 = 0xffffffffULL;
	const uint64_t mev = (q * e >= 64) ? ~0ULL : ((1ULL << (q * e)) - 1), mq = (1ULL << q) - 1;
	std::vector<uint32_t> kept_pos;
	std::vector<rh_seed_t> stream; /* one entry per kept event index >= e-1 */
	uint64_t packed = 0; uint32_t last = 0;
	for (uint32_t i = 0; i < len; ++i) {
		if (i && fabsf(ev[i] - ev[last]) < P.diff) continue;
		last = i;
		packed = ((packed << q) | (bucket_of(ev[i], P) & mq)) & mev;
		kept_pos.push_back(i);
		size_t m = kept_pos.size();
		if (m >= (size_t)e)
			stream.push_back({seed_mix(packed, m32) << 6 | span, idb | (uint64_t)(kept_pos[m - e] << 1) | (uint64_t)strand});
	}
	if (w == 0) { out.insert(out.end(), stream.begin(), stream.end()); return; }
	/* minimizers: ri_sketch_min, src/rsketch.c:94-140 (minimap2 window logic incl. equal minima) */
	std::vector<rh_seed_t> win(w, rh_seed_t{UINT64_MAX, UINT64_MAX});
	rh_seed_t mn = {UINT64_MAX, UINT64_MAX};
	int wp = 0, mp = 0;
	for (size_t t = 0; t < stream.size(); ++t) {
		const uint32_t l = (uint32_t)(t + e); /* kept events seen so far */
		const rh_seed_t cur = stream[t];
		win[wp] = cur;
		if (l == (uint32_t)(w + e - 1) && mn.x != UINT64_MAX) {
			for (int j = wp + 1; j < w; ++j) if (mn.x == win[j].x && win[j].y != mn.y) out.push_back(win[j]);
			for (int j = 0; j < wp; ++j) if (mn.x == win[j].x && win[j].y != mn.y) out.push_back(win[j]);
		}
		if (cur.x <= mn.x) {
			if (l >= (uint32_t)(w + e) && mn.x != UINT64_MAX) out.push_back(mn);
			mn = cur; mp = wp;
		} else if (wp == mp) {
			if (l >= (uint32_t)(w + e - 1) && mn.x != UINT64_MAX) out.push_back(mn);
			mn.x = UINT64_MAX;
			for (int j = wp + 1; j < w; ++j) if (mn.x >= win[j].x) { mn = win[j]; mp = j; }
			for (int j = 0; j <= wp; ++j) if (mn.x >= win[j].x) { mn = win[j]; mp = j; }
			if (l >= (uint32_t)(w + e - 1) && mn.x != UINT64_MAX) {
				for (int j = wp + 1; j < w; ++j) if (mn.x == win[j].x && mn.y != win[j].y) out.push_back(win[j]);
				for (int j = 0; j <= wp; ++j) if (mn.x == win[j].x && mn.y != win[j].y) out.push_back(win[j]);
			}
		}
		if (++wp == w) wp = 0;
	}
	if (mn.x != UINT64_MAX) out.push_back(mn);
}

/* ---- index ------------------------------------------------------------------------------------ */
void rh_index_from_seeds(rh_index_s *idx, std::vector<rh_seed_t> &seeds, int n_threads)
{ /* group by hash, positions ascending within a key (what worker_post leaves, src/rindex.c:311-363) */
	const size_t n = seeds.size();
	/* partition by the top 8 hash bits, sort partitions in parallel */
	std::vector<size_t> cnt(257, 0);
	for (size_t i = 0; i < n; ++i) ++cnt[(seeds[i].x >> 30) + 1];
	for (int b = 0; b < 256; ++b) cnt[b + 1] += cnt[b];
	std::vector<rh_seed_t> tmp(n);
	{
		std::vector<size_t> cur(cnt.begin(), cnt.end() - 1);
		for (size_t i = 0; i < n; ++i) tmp[cur[seeds[i].x >> 30]++] = seeds[i];
	}
	seeds.swap(tmp);
	std::vector<rh_seed_t>().swap(tmp);
	std::atomic<int> next(0);
	auto work = [&]() {
		for (int b; (b = next++) < 256;)
			std::sort(seeds.begin() + cnt[b], seeds.begin() + cnt[b + 1], [](const rh_seed_t &a, const rh_seed_t &c) {
				uint64_t ha = a.x >> 6, hc = c.x >> 6;
				return ha != hc ? ha < hc : a.y < c.y; });
	};
	std::vector<std::thread> th;
	for (int t = 0; t < std::max(1, n_threads); ++t) th.emplace_back(work);
	for (auto &t : th) t.join();
	idx->keys.clear(); idx->off.clear(); idx->pos.resize(n);
	for (size_t i = 0; i < n; ++i) {
		uint32_t h = (uint32_t)(seeds[i].x >> 6);
		if (i == 0 || h != idx->keys.back()) { idx->keys.push_back(h); idx->off.push_back(i); }
		idx->pos[i] = seeds[i].y;
	}
	idx->off.push_back(n);
}

static void index_set_params(rh_index_s *idx, const rh_params_t *p)
{
	idx->flag = p->idx_flag; idx->w = p->w; idx->e = p->e; idx->n = p->n; idx->q = p->q; idx->k = p->k;
	idx->diff = p->diff; idx->fine_min = p->fine_min; idx->fine_max = p->fine_max; idx->fine_range = p->fine_range;
}

static const unsigned char *base_code()
{
	static unsigned char t[256]; static std::once_flag once;
	std::call_once(once, []() { memset(t, 4, 256); t['A'] = t['a'] = 0; t['C'] = t['c'] = 1; t['G'] = t['g'] = 2; t['T'] = t['t'] = 3; });
	return t;
}

uint64_t rh_index_group_bases(uint64_t dflt)
{
	const char *e = getenv("RH_INDEX_GROUP_BASES");
	if (!e || !*e) return dflt;
	const unsigned long long v = strtoull(e, NULL, 10);
	return v ? (uint64_t)v : dflt;
}

rh_index_t *rh_index_build_grouped(rh_index_builder_fn one, uint64_t group_bases, const rh_params_t *p, const float *pore_vals, uint32_t n_pore_vals,
                                   uint32_t n_seq, const char *const *names, const char *const *seqs, const uint32_t *lens, int arg)
{
	std::vector<uint32_t> start; /* first sequence of every group */
	uint64_t in_group = 0;
	for (uint32_t i = 0; i < n_seq; ++i) {
		if (i == 0 || in_group + lens[i] > group_bases) { start.push_back(i); in_group = 0; }
		in_group += lens[i];
	}
	start.push_back(n_seq);
	const size_t G = start.size() - 1;
	std::vector<rh_index_s *> part(G, nullptr);
	auto drop = [&]() { for (rh_index_s *q : part) delete q; };
	for (size_t g = 0; g < G; ++g) {
		const uint32_t a = start[g], m = start[g + 1] - a;
		part[g] = one(p, pore_vals, n_pore_vals, m, names + a, seqs + a, lens + a, arg);
		if (!part[g]) { drop(); return NULL; }
	}
	rh_index_s *idx = new rh_index_s();
	index_set_params(idx, p);
	for (uint32_t i = 0; i < n_seq; ++i) { idx->names.emplace_back(names[i]); idx->lens.push_back(lens[i]); }
	/* sweep 1: union of the sorted key arrays with the summed list lengths */
	std::vector<size_t> at(G, 0);
	uint64_t total = 0;
	for (;;) {
		uint64_t key = UINT64_MAX;
		for (size_t g = 0; g < G; ++g) if (at[g] < part[g]->keys.size() && part[g]->keys[at[g]] < key) key = part[g]->keys[at[g]];
		if (key == UINT64_MAX) break;
		idx->keys.push_back((uint32_t)key); idx->off.push_back(total);
		for (size_t g = 0; g < G; ++g) if (at[g] < part[g]->keys.size() && part[g]->keys[at[g]] == key) { total += part[g]->off[at[g] + 1] - part[g]->off[at[g]]; ++at[g]; }
	}
	idx->off.push_back(total);
	idx->pos.resize(total);
	/* sweep 2: lists in group order, sequence ids rebased to the whole reference; key ranges are independent, so the
	 * copy (tens of GB at human size) runs on all host threads */
	{
		const size_t nk = idx->keys.size();
		const unsigned T = (unsigned)std::max<size_t>(1, std::min<size_t>(std::thread::hardware_concurrency(), nk / 4096 + 1));
		auto work = [&](unsigned t) {
			const size_t k0 = nk * t / T, k1 = nk * (t + 1) / T;
			if (k0 == k1) return;
			std::vector<size_t> cur(G);
			for (size_t g = 0; g < G; ++g) cur[g] = (size_t)(std::lower_bound(part[g]->keys.begin(), part[g]->keys.end(), idx->keys[k0]) - part[g]->keys.begin());
			for (size_t ki = k0; ki < k1; ++ki) {
				uint64_t w = idx->off[ki];
				for (size_t g = 0; g < G; ++g) {
					const rh_index_s *q = part[g];
					if (cur[g] >= q->keys.size() || q->keys[cur[g]] != idx->keys[ki]) continue;
					const uint64_t rebase = (uint64_t)start[g] << 32;
					for (uint64_t j = q->off[cur[g]]; j < q->off[cur[g] + 1]; ++j) idx->pos[w++] = q->pos[j] + rebase;
					++cur[g];
				}
			}
		};
		std::vector<std::thread> th;
		for (unsigned t = 1; t < T; ++t) th.emplace_back(work, t);
		work(0);
		for (auto &x : th) x.join();
	}
	drop();
	return idx;
}

static rh_index_t *index_build_host_one(const rh_params_t *p, const float *pore_vals, uint32_t n_pore_vals,
                                        uint32_t n_seq, const char *const *names, const char *const *seqs,
                                        const uint32_t *lens, int n_threads);

extern "C" rh_index_t *rh_index_build(const rh_params_t *p, const float *pore_vals, uint32_t n_pore_vals,
                                       uint32_t n_seq, const char *const *names, const char *const *seqs,
                                       const uint32_t *lens, int n_threads)
{
	if (!p || !pore_vals || n_pore_vals < (1u << (2 * p->k)) || (n_seq && (!names || !seqs || !lens))) { rh_set_error("rh_index_build: bad arguments"); return NULL; }
	uint64_t total = 0;
	for (uint32_t i = 0; i < n_seq; ++i) total += lens[i];
	const uint64_t group = rh_index_group_bases(UINT64_MAX); /* the host builder needs no grouping; the env knob exists to test the merge */
	if (n_seq > 1 && total > group) return rh_index_build_grouped(index_build_host_one, group, p, pore_vals, n_pore_vals, n_seq, names, seqs, lens, n_threads);
	return index_build_host_one(p, pore_vals, n_pore_vals, n_seq, names, seqs, lens, n_threads);
}

static rh_index_t *index_build_host_one(const rh_params_t *p, const float *pore_vals, uint32_t n_pore_vals,
                                        uint32_t n_seq, const char *const *names, const char *const *seqs,
                                        const uint32_t *lens, int n_threads)
{ /* ri_idx_gen, src/rindex.c:900-925: expected signal of both strands (ri_seq_to_sig,
     src/rsig.c:13-40) -> sketch -> grouped by hash */
	rh_index_s *idx = new rh_index_s();
	index_set_params(idx, p);
	std::vector<std::vector<rh_seed_t>> per(n_seq);
	std::atomic<uint32_t> next(0);
	const unsigned char *code = base_code();
	const int k = p->k;
	auto work = [&]() {
		std::vector<float> sv;
		for (uint32_t i; (i = next++) < n_seq;) {
			const int len = (int)lens[i];
			if (len <= 0) continue;
			const uint64_t mask = (1ULL << 2 * k) - 1;
			for (int strand = 0; strand < 2; ++strand) {
				uint64_t kmer = 0;
				sv.clear();
				for (int j = 0; j < len; ++j) {
					int c = code[(uint8_t)seqs[i][strand ? len - 1 - j : j]];
					if (c < 4) kmer = ((kmer << 2) | (uint64_t)(strand ? 3 - c : c)) & mask;
					if (j + 1 >= k) sv.push_back(pore_vals[kmer]);
				}
				rh_host_sketch(*p, sv.data(), (uint32_t)sv.size(), i, strand, per[i]);
			}
		}
	};
	std::vector<std::thread> th;
	for (int t = 0; t < std::max(1, n_threads); ++t) th.emplace_back(work);
	for (auto &t : th) t.join();
	size_t total = 0;
	for (auto &v : per) total += v.size();
	std::vector<rh_seed_t> all; all.reserve(total);
	for (uint32_t i = 0; i < n_seq; ++i) {
		idx->names.emplace_back(names[i]); idx->lens.push_back(lens[i]);
		all.insert(all.end(), per[i].begin(), per[i].end());
		std::vector<rh_seed_t>().swap(per[i]);
	}
	rh_index_from_seeds(idx, all, n_threads);
	return idx;
}

/* ---- layout of one chunk round --------------------------------------------------------------------------
 * Input: the anchor counts of ns chunks, the first n_mandatory of which must run in this round (their reads carry chain
 * anchors in the round's input arena); the rest are first chunks of waiting reads, in admission order, of which at most
 * max_optional may be admitted.  Output: the order the chunks take (mandatory ones heaviest first: CTAs are handed out in
 * slot order and a kernel lasts as long as its slowest chunk), the groups that fill the arena one after the other, each
 * chunk's region, and — when asked for — a heavy group of the largest chunks (more than 1.25 x the mean, at most a quarter
 * of the mandatory chunks and an eighth of the arena, at least two) that runs beside the ordinary groups in a slice at the
 * top of the arena.  Waiting reads only fill what the last ordinary group of mandatory chunks leaves free (a whole arena
 * when nothing is mandatory), in order, up to the first that does not fit. */
int rh_plan_round_impl(const uint32_t *n_anchors, uint32_t ns, uint32_t n_mandatory, uint32_t max_optional, uint64_t arena_bytes,
                       bool heaviest_first, bool heavy_lane, rh_round_plan_t *plan)
{
	const uint32_t n_mand = std::min(n_mandatory, ns);
	plan->order.resize(ns);
	for (uint32_t q = 0; q < ns; ++q) plan->order[q] = q;
	if (heaviest_first) std::stable_sort(plan->order.begin(), plan->order.begin() + n_mand, [&](uint32_t a, uint32_t b) { return n_anchors[a] > n_anchors[b]; });
	auto size_of = [&](uint32_t q) { return rh_slot_region_bytes(n_anchors[plan->order[q]]); };
	plan->a_off.assign(ns, 0);
	plan->groups.clear();
	uint32_t nh = 0; uint64_t heavy_bytes = 0;
	if (heavy_lane && heaviest_first && n_mand >= 64) {
		unsigned long long sum = 0;
		for (uint32_t q = 0; q < n_mand; ++q) sum += n_anchors[plan->order[q]];
		const double avg = (double)sum / n_mand;
		const uint64_t cap = arena_bytes / 8;
		while (nh < n_mand / 4 && n_anchors[plan->order[nh]] > 1.25 * avg && heavy_bytes + size_of(nh) <= cap) heavy_bytes += size_of(nh++);
		if (nh < 2) { nh = 0; heavy_bytes = 0; }
	}
	const uint64_t main_bytes = (arena_bytes - heavy_bytes) & ~(uint64_t)255; /* regions start 256-byte aligned */
	plan->main_bytes = main_bytes;
	if (nh) {
		uint64_t off = main_bytes;
		for (uint32_t q = 0; q < nh; ++q) { plan->a_off[q] = off; off += size_of(q); }
		plan->groups.push_back({0u, nh, 1u});
	}
	uint32_t n_total = n_mand, g0 = nh;
	while (g0 < n_total || (g0 == nh && n_mand == nh && ns > nh)) {
		uint64_t used = 0; uint32_t g1 = g0;
		while (g1 < n_total) {
			const uint64_t need = size_of(g1);
			if (need > main_bytes) { rh_set_error("a single chunk needs %llu bytes of anchor arena (have %llu)", (unsigned long long)need, (unsigned long long)main_bytes); return RH_ERR_NOMEM; }
			if (used + need > main_bytes) break;
			plan->a_off[g1] = used; used += need; ++g1;
		}
		if (g1 == n_mand && n_total == n_mand) {
			while (g1 < ns && g1 - n_mand < max_optional) {
				const uint64_t need = size_of(g1);
				if (need > main_bytes) { rh_set_error("a single chunk needs %llu bytes of anchor arena (have %llu)", (unsigned long long)need, (unsigned long long)main_bytes); return RH_ERR_NOMEM; }
				if (used + need > main_bytes) break;
				plan->a_off[g1] = used; used += need; ++g1;
			}
			n_total = g1;
			if (g1 == g0) break; /* nothing to run */
		}
		plan->groups.push_back({g0, g1 - g0, 0u});
		g0 = g1;
	}
	plan->n_run = n_total;
	return RH_OK;
}

extern "C" int rh_plan_round(const uint32_t *n_anchors, uint32_t n_chunks, uint32_t n_mandatory, uint32_t max_optional, uint64_t arena_bytes, int heavy_lane,
                             uint32_t *order, uint64_t *a_off, uint32_t *groups, uint32_t groups_cap, uint32_t *n_groups, uint32_t *n_run, uint64_t *main_bytes)
{
	if ((n_chunks && (!n_anchors || !order || !a_off)) || !n_groups || !n_run || (groups_cap && !groups)) { rh_set_error("rh_plan_round: null argument"); return RH_ERR_ARG; }
	rh_round_plan_t plan;
	const int rc = rh_plan_round_impl(n_anchors, n_chunks, n_mandatory, max_optional, arena_bytes, true, heavy_lane != 0, &plan);
	if (rc != RH_OK) return rc;
	for (uint32_t q = 0; q < n_chunks; ++q) { order[q] = plan.order[q]; a_off[q] = plan.a_off[q]; }
	*n_groups = (uint32_t)plan.groups.size(); *n_run = plan.n_run;
	if (main_bytes) *main_bytes = plan.main_bytes;
	for (uint32_t g = 0; g < plan.groups.size() && g < groups_cap; ++g) { groups[3 * g] = plan.groups[g].first; groups[3 * g + 1] = plan.groups[g].count; groups[3 * g + 2] = plan.groups[g].heavy; }
	return RH_OK;
}

extern "C" rh_index_t *rh_index_load(const char *path, rh_params_t *pp)
{ /* reader for the reference's `.ind` layout, src/rindex.c:650-776 (writer 545-648) */
	FILE *f = fopen(path, "rb");
	if (!f) { rh_set_error("cannot open %s", path); return NULL; }
	char magic[2];
	uint32_t hdr[7]; float fl[4];
	if (fread(magic, 1, 2, f) != 2 || memcmp(magic, "RI", 2) != 0 || fread(hdr, 4, 7, f) != 7 || fread(fl, 4, 4, f) != 4) {
		fclose(f); rh_set_error("%s: not a RawHash index", path); return NULL;
	}
	rh_index_s *idx = new rh_index_s();
	idx->w = hdr[0]; idx->e = hdr[1]; idx->n = hdr[2]; idx->q = hdr[3]; idx->k = hdr[4]; idx->flag = hdr[6];
	idx->diff = fl[0]; idx->fine_min = fl[1]; idx->fine_max = fl[2]; idx->fine_range = fl[3];
	const uint32_t n_seq = hdr[5];
	bool ok = true;
	/* ri_pore_t is written raw: two stale pointers, then n_pore_vals (u32), k (i16), pad, 2 floats */
	unsigned char pore_raw[32];
	ok = ok && fread(pore_raw, 1, 32, f) == 32;
	uint32_t n_pore = 0; memcpy(&n_pore, pore_raw + 16, 4);
	ok = ok && fseek(f, (long)n_pore * 4 + (long)n_pore * 12, SEEK_CUR) == 0; /* pore_vals + pore_inds */
	for (uint32_t i = 0; ok && i < n_seq; ++i) {
		uint8_t l; char name[256]; uint32_t len;
		ok = ok && fread(&l, 1, 1, f) == 1;
		if (ok && l) ok = fread(name, 1, l, f) == l;
		name[l] = 0;
		ok = ok && fread(&len, 4, 1, f) == 1;
		idx->names.emplace_back(name); idx->lens.push_back(len);
		if (ok && (idx->flag & 0x10)) { /* RI_I_STORE_SIG: skip stored expected signals */
			uint32_t nf; ok = fread(&nf, 4, 1, f) == 1 && fseek(f, (long)nf * 4, SEEK_CUR) == 0;
			if (ok && !(idx->flag & 0x40)) { ok = fread(&nf, 4, 1, f) == 1 && fseek(f, (long)nf * 4, SEEK_CUR) == 0; }
		}
	}
	/* 2^14 buckets (the bucket-bit count is not stored; the reference hard-codes 14, rindex.c:669).  Two passes so that a
	 * human-size index (≈42 GB of positions) is never held twice: pass 1 reads the (key, value) pairs and skips the
	 * position arrays, pass 2 reads one bucket's positions at a time and drops each list at its final place. */
	struct ent { uint32_t hash; uint32_t bucket; uint64_t val; bool single; };
	std::vector<ent> ents;
	std::vector<int64_t> p_at(1 << 14, 0); std::vector<int32_t> p_n(1 << 14, 0);
	std::vector<uint32_t> first_ent((1 << 14) + 1, 0);
	for (uint32_t b = 0; ok && b < (1u << 14); ++b) {
		int32_t n; uint32_t size;
		ok = ok && fread(&n, 4, 1, f) == 1 && n >= 0;
		if (!ok) break;
		p_n[b] = n; p_at[b] = (int64_t)ftello(f);
		ok = fseeko(f, (off_t)n * 8, SEEK_CUR) == 0 && fread(&size, 4, 1, f) == 1;
		first_ent[b] = (uint32_t)ents.size();
		for (uint32_t j = 0; ok && j < size; ++j) {
			uint64_t kv[2];
			ok = fread(kv, 8, 2, f) == 2;
			const bool single = (kv[0] & 1) != 0;
			if (ok && !single && ((kv[1] >> 32) + (uint32_t)kv[1] > (uint64_t)n || (uint32_t)kv[1] == 0)) ok = false; /* list outside the bucket's array */
			ents.push_back({(uint32_t)(((kv[0] >> 1) << 14) | b), b, kv[1], single});
		}
	}
	first_ent[1 << 14] = (uint32_t)ents.size();
	if (ok) {
		const size_t nk = ents.size();
		std::vector<uint32_t> order(nk);
		for (size_t i = 0; i < nk; ++i) order[i] = (uint32_t)i;
		std::sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return ents[x].hash < ents[y].hash; });
		idx->keys.resize(nk); idx->off.resize(nk + 1);
		std::vector<uint32_t> rank(nk);
		uint64_t total = 0;
		for (size_t i = 0; i < nk; ++i) {
			const ent &e = ents[order[i]];
			if (i && e.hash == ents[order[i - 1]].hash) { ok = false; break; } /* a key stored twice */
			idx->keys[i] = e.hash; idx->off[i] = total; rank[order[i]] = (uint32_t)i;
			total += e.single ? 1 : (uint32_t)e.val;
		}
		idx->off[nk] = total;
		if (ok) idx->pos.resize(total);
		std::vector<uint64_t> bucket_pos;
		for (uint32_t b = 0; ok && b < (1u << 14); ++b) {
			if (first_ent[b] == first_ent[b + 1]) continue;
			bucket_pos.resize((size_t)p_n[b]);
			if (p_n[b]) ok = fseeko(f, (off_t)p_at[b], SEEK_SET) == 0 && fread(bucket_pos.data(), 8, (size_t)p_n[b], f) == (size_t)p_n[b];
			for (uint32_t j = first_ent[b]; ok && j < first_ent[b + 1]; ++j) {
				const ent &e = ents[j];
				uint64_t *dst = &idx->pos[idx->off[rank[j]]];
				if (e.single) *dst = e.val;
				else memcpy(dst, &bucket_pos[e.val >> 32], (size_t)(uint32_t)e.val * 8);
			}
		}
	}
	if (ok) { /* ri_idx_reader_read writes one part per 4 G bases/samples and ri_idx_load is called until EOF (rindex.c:795-830):
	           * a second part would silently be a missing piece of the reference, so it is refused by name */
		if (fseeko(f, 0, SEEK_END) == 0) {
			const off_t fsize = ftello(f);
			/* end of the part = end of the last bucket's (key, value) pairs */
			const off_t at = p_at[(1 << 14) - 1] + (off_t)p_n[(1 << 14) - 1] * 8 + 4 + (off_t)(first_ent[1 << 14] - first_ent[(1 << 14) - 1]) * 16;
			if (fsize > at) { fclose(f); delete idx; rh_set_error("%s: multi-part index files (references above 4 G bases or samples) are not supported: %lld bytes follow the first part", path, (long long)(fsize - at)); return NULL; }
		}
	}
	fclose(f);
	if (!ok) { delete idx; rh_set_error("%s: truncated or inconsistent index", path); return NULL; }
	if (pp) {
		pp->w = idx->w; pp->e = idx->e; pp->n = idx->n; pp->q = idx->q; pp->k = idx->k; pp->idx_flag = idx->flag;
		pp->diff = idx->diff; pp->fine_min = idx->fine_min; pp->fine_max = idx->fine_max; pp->fine_range = idx->fine_range;
	}
	return idx;
}

rh_index_s::~rh_index_s() { rh_index_dev_release(this); }
int rh_host_threads(void) { const unsigned n = std::thread::hardware_concurrency(); return n ? (int)n : 1; }

extern "C" void rh_index_destroy(rh_index_t *idx) { delete idx; }
extern "C" uint32_t rh_index_n_seq(const rh_index_t *idx) { return (uint32_t)idx->names.size(); }
extern "C" const char *rh_index_seq_name(const rh_index_t *idx, uint32_t i) { return idx->names[i].c_str(); }
extern "C" uint32_t rh_index_seq_len(const rh_index_t *idx, uint32_t i) { return idx->lens[i]; }
extern "C" uint64_t rh_index_n_keys(const rh_index_t *idx) { return idx->host_valid ? idx->keys.size() : idx->dev.n_keys; }
extern "C" uint64_t rh_index_n_pos(const rh_index_t *idx) { return idx->host_valid ? idx->pos.size() : idx->dev.n_pos; }
extern "C" uint32_t rh_index_key(const rh_index_t *idx, uint64_t i) { if (rh_index_sync_host(idx)) return 0u; return i < idx->keys.size() ? idx->keys[i] : 0u; }
extern "C" int rh_index_on_device(const rh_index_t *idx) { return idx->dev.device; }
extern "C" int rh_index_flat(const rh_index_t *idx, const uint32_t **keys, const uint64_t **off, const uint64_t **pos)
{ /* host arrays of the flattened index (downloaded first when the index lives on a device) */
	const int rc = rh_index_sync_host(idx);
	if (rc) return rc;
	*keys = idx->keys.data(); *off = idx->off.data(); *pos = idx->pos.data();
	return RH_OK;
}

extern "C" const uint64_t *rh_index_get(const rh_index_t *idx, uint32_t hash, int *n)
{
	if (rh_index_sync_host(idx)) { *n = 0; return NULL; }
	auto it = std::lower_bound(idx->keys.begin(), idx->keys.end(), hash);
	if (it == idx->keys.end() || *it != hash) { *n = 0; return NULL; }
	size_t i = it - idx->keys.begin();
	*n = (int)(idx->off[i + 1] - idx->off[i]);
	return &idx->pos[idx->off[i]];
}

extern "C" void rh_index_update_mapopt(const rh_index_t *idx, rh_params_t *p)
{ /* ri_mapopt_update + ri_idx_cal_max_occ, src/rindex.c:1018-1053 */
	if (p->mid_occ > 0) return;
	int32_t thres = INT32_MAX;
	const size_t n = idx->host_valid ? idx->keys.size() : (size_t)idx->dev.n_keys;
	if (p->mid_occ_frac > 0. && n > 0 && !idx->host_valid) { /* device-resident index: radix select over the CSR offsets in HBM */
		size_t kth = (uint32_t)((1. - p->mid_occ_frac) * n);
		if (kth >= n) kth = n - 1;
		uint32_t v = 0;
		if (rh_index_dev_kth_occ(idx, kth, &v) == RH_OK) thres = (int32_t)std::min<uint32_t>(v, 0x7ffffffeu) + 1;
	} else if (p->mid_occ_frac > 0. && n > 0) {
		std::vector<uint32_t> occ(n);
		for (size_t i = 0; i < n; ++i) occ[i] = (uint32_t)(idx->off[i + 1] - idx->off[i]);
		size_t kth = (uint32_t)((1. - p->mid_occ_frac) * n);
		if (kth >= n) kth = n - 1;
		std::nth_element(occ.begin(), occ.begin() + kth, occ.end());
		thres = (int32_t)occ[kth] + 1;
	}
	if (thres < p->min_mid_occ) thres = p->min_mid_occ;
	if (p->max_mid_occ > p->min_mid_occ && thres > p->max_mid_occ) thres = p->max_mid_occ;
	p->mid_occ = thres;
}

/* ---- sharding ------------------------------------------------------------------------------------ */
extern "C" void rh_split_by_samples(uint32_t n, const uint64_t *raw_len, uint32_t parts, uint32_t *cut)
{ /* range r ends after the first read at which the running sample count reaches r/parts of the total */
	if (!cut || parts == 0) return;
	for (uint32_t r = 0; r <= parts; ++r) cut[r] = n;
	cut[0] = 0;
	unsigned __int128 total = 0;
	for (uint32_t i = 0; i < n; ++i) total += raw_len[i];
	unsigned __int128 acc = 0; uint32_t r = 1;
	for (uint32_t i = 0; i < n && r < parts; ++i) {
		acc += raw_len[i];
		while (r < parts && acc * parts >= total * r) cut[r++] = i + 1;
	}
}

/* ---- PAF ---------------------------------------------------------------------------------------- */
extern "C" char *rh_format_paf(const rh_index_t *idx, const rh_map_rec_t *recs, uint64_t n_recs, const char *const *names)
{ /* line formats of src/rmap.cpp:751-764 (mapped) and 768-772 (unmapped); tags of 527-570 */
	std::string out;
	char line[2048], tags[256];
	for (uint64_t i = 0; i < n_recs; ++i) {
		const rh_map_rec_t &m = recs[i];
		if (m.mapped || m.nc >= 1)
			snprintf(tags, sizeof(tags), "mt:f:%.6f\tci:i:%d\tsl:i:%d\tcm:i:%d\tnc:i:%d\ts1:i:%d\tsm:f:%.2f", (double)m.mt_ms, (int)m.ci, (int)m.sl, m.cm, m.nc, m.s1, 0.0);
		else
			snprintf(tags, sizeof(tags), "mt:f:%.6f\tci:i:%d\tsl:i:%d\tcm:i:0\tnc:i:0\ts1:i:0\tsm:f:0", (double)m.mt_ms, (int)m.ci, (int)m.sl);
		if (m.mapped) {
			if (m.ref_id >= idx->names.size()) continue;
			snprintf(line, sizeof(line), "%s\t%u\t%u\t%u\t%c\t%s\t%u\t%u\t%u\t%u\t%u\t%u\t%s\n", names[m.read_idx], m.read_length,
			         m.read_start_position, m.read_end_position, m.rev ? '-' : '+', idx->names[m.ref_id].c_str(), idx->lens[m.ref_id],
			         m.fragment_start_position, m.fragment_start_position + m.fragment_length,
			         m.read_end_position - m.read_start_position - 1, m.fragment_length, (unsigned)m.mapq, tags);
		} else {
			snprintf(line, sizeof(line), "%s\t%u\t*\t*\t*\t*\t*\t*\t*\t*\t*\t%u\t%s\n", names[m.read_idx], m.read_length, (unsigned)m.mapq, tags);
		}
		out += line;
	}
	char *ret = (char *)malloc(out.size() + 1);
	memcpy(ret, out.c_str(), out.size() + 1);
	return ret;
}
