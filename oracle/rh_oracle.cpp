/*
 * rh_oracle.cpp — TEST INFRASTRUCTURE ONLY (the parity oracle).
 *
 * A scalar CPU restatement of RawHash2's per-read mapping hot path, written from the
 * reference's behaviour (each function cites the reference file:line it follows; paths are
 * relative to /root/reference).  It exists to CHECK the CUDA path.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it;
 * the product (rawhash_b200/) never links, imports or calls anything in oracle/.
 *
 * PARITY PINNING: the reference ships no golden vectors.  This oracle is
 * pinned against the reference itself compiled in the build container
 * (oracle/_ref/libref_tap.so, see oracle/Makefile) by tests/test_oracle_vs_ref.py, and
 * against the fixtures that run generated under tests/golden/ (tests/golden/make_golden.py).
 *
 * Floating point: the reference is built with -O3 and FMA contraction on (upstream
 * -march=native; here x86-64-v3).  The contractions GCC 13.3 actually emits were read from
 * the object code and are written out below with explicit fma()/fmaf(); this file is
 * compiled with -ffp-contract=off so nothing else fuses.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#include <limits.h>
#include <string>
#include <vector>
#include <algorithm>
#include <thread>
#include <atomic>
#include <zlib.h>

#include "../include/rawhash_b200.h"

namespace {

struct xy_t { uint64_t x, y; };

/* ------------------------------------------------------------------------------------------
 * options (src/roptions.c:4-138, src/main.cpp:111-210,363-376)
 * ---------------------------------------------------------------------------------------- */
void params_defaults(rh_params_t *p)
{
	memset(p, 0, sizeof(*p));
	p->e = 8; p->w = 0; p->q = 4; p->n = 0; p->k = 6; p->lev_col = 1;
	p->diff = 0.35f; p->fine_min = -2.0f; p->fine_max = 2.0f; p->fine_range = 0.4;
	p->window_length1 = 3; p->window_length2 = 9; p->threshold1 = 4.0f; p->threshold2 = 3.5f; p->peak_height = 0.4f;
	p->bp_per_sec = 450; p->sample_rate = 4000; p->chunk_size = 4000;
	p->sample_per_base = (float)p->sample_rate / p->bp_per_sec;
	p->mid_occ_frac = 1e-2f; p->min_mid_occ = 50; p->max_mid_occ = 500000; p->mid_occ = 0;
	p->min_events = 50; p->bw = 500; p->max_target_gap_length = 2500; p->max_query_gap_length = 2500;
	p->max_chain_iter = 200; p->max_num_skips = 5; p->min_num_anchors = 2;
	p->min_chaining_score = 15; p->min_chaining_score2 = 0;
	p->chain_gap_scale = 0.8f; p->chain_skip_scale = 0.0f;
	p->mask_level = 0.5f; p->mask_len = INT_MAX; p->pri_ratio = 0.3f; p->best_n = 0; p->alt_drop = 0.15f;
	p->w_bestmq = 0.05f; p->w_bestmc = 0.6f; p->w_bestq = 0.35f; p->w_threshold = 0.45f;
	p->max_num_chunk = 10; p->min_mapq = 2; p->map_flag = 0;
}

int params_preset(rh_params_t *p, const char *name)
{
	std::string s(name ? name : "");
	auto ava_common = [&]() {
		p->diff = 0.45f; p->bw = 1000; p->max_target_gap_length = 2500; p->max_query_gap_length = 2500;
		p->min_num_anchors = 5; p->min_mapq = 5;
		p->idx_flag |= RH_I_SIG_TARGET; p->map_flag |= RH_M_ALL_CHAINS | RH_M_NO_ADAPTIVE; p->pri_ratio = 0.0f;
	};
	if (s.empty() || s == "sensitive" || s == "sequence-until") return 0;
	if (s == "viral") {
		p->e = 6; p->bw = 100; p->max_target_gap_length = 500; p->max_query_gap_length = 500;
		p->max_num_chunk = 5; p->min_chaining_score = 10; p->chain_gap_scale = 1.2f; p->chain_skip_scale = 0.3f;
	} else if (s == "fast") {
		p->fine_range = 0.6; p->min_mapq = 5; p->min_chaining_score = 10; p->chain_gap_scale = 0.6f;
	} else if (s == "faster") {
		p->e = 11; p->w = 3; p->fine_range = 0.6; p->max_num_chunk = 5; p->min_mapq = 5; p->min_chaining_score = 10; p->chain_gap_scale = 0.6f;
	} else if (s == "ava-viral") {
		ava_common(); p->e = 6; p->chain_gap_scale = 1.2f; p->chain_skip_scale = 0.3f; p->w = 0;
		p->min_chaining_score = 20; p->min_chaining_score2 = 30;
	} else if (s == "ava") {
		ava_common(); p->w = 3; p->min_chaining_score = 40; p->min_chaining_score2 = 75; p->bw = 5000;
	} else if (s == "ava-sensitive") {
		ava_common(); p->w = 0; p->min_chaining_score = 75; p->min_chaining_score2 = 100;
	} else if (s == "ava-large") {
		ava_common(); p->fine_range = 0.6; p->chain_gap_scale = 0.6f; p->w = 5;
		p->min_chaining_score = 20; p->min_chaining_score2 = 50; p->min_num_anchors = 2; p->min_mapq = 2; p->bw = 5000;
	} else return -1;
	return 0;
}

void params_r10(rh_params_t *p)
{
	p->k = 9; p->window_length1 = 3; p->window_length2 = 6; p->threshold1 = 6.5f; p->threshold2 = 4.0f;
	p->peak_height = 0.2f; p->chain_gap_scale = 1.2f;
}

/* ------------------------------------------------------------------------------------------
 * pore model (load_pore, src/rutils.c:133-178): column lev_col of every non-header line,
 * z-normalised with population mean/std; the variance is fma(-mean,mean,sum2/n) as compiled.
 * ---------------------------------------------------------------------------------------- */
bool pore_load(const char *path, int k, int lev_col, std::vector<float> &vals)
{
	FILE *fp = fopen(path, "r");
	if (!fp) return false;
	vals.assign((size_t)1 << (2 * k), 0.0f);
	char line[1024];
	size_t i = 0;
	double sum = 0, sum2 = 0;
	while (fgets(line, sizeof(line), fp)) {
		if (!strncmp(line, "kmer", 4)) continue;
		char *rest = line, *tok; int col = 0;
		while ((tok = strsep(&rest, "\t")) != NULL) {
			if (col++ == lev_col) {
				float v;
				if (sscanf(tok, "%f", &v) != 1 || i >= vals.size()) { fclose(fp); vals.clear(); return false; }
				vals[i] = v; sum += v; sum2 += v * v; /* float product, double accumulation */
				break;
			}
		}
		++i;
	}
	fclose(fp);
	double mean = sum / (int)i;
	double sd = sqrt(fma(-mean, mean, sum2 / (int)i));
	for (size_t j = 0; j < i && j < vals.size(); ++j) vals[j] = (vals[j] - mean) / sd;
	return true;
}

/* ------------------------------------------------------------------------------------------
 * sketching (src/rsketch.c)
 * ---------------------------------------------------------------------------------------- */
inline uint64_t mix64_masked(uint64_t key, uint64_t mask) /* rsketch.c:7-16 */
{
	key = (~key + (key << 21)) & mask;
	key ^= key >> 24;
	key = (key + (key << 3) + (key << 8)) & mask;
	key ^= key >> 14;
	key = (key + (key << 2) + (key << 4)) & mask;
	key ^= key >> 28;
	key = (key + (key << 31)) & mask;
	return key;
}

uint32_t quantize(float v, float fine_min, float fine_max, float fine_range, uint32_t n_buckets) /* rsketch.c:18-53 */
{
	const float lo = -3.0f, hi = 3.0f, span = hi - lo;
	float c1 = (1 - fine_range) / 2, c2 = fine_range + c1;
	float norm = (v - lo) / span;
	float a = (fine_min - lo) / span, b = (fine_max - lo) / span;
	float qv;
	if (v >= fine_min && v <= fine_max) qv = fine_range * ((norm - a) / (b - a));
	else if (norm < 0.5) qv = fine_range + c1 * norm; /* two roundings: no FMA in rsketch.o */
	else qv = c2 + c1 * norm;
	return (uint32_t)(qv * (n_buckets - 1));
}

/* w == 0: every kept event closes a window of e events (ri_sketch_reg, rsketch.c:143-204) */
void sketch_all(const rh_params_t &P, const float *ev, uint32_t len, uint32_t id, int strand, std::vector<xy_t> &out)
{
	const int e = P.e, q = P.q;
	const uint32_t span = P.k + e - 1;
	const uint64_t id_bits = (uint64_t)id << 32, m32 = (1ULL << 32) - 1;
	const uint64_t m_ev = (q * e >= 64) ? ~0ULL : (1ULL << (q * e)) - 1, m_q = (1ULL << q) - 1;
	std::vector<uint64_t> ring_y(e, 0);
	uint64_t packed = 0; uint32_t kept = 0, last = 0; int slot = 0;
	for (uint32_t i = 0; i < len; ++i) {
		if (i > 0 && fabsf(ev[i] - ev[last]) < P.diff) continue;
		last = i;
		uint64_t code = quantize(ev[i], P.fine_min, P.fine_max, P.fine_range, 1u << q) & m_q;
		ring_y[slot] = id_bits | (uint64_t)((uint32_t)i << 1) | (uint64_t)strand;
		slot = (slot + 1 == e) ? 0 : slot + 1;
		packed = (i == 0) ? (code & m_ev) : (((packed << q) | code) & m_ev);
		++kept;
		if (kept >= (uint32_t)e) /* ring slot now holds the y of the oldest event of the window */
			out.push_back({mix64_masked(packed, m32) << 6 | span, ring_y[slot]});
	}
}

/* w > 0: minimizer over w consecutive windows, minimap2 mm_sketch semantics (ri_sketch_min, rsketch.c:55-141) */
void sketch_min(const rh_params_t &P, const float *ev, uint32_t len, uint32_t id, int strand, std::vector<xy_t> &out)
{
	const int e = P.e, q = P.q, w = P.w;
	const uint32_t span = P.k + e - 1;
	const uint64_t id_bits = (uint64_t)id << 32, m32 = (1ULL << 32) - 1;
	const uint64_t m_ev = (1ULL << (q * e)) - 1, m_q = (1ULL << q) - 1;
	std::vector<xy_t> win(w, xy_t{UINT64_MAX, UINT64_MAX});
	std::vector<xy_t> ring(e, xy_t{0, 0});
	xy_t mn = {UINT64_MAX, UINT64_MAX};
	uint64_t packed = 0; uint32_t l = 0, last = 0; int slot = 0, full = 0, wp = 0, mp = 0;
	for (uint32_t i = 0; i < len; ++i) {
		if (i > 0 && fabsf(ev[i] - ev[last]) < P.diff) continue;
		++l; last = i;
		uint64_t code = quantize(ev[i], P.fine_min, P.fine_max, P.fine_range, 1u << q) & m_q;
		packed = ((packed << q) | code) & m_ev;
		ring[slot].y = id_bits | (uint64_t)((uint32_t)i << 1) | (uint64_t)strand;
		if (++slot == e) { full = 1; slot = 0; }
		ring[slot].x = mix64_masked(packed, m32) << 6 | span;
		if (!full) continue;
		xy_t cur = ring[slot];
		win[wp] = cur;
		if (l == (uint32_t)(w + e - 1) && mn.x != UINT64_MAX) { /* first full window: flush equal minima */
			for (int j = wp + 1; j < w; ++j) if (mn.x == win[j].x && win[j].y != mn.y) out.push_back(win[j]);
			for (int j = 0; j < wp; ++j) if (mn.x == win[j].x && win[j].y != mn.y) out.push_back(win[j]);
		}
		if (cur.x <= mn.x) {
			if (l >= (uint32_t)(w + e) && mn.x != UINT64_MAX) out.push_back(mn);
			mn = cur; mp = wp;
		} else if (wp == mp) { /* the minimum slid out of the window */
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

void sketch(const rh_params_t &P, const float *ev, uint32_t len, uint32_t id, int strand, std::vector<xy_t> &out) /* rsketch.c:271-290 */
{
	if (len == 0) return;
	if (P.w) sketch_min(P, ev, len, id, strand, out);
	else sketch_all(P, ev, len, id, strand, out);
}

/* ------------------------------------------------------------------------------------------
 * klib radix sort (src/ksort.h:98-151): in-place MSD byte sort ("American flag"), buckets
 * larger than 64 recurse on the next byte, the rest get a stable insertion sort.  The tie
 * order it leaves is observable downstream, so it is restated exactly: the
 * permutation cycles below visit slots in the same order as the reference's pointer walk.
 * ---------------------------------------------------------------------------------------- */
template <class T, class KeyF>
void insertion_by_key(T *b, T *e, KeyF key)
{
	for (T *i = b + 1; i < e; ++i) {
		if (key(*i) < key(*(i - 1))) {
			T tmp = *i, *j = i;
			while (j > b && key(tmp) < key(*(j - 1))) { *j = *(j - 1); --j; }
			*j = tmp;
		}
	}
}

template <class T, class KeyF>
void flag_sort_level(T *beg, T *end, int shift, KeyF key)
{
	size_t head[256], tail[256], cnt[256];
	memset(cnt, 0, sizeof(cnt));
	for (T *p = beg; p != end; ++p) ++cnt[(key(*p) >> shift) & 255];
	size_t acc = 0;
	for (int b = 0; b < 256; ++b) { head[b] = acc; acc += cnt[b]; tail[b] = acc; }
	for (int b = 0; b < 256;) {
		if (head[b] == tail[b]) { ++b; continue; }
		int d = (key(beg[head[b]]) >> shift) & 255;
		if (d == b) { ++head[b]; continue; }
		T carry = beg[head[b]];
		do { /* drop carry at the front of its bucket, pick up what was there */
			std::swap(carry, beg[head[d]]); ++head[d];
			d = (key(carry) >> shift) & 255;
		} while (d != b);
		beg[head[b]++] = carry;
	}
	if (shift == 0) return;
	int next = shift > 8 ? shift - 8 : 0;
	acc = 0;
	for (int b = 0; b < 256; ++b) {
		size_t n = cnt[b];
		if (n > 64) flag_sort_level(beg + acc, beg + acc + n, next, key);
		else if (n > 1) insertion_by_key(beg + acc, beg + acc + n, key);
		acc += n;
	}
}

template <class T, class KeyF>
void klib_radix_sort(T *beg, T *end, KeyF key)
{
	if (end - beg <= 64) insertion_by_key(beg, end, key);
	else flag_sort_level(beg, end, 56, key);
}

inline void sort_xy_by_x(xy_t *b, xy_t *e) { klib_radix_sort(b, e, [](const xy_t &a) { return a.x; }); }
inline void sort_u64(uint64_t *b, uint64_t *e) { klib_radix_sort(b, e, [](const uint64_t &a) { return a; }); }

/* ------------------------------------------------------------------------------------------
 * index: key = 32-bit hash, value = ascending list of (id<<32 | pos<<1 | strand)
 * (what ri_idx_get returns, src/rindex.c:497-514; built like src/rindex.c:100-192,311-363).
 * Layout here is a plain sorted array + offsets; only the key->list mapping is reference
 * behaviour.
 * ---------------------------------------------------------------------------------------- */
struct Index {
	std::vector<uint32_t> keys;      /* sorted distinct hashes */
	std::vector<uint64_t> off;       /* keys.size()+1 */
	std::vector<uint64_t> pos;
	std::vector<std::string> names;
	std::vector<uint32_t> lens;
	int flag = 0;
	void build(std::vector<xy_t> &seeds)
	{
		std::sort(seeds.begin(), seeds.end(), [](const xy_t &a, const xy_t &b) {
			uint64_t ha = a.x >> 6, hb = b.x >> 6; return ha != hb ? ha < hb : a.y < b.y; });
		keys.clear(); off.clear(); pos.resize(seeds.size());
		for (size_t i = 0; i < seeds.size(); ++i) {
			uint32_t h = (uint32_t)(seeds[i].x >> 6);
			if (i == 0 || h != keys.back()) { keys.push_back(h); off.push_back(i); }
			pos[i] = seeds[i].y;
		}
		off.push_back(seeds.size());
	}
	const uint64_t *get(uint32_t h, int *n) const
	{
		auto it = std::lower_bound(keys.begin(), keys.end(), h);
		if (it == keys.end() || *it != h) { *n = 0; return 0; }
		size_t i = it - keys.begin();
		*n = (int)(off[i + 1] - off[i]);
		return &pos[off[i]];
	}
	/* ri_idx_cal_max_occ (src/rindex.c:1018-1039): (k-th smallest occupancy)+1, k=(uint32)((1-f)*n) */
	int32_t max_occ(float f) const
	{
		if (f <= 0.) return INT32_MAX;
		size_t n = keys.size();
		if (n == 0) return 1;
		std::vector<uint32_t> a(n);
		for (size_t i = 0; i < n; ++i) a[i] = (uint32_t)(off[i + 1] - off[i]);
		size_t kth = (uint32_t)((1. - f) * n);
		if (kth >= n) kth = n - 1;
		std::nth_element(a.begin(), a.begin() + kth, a.end());
		return a[kth] + 1;
	}
};

const unsigned char *nt4()
{
	static unsigned char t[256]; static bool init = false;
	if (!init) { memset(t, 4, 256); t['A'] = t['a'] = 0; t['C'] = t['c'] = 1; t['G'] = t['g'] = 2; t['T'] = t['t'] = 3; init = true; }
	return t;
}

/* expected-signal of one strand (ri_seq_to_sig, src/rsig.c:13-40): a value is emitted for every
 * position >= k-1; an ambiguous base leaves the rolling k-mer untouched. */
void seq_to_sig(const char *s, int len, const float *pore, int k, int strand, std::vector<float> &out)
{
	const unsigned char *T = nt4();
	uint64_t mask = (1ULL << 2 * k) - 1, kmer = 0;
	out.clear();
	for (int i = 0; i < len; ++i) {
		int pos = strand ? len - i - 1 : i;
		int c = T[(uint8_t)s[pos]];
		if (c < 4) kmer = strand ? (((kmer << 2) | (3ULL ^ c)) & mask) : (((kmer << 2) | c) & mask);
		if (i + 1 < k) continue;
		out.push_back(pore[kmer]);
	}
}

bool read_fasta(const char *path, std::vector<std::string> &names, std::vector<std::string> &seqs)
{
	gzFile f = gzopen(path, "r");
	if (!f) return false;
	std::vector<char> buf(1 << 16);
	std::string cur; bool have = false;
	while (gzgets(f, buf.data(), (int)buf.size())) {
		size_t L = strlen(buf.data());
		bool eol = L && buf[L - 1] == '\n';
		while (L && (buf[L - 1] == '\n' || buf[L - 1] == '\r')) buf[--L] = 0;
		if (buf[0] == '>' && !have) { /* header start */
			std::string h(buf.data() + 1);
			size_t sp = h.find_first_of(" \t");
			names.push_back(sp == std::string::npos ? h : h.substr(0, sp));
			seqs.emplace_back();
			have = !eol; /* header longer than the buffer: skip continuation */
			continue;
		}
		if (have) { have = !eol; continue; }
		if (!seqs.empty()) seqs.back().append(buf.data(), L);
	}
	gzclose(f);
	return true;
}

/* ------------------------------------------------------------------------------------------
 * events (src/revent.c)
 * ---------------------------------------------------------------------------------------- */
struct NormState { double sum = 0, sum2 = 0; uint32_t n = 0; };

/* normalize_signal, revent.c:221-255 */
void znorm(const float *sig, uint32_t len, NormState &st, std::vector<float> &z)
{
	double s = st.sum, s2 = st.sum2;
	for (uint32_t i = 0; i < len; ++i) { s += sig[i]; s2 += sig[i] * sig[i]; }
	st.n += len; st.sum = s; st.sum2 = s2;
	double mean = s / st.n;
	double sd = sqrt(fma(-mean, mean, s2 / st.n));
	z.clear();
	for (uint32_t i = 0; i < len; ++i) {
		float v = (sig[i] - mean) / sd;
		if (v < 3 && v > -3) z.push_back(v);
	}
}

/* comp_tstat, revent.c:38-74 (with the contractions of the compiled object) */
void tstat(const std::vector<float> &ps, const std::vector<float> &pq, uint32_t n, uint32_t w, std::vector<float> &t)
{
	t.assign(n + 1, 0.0f);
	if (n < 2 * w || w < 2) return;
	const float fw = (float)w;
	for (uint32_t i = w; i <= n - w; ++i) {
		float s1 = ps[i], q1 = pq[i];
		if (i > w) { s1 -= ps[i - w]; q1 -= pq[i - w]; }
		float s2 = ps[i + w] - ps[i], q2 = pq[i + w] - pq[i];
		float m1 = s1 / fw, m2 = s2 / fw;
		float acc = fmaf(-m1, m1, q1 / fw);
		acc = acc + q2 / fw;
		acc = fmaf(-m2, m2, acc);
		float var = fmaxf(acc / fw, FLT_MIN);
		t[i] = fabsf(m2 - m1) / sqrtf(var);
	}
}

struct Detector { const float *t; float thr; uint32_t win; uint32_t masked_to = 0; int peak_pos = -1; float peak_val = FLT_MAX; int valid = 0; };

/* gen_peaks, revent.c:91-150: short detector first, then long, at every i */
uint32_t peaks_of(Detector *d, uint32_t n, float height, std::vector<uint32_t> &peaks)
{
	peaks.clear();
	for (uint32_t i = 0; i < n; ++i) {
		for (int k = 0; k < 2; ++k) {
			Detector &D = d[k];
			if (D.masked_to >= i) continue;
			float cur = D.t[i];
			if (D.peak_pos == -1) {
				if (cur < D.peak_val) D.peak_val = cur;
				else if (cur - D.peak_val > height) { D.peak_val = cur; D.peak_pos = (int)i; }
			} else {
				if (cur > D.peak_val) { D.peak_val = cur; D.peak_pos = (int)i; }
				if (D.peak_val > D.thr) {
					for (int m = k + 1; m < 2; ++m) {
						d[m].masked_to = D.peak_pos + d[0].win;
						d[m].peak_pos = -1; d[m].peak_val = FLT_MAX; d[m].valid = 0;
					}
				}
				if (D.peak_val - cur > height && D.peak_val > D.thr) D.valid = 1;
				if (D.valid && (i - D.peak_pos) > D.win / 2) {
					peaks.push_back((uint32_t)D.peak_pos);
					D.peak_pos = -1; D.peak_val = cur; D.valid = 0;
				}
			}
		}
	}
	return (uint32_t)peaks.size();
}

/* calculate_mean_of_filtered_segment, revent.c:158-180 (sorts the segment in place) */
float seg_mean(float *seg, uint32_t len)
{
	std::sort(seg, seg + len);
	float q1 = seg[len / 4], q3 = seg[3 * len / 4], iqr = q3 - q1;
	float lo = q1 - iqr, hi = q3 + iqr, sum = 0.0f; uint32_t c = 0;
	for (uint32_t i = 0; i < len; ++i) if (seg[i] >= lo && seg[i] <= hi) { sum += seg[i]; ++c; }
	return c > 0 ? sum / c : 0;
}

/* detect_events, revent.c:257-316.  Returns events; n_sig_out = samples surviving |z|<3. */
void events_of(const rh_params_t &P, const float *sig, uint32_t len, NormState &st, std::vector<float> &ev, uint32_t *n_sig_out)
{
	std::vector<float> z;
	ev.clear();
	znorm(sig, len, st, z);
	uint32_t n = (uint32_t)z.size();
	if (n_sig_out) *n_sig_out = n;
	if (n == 0) return;
	std::vector<float> ps(n + 1), pq(n + 1);
	ps[0] = pq[0] = 0.0f;
	for (uint32_t i = 0; i < n; ++i) { ps[i + 1] = ps[i] + z[i]; pq[i + 1] = fmaf(z[i], z[i], pq[i]); }
	std::vector<float> t1, t2;
	tstat(ps, pq, n, P.window_length1, t1);
	tstat(ps, pq, n, P.window_length2, t2);
	Detector d[2];
	d[0].t = t1.data(); d[0].thr = P.threshold1; d[0].win = P.window_length1;
	d[1].t = t2.data(); d[1].thr = P.threshold2; d[1].win = P.window_length2;
	std::vector<uint32_t> peaks;
	if (peaks_of(d, n, P.peak_height, peaks) == 0) return;
	uint32_t start = 0;
	for (uint32_t p : peaks) { /* gen_events, revent.c:193-219 */
		if (!(p > 0 && p < n)) continue;
		/* the reference assumes increasing peaks (it would index out of bounds otherwise) */
		uint32_t seglen = p > start ? p - start : 0;
		ev.push_back(seglen ? seg_mean(z.data() + start, seglen) : 0.0f);
		start = p;
	}
}

/* ------------------------------------------------------------------------------------------
 * seeding (src/rseed.c:60-154, src/rmap.cpp:51-126)
 * ---------------------------------------------------------------------------------------- */
struct Carry { std::vector<xy_t> prev; uint32_t offset = 0; };

void seed_hits(const rh_params_t &P, const Index &idx, const char *qname, const std::vector<xy_t> &seeds, Carry &cy,
               std::vector<xy_t> &anchors, int *rep_len_out)
{
	struct M { const uint64_t *cr; int n; uint32_t q_pos, q_span, seg; bool tandem; };
	std::vector<M> ms;
	size_t ns = seeds.size();
	for (size_t i = 0; i < ns; ++i) {
		int n; const uint64_t *cr = idx.get((uint32_t)(seeds[i].x >> 6), &n);
		if (n == 0) continue;
		M m{cr, n, (uint32_t)seeds[i].y, (uint32_t)(seeds[i].x & 63), (uint32_t)(seeds[i].y >> 32), false};
		if (i > 0 && seeds[i].x >> 6 == seeds[i - 1].x >> 6) m.tandem = true;
		if (i + 1 < ns && seeds[i].x >> 6 == seeds[i + 1].x >> 6) m.tandem = true;
		ms.push_back(m);
	}
	int rep_st = 0, rep_en = 0, rep_len = 0;
	anchors.clear();
	const bool ava = (P.map_flag & RH_M_ALL_CHAINS) != 0;
	const uint64_t keep_id = ((((1ULL << 32) - 1) << 32) >> 1); /* bits 31..62, rmap.cpp:70 */
	for (const M &m : ms) {
		if (m.n > P.mid_occ) { /* occurrence filter + repeat length, rseed.c:130-151 */
			int st = (int)(m.q_pos >> 1) + 1, en = st + (int)m.q_span + 1;
			if (st > rep_en) { rep_len += rep_en - rep_st; rep_st = st; rep_en = en; } else rep_en = en;
			continue;
		}
		for (int k = 0; k < m.n; ++k) {
			uint64_t h = m.cr[k];
			if (ava && strcmp(qname, idx.names[h >> 32].c_str()) >= 0) continue; /* rmap.cpp:86 */
			xy_t a;
			a.x = (h & keep_id) | (uint32_t)((h >> 1) & 0x7fffffffu);
			if (h & 1) a.x |= 1ULL << 63;
			a.y = (uint64_t)m.seg << 40 | (uint64_t)m.q_span << 32 | (uint32_t)((m.q_pos >> 1) + cy.offset);
			if (m.tandem) a.y |= 1ULL << 38;
			anchors.push_back(a);
		}
	}
	rep_len += rep_en - rep_st;
	*rep_len_out = rep_len;
	anchors.insert(anchors.end(), cy.prev.begin(), cy.prev.end()); /* rmap.cpp:111-116 */
	cy.prev.clear();
	if (!anchors.empty()) sort_xy_by_x(anchors.data(), anchors.data() + anchors.size());
}

/* ------------------------------------------------------------------------------------------
 * chaining (src/lchain.c)
 * ---------------------------------------------------------------------------------------- */
inline float approx_log2(float x) /* mg_log2, lchain.c:23-31 */
{
	union { float f; uint32_t i; } z = {x};
	float r = (float)(int)(((z.i >> 23) & 255) - 128);
	z.i &= ~(255u << 23); z.i += 127u << 23;
	r += fmaf(fmaf(-0.34484843f, z.f, 2.02466578f), z.f, -0.67487759f);
	return r;
}

inline int32_t pair_score(const xy_t &ai, const xy_t &aj, int32_t max_t, int32_t max_q, int32_t bw, float pen_gap, float pen_skip)
{ /* compute_score, lchain.c:297-356 */
	int32_t dq = (int32_t)ai.y - (int32_t)aj.y;
	if (dq <= 0 || dq > max_q) return INT32_MIN;
	int32_t dr = (int32_t)(ai.x - aj.x);
	if (dr == 0 || dr > max_t) return INT32_MIN;
	int32_t dd = dr > dq ? dr - dq : dq - dr;
	if (dd > bw || dr > max_q) return INT32_MIN;
	int32_t dg = dr < dq ? dr : dq;
	int32_t qs = (int32_t)((aj.y >> 32) & 63);
	int32_t sc = qs < dg ? qs : dg;
	if (dd || dg > qs) {
		float lin = fmaf(pen_gap, (float)dd, pen_skip * (float)dg);
		float lg = dd >= 1 ? approx_log2((float)(dd + 1)) : 0.0f;
		sc -= (int)(lin + .5f * lg);
	}
	return sc;
}

struct ChainOut { std::vector<uint64_t> u; std::vector<xy_t> a; };

/* mg_lchain_dp + mg_chain_backtrack + compact_a (lchain.c:385-530, 95-194, 214-281).
 * `a` is consumed; `prev` receives the backtrack-order copy (next chunk's prev_anchors). */
void chain(const rh_params_t &P, float pen_gap, float pen_skip, std::vector<xy_t> &a, std::vector<xy_t> &prev, ChainOut &out)
{
	out.u.clear(); out.a.clear(); prev.clear();
	const int64_t n = (int64_t)a.size();
	if (n == 0) return;
	int32_t max_t = P.max_target_gap_length, max_q = P.max_query_gap_length, bw = P.bw;
	const int32_t max_skip = P.max_num_skips, max_iter = P.max_chain_iter, min_cnt = P.min_num_anchors, min_sc = P.min_chaining_score;
	const int32_t max_drop = bw;
	if (max_t < bw) max_t = bw;
	if (max_q < bw) max_q = bw;
	std::vector<int32_t> f(n), t(n, 0), v(n);
	std::vector<int64_t> p(n);
	int64_t st = 0, best_in_band = -1;
	for (int64_t i = 0; i < n; ++i) {
		int64_t best_j = -1, j;
		int32_t best = (int32_t)((a[i].y >> 32) & 63), skipped = 0;
		while (st < i && (a[i].x >> 32 != a[st].x >> 32 || a[i].x > a[st].x + max_t)) ++st;
		if (i - st > max_iter) st = i - max_iter;
		for (j = i - 1; j >= st; --j) {
			int32_t sc = pair_score(a[i], a[j], max_t, max_q, bw, pen_gap, pen_skip);
			if (sc == INT32_MIN) continue;
			sc += f[j];
			if (sc > best) { best = sc; best_j = j; if (skipped > 0) --skipped; }
			else if (t[j] == (int32_t)i) { if (++skipped > max_skip) break; }
			if (p[j] >= 0) t[p[j]] = (int32_t)i;
		}
		int64_t end_j = j;
		if (best_in_band < 0 || a[i].x - a[best_in_band].x > (uint64_t)(int64_t)max_t) {
			int32_t mx = INT32_MIN; best_in_band = -1;
			for (j = i - 1; j >= st; --j) if (mx < f[j]) { mx = f[j]; best_in_band = j; }
		}
		if (best_in_band >= 0 && best_in_band < end_j) {
			int32_t sc = pair_score(a[i], a[best_in_band], max_t, max_q, bw, pen_gap, pen_skip);
			if (sc != INT32_MIN && best < sc + f[best_in_band]) { best = sc + f[best_in_band]; best_j = best_in_band; }
		}
		f[i] = best; p[i] = best_j;
		v[i] = (best_j >= 0 && v[best_j] > best) ? v[best_j] : best;
		if (best_in_band < 0 || (a[i].x - a[best_in_band].x <= (uint64_t)(int64_t)max_t && f[best_in_band] < f[i])) best_in_band = i;
	}
	/* backtrack */
	std::vector<xy_t> z;
	for (int64_t i = 0; i < n; ++i) if (f[i] >= min_sc) z.push_back({(uint64_t)(int64_t)f[i], (uint64_t)i});
	if (z.empty()) return;
	sort_xy_by_x(z.data(), z.data() + z.size());
	std::fill(t.begin(), t.end(), 0);
	std::vector<int32_t> order; /* anchor indices, chain after chain, end -> start */
	std::vector<uint64_t> u;
	for (int64_t k = (int64_t)z.size() - 1; k >= 0; --k) {
		int64_t i = (int64_t)z[k].y;
		if (t[i] != 0) continue;
		/* mg_chain_bk_end, lchain.c:47-75 */
		int64_t end_i = -1, max_i = i, c = i; int32_t max_s = 0;
		do {
			t[c] = 2;
			end_i = c = p[c];
			int32_t s = c < 0 ? (int32_t)z[k].x : (int32_t)z[k].x - f[c];
			if (s > max_s) { max_s = s; max_i = c; }
			else if (max_s - s > max_drop) break;
		} while (c >= 0 && t[c] == 0);
		for (c = i; c >= 0 && c != end_i; c = p[c]) t[c] = 0;
		size_t n0 = order.size();
		for (c = i; c != max_i; c = p[c]) { order.push_back((int32_t)c); t[c] = 1; }
		int32_t sc = c < 0 ? (int32_t)z[k].x : (int32_t)z[k].x - f[c];
		size_t cnt = order.size() - n0;
		if (sc >= min_sc && cnt > 0 && (int64_t)cnt >= min_cnt) u.push_back((uint64_t)sc << 32 | cnt);
		else order.resize(n0);
	}
	if (u.empty()) return;
	/* compact_a */
	size_t n_v = order.size();
	std::vector<xy_t> b(n_v);
	size_t k = 0;
	for (size_t ci = 0; ci < u.size(); ++ci) {
		size_t ni = (uint32_t)u[ci], k0 = k;
		for (size_t j = 0; j < ni; ++j) b[k++] = a[order[k0 + (ni - j - 1)]];
	}
	prev = b;
	std::vector<xy_t> w(u.size());
	k = 0;
	for (size_t ci = 0; ci < u.size(); ++ci) { w[ci].x = b[k].x; w[ci].y = (uint64_t)k << 32 | ci; k += (uint32_t)u[ci]; }
	sort_xy_by_x(w.data(), w.data() + w.size());
	out.u.resize(u.size()); out.a.resize(n_v);
	k = 0;
	for (size_t ci = 0; ci < u.size(); ++ci) {
		size_t j = (uint32_t)w[ci].y, cnt = (uint32_t)u[j];
		out.u[ci] = u[j];
		memcpy(&out.a[k], &b[w[ci].y >> 32], cnt * sizeof(xy_t));
		k += cnt;
	}
}

/* ------------------------------------------------------------------------------------------
 * regions (src/hit.c)
 * ---------------------------------------------------------------------------------------- */
struct Reg {
	int32_t id, cnt, rid, score, qs, qe, rs, re, parent, subsc, as, n_sub, score0;
	uint32_t mapq, rev, hash;
};

inline uint64_t mix64(uint64_t key) /* hit.c:73-83 */
{
	key = ~key + (key << 21); key ^= key >> 24;
	key = key + (key << 3) + (key << 8); key ^= key >> 14;
	key = key + (key << 2) + (key << 4); key ^= key >> 28;
	key = key + (key << 31);
	return key;
}

inline uint32_t wang32(uint32_t key) /* __ac_Wang_hash, khash.h:400-409 */
{
	key += ~(key << 15); key ^= (key >> 10); key += (key << 3);
	key ^= (key >> 6); key += ~(key << 11); key ^= (key >> 16);
	return key;
}

void gen_regs(uint32_t hash, const std::vector<uint64_t> &u, const std::vector<xy_t> &a, std::vector<Reg> &r) /* hit.c:100-150 */
{
	size_t n_u = u.size();
	r.clear();
	if (n_u == 0) return;
	std::vector<xy_t> z(n_u);
	size_t k = 0;
	for (size_t i = 0; i < n_u; ++i) {
		uint32_t h = (uint32_t)mix64((mix64(a[k].x) + mix64(a[k].y)) ^ hash);
		z[i].x = u[i] ^ h;
		z[i].y = (uint64_t)k << 32 | (uint32_t)u[i];
		k += (uint32_t)u[i];
	}
	sort_xy_by_x(z.data(), z.data() + n_u);
	std::reverse(z.begin(), z.end());
	r.resize(n_u);
	for (size_t i = 0; i < n_u; ++i) {
		Reg &g = r[i];
		memset(&g, 0, sizeof(g));
		g.id = (int32_t)i; g.parent = -1;
		g.score = g.score0 = (int32_t)(z[i].x >> 32);
		g.hash = (uint32_t)z[i].x;
		g.cnt = (int32_t)z[i].y; g.as = (int32_t)(z[i].y >> 32);
		const xy_t &f = a[g.as], &l = a[g.as + g.cnt - 1]; /* mm_reg_set_coor, hit.c:40-64 */
		g.rev = (uint32_t)(f.x >> 63); g.rid = (int32_t)(f.x << 1 >> 33);
		g.rs = (int32_t)f.x; g.re = (int32_t)l.x + 1;
		g.qs = (int32_t)f.y; g.qe = (int32_t)l.y + 1;
	}
}

void set_parent(float mask_level, int mask_len, std::vector<Reg> &r) /* hit.c:195-263, hard_mask_level = 0 */
{
	int n = (int)r.size();
	if (n <= 0) return;
	for (int i = 0; i < n; ++i) r[i].id = i;
	std::vector<int> w(n); std::vector<uint64_t> cov(n);
	w[0] = 0; r[0].parent = 0;
	int k = 1;
	for (int i = 1; i < n; ++i) {
		Reg &ri = r[i];
		int si = ri.qs, ei = ri.qe, n_cov = 0, uncov = 0, j;
		for (j = 0; j < k; ++j) {
			const Reg &rp = r[w[j]];
			int sj = rp.qs, ej = rp.qe;
			if (ej <= si || sj >= ei) continue;
			if (sj < si) sj = si;
			if (ej > ei) ej = ei;
			cov[n_cov++] = (uint64_t)sj << 32 | (uint32_t)ej;
		}
		j = k; /* default: primary when nothing overlaps */
		if (n_cov > 0) {
			int x = si;
			sort_u64(cov.data(), cov.data() + n_cov);
			for (int c = 0; c < n_cov; ++c) {
				if ((int)(cov[c] >> 32) > x) uncov += (int)(cov[c] >> 32) - x;
				x = (int32_t)cov[c] > x ? (int32_t)cov[c] : x;
			}
			if (ei > x) uncov += ei - x;
			for (j = 0; j < k; ++j) {
				Reg &rp = r[w[j]];
				int sj = rp.qs, ej = rp.qe;
				if (ej <= si || sj >= ei) continue;
				int mn = ej - sj < ei - si ? ej - sj : ei - si;
				int mx = ej - sj > ei - si ? ej - sj : ei - si;
				int ol = si < sj ? (ei < sj ? 0 : ei < ej ? ei - sj : ej - sj) : (ej < si ? 0 : ej < ei ? ej - si : ei - si);
				if ((float)ol / mn - (float)uncov / mx > mask_level && uncov <= mask_len) {
					ri.parent = rp.parent;
					rp.subsc = rp.subsc > ri.score ? rp.subsc : ri.score;
					if (ri.cnt >= rp.cnt) ++rp.n_sub;
					break;
				}
			}
		}
		if (j == k) { w[k++] = i; ri.parent = i; ri.n_sub = 0; }
	}
}

/* mm_select_sub with best_n handling + mm_sync_regs (hit.c:338-367, 312-336); inv/p never set here */
void select_sub(float pri_ratio, int best_n, int min_strand_sc, std::vector<Reg> &r)
{
	if (pri_ratio <= 0.0f || r.empty()) return;
	int n = (int)r.size(), k = 0, n_2nd = 0;
	for (int i = 0; i < n; ++i) {
		int p = r[i].parent;
		if (p == i) r[k++] = r[i];
		else if (r[i].score >= r[p].score * pri_ratio && n_2nd < best_n) {
			if (!(r[i].qs == r[p].qs && r[i].qe == r[p].qe && r[i].rid == r[p].rid && r[i].rs == r[p].rs && r[i].re == r[p].re)) { r[k++] = r[i]; ++n_2nd; }
		} else if (n_2nd < best_n && r[i].score > min_strand_sc && r[i].rev != r[p].rev) { r[k++] = r[i]; ++n_2nd; }
	}
	if (k != n) {
		int max_id = -1;
		for (int i = 0; i < k; ++i) max_id = std::max(max_id, r[i].id);
		std::vector<int> tmp(max_id + 1, -1);
		for (int i = 0; i < k; ++i) if (r[i].id >= 0) tmp[r[i].id] = i;
		for (int i = 0; i < k; ++i) {
			Reg &g = r[i];
			g.id = i;
			if (g.parent == -2) g.parent = i;
			else if (g.parent >= 0 && g.parent <= max_id && tmp[g.parent] >= 0) g.parent = tmp[g.parent];
			else g.parent = -1;
		}
	}
	r.resize(k);
}

void set_mapq(std::vector<Reg> &r, int min_chain_sc, int rep_len) /* hit.c:502-539, non-DTW branch */
{
	if (r.empty()) return;
	int64_t sum_sc = 0;
	for (const Reg &g : r) if (g.parent == g.id) sum_sc += g.score;
	float uniq = (float)sum_sc / (float)(sum_sc + rep_len);
	for (Reg &g : r) {
		float pen_s1 = (float)((g.score > 100 ? 1.0 : 0.01 * g.score) * (double)uniq);
		float pen_cm = g.cnt > 10 ? 1.0f : 0.1f * g.cnt;
		pen_cm = pen_s1 < pen_cm ? pen_s1 : pen_cm;
		int subsc = g.subsc > min_chain_sc ? g.subsc : min_chain_sc;
		float x = (float)subsc / g.score0;
		int mapq = (int)(pen_cm * 40.0f * (1.0f - x) * logf((float)g.score));
		mapq -= (int)fmaf(logf((float)(g.n_sub + 1)), 4.343f, .499f);
		mapq = mapq > 0 ? mapq : 0;
		g.mapq = mapq < 60 ? mapq : 60;
	}
}

/* ------------------------------------------------------------------------------------------
 * per-chunk glue (ri_map_frag, src/rmap.cpp:210-387)
 * ---------------------------------------------------------------------------------------- */
struct ChunkTap {
	uint32_t n_sig = 0; int rep_len = 0;
	std::vector<float> events; std::vector<xy_t> seeds, anchors, chain_a, prev_a; std::vector<uint64_t> u; std::vector<Reg> regs;
	bool gated = false;
};

struct ReadState { NormState norm; Carry carry; std::vector<Reg> regs; };

void map_chunk(const rh_params_t &P, const Index &idx, const float *sig, uint32_t len, const char *qname, ReadState &rs, ChunkTap *tap)
{
	std::vector<float> ev; uint32_t n_sig = 0;
	rs.regs.clear();
	events_of(P, sig, len, rs.norm, ev, &n_sig);
	uint32_t n_events = (uint32_t)ev.size();
	if (tap) { tap->n_sig = n_sig; tap->events = ev; }
	if (n_events < P.min_events) { if (tap) tap->gated = true; return; }
	std::vector<xy_t> seeds;
	sketch(P, ev.data(), n_events, 0, 0, seeds);
	std::vector<xy_t> anchors; int rep_len = 0;
	seed_hits(P, idx, qname, seeds, rs.carry, anchors, &rep_len);
	if (tap) { tap->seeds = seeds; tap->anchors = anchors; tap->rep_len = rep_len; }
	float pen_gap = P.chain_gap_scale * 0.01 * (P.e + P.k - 1), pen_skip = P.chain_skip_scale * 0.01 * (P.e + P.k - 1);
	ChainOut co;
	chain(P, pen_gap, pen_skip, anchors, rs.carry.prev, co);
	uint32_t hash = wang32(wang32(rs.carry.offset + n_events) + wang32(11));
	gen_regs(hash, co.u, co.a, rs.regs);
	set_parent(P.mask_level, P.mask_len, rs.regs);
	if (!(P.map_flag & RH_M_ALL_CHAINS)) select_sub(P.pri_ratio, P.best_n, (int)(P.max_target_gap_length * 0.8), rs.regs);
	set_mapq(rs.regs, P.min_chaining_score, rep_len);
	if (tap) { tap->u = co.u; tap->chain_a = co.a; tap->prev_a = rs.carry.prev; tap->regs = rs.regs; }
	rs.carry.offset += n_events;
}

/* map_worker_for, src/rmap.cpp:389-599: chunk loop, stop rules, PAF fields */
void map_read(const rh_params_t &P, const Index &idx, const float *sig, uint32_t qlen, const char *qname, uint32_t read_idx, std::vector<rh_map_rec_t> &out)
{
	const bool ava = (P.map_flag & RH_M_ALL_CHAINS) != 0, noadapt = (P.map_flag & RH_M_NO_ADAPTIVE) != 0;
	const bool sig_target = (idx.flag & RH_I_SIG_TARGET) != 0;
	uint32_t l_chunk = (P.chunk_size > qlen || noadapt) ? qlen : P.chunk_size;
	uint32_t max_chunk = noadapt ? 1 : P.max_num_chunk;
	ReadState rs;
	std::vector<uint32_t> maps;
	uint32_t s_qs, c_count;
	for (s_qs = c_count = 0; s_qs < qlen && c_count < max_chunk; s_qs += l_chunk, ++c_count) {
		uint32_t s_qe = std::min(s_qs + l_chunk, qlen);
		map_chunk(P, idx, sig + s_qs, s_qe - s_qs, qname, rs, nullptr);
		const std::vector<Reg> &R = rs.regs;
		int n_regs = (int)R.size();
		int n_chains = (ava || n_regs < 1) ? n_regs : 1;
		if (n_regs == 1 && (int)R[0].mapq >= P.min_mapq) { maps.push_back(0); break; }
		float meanC = 0, meanQ = 0;
		for (const Reg &g : R) { meanC += g.score; meanQ += g.mapq; }
		if (n_regs > 0) { meanC /= n_regs; meanQ /= n_regs; }
		for (int ic = 0; ic < n_chains; ++ic) {
			float weighted = 0.0f;
			float bestQ = R[ic].mapq, bestC = R[ic].score;
			if (!ava) {
				float r_q = bestQ > 0 ? bestQ / 30.0f : 0.0f; if (r_q > 1) r_q = 1.0f;
				float r_mq = bestQ > 0 ? 1.0f - meanQ / bestQ : 0.0f; if (r_mq < 0) r_mq = 0.0f;
				float r_mc = bestC > 0 ? 1.0f - meanC / bestC : 0.0f; if (r_mc < 0) r_mc = 0.0f;
				/* rmap.cpp:487 as compiled: fma(r_mc,w_bestmc, fma(r_q,w_bestq, w_bestmq*r_mq)) */
				weighted = fmaf(r_mc, P.w_bestmc, fmaf(r_q, P.w_bestq, P.w_bestmq * r_mq));
			}
			if (weighted >= P.w_threshold || (ava && R[ic].score >= P.min_chaining_score2)) maps.push_back(ic);
		}
		if (!maps.empty()) break;
	}
	if (c_count > 0 && (s_qs >= qlen || c_count == max_chunk)) --c_count;
	uint32_t offset = rs.carry.offset;
	float scale = offset == 0 ? 0.0f : (P.sample_per_base == 0 ? 0.0f : ((float)(c_count + 1) * l_chunk / offset) / P.sample_per_base);
	const std::vector<Reg> &R = rs.regs;
	int n_regs = (int)R.size();
	if (maps.empty() && n_regs > 0 && (int)R[0].mapq > P.min_mapq) maps.push_back(0);
	rh_map_rec_t rec; memset(&rec, 0, sizeof(rec));
	rec.read_idx = read_idx; rec.ci = c_count + 1; rec.sl = qlen;
	if (maps.empty()) {
		rec.read_length = sig_target ? offset : (uint32_t)(scale * offset);
		if (n_regs >= 1) { rec.cm = R[0].cnt; rec.nc = n_regs; rec.s1 = R[0].score; }
		out.push_back(rec);
		return;
	}
	for (uint32_t c_id : maps) {
		const Reg &g = R[c_id];
		rh_map_rec_t m = rec;
		m.c_id = c_id; m.cm = g.cnt; m.nc = n_regs; m.s1 = g.score;
		m.read_length = sig_target ? offset : (uint32_t)(scale * g.qe);
		m.ref_id = g.rid;
		m.read_start_position = sig_target ? g.qs : (uint32_t)(scale * g.qs);
		m.read_end_position = sig_target ? g.qe : (uint32_t)(scale * g.qe);
		m.fragment_start_position = g.rev ? (uint32_t)(idx.lens[g.rid] + 1 - g.re) : g.rs;
		m.fragment_length = (uint32_t)(g.re - g.rs + 1);
		m.mapq = (uint8_t)g.mapq; m.rev = g.rev ? 1 : 0; m.mapped = 1;
		out.push_back(m);
	}
}

std::string format_paf(const Index &idx, const std::vector<rh_map_rec_t> &recs, const char *const *names) /* rmap.cpp:751-772 + tags 527-570 */
{
	std::string s; char line[4096];
	for (const rh_map_rec_t &m : recs) {
		char tags[256];
		if (m.mapped || m.nc >= 1)
			snprintf(tags, sizeof(tags), "mt:f:%.6f\tci:i:%d\tsl:i:%d\tcm:i:%d\tnc:i:%d\ts1:i:%d\tsm:f:%.2f", 0.0, m.ci, m.sl, m.cm, m.nc, m.s1, 0.0);
		else
			snprintf(tags, sizeof(tags), "mt:f:%.6f\tci:i:%d\tsl:i:%d\tcm:i:0\tnc:i:0\ts1:i:0\tsm:f:0", 0.0, m.ci, m.sl);
		if (m.mapped) {
			if (m.ref_id >= idx.names.size()) continue;
			snprintf(line, sizeof(line), "%s\t%u\t%u\t%u\t%c\t%s\t%u\t%u\t%u\t%u\t%u\t%u\t%s\n", names[m.read_idx], m.read_length,
			         m.read_start_position, m.read_end_position, m.rev ? '-' : '+', idx.names[m.ref_id].c_str(), idx.lens[m.ref_id],
			         m.fragment_start_position, m.fragment_start_position + m.fragment_length,
			         m.read_end_position - m.read_start_position - 1, m.fragment_length, m.mapq, tags);
		} else
			snprintf(line, sizeof(line), "%s\t%u\t*\t*\t*\t*\t*\t*\t*\t*\t*\t%u\t%s\n", names[m.read_idx], m.read_length, m.mapq, tags);
		s += line;
	}
	return s;
}

struct Ctx { rh_params_t P; std::vector<float> pore; Index idx; bool have_idx = false; };

} /* namespace */

/* ==========================================================================================
 * C surface (same shape as oracle/ref_tap.cpp so tests can swap one for the other)
 * ======================================================================================== */
extern "C" {

void *orc_open(const char *preset, int r10, const char *pore_path)
{
	Ctx *c = new Ctx();
	params_defaults(&c->P);
	if (preset && preset[0] && params_preset(&c->P, preset) != 0) { delete c; return 0; }
	if (r10) params_r10(&c->P);
	if (pore_path && pore_path[0]) pore_load(pore_path, c->P.k, c->P.lev_col, c->pore);
	return c;
}
void orc_set_sampling(void *h, uint32_t sample_rate, uint32_t bp_per_sec)
{
	Ctx *c = (Ctx *)h;
	c->P.bp_per_sec = bp_per_sec; c->P.sample_rate = sample_rate; c->P.sample_per_base = (float)sample_rate / bp_per_sec;
}
void orc_set_chunks(void *h, uint32_t chunk_size, uint32_t max_num_chunk)
{
	Ctx *c = (Ctx *)h;
	if (chunk_size) c->P.chunk_size = chunk_size;
	if (max_num_chunk) c->P.max_num_chunk = max_num_chunk;
}
void orc_get_params(void *h, rh_params_t *p) { *p = ((Ctx *)h)->P; }
void orc_set_params(void *h, const rh_params_t *p) { ((Ctx *)h)->P = *p; }
uint32_t orc_pore_vals(void *h, float *out, uint32_t cap)
{
	Ctx *c = (Ctx *)h; uint32_t n = (uint32_t)c->pore.size();
	if (out) memcpy(out, c->pore.data(), std::min(n, cap) * sizeof(float));
	return n;
}

int orc_build_index(void *h, const char *fasta, const char *dump_path, int n_threads)
{
	(void)dump_path;
	Ctx *c = (Ctx *)h;
	if (c->pore.empty()) return -5;
	std::vector<std::string> names, seqs;
	if (!read_fasta(fasta, names, seqs)) return -1;
	c->idx = Index(); c->idx.flag = c->P.idx_flag;
	size_t ns = seqs.size();
	std::vector<std::vector<xy_t>> per(ns);
	std::atomic<size_t> next(0);
	auto work = [&]() {
		std::vector<float> sv;
		for (size_t i; (i = next++) < ns;) {
			if (seqs[i].empty()) continue;
			for (int strand = 0; strand < 2; ++strand) {
				seq_to_sig(seqs[i].data(), (int)seqs[i].size(), c->pore.data(), c->P.k, strand, sv);
				sketch(c->P, sv.data(), (uint32_t)sv.size(), (uint32_t)i, strand, per[i]);
			}
		}
	};
	std::vector<std::thread> th;
	for (int t = 0; t < std::max(1, n_threads); ++t) th.emplace_back(work);
	for (auto &t : th) t.join();
	std::vector<xy_t> all;
	for (size_t i = 0; i < ns; ++i) {
		c->idx.names.push_back(names[i]); c->idx.lens.push_back((uint32_t)seqs[i].size());
		all.insert(all.end(), per[i].begin(), per[i].end());
		std::vector<xy_t>().swap(per[i]);
	}
	c->idx.build(all);
	c->have_idx = true;
	return 0;
}

int orc_build_index_sig(void *h, uint32_t n, const float *const *sig, const uint32_t *lens, const char *const *names, int n_threads)
{ /* worker_sig_pipeline, src/rindex.c:239-309: whole-read event detection with fresh sums */
	(void)n_threads;
	Ctx *c = (Ctx *)h;
	c->idx = Index(); c->idx.flag = c->P.idx_flag;
	std::vector<xy_t> all; std::vector<float> ev;
	for (uint32_t i = 0; i < n; ++i) {
		c->idx.names.push_back(names[i]); c->idx.lens.push_back(lens[i]);
		if (lens[i] == 0) continue;
		NormState st;
		events_of(c->P, sig[i], lens[i], st, ev, nullptr);
		sketch(c->P, ev.data(), (uint32_t)ev.size(), i, 0, all);
	}
	c->idx.build(all);
	c->have_idx = true;
	return 0;
}

int orc_mapopt_update(void *h) /* ri_mapopt_update, src/rindex.c:1041-1053 */
{
	Ctx *c = (Ctx *)h;
	if (!c->have_idx) return -1;
	int m = c->idx.max_occ(c->P.mid_occ_frac);
	if (m < c->P.min_mid_occ) m = c->P.min_mid_occ;
	if (c->P.max_mid_occ > c->P.min_mid_occ && m > c->P.max_mid_occ) m = c->P.max_mid_occ;
	c->P.mid_occ = m;
	return m;
}
void orc_set_mid_occ(void *h, int v) { ((Ctx *)h)->P.mid_occ = v; }
void orc_set_best_n(void *h, int v) { ((Ctx *)h)->P.best_n = v; }
uint32_t orc_n_seq(void *h) { return (uint32_t)((Ctx *)h)->idx.names.size(); }
const uint64_t *orc_idx_get(void *h, uint64_t hash, int *n) { return ((Ctx *)h)->idx.get((uint32_t)hash, n); }

/* slow5 pA conversion + outlier drop (src/rsig.c:488-503) */
uint32_t orc_raw_to_pa(const int16_t *raw, uint64_t len, double offset, double range, double digitisation, float *out)
{
	uint32_t n = 0;
	float scale = (float)(range / digitisation);
	for (uint64_t i = 0; i < len; ++i) {
		float pa = (float)((raw[i] + offset) * scale);
		if (pa > 30.0f && pa < 200.0f) out[n++] = pa;
	}
	return n;
}

uint32_t orc_detect_events(void *h, const float *sig, uint32_t s_len, double *mean_sum, double *std_dev_sum, uint32_t *n_events_sum, float *out, uint32_t cap)
{
	Ctx *c = (Ctx *)h;
	NormState st; st.sum = *mean_sum; st.sum2 = *std_dev_sum; st.n = *n_events_sum;
	std::vector<float> ev;
	events_of(c->P, sig, s_len, st, ev, nullptr);
	*mean_sum = st.sum; *std_dev_sum = st.sum2; *n_events_sum = st.n;
	memcpy(out, ev.data(), std::min((size_t)cap, ev.size()) * sizeof(float));
	return (uint32_t)ev.size();
}

uint32_t orc_sketch(void *h, const float *events, uint32_t n_events, uint32_t id, int strand, uint64_t *out_xy, uint32_t cap)
{
	Ctx *c = (Ctx *)h;
	std::vector<xy_t> s;
	sketch(c->P, events, n_events, id, strand, s);
	for (size_t i = 0; i < s.size() && i < cap; ++i) { out_xy[2 * i] = s[i].x; out_xy[2 * i + 1] = s[i].y; }
	return (uint32_t)s.size();
}

uint32_t orc_dynamic_quantize(float v, float fine_min, float fine_max, float fine_range, uint32_t n_buckets) { return quantize(v, fine_min, fine_max, fine_range, n_buckets); }
void orc_radix_sort_128x(uint64_t *xy, uint64_t n) { sort_xy_by_x((xy_t *)xy, (xy_t *)xy + n); }
void orc_radix_sort_64(uint64_t *x, uint64_t n) { sort_u64(x, x + n); }

char *orc_map_paf(void *h, uint32_t n, const float *const *sig, const uint32_t *lens, const char *const *names, int n_threads, double *map_seconds)
{
	Ctx *c = (Ctx *)h;
	if (!c->have_idx) return 0;
	std::vector<std::vector<rh_map_rec_t>> per(n);
	std::atomic<uint32_t> next(0);
	auto work = [&]() { for (uint32_t i; (i = next++) < n;) map_read(c->P, c->idx, sig[i], lens[i], names[i], i, per[i]); };
	struct timespec t0, t1; clock_gettime(CLOCK_MONOTONIC, &t0);
	std::vector<std::thread> th;
	for (int t = 0; t < std::max(1, n_threads); ++t) th.emplace_back(work);
	for (auto &t : th) t.join();
	clock_gettime(CLOCK_MONOTONIC, &t1);
	if (map_seconds) *map_seconds = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
	std::vector<rh_map_rec_t> all;
	for (auto &v : per) all.insert(all.end(), v.begin(), v.end());
	std::string s = format_paf(c->idx, all, names);
	char *ret = (char *)malloc(s.size() + 1);
	memcpy(ret, s.c_str(), s.size() + 1);
	return ret;
}

/* records instead of text (used to check the GPU records field by field) */
int64_t orc_map_recs(void *h, uint32_t n, const float *const *sig, const uint32_t *lens, const char *const *names, rh_map_rec_t *out, uint64_t cap)
{
	Ctx *c = (Ctx *)h;
	std::vector<rh_map_rec_t> all;
	for (uint32_t i = 0; i < n; ++i) map_read(c->P, c->idx, sig[i], lens[i], names[i], i, all);
	if (all.size() > cap) return -(int64_t)all.size();
	memcpy(out, all.data(), all.size() * sizeof(rh_map_rec_t));
	return (int64_t)all.size();
}

void orc_free(void *p) { free(p); }

int orc_tap_read(void *h, const float *sig, uint32_t qlen, const char *qname, rh_tap_t *tap)
{
	Ctx *c = (Ctx *)h; const rh_params_t &P = c->P;
	const bool noadapt = (P.map_flag & RH_M_NO_ADAPTIVE) != 0;
	uint32_t l_chunk = (P.chunk_size > qlen || noadapt) ? qlen : P.chunk_size;
	uint32_t max_chunk = noadapt ? 1 : P.max_num_chunk;
	ReadState rs;
	uint64_t o_ev = 0, o_seed = 0, o_anc = 0, o_u = 0, o_ca = 0, o_reg = 0;
	uint32_t cc = 0;
	for (uint32_t s_qs = 0; s_qs < qlen && cc < max_chunk; s_qs += l_chunk, ++cc) {
		uint32_t s_qe = std::min(s_qs + l_chunk, qlen);
		if (cc >= tap->cap_chunks) return RH_ERR_NOMEM;
		ChunkTap ct;
		map_chunk(P, c->idx, sig + s_qs, s_qe - s_qs, qname, rs, &ct);
		int32_t *cnt = tap->cnt + (size_t)cc * RH_TAP_NCNT;
		memset(cnt, 0, sizeof(int32_t) * RH_TAP_NCNT);
		cnt[RH_TAP_NSIG] = ct.n_sig; cnt[RH_TAP_NEVENTS] = (int32_t)ct.events.size();
		if (o_ev + ct.events.size() > tap->cap_events) return RH_ERR_NOMEM;
		memcpy(tap->events + o_ev, ct.events.data(), ct.events.size() * 4); o_ev += ct.events.size();
		if (ct.gated) continue;
		cnt[RH_TAP_NSEEDS] = (int32_t)ct.seeds.size(); cnt[RH_TAP_NANCHORS] = (int32_t)ct.anchors.size();
		cnt[RH_TAP_NU] = (int32_t)ct.u.size(); cnt[RH_TAP_NV] = (int32_t)ct.chain_a.size();
		cnt[RH_TAP_NREGS] = (int32_t)ct.regs.size(); cnt[RH_TAP_REPLEN] = ct.rep_len;
		if (o_seed + ct.seeds.size() > tap->cap_seeds || o_anc + ct.anchors.size() > tap->cap_anchors || o_u + ct.u.size() > tap->cap_u ||
		    o_ca + ct.chain_a.size() > tap->cap_chain_a || o_ca + ct.chain_a.size() > tap->cap_prev_a || o_reg + ct.regs.size() > tap->cap_regs) return RH_ERR_NOMEM;
		memcpy(tap->seeds + 2 * o_seed, ct.seeds.data(), ct.seeds.size() * 16); o_seed += ct.seeds.size();
		memcpy(tap->anchors + 2 * o_anc, ct.anchors.data(), ct.anchors.size() * 16); o_anc += ct.anchors.size();
		memcpy(tap->u + o_u, ct.u.data(), ct.u.size() * 8); o_u += ct.u.size();
		memcpy(tap->chain_a + 2 * o_ca, ct.chain_a.data(), ct.chain_a.size() * 16);
		memcpy(tap->prev_a + 2 * o_ca, ct.prev_a.data(), ct.prev_a.size() * 16); o_ca += ct.chain_a.size();
		for (size_t i = 0; i < ct.regs.size(); ++i) {
			const Reg &g = ct.regs[i];
			int32_t *f = tap->regs + (o_reg + i) * RH_TAP_REG_NF;
			f[0] = g.score; f[1] = g.cnt; f[2] = g.rid; f[3] = g.rev; f[4] = g.qs; f[5] = g.qe; f[6] = g.rs; f[7] = g.re;
			f[8] = g.parent; f[9] = g.subsc; f[10] = g.n_sub; f[11] = g.mapq; f[12] = g.as; f[13] = g.score0;
		}
		o_reg += ct.regs.size();
	}
	tap->n_chunks = cc;
	return 0;
}

} /* extern "C" */
