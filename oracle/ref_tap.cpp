/*
 * ref_tap.cpp — TEST INFRASTRUCTURE ONLY.
 *
 * Harness around the UNMODIFIED reference (RawHash2 @ /root/reference/src).
 * It is compiled by oracle/Makefile together with the reference sources *where
 * they lie* (nothing is copied into this repo) into oracle/_ref/libref_tap.so.
 * The reference's translation units rmap.cpp and main.cpp are #included so
 * that their `static` functions (map_worker_for, collect_seed_hits,
 * ri_set_opt) can be called directly; main() itself is renamed and never run.
 *
 * Exposes a small C API (ctypes-friendly) to:
 *   - build / dump / load the reference index,
 *   - run the reference's own kt_for(map_worker_for) loop on in-memory pA
 *     signals and return the PAF text (the CPU baseline, kind "reference"),
 *   - tap every stage of ri_map_frag for one read, chunk by chunk.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's reference / cpu_baseline
 * legs may load this library.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "khash.h" /* __ac_Wang_hash, normally pulled in through slow5lib */
#define main rawhash_ref_cli_main_unused
#include "main.cpp" /* ri_set_opt + presets (src/main.cpp:111-210) */
#undef main
#include "rmap.cpp" /* map_worker_for, collect_seed_hits, ri_map_frag (static/internal) */
#include "kalloc.h"

#include "../include/rawhash_b200.h"

void ri_idx_sort(ri_idx_t *ri, int n_threads); /* C++ linkage: defined in rindex.c outside any extern "C" block */

struct ref_ctx {
	ri_idxopt_t ipt;
	ri_mapopt_t opt;
	ri_pore_t pore;
	ri_idx_t *ri;
	int have_pore;
};

/* The signal readers are compiled out (-DNHDF5RH -DNPOD5RH -DNSLOW5RH): the harness
 * feeds signals from memory, so file discovery/reading is never reached. */

extern "C" {

void *ref_open(const char *preset, int r10, const char *pore_path)
{
	ref_ctx *c = (ref_ctx *)calloc(1, sizeof(ref_ctx));
	ri_verbose = 0;
	ri_set_opt(0, &c->ipt, &c->opt);
	if (preset && preset[0] && ri_set_opt(preset, &c->ipt, &c->opt) != 0) { free(c); return 0; }
	if (r10) { /* src/main.cpp:363-376 */
		c->ipt.k = 9;
		c->ipt.window_length1 = 3; c->ipt.window_length2 = 6;
		c->ipt.threshold1 = 6.5f; c->ipt.threshold2 = 4.0f;
		c->ipt.peak_height = 0.2f;
		c->opt.window_length1 = 3; c->opt.window_length2 = 6;
		c->opt.threshold1 = 6.5f; c->opt.threshold2 = 4.0f;
		c->opt.peak_height = 0.2f;
		c->opt.chain_gap_scale = 1.2f;
	}
	c->pore.pore_vals = NULL; c->pore.pore_inds = NULL;
	c->pore.max_val = -5000.0; c->pore.min_val = 5000.0;
	if (pore_path && pore_path[0]) {
		load_pore(pore_path, c->ipt.k, c->ipt.lev_col, &c->pore);
		c->have_pore = c->pore.pore_vals != NULL;
	}
	return c;
}

void ref_set_sampling(void *h, uint32_t sample_rate, uint32_t bp_per_sec)
{ /* --sample-rate / --bp-per-sec, src/main.cpp:343-350 */
	ref_ctx *c = (ref_ctx *)h;
	c->opt.bp_per_sec = bp_per_sec; c->opt.sample_rate = sample_rate;
	c->opt.sample_per_base = (float)c->opt.sample_rate / c->opt.bp_per_sec;
	c->ipt.bp_per_sec = bp_per_sec; c->ipt.sample_rate = sample_rate;
	c->ipt.sample_per_base = (float)c->ipt.sample_rate / c->ipt.bp_per_sec;
}

void ref_set_chunks(void *h, uint32_t chunk_size, uint32_t max_num_chunk)
{
	ref_ctx *c = (ref_ctx *)h;
	if (chunk_size) c->opt.chunk_size = chunk_size;
	if (max_num_chunk) c->opt.max_num_chunk = max_num_chunk;
}

void ref_get_params(void *h, rh_params_t *p)
{
	ref_ctx *c = (ref_ctx *)h;
	memset(p, 0, sizeof(*p));
	const ri_idxopt_t &i = c->ipt; const ri_mapopt_t &o = c->opt;
	p->w = i.w; p->e = i.e; p->n = i.n; p->q = i.q; p->k = i.k; p->idx_flag = i.flag; p->lev_col = i.lev_col;
	p->diff = i.diff; p->fine_min = i.fine_min; p->fine_max = i.fine_max; p->fine_range = i.fine_range;
	p->window_length1 = o.window_length1; p->window_length2 = o.window_length2;
	p->threshold1 = o.threshold1; p->threshold2 = o.threshold2; p->peak_height = o.peak_height;
	p->bp_per_sec = o.bp_per_sec; p->sample_rate = o.sample_rate; p->chunk_size = o.chunk_size;
	p->sample_per_base = o.sample_per_base;
	p->mid_occ_frac = o.mid_occ_frac; p->min_mid_occ = o.min_mid_occ; p->max_mid_occ = o.max_mid_occ; p->mid_occ = o.mid_occ;
	p->min_events = o.min_events; p->bw = o.bw;
	p->max_target_gap_length = o.max_target_gap_length; p->max_query_gap_length = o.max_query_gap_length;
	p->max_chain_iter = o.max_chain_iter; p->max_num_skips = o.max_num_skips; p->min_num_anchors = o.min_num_anchors;
	p->min_chaining_score = o.min_chaining_score; p->min_chaining_score2 = o.min_chaining_score2;
	p->chain_gap_scale = o.chain_gap_scale; p->chain_skip_scale = o.chain_skip_scale;
	p->mask_level = o.mask_level; p->mask_len = o.mask_len; p->pri_ratio = o.pri_ratio; p->best_n = o.best_n; p->alt_drop = o.alt_drop;
	p->w_bestq = o.w_bestq; p->w_bestmq = o.w_bestmq; p->w_bestmc = o.w_bestmc; p->w_threshold = o.w_threshold;
	p->max_num_chunk = o.max_num_chunk; p->min_mapq = o.min_mapq; p->map_flag = o.flag;
}

/* normalised pore levels as load_pore left them (src/rutils.c:133-178) */
uint32_t ref_pore_vals(void *h, float *out, uint32_t cap)
{
	ref_ctx *c = (ref_ctx *)h;
	if (!c->have_pore) return 0;
	uint32_t n = c->pore.n_pore_vals;
	if (out) memcpy(out, c->pore.pore_vals, (n < cap ? n : cap) * sizeof(float));
	return n;
}

int ref_build_index(void *h, const char *fasta, const char *dump_path, int n_threads)
{
	ref_ctx *c = (ref_ctx *)h;
	if (c->ri) { ri_idx_destroy(c->ri); c->ri = 0; }
	ri_idx_reader_t *r = ri_idx_reader_open(fasta, &c->ipt, (dump_path && dump_path[0]) ? dump_path : 0);
	if (!r) return -1;
	c->ri = ri_idx_reader_read(r, &c->pore, n_threads, 1);
	ri_idx_reader_close(r);
	if (!c->ri) return -2;
	if (dump_path && dump_path[0]) {
		/* ri_idx_dump frees pore arrays of the in-memory index (src/rindex.c:558-566); reload for a clean state */
		ri_idx_destroy(c->ri);
		r = ri_idx_reader_open(dump_path, &c->ipt, 0);
		if (!r) return -3;
		c->ri = ri_idx_reader_read(r, &c->pore, n_threads, 1);
		ri_idx_reader_close(r);
		if (!c->ri) return -4;
	}
	return 0;
}

/* The reference's in-memory index (ri_idx_t: 2^14 buckets of khash + position arrays, src/rindex.c:311-363) filled
 * from a flattened key -> ascending-positions table instead of from a FASTA or an `.ind` file.  Every bucket is
 * built with the reference's own khash instantiation and the same calls worker_post makes (kh_resize, kh_put,
 * singleton keys tagged with bit 0, lists as start<<32|n into b->p), so ri_idx_get answers from the structure it
 * would have after ri_idx_load.  Used for human-size worlds, where a 13-minute host index build (or a 42 GB file
 * round trip) per bench run is not an option; tests/test_oracle_vs_ref.py checks it against ri_idx_gen. */
#define rt_idx_hash(a) ((a)>>1)
#define rt_idx_eq(a, b) ((a)>>1 == (b)>>1)
KHASH_INIT(idx, uint64_t, uint64_t, 1, rt_idx_hash, rt_idx_eq) /* same parameters as src/rindex.c:17-19 */

struct flat_fill_t {
	ri_idx_t *ri; const uint32_t *keys; const uint64_t *off, *pos;
	const uint32_t *order; const uint64_t *bstart; /* keys of bucket b: order[bstart[b] .. bstart[b+1]) */
	int bad;
};

static void flat_fill_bucket(void *g, long b, int tid)
{
	flat_fill_t *F = (flat_fill_t *)g;
	ri_idx_bucket_t *B = &F->ri->B[b];
	const uint64_t k0 = F->bstart[b], k1 = F->bstart[b + 1];
	if (k0 == k1) return;
	uint64_t np = 0;
	for (uint64_t j = k0; j < k1; ++j) { const uint64_t c = F->off[F->order[j] + 1] - F->off[F->order[j]]; if (c > 1) np += c; }
	if (np > 0x7fffffffULL) { F->bad = 1; return; }
	khash_t(idx) *h = kh_init(idx);
	kh_resize(idx, h, (khint_t)(k1 - k0));
	B->n = (int32_t)np;
	B->p = (uint64_t *)calloc(np ? np : 1, 8);
	uint64_t start_p = 0;
	for (uint64_t j = k0; j < k1; ++j) {
		const uint32_t ki = F->order[j];
		const uint64_t c = F->off[ki + 1] - F->off[ki];
		int absent;
		khint_t itr = kh_put(idx, h, (uint64_t)(F->keys[ki] >> F->ri->b) << 1, &absent);
		if (c == 1) { kh_key(h, itr) |= 1; kh_val(h, itr) = F->pos[F->off[ki]]; }
		else { memcpy(B->p + start_p, F->pos + F->off[ki], c * 8); kh_val(h, itr) = start_p << 32 | c; start_p += c; }
	}
	B->h = h;
}

int ref_index_from_flat(void *hd, uint32_t n_seq, const char *const *names, const uint32_t *lens,
                        uint64_t n_keys, const uint32_t *keys, const uint64_t *off, const uint64_t *pos, int n_threads)
{
	ref_ctx *c = (ref_ctx *)hd;
	if (c->ri) { ri_idx_destroy(c->ri); c->ri = 0; }
	const ri_idxopt_t &o = c->ipt;
	ri_idx_t *ri = ri_idx_init(o.diff, o.b, o.w, o.e, o.n, o.q, o.k, o.fine_min, o.fine_max, o.fine_range, o.flag);
	ri->window_length1 = o.window_length1; ri->window_length2 = o.window_length2;
	ri->threshold1 = o.threshold1; ri->threshold2 = o.threshold2; ri->peak_height = o.peak_height;
	ri->seq = (ri_idx_seq_t *)ri_kcalloc(ri->km, n_seq ? n_seq : 1, sizeof(ri_idx_seq_t));
	uint64_t sum_len = 0;
	for (uint32_t i = 0; i < n_seq; ++i) {
		ri->seq[i].name = (char *)ri_kmalloc(ri->km, strlen(names[i]) + 1);
		strcpy(ri->seq[i].name, names[i]);
		ri->seq[i].len = lens[i]; ri->seq[i].offset = sum_len; sum_len += lens[i];
	}
	ri->n_seq = n_seq;
	const uint32_t nb = 1u << ri->b, mask = nb - 1;
	std::vector<uint64_t> bstart(nb + 1, 0);
	for (uint64_t i = 0; i < n_keys; ++i) ++bstart[(keys[i] & mask) + 1];
	for (uint32_t b = 0; b < nb; ++b) bstart[b + 1] += bstart[b];
	std::vector<uint32_t> order(n_keys);
	{ std::vector<uint64_t> cur(bstart.begin(), bstart.end() - 1); for (uint64_t i = 0; i < n_keys; ++i) order[cur[keys[i] & mask]++] = (uint32_t)i; }
	flat_fill_t F = {ri, keys, off, pos, order.data(), bstart.data(), 0};
	kt_for(n_threads > 0 ? n_threads : 1, flat_fill_bucket, &F, nb);
	if (F.bad) { ri_idx_destroy(ri); return -2; }
	c->ri = ri;
	return 0;
}

int ref_load_index(void *h, const char *path)
{
	ref_ctx *c = (ref_ctx *)h;
	if (c->ri) { ri_idx_destroy(c->ri); c->ri = 0; }
	ri_idx_reader_t *r = ri_idx_reader_open(path, &c->ipt, 0);
	if (!r || !r->is_idx) return -1;
	c->ri = ri_idx_reader_read(r, &c->pore, 1, 1);
	ri_idx_reader_close(r);
	return c->ri ? 0 : -2;
}

/* Rawsamble: index built from pA signals; the flow of worker_sig_pipeline (src/rindex.c:239-309)
 * with the file reader replaced by in-memory signals. */
int ref_build_index_sig(void *h, uint32_t n, const float *const *sig, const uint32_t *lens, const char *const *names, int n_threads)
{
	ref_ctx *c = (ref_ctx *)h;
	if (c->ri) { ri_idx_destroy(c->ri); c->ri = 0; }
	const ri_idxopt_t &o = c->ipt;
	ri_idx_t *ri = ri_idx_init(o.diff, o.b, o.w, o.e, o.n, o.q, o.k, o.fine_min, o.fine_max, o.fine_range, o.flag);
	ri->window_length1 = o.window_length1; ri->window_length2 = o.window_length2;
	ri->threshold1 = o.threshold1; ri->threshold2 = o.threshold2; ri->peak_height = o.peak_height;
	ri->sig = (ri_sig_t *)ri_kcalloc(ri->km, n ? n : 1, sizeof(ri_sig_t));
	mm128_v a = {0, 0, 0};
	uint64_t sum_len = 0;
	for (uint32_t i = 0; i < n; ++i) {
		ri_sig_t *s = &ri->sig[ri->n_seq];
		s->name = (char *)ri_kmalloc(ri->km, strlen(names[i]) + 1);
		strcpy(s->name, names[i]);
		s->l_sig = lens[i]; s->offset = sum_len; sum_len += lens[i];
		uint32_t rid = ri->n_seq++;
		if (lens[i] > 0) {
			uint32_t s_len = 0, n_events_sum = 0; double s_sum = 0, s_std = 0;
			float *ev = detect_events(0, lens[i], sig[i], ri->window_length1, ri->window_length2, ri->threshold1, ri->threshold2, ri->peak_height, &s_sum, &s_std, &n_events_sum, &s_len);
			if (ev && s_len > 0) ri_sketch(0, ev, rid, 0, s_len, ri->diff, ri->w, ri->e, ri->n, ri->q, ri->k, ri->fine_min, ri->fine_max, ri->fine_range, &a, 0);
			if (ev) free(ev);
		}
	}
	ri_idx_add(ri, a.n, a.a);
	ri_kfree(0, a.a);
	ri_idx_sort(ri, n_threads);
	c->ri = ri;
	return 0;
}

int ref_mapopt_update(void *h)
{
	ref_ctx *c = (ref_ctx *)h;
	if (!c->ri) return -1;
	c->opt.mid_occ = 0;
	ri_mapopt_update(&c->opt, c->ri);
	return c->opt.mid_occ;
}

void ref_set_mid_occ(void *h, int mid_occ) { ((ref_ctx *)h)->opt.mid_occ = mid_occ; }
void ref_set_best_n(void *h, int best_n) { ((ref_ctx *)h)->opt.best_n = best_n; } /* --best-chains, src/main.cpp */

uint32_t ref_n_seq(void *h) { ref_ctx *c = (ref_ctx *)h; return c->ri ? c->ri->n_seq : 0; }

const uint64_t *ref_idx_get(void *h, uint64_t hash, int *n)
{
	ref_ctx *c = (ref_ctx *)h;
	return ri_idx_get(c->ri, hash, n);
}

/* pA conversion + outlier drop, restated from src/rsig.c:488-503 (that function is compiled
 * out together with slow5lib; slow5_rec_t's offset/range/digitisation are doubles). */
uint32_t ref_raw_to_pa(const int16_t *raw, uint64_t len, double offset, double range, double digitisation, float *out)
{
	uint32_t l_sig = 0;
	float pa = 0.0f;
	float scale = range / digitisation;
	for (uint64_t i = 0; i < len; ++i) {
		pa = (raw[i] + offset) * scale;
		if (pa > 30.0f && pa < 200.0f) out[l_sig++] = pa;
	}
	return l_sig;
}

/* --- stage functions exported one by one (called through the reference's own symbols) --- */
uint32_t ref_detect_events(void *h, const float *sig, uint32_t s_len, double *mean_sum, double *std_dev_sum, uint32_t *n_events_sum, float *out, uint32_t cap)
{
	ref_ctx *c = (ref_ctx *)h;
	uint32_t n_events = 0;
	float *ev = detect_events(0, s_len, sig, c->opt.window_length1, c->opt.window_length2, c->opt.threshold1, c->opt.threshold2, c->opt.peak_height, mean_sum, std_dev_sum, n_events_sum, &n_events);
	if (ev) { memcpy(out, ev, (n_events < cap ? n_events : cap) * sizeof(float)); free(ev); }
	return n_events;
}

uint32_t ref_sketch(void *h, const float *events, uint32_t n_events, uint32_t id, int strand, uint64_t *out_xy, uint32_t cap)
{
	ref_ctx *c = (ref_ctx *)h;
	const ri_idxopt_t &o = c->ipt;
	mm128_v riv = {0, 0, 0};
	if (n_events == 0) return 0;
	ri_sketch(0, events, id, strand, n_events, o.diff, o.w, o.e, o.n, o.q, o.k, o.fine_min, o.fine_max, o.fine_range, &riv, 0);
	uint32_t n = riv.n;
	for (uint32_t i = 0; i < n && i < cap; ++i) { out_xy[2 * i] = riv.a[i].x; out_xy[2 * i + 1] = riv.a[i].y; }
	free(riv.a);
	return n;
}

uint32_t ref_dynamic_quantize(float v, float fine_min, float fine_max, float fine_range, uint32_t n_buckets)
{
	return dynamic_quantize(v, fine_min, fine_max, fine_range, n_buckets);
}

void ref_radix_sort_128x(uint64_t *xy, uint64_t n) { radix_sort_128x((mm128_t *)xy, (mm128_t *)xy + n); }
void ref_radix_sort_64(uint64_t *x, uint64_t n) { radix_sort_64(x, x + n); }

/* run the reference's own worker loop: kt_for(n_threads, map_worker_for, step, n) (src/rmap.cpp:700),
 * then format like step 2 (src/rmap.cpp:736-776) into a string. */
char *ref_map_paf(void *h, uint32_t n, const float *const *sig, const uint32_t *lens, const char *const *names, int n_threads, double *map_seconds)
{
	ref_ctx *c = (ref_ctx *)h;
	if (!c->ri) return 0;
	pipeline_mt pl; memset(&pl, 0, sizeof(pl));
	pl.opt = &c->opt; pl.ri = c->ri; pl.n_threads = n_threads > 0 ? n_threads : 1;
	step_mt s; memset(&s, 0, sizeof(s));
	s.p = &pl; s.n_sig = (int)n;
	std::vector<ri_sig_t> sigs(n);
	std::vector<ri_sig_t *> sigp(n);
	std::vector<ri_reg1_t *> regs(n);
	std::vector<ri_tbuf_t *> bufs(pl.n_threads);
	for (uint32_t i = 0; i < n; ++i) {
		sigs[i].rid = i; sigs[i].l_sig = lens[i]; sigs[i].name = (char *)names[i]; sigs[i].offset = 0; sigs[i].sig = (float *)sig[i];
		sigp[i] = &sigs[i];
		regs[i] = (ri_reg1_t *)calloc(1, sizeof(ri_reg1_t));
	}
	for (int t = 0; t < pl.n_threads; ++t) bufs[t] = ri_tbuf_init();
	s.sig = sigp.data(); s.reg = regs.data(); s.buf = bufs.data();
	double t0 = ri_realtime();
	kt_for(pl.n_threads, map_worker_for, &s, n);
	if (map_seconds) *map_seconds = ri_realtime() - t0;

	std::string out;
	char line[4096];
	const ri_idx_t *ri = c->ri;
	for (uint32_t k = 0; k < n; ++k) {
		ri_reg1_t *reg0 = regs[k];
		if (reg0->read_name) {
			if (reg0->n_maps > 0) {
				for (uint32_t m = 0; m < reg0->n_maps; ++m) {
					if (reg0->maps[m].ref_id < ri->n_seq) {
						snprintf(line, sizeof(line), "%s\t%u\t%u\t%u\t%c\t%s\t%u\t%u\t%u\t%u\t%u\t%u\t%s\n",
								 reg0->read_name, reg0->maps[m].read_length, reg0->maps[m].read_start_position, reg0->maps[m].read_end_position,
								 reg0->maps[m].rev ? '-' : '+',
								 (ri->flag & RI_I_SIG_TARGET) ? ri->sig[reg0->maps[m].ref_id].name : ri->seq[reg0->maps[m].ref_id].name,
								 (ri->flag & RI_I_SIG_TARGET) ? ri->sig[reg0->maps[m].ref_id].l_sig : ri->seq[reg0->maps[m].ref_id].len,
								 reg0->maps[m].fragment_start_position, reg0->maps[m].fragment_start_position + reg0->maps[m].fragment_length,
								 reg0->maps[m].read_end_position - reg0->maps[m].read_start_position - 1,
								 reg0->maps[m].fragment_length, reg0->maps[m].mapq, reg0->maps[m].tags);
						out += line;
					}
					if (reg0->maps[m].tags) free(reg0->maps[m].tags);
				}
			} else {
				snprintf(line, sizeof(line), "%s\t%u\t*\t*\t*\t*\t*\t*\t*\t*\t*\t%u\t%s\n", reg0->read_name, reg0->maps[0].read_length, reg0->maps[0].mapq, reg0->maps[0].tags);
				out += line;
				if (reg0->maps[0].tags) free(reg0->maps[0].tags);
			}
		}
		if (reg0->maps) free(reg0->maps);
		free(reg0);
	}
	for (int t = 0; t < pl.n_threads; ++t) ri_tbuf_destroy(bufs[t]);
	char *ret = (char *)malloc(out.size() + 1);
	memcpy(ret, out.c_str(), out.size() + 1);
	return ret;
}

void ref_free(void *p) { free(p); }

/* Stage tap of one read: every chunk up to max_num_chunk, no stop rules.  The sequence of
 * calls is that of ri_map_frag (src/rmap.cpp:210-387); each stage is the reference's function. */
int ref_tap_read(void *h, const float *sig, uint32_t qlen, const char *qname, rh_tap_t *tap)
{
	ref_ctx *c = (ref_ctx *)h;
	const ri_mapopt_t *opt = &c->opt; const ri_idx_t *ri = c->ri;
	if (!ri) return -1;
	ri_reg1_t reg; memset(&reg, 0, sizeof(reg));
	void *km = ri_km_init();
	uint32_t l_chunk = (opt->chunk_size > qlen || (opt->flag & RI_M_NO_ADAPTIVE)) ? qlen : opt->chunk_size;
	uint32_t max_chunk = (opt->flag & RI_M_NO_ADAPTIVE) ? 1 : opt->max_num_chunk;
	double mean_sum = 0, std_dev_sum = 0; uint32_t n_events_sum = 0;
	uint64_t o_ev = 0, o_seed = 0, o_anc = 0, o_u = 0, o_ca = 0, o_pa = 0, o_reg = 0;
	uint32_t c_count = 0;
	int rc = 0;
	for (uint32_t s_qs = 0; s_qs < qlen && c_count < max_chunk; s_qs += l_chunk, ++c_count) {
		uint32_t s_qe = s_qs + l_chunk; if (s_qe > qlen) s_qe = qlen;
		if (c_count >= tap->cap_chunks) { rc = RH_ERR_NOMEM; break; }
		int32_t *cnt = tap->cnt + (size_t)c_count * RH_TAP_NCNT;
		memset(cnt, 0, sizeof(int32_t) * RH_TAP_NCNT);
		if (reg.creg) { free(reg.creg); reg.creg = NULL; reg.n_cregs = 0; }
		uint32_t n_events = 0;
		float *events = detect_events(km, s_qe - s_qs, sig + s_qs, opt->window_length1, opt->window_length2, opt->threshold1, opt->threshold2, opt->peak_height, &mean_sum, &std_dev_sum, &n_events_sum, &n_events);
		cnt[RH_TAP_NEVENTS] = n_events;
		if (o_ev + n_events > tap->cap_events) { rc = RH_ERR_NOMEM; break; }
		if (events) memcpy(tap->events + o_ev, events, n_events * sizeof(float));
		o_ev += n_events;
		if (n_events < opt->min_events) { if (events) ri_kfree(km, events); continue; }
		mm128_v riv = {0, 0, 0};
		ri_sketch(km, events, 0, 0, n_events, ri->diff, ri->w, ri->e, ri->n, ri->q, ri->k, ri->fine_min, ri->fine_max, ri->fine_range, &riv, 0);
		ri_kfree(km, events);
		cnt[RH_TAP_NSEEDS] = riv.n;
		if (o_seed + riv.n > tap->cap_seeds) { rc = RH_ERR_NOMEM; break; }
		memcpy(tap->seeds + 2 * o_seed, riv.a, riv.n * 16); o_seed += riv.n;
		int rep_len; int64_t n_seed_pos; uint64_t *u = 0;
		mm128_t *seed_hits = collect_seed_hits(km, (opt->flag & RI_M_ALL_CHAINS) ? 1 : 0, opt->mid_occ, opt->max_max_occ, opt->occ_dist, ri, qname, &reg, &riv, n_events, &n_seed_pos, &rep_len);
		if (riv.a) ri_kfree(km, riv.a);
		cnt[RH_TAP_NANCHORS] = (int32_t)n_seed_pos; cnt[RH_TAP_REPLEN] = rep_len;
		if (o_anc + n_seed_pos > tap->cap_anchors) { rc = RH_ERR_NOMEM; break; }
		if (n_seed_pos) memcpy(tap->anchors + 2 * o_anc, seed_hits, n_seed_pos * 16);
		o_anc += n_seed_pos;
		float chn_pen_gap = opt->chain_gap_scale * 0.01 * (ri->e + ri->k - 1), chn_pen_skip = opt->chain_skip_scale * 0.01 * (ri->e + ri->k - 1);
		seed_hits = mg_lchain_dp(opt->max_target_gap_length, opt->max_query_gap_length, opt->bw, opt->max_num_skips, opt->max_chain_iter, opt->min_num_anchors, opt->min_chaining_score, chn_pen_gap, chn_pen_skip, &n_seed_pos, seed_hits, &(reg.prev_anchors), &(reg.n_cregs), &u, km);
		reg.n_prev_anchors = 0;
		if (n_seed_pos > 0) reg.n_prev_anchors = n_seed_pos;
		else if (reg.prev_anchors) { ri_kfree(km, reg.prev_anchors); reg.prev_anchors = NULL; }
		cnt[RH_TAP_NU] = reg.n_cregs; cnt[RH_TAP_NV] = (int32_t)n_seed_pos;
		if (o_u + reg.n_cregs > tap->cap_u || o_ca + n_seed_pos > tap->cap_chain_a || o_pa + n_seed_pos > tap->cap_prev_a) { rc = RH_ERR_NOMEM; break; }
		if (reg.n_cregs) memcpy(tap->u + o_u, u, reg.n_cregs * 8);
		o_u += reg.n_cregs;
		if (n_seed_pos) { memcpy(tap->chain_a + 2 * o_ca, seed_hits, n_seed_pos * 16); memcpy(tap->prev_a + 2 * o_pa, reg.prev_anchors, n_seed_pos * 16); }
		o_ca += n_seed_pos; o_pa += n_seed_pos;
		uint32_t hash = 0;
		hash ^= __ac_Wang_hash(reg.offset + n_events) + __ac_Wang_hash(11);
		hash = __ac_Wang_hash(hash);
		reg.creg = mm_gen_regs(km, hash, reg.offset + n_events, reg.n_cregs, u, seed_hits);
		mm_set_parent(km, opt->mask_level, opt->mask_len, reg.n_cregs, reg.creg, opt->flag & RI_M_HARD_MLEVEL, opt->alt_drop);
		if (!(opt->flag & RI_M_ALL_CHAINS))
			mm_select_sub(km, opt->pri_ratio, opt->best_n, 1, opt->max_target_gap_length * 0.8, &(reg.n_cregs), reg.creg);
		mm_set_mapq(km, reg.n_cregs, reg.creg, opt->min_chaining_score, rep_len, 0);
		cnt[RH_TAP_NREGS] = reg.n_cregs;
		if (o_reg + reg.n_cregs > tap->cap_regs) { rc = RH_ERR_NOMEM; break; }
		for (int i = 0; i < reg.n_cregs; ++i) {
			const mm_reg1_t *r = &reg.creg[i];
			int32_t *f = tap->regs + (o_reg + i) * RH_TAP_REG_NF;
			f[0] = r->score; f[1] = r->cnt; f[2] = r->rid; f[3] = r->rev; f[4] = r->qs; f[5] = r->qe; f[6] = r->rs; f[7] = r->re;
			f[8] = r->parent; f[9] = r->subsc; f[10] = r->n_sub; f[11] = r->mapq; f[12] = r->as; f[13] = r->score0;
		}
		o_reg += reg.n_cregs;
		if (seed_hits) ri_kfree(km, seed_hits);
		if (u) ri_kfree(km, u);
		reg.offset += n_events;
	}
	tap->n_chunks = c_count;
	if (reg.creg) free(reg.creg);
	ri_km_destroy(km);
	return rc;
}

} /* extern "C" */
