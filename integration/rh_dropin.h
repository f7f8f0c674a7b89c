/*
 * rh_dropin.h — the reference-side binding of librawhash_b200.so, as real code (INTEGRATION.md §2-3 describe it).
 *
 * This header is written AGAINST the reference's own structures (step_mt, pipeline_mt, ri_sig_t, ri_reg1_t, ri_map_t,
 * ri_idx_t, ri_mapopt_t: src/rmap.h:12-67, src/rsig.h:19-28, src/rindex.h:29-60, src/roptions.h:50-143) and is meant
 * to be #included into the reference's src/rmap.cpp.  integration/build_dropin.py copies the reference sources to a
 * scratch directory, makes the four small edits listed there (keep the raw samples in ri_sig_t, include this file,
 * replace the kt_for line, initialise after ri_mapopt_update) and builds oracle/_ref/rawhash2_gpu: the reference's
 * own main(), slow5lib reader, three-step pipeline and PAF printer, with step 1 running on the GPU.
 * tests/test_zz_dropin_gpu.py runs that binary against the golden PAFs.
 */
#ifndef RH_DROPIN_H
#define RH_DROPIN_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "rawhash_b200.h"

static rh_gpu_ctx *g_rh_ctx = 0;
static rh_index_t *g_rh_idx = 0;

/* rh_params_t from what the reference CLI already built: sketch parameters as stored in the index
 * (src/rindex.h:29-60), everything else from ri_mapopt_t — same field names on both sides. */
static void rh_dropin_params(const ri_idx_t *ri, const ri_mapopt_t *o, rh_params_t *p)
{
	rh_params_init(p);
	p->w = ri->w; p->e = ri->e; p->n = ri->n; p->q = ri->q; p->k = ri->k; p->idx_flag = ri->flag;
	p->diff = ri->diff; p->fine_min = ri->fine_min; p->fine_max = ri->fine_max; p->fine_range = ri->fine_range;
	p->window_length1 = o->window_length1; p->window_length2 = o->window_length2;
	p->threshold1 = o->threshold1; p->threshold2 = o->threshold2; p->peak_height = o->peak_height;
	p->bp_per_sec = o->bp_per_sec; p->sample_rate = o->sample_rate; p->chunk_size = o->chunk_size; p->sample_per_base = o->sample_per_base;
	p->mid_occ_frac = o->mid_occ_frac; p->min_mid_occ = o->min_mid_occ; p->max_mid_occ = o->max_mid_occ; p->mid_occ = o->mid_occ;
	p->min_events = o->min_events; p->bw = o->bw;
	p->max_target_gap_length = o->max_target_gap_length; p->max_query_gap_length = o->max_query_gap_length;
	p->max_chain_iter = o->max_chain_iter; p->max_num_skips = o->max_num_skips; p->min_num_anchors = o->min_num_anchors;
	p->min_chaining_score = o->min_chaining_score; p->min_chaining_score2 = o->min_chaining_score2;
	p->chain_gap_scale = o->chain_gap_scale; p->chain_skip_scale = o->chain_skip_scale;
	p->mask_level = o->mask_level; p->mask_len = o->mask_len; p->pri_ratio = o->pri_ratio; p->best_n = o->best_n; p->alt_drop = o->alt_drop;
	p->w_bestq = o->w_bestq; p->w_bestmq = o->w_bestmq; p->w_bestmc = o->w_bestmc; p->w_threshold = o->w_threshold;
	p->max_num_chunk = o->max_num_chunk; p->min_mapq = o->min_mapq; p->map_flag = o->flag;
}

/* Called from main() once the index is in memory and ri_mapopt_update has set mid_occ (src/main.cpp:573).
 * `ind_path` is the `.ind` file the index came from or was dumped to (-d); with neither, the in-memory index is
 * written to a scratch file with the reference's own ri_idx_dump and read back. */
static int rh_dropin_init(const ri_idx_t *ri, const ri_mapopt_t *opt, const char *ind_path)
{
	rh_params_t P, from_file;
	rh_dropin_params(ri, opt, &P);
	char scratch[64] = "";
	if (!ind_path) {
		snprintf(scratch, sizeof(scratch), "/tmp/rh_dropin_%d.ind", (int)getpid());
		FILE *f = fopen(scratch, "wb");
		if (!f) { fprintf(stderr, "[rawhash_b200] cannot create %s\n", scratch); return -1; }
		ri_idx_dump(f, ri);
		fclose(f);
		ind_path = scratch;
	}
	from_file = P;
	g_rh_idx = rh_index_load(ind_path, &from_file);
	if (scratch[0]) unlink(scratch);
	if (!g_rh_idx) { fprintf(stderr, "[rawhash_b200] %s\n", rh_gpu_last_error()); return -1; }
	const char *dev = getenv("RH_DEVICE");
	g_rh_ctx = rh_gpu_init(g_rh_idx, &P, dev ? atoi(dev) : 0, 0);
	if (!g_rh_ctx) { fprintf(stderr, "[rawhash_b200] %s (there is no CPU fallback behind this call)\n", rh_gpu_last_error()); return -1; }
	return 0;
}

static void rh_dropin_destroy(void)
{
	if (g_rh_ctx) rh_gpu_destroy(g_rh_ctx);
	if (g_rh_idx) rh_index_destroy(g_rh_idx);
	g_rh_ctx = 0; g_rh_idx = 0;
}

/* Stands where `kt_for(p->n_threads, map_worker_for, in, s->n_sig)` stood (src/rmap.cpp:700): fills reg[i] for
 * every read of the mini-batch exactly as map_worker_for leaves it (src/rmap.cpp:521-590). */
static void rh_dropin_map_step(step_mt *s)
{
	const uint32_t n = (uint32_t)s->n_sig;
	const int16_t **raw = (const int16_t **)malloc(n * sizeof(*raw));
	uint64_t *len = (uint64_t *)malloc(n * sizeof(*len));
	double *off = (double *)malloc(n * sizeof(double)), *rng = (double *)malloc(n * sizeof(double)), *dig = (double *)malloc(n * sizeof(double));
	const char **names = (const char **)malloc(n * sizeof(*names));
	for (uint32_t i = 0; i < n; ++i) {
		raw[i] = s->sig[i]->raw; len[i] = s->sig[i]->l_raw; names[i] = s->sig[i]->name;
		off[i] = s->sig[i]->cal_offset; rng[i] = s->sig[i]->cal_range; dig[i] = s->sig[i]->cal_digitisation;
	}
	rh_map_rec_t *recs = 0; uint64_t n_recs = 0;
	if (rh_gpu_map_batch_raw(g_rh_ctx, n, raw, len, off, rng, dig, names, &recs, &n_recs) != RH_OK) {
		fprintf(stderr, "[rawhash_b200] %s\n", rh_gpu_last_error());
		abort(); /* map_worker_for has no error path either */
	}
	for (uint64_t k = 0; k < n_recs;) { /* records arrive grouped by read, in read order */
		const uint32_t i = recs[k].read_idx; uint64_t e = k;
		while (e < n_recs && recs[e].read_idx == i) ++e;
		ri_reg1_t *reg = s->reg[i];
		reg->read_id = s->sig[i]->rid; reg->read_name = s->sig[i]->name;
		const uint32_t n_rec = (uint32_t)(e - k);
		/* an unmapped read keeps n_maps == 0 but owns ONE zeroed record carrying its tags (src/rmap.cpp:521-556);
		 * step 2 prints maps[0] in the unmapped format when n_maps == 0 (src/rmap.cpp:766-772) */
		reg->n_maps = recs[k].mapped ? n_rec : 0;
		reg->maps = (ri_map_t *)calloc(n_rec, sizeof(ri_map_t));
		for (uint32_t m = 0; m < n_rec; ++m) {
			const rh_map_rec_t *r = &recs[k + m]; ri_map_t *o = &reg->maps[m];
			o->c_id = r->c_id; o->read_length = r->read_length; o->ref_id = r->ref_id;
			o->read_start_position = r->read_start_position; o->read_end_position = r->read_end_position;
			o->fragment_start_position = r->fragment_start_position; o->fragment_length = r->fragment_length;
			o->mapq = r->mapq; o->rev = r->rev; o->mapped = r->mapped;
			o->tags = (char *)malloc(1024);
			if (r->mapped || r->nc >= 1) /* src/rmap.cpp:527-570 */
				snprintf(o->tags, 1024, "mt:f:%.6f\tci:i:%u\tsl:i:%u\tcm:i:%d\tnc:i:%d\ts1:i:%d\tsm:f:%.2f", (double)r->mt_ms, r->ci, r->sl, r->cm, r->nc, r->s1, 0.0);
			else
				snprintf(o->tags, 1024, "mt:f:%.6f\tci:i:%u\tsl:i:%u\tcm:i:0\tnc:i:0\ts1:i:0\tsm:f:0", (double)r->mt_ms, r->ci, r->sl);
		}
		k = e;
	}
	for (uint32_t i = 0; i < n; ++i) { free(s->sig[i]->raw); s->sig[i]->raw = 0; } /* the raw copy was made for this step only */
	rh_free(recs); free(raw); free(len); free(off); free(rng); free(dig); free(names);
}

#endif /* RH_DROPIN_H */
