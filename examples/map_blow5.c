/*
 * map_blow5.c — the C-ABI from plain C99: load a reference-format `.ind`, read a BLOW5/SLOW5 file mini-batch by
 * mini-batch, map every batch on the GPU, print PAF.  The sequential skeleton of what rawhash2_b200
 * (rawhash_b200/csrc/rh_main.cpp) runs as a three-thread pipeline, and of steps 0-2 of the reference's
 * map_worker_pipeline (src/rmap.cpp:662-800).
 *
 *   gcc -std=c99 -O2 -Iinclude examples/map_blow5.c -Lrawhash_b200 -lrawhash_b200 -Wl,-rpath,$PWD/rawhash_b200 -o map_blow5
 *   ./map_blow5 [-x preset] target.ind reads.blow5 > out.paf
 */
#include <stdio.h>
#include <string.h>
#include "rawhash_b200.h"

int main(int argc, char **argv)
{
	const char *preset = "sensitive";
	int a = 1;
	if (argc > 2 && strcmp(argv[1], "-x") == 0) { preset = argv[2]; a = 3; }
	if (argc - a != 2) { fprintf(stderr, "usage: %s [-x preset] target.ind reads.blow5\n", argv[0]); return 2; }

	rh_params_t P;
	rh_params_init(&P);                                   /* ri_idxopt_init + ri_mapopt_init */
	if (rh_params_preset(&P, preset) != RH_OK) { fprintf(stderr, "unknown preset %s\n", preset); return 2; }
	rh_index_t *idx = rh_index_load(argv[a], &P);         /* ri_idx_load; sketch parameters come from the file */
	if (!idx) { fprintf(stderr, "%s\n", rh_gpu_last_error()); return 1; }
	rh_index_update_mapopt(idx, &P);                      /* ri_mapopt_update */
	rh_gpu_ctx *ctx = rh_gpu_init(idx, &P, 0, 0);         /* index -> HBM; fails without a GPU: there is no CPU path */
	if (!ctx) { fprintf(stderr, "%s\n", rh_gpu_last_error()); rh_index_destroy(idx); return 1; }
	rh_sigfile_t *f = rh_sigfile_open(argv[a + 1], 8);
	if (!f) { fprintf(stderr, "%s\n", rh_gpu_last_error()); rh_gpu_destroy(ctx); rh_index_destroy(idx); return 1; }

	int rc = RH_OK;
	for (;;) {
		rh_sigbatch_t *b = NULL;                          /* step 0: ri_sig_read_frag */
		if ((rc = rh_sigfile_next_batch(f, 0, 0, &b)) != RH_OK || !b) break;
		rh_map_rec_t *recs = NULL; uint64_t n_recs = 0;  /* step 1: kt_for(map_worker_for) */
		rc = rh_gpu_map_batch_raw(ctx, b->n, b->raw, b->raw_len, b->offset, b->range, b->digitisation, b->names, &recs, &n_recs);
		if (rc == RH_OK) {
			char *paf = rh_format_paf(idx, recs, n_recs, b->names); /* step 2 */
			fputs(paf, stdout);
			rh_free(paf);
			rh_free(recs);
		}
		rh_sigbatch_free(b);
		if (rc != RH_OK) break;
	}
	if (rc != RH_OK) fprintf(stderr, "%s\n", rh_gpu_last_error());
	rh_sigfile_close(f);
	rh_gpu_destroy(ctx);
	rh_index_destroy(idx);
	return rc == RH_OK ? 0 : 1;
}
