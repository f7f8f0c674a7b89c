/*
 * rawhash_b200.h — C-ABI of the B200-native RawHash2 mapping hot path.
 *
 * This header is the drop-in boundary.  Everything here is
 * `extern "C"`, plain pointers and sizes; no torch / CUDA types cross it.
 *
 * What it replaces in the reference (paths relative to the RawHash tree):
 *   - src/rmap.cpp:700   `kt_for(p->n_threads, map_worker_for, in, s->n_sig)`
 *                        -> rh_gpu_map_batch_raw()
 *   - src/rmap.cpp:389   map_worker_for (per-read chunk loop, stop rules, PAF fields)
 *   - src/rsig.c:496-503 raw int16 -> pA conversion + (30,200) outlier drop
 *                        (moves onto the GPU: the batch call takes raw int16)
 *   - src/rindex.c:497   ri_idx_get (khash lookup) -> flattened device index
 *                        built by rh_index_* below
 *   - src/roptions.c:4-138, src/main.cpp:111-210,363-376
 *                        option defaults / presets -> rh_params_*
 *   - src/rmap.cpp:751-772 PAF line formatting -> rh_format_paf()
 *
 * All mapping compute runs in CUDA kernels (sm_100a).  There is no CPU
 * fallback: every rh_gpu_* entry point returns RH_ERR_CUDA when no device
 * is usable.
 */
#ifndef RAWHASH_B200_H
#define RAWHASH_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- flags (values identical to src/roptions.h:6-41) ------------------- */
#define RH_I_SIG_TARGET     0x20      /* RI_I_SIG_TARGET  */
#define RH_M_NO_ADAPTIVE    0x20      /* RI_M_NO_ADAPTIVE */
#define RH_M_ALL_CHAINS     0x2000    /* RI_M_ALL_CHAINS  */

/* ---- error codes -------------------------------------------------------- */
#define RH_OK            0
#define RH_ERR_ARG      -1
#define RH_ERR_IO       -2
#define RH_ERR_CUDA     -3
#define RH_ERR_NOMEM    -4
#define RH_ERR_FORMAT   -5

/*
 * Flat mirror of the fields of ri_idxopt_t / ri_mapopt_t (src/roptions.h:50-143)
 * that the hot path reads.  Same names, same meaning.
 */
typedef struct rh_params_s {
	/* sketch / index (ri_idxopt_t; stored in ri_idx_t, src/rindex.h:29-60) */
	int32_t w, e, n, q, k, idx_flag, lev_col;
	float diff, fine_min, fine_max, fine_range;
	/* event segmentation (--seg-*) */
	uint32_t window_length1, window_length2;
	float threshold1, threshold2, peak_height;
	/* device */
	uint32_t bp_per_sec, sample_rate, chunk_size;
	float sample_per_base;
	/* seeding */
	float mid_occ_frac;
	int32_t min_mid_occ, max_mid_occ, mid_occ;
	/* chaining */
	uint32_t min_events;
	int32_t bw, max_target_gap_length, max_query_gap_length, max_chain_iter;
	int32_t max_num_skips, min_num_anchors, min_chaining_score, min_chaining_score2;
	float chain_gap_scale, chain_skip_scale;
	/* primary/secondary + mapq */
	float mask_level;
	int32_t mask_len;
	float pri_ratio;
	int32_t best_n;
	float alt_drop;
	/* decision */
	float w_bestq, w_bestmq, w_bestmc, w_threshold;
	uint32_t max_num_chunk;
	int32_t min_mapq;
	int64_t map_flag;
} rh_params_t;

/*
 * One output record = one PAF line (mirrors ri_map_t, src/rmap.h:12-22, plus
 * the integer tag fields written at src/rmap.cpp:527-570).  Unmapped reads
 * get exactly one record with mapped == 0 (src/rmap.cpp:521-556).
 */
typedef struct rh_map_rec_s {
	uint32_t read_idx;       /* index of the read in the batch */
	uint32_t c_id;
	uint32_t read_length;
	uint32_t ref_id;
	uint32_t read_start_position;
	uint32_t read_end_position;
	uint32_t fragment_start_position;
	uint32_t fragment_length;
	uint8_t  mapq, rev, mapped, _pad;
	uint32_t ci;             /* chunks consumed  (ci:i) */
	uint32_t sl;             /* filtered signal length (sl:i) */
	int32_t  cm, nc, s1;     /* cm:i nc:i s1:i */
	float    mt_ms;          /* mt:f: the reference prints the wall time map_worker_for spent on the read (rmap.cpp:527).
	                          * Reads of a batch are mapped together here, so a read is charged its share of the batch's
	                          * stream time: batch_ms * (chunks this read consumed) / (chunks all reads consumed).  The shares
	                          * add up to the batch time, and 1000 * bases / mt (test/scripts/pafstats.py:88) stays defined. */
} rh_map_rec_t;

/* ---- parameters / presets ---------------------------------------------- */
/* defaults of ri_idxopt_init + ri_mapopt_init (src/roptions.c:4-138) */
void rh_params_init(rh_params_t *p);
/* presets of ri_set_opt (src/main.cpp:111-210); returns RH_ERR_ARG on unknown name */
int  rh_params_preset(rh_params_t *p, const char *preset);
/* the --r10 macro flag (src/main.cpp:363-376) */
void rh_params_r10(rh_params_t *p);

/* ---- pore model (load_pore, src/rutils.c:133-178) ----------------------- */
/* Parses a k-mer model file; returns malloc'ed z-normalised levels (4^k floats) */
int rh_pore_load(const char *path, int k, int lev_col, float **vals, uint32_t *n_vals);
void rh_free(void *p);

/* ---- index -------------------------------------------------------------- */
typedef struct rh_index_s rh_index_t;   /* flattened host-side index */

/* Build from sequences (semantics of ri_idx_gen, src/rindex.c:900-925:
 * ri_seq_to_sig + ri_sketch on both strands, positions ascending per key). */
rh_index_t *rh_index_build(const rh_params_t *p, const float *pore_vals, uint32_t n_pore_vals,
                           uint32_t n_seq, const char *const *names, const char *const *seqs,
                           const uint32_t *lens, int n_threads);
/* Same index, built on the GPU (expected-signal generation, sketching of both strands,
 * sort, key table).  ACGT-only references and w == 0; other inputs are handed to rh_index_build. */
rh_index_t *rh_index_build_gpu(const rh_params_t *p, const float *pore_vals, uint32_t n_pore_vals,
                               uint32_t n_seq, const char *const *names, const char *const *seqs,
                               const uint32_t *lens, int device);
/* Same index from 2-bit base codes (one per byte, 0..3 = A,C,G,T; sequences back to back) that already lie in the
 * memory of `device`.  The result is device-resident: rh_gpu_init on that device maps from it without a copy, and the
 * host accessors below download it on first use.  (ri_idx_gen, src/rindex.c:900-925; w == 0 only.) */
rh_index_t *rh_index_build_dev(const rh_params_t *p, const float *pore_vals, uint32_t n_pore_vals,
                               uint32_t n_seq, const char *const *names, const void *d_codes,
                               const uint32_t *lens, int device);
/* device the flattened index currently lives on, or -1 */
int         rh_index_on_device(const rh_index_t *idx);
/* host arrays of the flattened index: keys[n_keys] ascending, off[n_keys+1], pos[n_pos] (valid until rh_index_destroy) */
int         rh_index_flat(const rh_index_t *idx, const uint32_t **keys, const uint64_t **off, const uint64_t **pos);
/* Build from raw signals for Rawsamble (semantics of ri_idx_siggen, src/rindex.c:927-969).
 * Event detection runs on the GPU. */
rh_index_t *rh_index_build_sig(const rh_params_t *p, uint32_t n_reads, const char *const *names,
                               const int16_t *const *raw, const uint64_t *raw_len,
                               const double *offset, const double *range, const double *digitisation);
/* Load a reference-format `.ind` file (reader of src/rindex.c:650-776). */
rh_index_t *rh_index_load(const char *path, rh_params_t *p_inout);
void        rh_index_destroy(rh_index_t *idx);
uint32_t    rh_index_n_seq(const rh_index_t *idx);
const char *rh_index_seq_name(const rh_index_t *idx, uint32_t i);
uint32_t    rh_index_seq_len(const rh_index_t *idx, uint32_t i);
uint64_t    rh_index_n_keys(const rh_index_t *idx);
uint64_t    rh_index_n_pos(const rh_index_t *idx);
uint32_t    rh_index_key(const rh_index_t *idx, uint64_t i);   /* i-th distinct hash, ascending */
/* ri_idx_cal_max_occ + ri_mapopt_update (src/rindex.c:1018-1053): sets p->mid_occ */
void        rh_index_update_mapopt(const rh_index_t *idx, rh_params_t *p);
/* ri_idx_get (src/rindex.c:497-514): returns pointer to ascending position list */
const uint64_t *rh_index_get(const rh_index_t *idx, uint32_t hash, int *n);

/* ---- GPU context --------------------------------------------------------- */
typedef struct rh_gpu_ctx_s rh_gpu_ctx;

/* Uploads the index to device `device` and allocates work arenas.
 * arena_bytes == 0 picks a default from free device memory. */
rh_gpu_ctx *rh_gpu_init(const rh_index_t *idx, const rh_params_t *p, int device, size_t arena_bytes);
void        rh_gpu_destroy(rh_gpu_ctx *ctx);
const char *rh_gpu_last_error(void);
/* Launch on the caller's CUDA stream (a cudaStream_t passed as void*; NULL = the context's own). */
void        rh_gpu_set_stream(rh_gpu_ctx *ctx, void *cuda_stream);
/* A batch is cut into contiguous read ranges that run concurrently on their own CUDA streams (the
 * multi-threaded side of kt_for, src/kthread.c:47-65): the slow tail of one range overlaps the bulk of
 * the others and the host->device copy overlaps compute.  The number of workers is fixed at rh_gpu_init
 * (env RH_WORKERS, default 1: on B200 the kernels of one range already fill the GPU, and splitting the
 * work arena costs more than the overlap wins — measured 118.6k / 115.0k / 113.5k reads/s at 1 / 2 / 4);
 * this call lowers or restores the count in use and returns it.  With a caller stream set, one range is used. */
int         rh_gpu_set_workers(rh_gpu_ctx *ctx, int n_workers);

/*
 * The drop-in for `kt_for(n_threads, map_worker_for, step, n_sig)` with the
 * signal payload moved one step earlier (raw int16 + calibration instead of
 * pA floats).  For each read i in [0,n): raw[i] points to raw_len[i] samples
 * in HOST memory.  Output: *recs is a malloc'ed array of *n_recs records in
 * read order (free with rh_free).  Returns RH_OK or an error code.
 */
int rh_gpu_map_batch_raw(rh_gpu_ctx *ctx, uint32_t n,
                         const int16_t *const *raw, const uint64_t *raw_len,
                         const double *offset, const double *range, const double *digitisation,
                         const char *const *names,
                         rh_map_rec_t **recs, uint64_t *n_recs);

/*
 * Same computation with the raw samples already resident in device memory as
 * one concatenated int16 buffer (d_raw, device pointer) with per-read start
 * offsets raw_off[n+1] (host).  Used by bench.py for the HBM-resident number.
 */
int rh_gpu_map_batch_dev(rh_gpu_ctx *ctx, uint32_t n,
                         const void *d_raw, const uint64_t *raw_off,
                         const double *offset, const double *range, const double *digitisation,
                         const char *const *names,
                         rh_map_rec_t **recs, uint64_t *n_recs);

/* Multi-GPU sharding of a mini-batch: boundaries cut[0..parts] of contiguous read ranges with nearly
 * equal raw sample counts; range r = reads [cut[r], cut[r+1]).  Order preserving, every read in exactly one range.
 * The same rule cuts a batch into worker ranges inside one context. */
void rh_split_by_samples(uint32_t n, const uint64_t *raw_len, uint32_t parts, uint32_t *cut);

/* Per-stage statistics of the last rh_gpu_map_batch_* call. */
typedef struct rh_gpu_stats_s {
	uint64_t n_reads, n_chunks, n_rounds;
	uint64_t raw_samples_consumed;   /* int16 samples read by the event kernel   */
	uint64_t n_events, n_seeds, n_anchors, n_chains;
	uint64_t kernel_launches;        /* launches of our own kernels              */
	double   ms_total;               /* CUDA-event time of the whole call        */
	double   ms_event_kernel;        /* signal->seeds kernel(s) only             */
	double   ms_seed, ms_sort, ms_chain, ms_post;
	uint64_t event_kernel_launches;
	uint64_t h2d_bytes, d2h_bytes;
	double   ms_sort_ties;           /* part of ms_sort spent replaying klib's tie order */
	uint64_t event_stage_samples;    /* raw samples / seeds the event-stage launches produced, including chunks computed */
	uint64_t event_stage_seeds;      /* ahead of the read's stop decision and never mapped                              */
} rh_gpu_stats_t;
void rh_gpu_get_stats(const rh_gpu_ctx *ctx, rh_gpu_stats_t *st);

/*
 * Stage tap (parity tests only): map ONE read through every chunk up to
 * max_num_chunk WITHOUT the stop rules and copy each stage's device output
 * back.  Arrays are concatenated over chunks; cnt[c*RH_TAP_NCNT + j] holds
 * the per-chunk counts.  Capacities are in elements; returns RH_ERR_NOMEM if
 * a buffer is too small.
 */
#define RH_TAP_NCNT 8
enum { RH_TAP_NSIG = 0, RH_TAP_NEVENTS, RH_TAP_NSEEDS, RH_TAP_NANCHORS, RH_TAP_NU, RH_TAP_NV, RH_TAP_NREGS, RH_TAP_REPLEN };
#define RH_TAP_REG_NF 14  /* score cnt rid rev qs qe rs re parent subsc n_sub mapq as score0 */
typedef struct rh_tap_s {
	uint32_t n_chunks;
	int32_t  *cnt;      uint64_t cap_chunks;
	float    *events;   uint64_t cap_events;
	uint64_t *seeds;    uint64_t cap_seeds;    /* x,y pairs */
	uint64_t *anchors;  uint64_t cap_anchors;  /* x,y pairs, sorted list fed to chaining */
	uint64_t *u;        uint64_t cap_u;
	uint64_t *chain_a;  uint64_t cap_chain_a;  /* x,y pairs, compacted target-sorted anchors */
	uint64_t *prev_a;   uint64_t cap_prev_a;   /* x,y pairs, backtrack-order copy carried to next chunk */
	int32_t  *regs;     uint64_t cap_regs;     /* RH_TAP_REG_NF ints per region */
} rh_tap_t;
int rh_gpu_tap_read(rh_gpu_ctx *ctx, const int16_t *raw, uint64_t raw_len,
                    double offset, double range, double digitisation,
                    const char *name, rh_tap_t *tap);

/* How the library lays one chunk round out in its anchor arena (no device needed; the streaming scheduler calls the same
 * code): n_anchors[n_chunks] = anchors per chunk, the first n_mandatory of which must run, the rest being first chunks of
 * waiting reads in admission order (at most max_optional admitted).  Out: order[q] = input chunk at slot q (mandatory
 * chunks heaviest first), a_off[q] = byte offset of slot q's region, groups = (first slot, count, heavy lane?) triples —
 * the heavy group (largest chunks, second stream, top of the arena) first if there is one —, n_run = slots that run,
 * main_bytes = arena bytes of the ordinary groups.  RH_ERR_NOMEM if one chunk exceeds the arena. */
int rh_plan_round(const uint32_t *n_anchors, uint32_t n_chunks, uint32_t n_mandatory, uint32_t max_optional, uint64_t arena_bytes, int heavy_lane,
                  uint32_t *order, uint64_t *a_off, uint32_t *groups, uint32_t groups_cap, uint32_t *n_groups, uint32_t *n_run, uint64_t *main_bytes);

/* ---- files either side of the path (SURVEY.md §8b "file surfaces kept", §8f rank 3) -------------- */
/* Write the index in the reference's `.ind` layout (ri_idx_dump, src/rindex.c:545-648): header, pore table,
 * sequence names/lengths, 2^14 buckets of (position array, key/value pairs).  The reference's ri_idx_load
 * (src/rindex.c:650-776) and rh_index_load read it back.  pore_vals (z-normalised, as rh_pore_load returns
 * them) may be NULL for a signal index. */
int rh_index_dump(const rh_index_t *idx, const char *path, const float *pore_vals, uint32_t n_pore_vals);

/* FASTA / gzipped FASTA reader (mm_bseq_open/mm_bseq_read, src/bseq.c:28-110: name = text up to the first
 * white space).  Arrays are malloc'ed; release with rh_fasta_free. */
int  rh_fasta_load(const char *path, uint32_t *n_seq, char ***names, char ***seqs, uint32_t **lens);
void rh_fasta_free(uint32_t n_seq, char **names, char **seqs, uint32_t *lens);

/* SLOW5 / BLOW5 signal files (ri_sig_open_slow5 + ri_read_sig_slow5, src/rsig.c:170-207,478-533; slow5lib
 * 0.2.0 file format: ASCII, or binary with record compression none|zlib|zstd and signal compression none|svb-zd;
 * zstd is bound from libzstd.so.1 at run time and refused by name when that library is absent).
 * A batch keeps the raw int16 samples of its reads back to back in ONE arena (page-locked when a CUDA device
 * is usable) so that rh_gpu_map_batch_raw uploads it in a single copy; the pA conversion stays on the GPU. */
typedef struct rh_sigfile_s rh_sigfile_t;
typedef struct rh_sigbatch_s {
	uint32_t n;                       /* reads in the batch                          */
	uint64_t n_samples;               /* sum of raw_len                              */
	const int16_t *const *raw;        /* raw[i] -> raw_len[i] samples in the arena   */
	const uint64_t *raw_len;
	const double *offset, *range, *digitisation, *sampling_rate;
	const char *const *names;         /* read ids                                    */
	int32_t arena_pinned;             /* 1 if the arena is page-locked (cudaHostAlloc) */
	void *priv;
} rh_sigbatch_t;
rh_sigfile_t *rh_sigfile_open(const char *path, int n_threads);
void          rh_sigfile_close(rh_sigfile_t *f);
/* Reads records until the batch holds >= max_samples raw samples (the -K mini-batch rule of
 * ri_sig_read_frag, src/rmap.cpp:600-660, counted on raw samples) or max_reads reads (0 = no limit) or the
 * file ends.  *out == NULL with RH_OK at end of file. */
int  rh_sigfile_next_batch(rh_sigfile_t *f, uint64_t max_samples, uint32_t max_reads, rh_sigbatch_t **out);
void rh_sigbatch_free(rh_sigbatch_t *b);
/* One zlib stream -> malloc'ed bytes (rh_free): the record decompression of slow5lib's ptr_depress_zlib_solo
 * (extern/slow5lib/src/slow5_press.c:945-982) through this library's own inflate (csrc/rh_inflate.h, Adler-32 checked). */
int  rh_zlib_inflate(const void *in, size_t in_bytes, void **out, size_t *out_bytes);
/* find_sfiles (src/rsig.c:286-330): `path` itself or, for a directory, every *.slow5 / *.blow5 below it.
 * Returns a malloc'ed array of malloc'ed strings (rh_free each, then the array). */
int  rh_find_sigfiles(const char *path, char ***files, uint32_t *n_files);
/* Writer used by the synthetic-data generator, bench and tests: `.slow5` (ASCII) or `.blow5`
 * (record_press 0 none | 1 zlib | 2 zstd, signal_press 0 none | 1 svb-zd); slow5lib reads the result. */
int  rh_slow5_write(const char *path, uint32_t n, const char *const *names,
                    const int16_t *const *raw, const uint64_t *raw_len,
                    const double *offset, const double *range, const double *digitisation, double sampling_rate,
                    int record_press, int signal_press);

/* ---- PAF ------------------------------------------------------------------ */
/* Formats records exactly like src/rmap.cpp:751-772 (mt:f: carries rh_map_rec_t::mt_ms).
 * Returns a malloc'ed NUL-terminated string (free with rh_free). */
char *rh_format_paf(const rh_index_t *idx, const rh_map_rec_t *recs, uint64_t n_recs,
                    const char *const *names);

#ifdef __cplusplus
}
#endif
#endif /* RAWHASH_B200_H */
