/*
 * rh_main.cpp — `rawhash2_b200`: the rawhash2 command line (reference src/main.cpp) on top of the C-ABI library.
 *
 *   rawhash2_b200 [options] <target.fa>|<target.ind> [query.blow5|dir ...]
 *
 * Same option names, presets and order of application as src/main.cpp:256-418 (-x first, then everything else), same
 * three-step pipeline as map_worker_pipeline (src/rmap.cpp:662-800): step 0 reads a mini-batch of signals (here: raw
 * int16 straight out of SLOW5/BLOW5 into a page-locked arena), step 1 maps it (here: ONE call of rh_gpu_map_batch_raw
 * per GPU instead of kt_for over reads), step 2 prints PAF in input order — the three steps of consecutive batches
 * overlap on their own threads.  Everything that computes lives in librawhash_b200.so; this file only parses options
 * and moves batches.  Options that select reference code outside the mapping path this library implements (RMQ
 * chaining, DTW, Sequence-Until, --store-sig ...) are refused by name instead of being silently ignored.
 *
 * Extra options (not in the reference): --gpus INT (GPUs of this node to spread each mini-batch over, default 1),
 * --device INT (first device), --index-on-host (build a FASTA index with the threaded host builder).
 */
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rawhash_b200.h"

#define RH_VERSION "2.1-b200"

namespace {

double now_s() { struct timeval tv; gettimeofday(&tv, NULL); return tv.tv_sec + tv.tv_usec * 1e-6; }
double g_t0;

struct long_opt { const char *name; int has_arg; int id; };
/* ids as in src/main.cpp:11-86 */
const long_opt LONG_OPTS[] = {
	{"level_column", 1, 300}, {"q-mid-occ", 1, 301}, {"mid_occ_frac", 1, 302}, {"min-events", 1, 303}, {"bw", 1, 304},
	{"max-target-gap", 1, 305}, {"max-query-gap", 1, 306}, {"min-anchors", 1, 307}, {"min-score", 1, 308},
	{"chain-gap-scale", 1, 309}, {"chain-skip-scale", 1, 310}, {"best-chains", 1, 311}, {"primary-ratio", 1, 312},
	{"primary-length", 1, 313}, {"max-skips", 1, 314}, {"max-iterations", 1, 315}, {"rmq", 0, 316},
	{"rmq-inner-dist", 1, 317}, {"rmq-size-cap", 1, 318}, {"bw-long", 1, 319}, {"max-chunks", 1, 320},
	{"min-mapq", 1, 321}, {"alt-drop", 1, 322}, {"w-besta", 1, 323}, {"w-bestma", 1, 324}, {"w-bestq", 1, 325},
	{"w-bestmq", 1, 326}, {"w-bestmc", 1, 327}, {"w-threshold", 1, 328}, {"bp-per-sec", 1, 329}, {"sample-rate", 1, 330},
	{"chunk-size", 1, 331}, {"seg-window-length1", 1, 332}, {"seg-window-length2", 1, 333}, {"seg-threshold1", 1, 334},
	{"seg-threshold2", 1, 335}, {"seg-peak-height", 1, 336}, {"sequence-until", 0, 337}, {"threshold", 1, 338},
	{"n-samples", 1, 339}, {"test-frequency", 1, 340}, {"min-reads", 1, 341}, {"occ-frac", 1, 342}, {"depletion", 0, 343},
	{"store-sig", 0, 344}, {"sig-target", 0, 345}, {"disable-adaptive", 0, 346}, {"sig-diff", 1, 347}, {"align", 0, 348},
	{"dtw-evaluate-chains", 0, 349}, {"dtw-output-cigar", 0, 350}, {"dtw-border-constraint", 1, 351}, {"dtw-log-scores", 0, 352},
	{"no-chainingscore-filtering", 0, 353}, {"dtw-match-bonus", 1, 354}, {"output-chains", 0, 355}, {"dtw-fill-method", 1, 356},
	{"dtw-min-score", 1, 357}, {"log-anchors", 0, 358}, {"log-num-anchors", 0, 359}, {"rev-collision-count", 1, 360},
	{"chn-rev-bump", 1, 361}, {"no-rev-target", 0, 362}, {"r10", 0, 363}, {"fine-min", 1, 364}, {"fine-max", 1, 365},
	{"fine-range", 1, 366}, {"out-quantize", 0, 367}, {"no-event-detection", 0, 368}, {"io-thread", 1, 369},
	{"min-score2", 1, 370}, {"version", 0, 371},
	/* not in the reference */
	{"gpus", 1, 900}, {"device", 1, 901}, {"index-on-host", 0, 902},
	{0, 0, 0}
};
const char *SHORT_WITH_ARG = "kdpeqwnotKx";

struct parsed_opt { int id; const char *arg; const char *spelled; };

/* options may be mixed with the positional arguments (ketopt is called with permute = 1, src/main.cpp:275) */
int split_args(int argc, char **argv, std::vector<parsed_opt> &opts, std::vector<const char *> &pos)
{
	for (int i = 1; i < argc; ++i) {
		const char *a = argv[i];
		if (a[0] != '-' || a[1] == 0) { pos.push_back(a); continue; }
		if (a[1] == '-' && a[2] == 0) { for (++i; i < argc; ++i) pos.push_back(argv[i]); break; }
		if (a[1] == '-') {
			const char *eq = strchr(a + 2, '=');
			const std::string name = eq ? std::string(a + 2, eq - a - 2) : std::string(a + 2);
			const long_opt *lo = LONG_OPTS;
			while (lo->name && name != lo->name) ++lo;
			if (!lo->name) { fprintf(stderr, "[ERROR] unknown option in \"%s\"\n", a); return -1; }
			const char *arg = NULL;
			if (lo->has_arg) {
				if (eq) arg = eq + 1;
				else if (i + 1 < argc) arg = argv[++i];
				else { fprintf(stderr, "[ERROR] missing option argument\n"); return -1; }
			}
			opts.push_back({lo->id, arg, a});
			continue;
		}
		for (const char *c = a + 1; *c; ++c) { /* bundled short options; an argument ends the bundle */
			if (*c == 'h') { opts.push_back({'h', NULL, a}); continue; }
			if (!strchr(SHORT_WITH_ARG, *c)) { fprintf(stderr, "[ERROR] unknown option in \"%s\"\n", a); return -1; }
			const char *arg = c[1] ? c + 1 : (i + 1 < argc ? argv[++i] : NULL);
			if (!arg) { fprintf(stderr, "[ERROR] missing option argument\n"); return -1; }
			opts.push_back({*c, arg, a});
			break;
		}
	}
	return 0;
}

int64_t parse_num(const char *s)
{ /* mm_parse_num, src/main.cpp:88-97 */
	char *p; double x = strtod(s, &p);
	if (*p == 'G' || *p == 'g') x *= 1e9; else if (*p == 'M' || *p == 'm') x *= 1e6; else if (*p == 'K' || *p == 'k') x *= 1e3;
	return (int64_t)(x + .499);
}

struct settings {
	rh_params_t P;
	const char *dump = NULL, *pore = NULL;
	/* -t: the reference defaults to 3 mapping threads; here -t only sizes the host side (signal decoding, host index
	 * builder), which has to keep up with a GPU, so the default is every hardware thread */
	int n_threads = (int)std::max(3u, std::thread::hardware_concurrency()), io_threads = 1, n_gpus = 1, device = 0;
	bool index_on_host = false, help = false;
	int64_t mini_batch = 500000000; /* src/roptions.c:88 */
};

int unsupported(const parsed_opt &o, const char *why)
{
	fprintf(stderr, "[ERROR] option '%s' selects %s, which is outside the mapping path rawhash2_b200 implements\n", o.spelled, why);
	return -1;
}

int apply_options(const std::vector<parsed_opt> &opts, settings &S)
{
	rh_params_t &P = S.P;
	rh_params_init(&P);
	for (const parsed_opt &o : opts) /* presets are applied before every other option (src/main.cpp:274-290) */
		if (o.id == 'x' && rh_params_preset(&P, o.arg) != RH_OK) { fprintf(stderr, "[ERROR] unknown preset '%s'\n", o.arg); return -1; }
	for (const parsed_opt &o : opts) {
		const char *a = o.arg; char *s;
		switch (o.id) {
		case 'x': break;
		case 'd': S.dump = a; break;
		case 'p': S.pore = a; break;
		case 'k': P.k = atoi(a); break;
		case 'e': P.e = atoi(a); break;
		case 'q': P.q = atoi(a); break;
		case 'w': P.w = atoi(a); break;
		case 'n': P.n = atoi(a); break;
		case 't': S.n_threads = atoi(a); break;
		case 'K': S.mini_batch = parse_num(a); break;
		case 'h': S.help = true; break;
		case 'o':
			if (strcmp(a, "-") != 0 && freopen(a, "wb", stdout) == NULL) { fprintf(stderr, "[ERROR] failed to write the output to file '%s': %s\n", a, strerror(errno)); return -1; }
			break;
		case 300: P.lev_col = atoi(a); break;
		case 301: P.min_mid_occ = (int32_t)strtol(a, &s, 10); if (*s == ',') P.max_mid_occ = (int32_t)strtol(s + 1, &s, 10); break;
		case 302: case 342: P.mid_occ_frac = (float)atof(a); break;
		case 303: P.min_events = (uint32_t)atoi(a); break;
		case 304: P.bw = atoi(a); break;
		case 305: P.max_target_gap_length = atoi(a); break;
		case 306: P.max_query_gap_length = atoi(a); break;
		case 307: P.min_num_anchors = atoi(a); break;
		case 308: P.min_chaining_score = atoi(a); break;
		case 309: P.chain_gap_scale = (float)atof(a); break;
		case 310: P.chain_skip_scale = (float)atof(a); break;
		case 311: P.best_n = atoi(a); break;
		case 312: P.mask_level = (float)atof(a); break;
		case 313: P.mask_len = atoi(a); break;
		case 314: P.max_num_skips = atoi(a); break;
		case 315: P.max_chain_iter = atoi(a); break;
		case 316: case 317: case 318: case 319: return unsupported(o, "RMQ chaining / long-gap re-chaining (src/lchain.c:606-756)");
		case 320: P.max_num_chunk = (uint32_t)atoi(a); break;
		case 321: P.min_mapq = atoi(a); break;
		case 322: P.alt_drop = (float)atof(a); break;
		case 323: case 324: break; /* --w-besta / --w-bestma only weigh DTW scores (src/rmap.cpp:474-481) */
		case 325: P.w_bestq = (float)atof(a); break;
		case 326: P.w_bestmq = (float)atof(a); break;
		case 327: P.w_bestmc = (float)atof(a); break;
		case 328: P.w_threshold = (float)atof(a); break;
		case 329: P.bp_per_sec = (uint32_t)atoi(a); P.sample_per_base = (float)P.sample_rate / P.bp_per_sec; break;
		case 330: P.sample_rate = (uint32_t)atoi(a); P.sample_per_base = (float)P.sample_rate / P.bp_per_sec; break;
		case 331: P.chunk_size = (uint32_t)atoi(a); break;
		case 332: P.window_length1 = (uint32_t)atoi(a); break;
		case 333: P.window_length2 = (uint32_t)atoi(a); break;
		case 334: P.threshold1 = (float)atof(a); break;
		case 335: P.threshold2 = (float)atof(a); break;
		case 336: P.peak_height = (float)atof(a); break;
		case 337: case 338: case 339: case 340: case 341: return unsupported(o, "Sequence-Until (src/sequence_until.c)");
		case 343: /* --depletion, src/main.cpp:357-360 */
			P.best_n = 5; P.min_mapq = 10; P.w_threshold = 0.50f; P.min_num_anchors = 2; P.min_chaining_score = 15; P.chain_skip_scale = 0.0f;
			break;
		case 344: return unsupported(o, "stored target signals for DTW");
		case 345: P.idx_flag |= RH_I_SIG_TARGET; break;
		case 346: P.map_flag |= RH_M_NO_ADAPTIVE; break;
		case 347: P.diff = (float)atof(a); break;
		case 348: case 349: case 350: case 351: case 352: case 353: case 354: case 356: case 357: return unsupported(o, "DTW re-scoring (src/dtw.cpp)");
		case 355: case 358: case 359: return unsupported(o, "the reference's debugging output");
		case 360: case 361: break; /* parsed and never read by the reference either */
		case 362: return unsupported(o, "the experimental forward-only index");
		case 363: rh_params_r10(&P); break;
		case 364: P.fine_min = (float)atof(a); break;
		case 365: P.fine_max = (float)atof(a); break;
		case 366: P.fine_range = (float)atof(a); break;
		case 367: case 368: return unsupported(o, "an experimental mode without mapping / without event detection");
		case 369: S.io_threads = atoi(a); break;
		case 370: P.min_chaining_score2 = atoi(a); break;
		case 371: puts(RH_VERSION); exit(0);
		case 900: S.n_gpus = atoi(a); break;
		case 901: S.device = atoi(a); break;
		case 902: S.index_on_host = true; break;
		default: break;
		}
	}
	return 0;
}

void usage(FILE *fp, const settings &S)
{
	const rh_params_t &P = S.P;
	fprintf(fp, "Usage: rawhash2_b200 [options] <target.fa>|<target.ind> [query.blow5|query.slow5|dir] [...]\n");
	fprintf(fp, "Options (names and defaults of rawhash2 %s; mapping runs on the GPU):\n", "2.1");
	fprintf(fp, "    --version     show version number\n");
	fprintf(fp, "  K-mer (pore) Model:\n    -p FILE      pore model FILE\n    -k INT       k-mer size of the pore model [%d]\n    --level_column INT   0-based column of the level mean [%d]\n", P.k, P.lev_col);
	fprintf(fp, "  Indexing:\n    -d FILE      dump index to FILE (reference `.ind` format)\n    -e INT       events per hash [%d]\n    -q INT       quantisation bits [%d]\n    -w INT       minimizer window [%d]\n", P.e, P.q, P.w);
	fprintf(fp, "    --sig-target    the target is a set of signal files (Rawsamble)\n    --sig-diff FLOAT   minimum difference between consecutive kept events [%g]\n", P.diff);
	fprintf(fp, "    --fine-min/--fine-max/--fine-range FLOAT   dynamic quantisation [%g %g %g]\n", P.fine_min, P.fine_max, P.fine_range);
	fprintf(fp, "  Seeding:\n    --q-mid-occ INT1[,INT2]   bounds of the occurrence threshold [%d, %d]\n    --occ-frac FLOAT   top fraction of repetitive seeds dropped [%g]\n", P.min_mid_occ, P.max_mid_occ, P.mid_occ_frac);
	fprintf(fp, "  Chaining:\n    --min-events INT [%u]   --bw INT [%d]   --max-target-gap INT [%d]   --max-query-gap INT [%d]\n", P.min_events, P.bw, P.max_target_gap_length, P.max_query_gap_length);
	fprintf(fp, "    --min-anchors INT [%d]   --best-chains INT [%d]   --min-score INT [%d]   --min-score2 INT [%d]\n", P.min_num_anchors, P.best_n, P.min_chaining_score, P.min_chaining_score2);
	fprintf(fp, "    --chain-gap-scale FLOAT [%g]   --chain-skip-scale FLOAT [%g]   --max-skips INT [%d]   --max-iterations INT [%d]\n", P.chain_gap_scale, P.chain_skip_scale, P.max_num_skips, P.max_chain_iter);
	fprintf(fp, "    --primary-ratio FLOAT [%g]   --primary-length INT [%d]   --alt-drop FLOAT [%g]\n", P.mask_level, P.mask_len, P.alt_drop);
	fprintf(fp, "  Mapping decisions:\n    --max-chunks INT [%u]   --min-mapq INT [%d]   --disable-adaptive\n    --w-bestq [%g] --w-bestmq [%g] --w-bestmc [%g] --w-threshold [%g]\n", P.max_num_chunk, P.min_mapq, P.w_bestq, P.w_bestmq, P.w_bestmc, P.w_threshold);
	fprintf(fp, "  Nanopore:\n    --bp-per-sec INT [%u]   --sample-rate INT [%u]   --chunk-size INT [%u]\n", P.bp_per_sec, P.sample_rate, P.chunk_size);
	fprintf(fp, "    --seg-window-length1 INT [%u]  --seg-window-length2 INT [%u]  --seg-threshold1 FLOAT [%g]  --seg-threshold2 FLOAT [%g]  --seg-peak-height FLOAT [%g]\n", P.window_length1, P.window_length2, P.threshold1, P.threshold2, P.peak_height);
	fprintf(fp, "  Input/Output:\n    -o FILE      output mappings to FILE [stdout]\n    -t INT       host threads for index construction and signal decoding [%d]\n    --io-thread INT   of which for reading signal files [%d]\n    -K NUM       mini-batch size in samples [500M]\n", S.n_threads, S.io_threads);
	fprintf(fp, "    --gpus INT   GPUs of this node to spread each mini-batch over [1]   --device INT   first device [0]\n    --index-on-host   build a FASTA index with the host builder instead of the GPU builder\n");
	fprintf(fp, "  Presets:\n    --depletion   --r10   -x STR  (viral, sensitive, fast, faster; Rawsamble: ava, ava-sensitive, ava-viral, ava-large)\n");
	fprintf(fp, "  Refused (reference code outside the GPU mapping path): --rmq* --bw-long --dtw-* --align --store-sig --sequence-until ... --no-rev-target --out-quantize --no-event-detection\n");
}

bool is_index_file(const char *path)
{ /* ri_idx_is_idx, src/rindex.c:994-1016 */
	FILE *f = fopen(path, "rb");
	if (!f) return false;
	char m[2]; const bool yes = fread(m, 1, 2, f) == 2 && m[0] == 'R' && m[1] == 'I';
	fclose(f);
	return yes;
}

/* ---- a bounded hand-over queue between pipeline steps (kt_pipeline keeps at most n_threads steps in flight) ----- */
template <class T> struct handoff {
	std::mutex mu; std::condition_variable cv; std::deque<T> q; bool done = false; size_t depth;
	explicit handoff(size_t d) : depth(d) {}
	void push(T v) { std::unique_lock<std::mutex> l(mu); cv.wait(l, [&] { return q.size() < depth; }); q.push_back(v); cv.notify_all(); }
	bool pop(T &v) { std::unique_lock<std::mutex> l(mu); cv.wait(l, [&] { return !q.empty() || done; }); if (q.empty()) return false; v = q.front(); q.pop_front(); cv.notify_all(); return true; }
	void close() { std::lock_guard<std::mutex> l(mu); done = true; cv.notify_all(); }
};

struct mapped_batch { rh_sigbatch_t *in; rh_map_rec_t *recs; uint64_t n_recs; };

/* step 1 on several GPUs: contiguous read ranges balanced by sample count, one host thread per GPU, records
 * concatenated in range order = input order */
int map_on_gpus(std::vector<rh_gpu_ctx *> &ctx, const rh_sigbatch_t *b, rh_map_rec_t **recs, uint64_t *n_recs)
{
	const size_t G = ctx.size();
	if (G == 1) return rh_gpu_map_batch_raw(ctx[0], b->n, b->raw, b->raw_len, b->offset, b->range, b->digitisation, b->names, recs, n_recs);
	std::vector<uint32_t> cut(G + 1, b->n);
	rh_split_by_samples(b->n, b->raw_len, (uint32_t)G, cut.data());
	std::vector<rh_map_rec_t *> part(G, nullptr); std::vector<uint64_t> pn(G, 0); std::vector<int> rc(G, RH_OK); std::vector<std::string> err(G);
	std::vector<std::thread> th;
	for (size_t g = 0; g < G; ++g) th.emplace_back([&, g]() {
		const uint32_t lo = cut[g], n = cut[g + 1] - cut[g];
		if (!n) return;
		rc[g] = rh_gpu_map_batch_raw(ctx[g], n, b->raw + lo, b->raw_len + lo, b->offset + lo, b->range + lo, b->digitisation + lo, b->names + lo, &part[g], &pn[g]);
		if (rc[g] != RH_OK) err[g] = rh_gpu_last_error();
	});
	for (auto &t : th) t.join();
	uint64_t total = 0; int bad = RH_OK;
	for (size_t g = 0; g < G; ++g) { total += pn[g]; if (rc[g] != RH_OK && bad == RH_OK) { bad = rc[g]; fprintf(stderr, "[ERROR] GPU %zu: %s\n", g, err[g].c_str()); } }
	rh_map_rec_t *all = (rh_map_rec_t *)malloc(sizeof(rh_map_rec_t) * (total ? total : 1));
	uint64_t at = 0;
	for (size_t g = 0; g < G; ++g) {
		for (uint64_t i = 0; i < pn[g]; ++i) { all[at] = part[g][i]; all[at].read_idx += cut[g]; ++at; }
		rh_free(part[g]);
	}
	*recs = all; *n_recs = total;
	return bad;
}

/* One small batch of noise through every GPU while step 0 decodes the first real mini-batch: first-use costs of
 * the CUDA runtime (scratch allocations, stream/event pools) are paid here instead of inside the first batch. */
void warm_up(std::vector<rh_gpu_ctx *> &ctx)
{
	const uint32_t n = 256, len = 9000;
	std::vector<int16_t> raw((size_t)n * len);
	uint32_t x = 12345;
	for (size_t i = 0; i < raw.size(); ++i) {
		x = x * 1664525u + 1013904223u;
		if (i % 9 == 0 || i % len == 0) raw[i] = (int16_t)(350 + (x >> 16) % 350); /* a new level about every 9 samples */
		else raw[i] = (int16_t)(raw[i - 1] + (int)((x >> 24) % 9) - 4);
	}
	std::vector<const int16_t *> ptr(n); std::vector<uint64_t> l(n, len); std::vector<double> off(n, 10.0), rng(n, 1400.0), dig(n, 8192.0);
	std::vector<std::string> nm(n); std::vector<const char *> names(n);
	for (uint32_t i = 0; i < n; ++i) { ptr[i] = raw.data() + (size_t)i * len; nm[i] = "warmup_" + std::to_string(i); names[i] = nm[i].c_str(); }
	for (rh_gpu_ctx *c : ctx) {
		rh_map_rec_t *recs = NULL; uint64_t nr = 0;
		rh_gpu_map_batch_raw(c, n, ptr.data(), l.data(), off.data(), rng.data(), dig.data(), names.data(), &recs, &nr);
		rh_free(recs);
	}
}

} // namespace

int main(int argc, char **argv)
{
	g_t0 = now_s();
	std::vector<parsed_opt> opts; std::vector<const char *> pos;
	settings S;
	if (split_args(argc, argv, opts, pos) < 0) return 1;
	if (apply_options(opts, S) < 0) return 1;
	rh_params_t &P = S.P;
	if (pos.empty() || S.help) { usage(S.help ? stdout : stderr, S); return S.help ? 0 : 1; }
	if (S.n_threads < S.io_threads) { fprintf(stderr, "[ERROR] The overall number of threads (-t [%d]) must NOT be smaller than the number of IO threads (--io-thread [%d]).\n", S.n_threads, S.io_threads); return 1; }
	if (P.w && P.n) { fprintf(stderr, "[ERROR] minimizer window 'w' ('%d') and BLEND 'neighbor' ('%d') values cannot be set together.\n", P.w, P.n); return 1; }
	if (P.n) { fprintf(stderr, "[ERROR] BLEND seeding (-n) is disabled in the reference and not implemented here\n"); return 1; }
	if (S.n_gpus < 1) S.n_gpus = 1;

	const char *target = pos[0];
	const bool sig_target = (P.idx_flag & RH_I_SIG_TARGET) != 0;
	const bool target_is_idx = is_index_file(target);
	if (!target_is_idx && strcmp(target, "-") != 0) { FILE *f = fopen(target, "rb"); if (!f) { fprintf(stderr, "[ERROR] failed to open file '%s': %s\n", target, strerror(errno)); return 1; } fclose(f); }
	if (!target_is_idx && !S.dump && pos.size() < 2) {
		fprintf(stderr, "[ERROR] missing input: please specify a query SLOW5/BLOW5 file(s) to map or option -d to store the index in a file before running the mapping\n");
		return 1;
	}
	if (!target_is_idx && !S.pore && !sig_target) {
		fprintf(stderr, "[ERROR] missing input: please specify a pore model file with -p when generating the index from a sequence file\n");
		return 1;
	}

	/* ---- pipeline state; step 0 starts right away so that decoding the first mini-batches (and page-locking their
	 * arenas) overlaps index loading and GPU initialisation ------------------------------------------------------ */
	const bool have_queries = pos.size() >= 2;
	handoff<rh_sigbatch_t *> to_map(2);
	handoff<mapped_batch> to_print(2);
	std::atomic<int> status(0); std::mutex status_mu;
	auto fail = [&](const std::string &msg) { std::lock_guard<std::mutex> l(status_mu); if (!status) fprintf(stderr, "[ERROR] %s\n", msg.c_str()); status = 1; };
	uint64_t n_reads = 0, n_mapped = 0, n_samples = 0;
	double t_map = 0, t_read = 0;
	uint64_t n_batches = 0;
	const double t_pipe0 = now_s();
	const int decode_threads = std::max(1, S.n_threads - 2); /* one thread drives the GPU, one prints */
	std::thread reader;
	if (have_queries) reader = std::thread([&]() { /* step 0: ri_sig_read_frag, src/rmap.cpp:600-660 */
		for (size_t q = 1; q < pos.size() && !status; ++q) {
			char **files = NULL; uint32_t nf = 0;
			rh_find_sigfiles(pos[q], &files, &nf);
			if (nf == 0) fail(std::string("failed to open file '") + pos[q] + "': no .slow5/.blow5 signal file");
			for (uint32_t i = 0; i < nf && !status; ++i) {
				rh_sigfile_t *f = rh_sigfile_open(files[i], decode_threads);
				if (!f) { fail(rh_gpu_last_error()); break; }
				for (;;) {
					rh_sigbatch_t *b = NULL;
					const double t1 = now_s();
					if (rh_sigfile_next_batch(f, (uint64_t)S.mini_batch, 0, &b) != RH_OK) { fail(rh_gpu_last_error()); break; }
					if (!b || status) { if (b) rh_sigbatch_free(b); break; }
					t_read += now_s() - t1;
					to_map.push(b);
				}
				rh_sigfile_close(f);
			}
			for (uint32_t i = 0; i < nf; ++i) rh_free(files[i]);
			rh_free(files);
		}
		to_map.close();
	});
	auto quit = [&](int code) { /* leave before the mapping loop: stop step 0 and drop what it has queued */
		status = 1;
		rh_sigbatch_t *q;
		if (reader.joinable()) { while (to_map.pop(q)) rh_sigbatch_free(q); reader.join(); }
		return code;
	};

	/* ---- index ---------------------------------------------------------------------------------------------------- */
	float *pore = NULL; uint32_t n_pore = 0;
	if (!target_is_idx && S.pore && rh_pore_load(S.pore, P.k, P.lev_col, &pore, &n_pore) != RH_OK) {
		fprintf(stderr, "[ERROR] cannot parse the k-mer pore model file: %s\n", rh_gpu_last_error());
		return quit(1);
	}
	rh_index_t *idx = NULL;
	if (target_is_idx) {
		idx = rh_index_load(target, &P); /* sketch parameters come from the file (ri_idx_load, src/rindex.c:650-776) */
	} else if (sig_target) {
		/* ri_idx_siggen, src/rindex.c:927-969: every read of the target signal files becomes a target */
		char **files = NULL; uint32_t nf = 0;
		rh_find_sigfiles(target, &files, &nf);
		std::vector<rh_sigbatch_t *> held; std::vector<rh_sigfile_t *> open_files;
		std::vector<const char *> names; std::vector<const int16_t *> raw; std::vector<uint64_t> len; std::vector<double> off, rng, dig;
		bool ok = nf > 0;
		for (uint32_t i = 0; i < nf && ok; ++i) {
			rh_sigfile_t *f = rh_sigfile_open(files[i], S.n_threads);
			if (!f) { ok = false; break; }
			open_files.push_back(f);
			for (;;) {
				rh_sigbatch_t *b = NULL;
				if (rh_sigfile_next_batch(f, (uint64_t)S.mini_batch, 0, &b) != RH_OK) { ok = false; break; }
				if (!b) break;
				held.push_back(b);
				for (uint32_t r = 0; r < b->n; ++r) { names.push_back(b->names[r]); raw.push_back(b->raw[r]); len.push_back(b->raw_len[r]); off.push_back(b->offset[r]); rng.push_back(b->range[r]); dig.push_back(b->digitisation[r]); }
			}
		}
		if (ok) idx = rh_index_build_sig(&P, (uint32_t)names.size(), names.data(), raw.data(), len.data(), off.data(), rng.data(), dig.data());
		else if (nf == 0) fprintf(stderr, "[ERROR] no .slow5/.blow5 file under '%s'\n", target);
		for (rh_sigbatch_t *b : held) rh_sigbatch_free(b);
		for (rh_sigfile_t *f : open_files) rh_sigfile_close(f);
		for (uint32_t i = 0; i < nf; ++i) rh_free(files[i]);
		rh_free(files);
	} else {
		uint32_t n_seq = 0; char **names = NULL, **seqs = NULL; uint32_t *lens = NULL;
		if (rh_fasta_load(target, &n_seq, &names, &seqs, &lens) != RH_OK) { fprintf(stderr, "[ERROR] %s\n", rh_gpu_last_error()); return quit(1); }
		if (!S.index_on_host) {
			idx = rh_index_build_gpu(&P, pore, n_pore, n_seq, names, seqs, lens, S.device);
			if (!idx) { fprintf(stderr, "[M::%s] GPU index build not possible (%s); using the host builder with %d threads\n", __func__, rh_gpu_last_error(), S.n_threads); }
		}
		if (!idx) idx = rh_index_build(&P, pore, n_pore, n_seq, names, seqs, lens, S.n_threads);
		rh_fasta_free(n_seq, names, seqs, lens);
	}
	if (!idx) { fprintf(stderr, "[ERROR] failed to load/build the index from '%s': %s\n", target, rh_gpu_last_error()); return quit(1); }
	fprintf(stderr, "[M::%s::%.3f] loaded/built the index for %u target sequence(s)\n", __func__, now_s() - g_t0, rh_index_n_seq(idx));
	if (S.dump && !target_is_idx) {
		if (rh_index_dump(idx, S.dump, pore, n_pore) != RH_OK) { fprintf(stderr, "[ERROR] %s\n", rh_gpu_last_error()); return quit(1); }
	}
	rh_free(pore);
	if (pos.size() < 2) {
		fprintf(stderr, "[INFO] No files to query index on. Only the index is constructed.\n");
		rh_index_destroy(idx);
		return quit(0);
	}
	rh_index_update_mapopt(idx, &P); /* ri_mapopt_update, src/main.cpp:606 */
	fprintf(stderr, "[M::%s::%.3f] mid_occ = %d, min_mid_occ = %d, max_mid_occ = %d; distinct seeds: %llu, positions: %llu\n", __func__, now_s() - g_t0,
	        P.mid_occ, P.min_mid_occ, P.max_mid_occ, (unsigned long long)rh_index_n_keys(idx), (unsigned long long)rh_index_n_pos(idx));

	/* ---- GPUs ------------------------------------------------------------------------------------------------------ */
	std::vector<rh_gpu_ctx *> ctx;
	for (int g = 0; g < S.n_gpus; ++g) {
		rh_gpu_ctx *c = rh_gpu_init(idx, &P, S.device + g, 0);
		if (!c) { fprintf(stderr, "[ERROR] cannot initialise GPU %d: %s (there is no CPU mapping path)\n", S.device + g, rh_gpu_last_error()); return quit(1); }
		ctx.push_back(c);
	}
	if (!getenv("RH_NO_WARMUP")) warm_up(ctx);
	const double t_ready = now_s() - g_t0;

	std::thread printer([&]() { /* step 2: src/rmap.cpp:745-800 */
		mapped_batch m;
		while (to_print.pop(m)) {
			char *paf = rh_format_paf(idx, m.recs, m.n_recs, m.in->names);
			if (paf) { fputs(paf, stdout); rh_free(paf); }
			for (uint64_t i = 0; i < m.n_recs; ++i) n_mapped += m.recs[i].mapped && (i == 0 || m.recs[i].read_idx != m.recs[i - 1].read_idx);
			rh_free(m.recs);
			rh_sigbatch_free(m.in);
		}
	});
	rh_sigbatch_t *b;
	while (to_map.pop(b)) { /* step 1: the kt_for(map_worker_for) line, src/rmap.cpp:700 */
		if (status) { rh_sigbatch_free(b); continue; }
		const double t1 = now_s();
		rh_map_rec_t *recs = NULL; uint64_t n_recs = 0;
		const int rc = map_on_gpus(ctx, b, &recs, &n_recs);
		t_map += now_s() - t1;
		if (rc != RH_OK) { fail(std::string("mapping failed: ") + rh_gpu_last_error()); rh_free(recs); rh_sigbatch_free(b); continue; }
		n_reads += b->n; n_samples += b->n_samples;
		if (getenv("RH_CLI_VERBOSE")) {
			rh_gpu_stats_t st; rh_gpu_get_stats(ctx[0], &st);
			fprintf(stderr, "[M::batch %llu] %u reads, %llu samples (arena %s): map %.3f sec; GPU 0: %.1f ms on the stream (events %.1f seed %.1f sort %.1f chain %.1f post %.1f), %llu chunks in %llu rounds, H2D %.1f MB, %llu launches\n",
			        (unsigned long long)++n_batches, b->n, (unsigned long long)b->n_samples, b->arena_pinned ? "page-locked" : "pageable", now_s() - t1,
			        st.ms_total, st.ms_event_kernel, st.ms_seed, st.ms_sort, st.ms_chain, st.ms_post, (unsigned long long)st.n_chunks, (unsigned long long)st.n_rounds, st.h2d_bytes / 1e6, (unsigned long long)st.kernel_launches);
		}
		to_print.push({b, recs, n_recs});
	}
	to_print.close();
	reader.join(); printer.join();
	const double t_pipe = now_s() - t_pipe0;
	for (rh_gpu_ctx *c : ctx) rh_gpu_destroy(c);
	rh_index_destroy(idx);
	if (status) { fprintf(stderr, "ERROR: failed to map the query file\n"); return 1; }
	if (fflush(stdout) == EOF) { perror("[ERROR] failed to write the results"); return 1; }
	fprintf(stderr, "[M::%s] Version: %s\n[M::%s] mapped %llu of %llu reads (%llu raw samples); read+map+print pipeline: %.3f sec (%.0f reads/s); mapping step alone: %.3f sec (%.0f reads/s); file decode alone: %.3f sec; index + GPU ready after %.3f sec; real time: %.3f sec\n", __func__, RH_VERSION, __func__,
	        (unsigned long long)n_mapped, (unsigned long long)n_reads, (unsigned long long)n_samples, t_pipe, t_pipe > 0 ? n_reads / t_pipe : 0.0, t_map, t_map > 0 ? n_reads / t_map : 0.0, t_read, t_ready, now_s() - g_t0);
	return 0;
}
