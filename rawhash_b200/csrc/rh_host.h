/* rh_host.h — host-side structures shared by rh_host.cpp and rh_gpu.cu (product code). */
#ifndef RH_HOST_H
#define RH_HOST_H

#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/rawhash_b200.h"

/* Flattened index: key -> ascending position list (the mapping ri_idx_get returns,
 * reference src/rindex.c:497-514), stored as sorted keys + CSR offsets so it can be copied
 * to HBM verbatim. */
/* device-resident copy of the flattened index (rh_index_build_dev leaves the index here; rh_gpu_init on the same
 * device maps straight from it) */
struct rh_index_dev_t {
	int device = -1;
	uint32_t *keys = nullptr; uint64_t *off = nullptr; uint64_t *pos = nullptr;
	uint32_t *bucket = nullptr; int bucket_bits = 0;
	uint64_t n_keys = 0, n_pos = 0;
};

struct rh_index_s {
	~rh_index_s();
	rh_index_dev_t dev;
	bool host_valid = true;       /* false: keys/off/pos below are empty until rh_index_sync_host() */
	std::vector<uint32_t> keys;   /* distinct 32-bit seed hashes, ascending            */
	std::vector<uint64_t> off;    /* keys.size()+1 offsets into pos                    */
	std::vector<uint64_t> pos;    /* id<<32 | pos<<1 | strand, ascending within a key  */
	std::vector<std::string> names;
	std::vector<uint32_t> lens;   /* sequence length, or l_sig for a signal index      */
	int32_t flag = 0;
	int32_t w = 0, e = 0, n = 0, q = 0, k = 0;
	float diff = 0, fine_min = 0, fine_max = 0, fine_range = 0;
};

struct rh_seed_t { uint64_t x, y; };

/* ---- layout of one chunk round in the anchor arena (pure host logic: rh_plan_round, CPU-tested) ---- */
#define RH_SLOT_BYTES_PER_ANCHOR 144
static inline uint64_t rh_slot_region_bytes(uint64_t n_anchors) { return ((n_anchors * RH_SLOT_BYTES_PER_ANCHOR + 1024 + 255) / 256) * 256; }
struct rh_round_group_t { uint32_t first, count, heavy; };   /* plan slots [first, first + count); heavy: the second-stream lane */
struct rh_round_plan_t {
	std::vector<uint32_t> order;            /* plan slot q = input slot order[q] (all ns slots; only [0, n_run) run) */
	std::vector<uint64_t> a_off;            /* byte offset of plan slot q's region in the arena, q < n_run            */
	std::vector<rh_round_group_t> groups;   /* the heavy group first (if any), then the ordinary groups in launch order */
	uint32_t n_run = 0;                     /* mandatory slots + admitted optional slots                              */
	uint64_t main_bytes = 0;                /* arena bytes of the ordinary groups; the heavy group sits above          */
};
int rh_plan_round_impl(const uint32_t *n_anchors, uint32_t ns, uint32_t n_mandatory, uint32_t max_optional, uint64_t arena_bytes,
                       bool heaviest_first, bool heavy_lane, rh_round_plan_t *plan);

/* sketch of one event array on the host (index build side; the read side runs on the GPU) */
void rh_host_sketch(const rh_params_t &P, const float *ev, uint32_t len, uint32_t id, int strand, std::vector<rh_seed_t> &out);
void rh_index_from_seeds(rh_index_s *idx, std::vector<rh_seed_t> &seeds, int n_threads);
void rh_set_error(const char *fmt, ...);
int  rh_host_threads(void);                           /* hardware threads of the host (>= 1) */
int  rh_index_sync_host(const rh_index_s *idx);       /* download the device-resident index once */
void rh_index_dev_release(rh_index_s *idx);
int  rh_index_dev_make_buckets(rh_index_s *idx);
int  rh_index_dev_kth_occ(const rh_index_s *idx, uint64_t kth, uint32_t *out);

/* Index construction over contig groups (human-size references): consecutive sequences are grouped up to `group_bases`
 * bases (at least one sequence per group), `one` builds the index of a group with local sequence ids, and the parts
 * are merged key by key with ids rebased — position lists stay ascending because groups are in sequence order. */
typedef rh_index_t *(*rh_index_builder_fn)(const rh_params_t *, const float *, uint32_t, uint32_t, const char *const *, const char *const *, const uint32_t *, int);
rh_index_t *rh_index_build_grouped(rh_index_builder_fn one, uint64_t group_bases, const rh_params_t *p, const float *pore_vals, uint32_t n_pore_vals,
                                   uint32_t n_seq, const char *const *names, const char *const *seqs, const uint32_t *lens, int arg);
uint64_t rh_index_group_bases(uint64_t dflt); /* env RH_INDEX_GROUP_BASES or the default */

#endif
