/*
 * rh_io.cpp — the file surfaces either side of the mapping path (SURVEY.md §8b "file surfaces kept", §8f rank 3):
 *
 *   rh_index_dump      `.ind` writer            (ri_idx_dump,  src/rindex.c:545-648; read back by ri_idx_load 650-776)
 *   rh_fasta_load      FASTA / FASTA.gz reader  (mm_bseq_open/mm_bseq_read, src/bseq.c:38-128)
 *   rh_sigfile_*       SLOW5 / BLOW5 reader     (ri_sig_open_slow5 / ri_read_sig_slow5, src/rsig.c:170-207,478-533;
 *                                                file format of slow5lib 0.2.0, extern/slow5lib/src/slow5.c)
 *   rh_find_sigfiles   directory scan           (find_sfiles, src/rsig.c:286-330)
 *   rh_slow5_write     SLOW5 / BLOW5 writer     (synthetic data for bench and tests)
 *
 * Unlike the reference's reader this one never converts to pA: a batch is the raw int16 samples of its reads laid
 * back to back in one page-locked arena plus the calibration triples, which is what rh_gpu_map_batch_raw uploads
 * in one copy (the conversion of rsig.c:496-503 runs on the GPU).  Records are inflated and their signals decoded
 * by a pool of threads; the file itself is read sequentially.
 *
 * Host code only; nothing here is on the per-read hot path.  Citations are relative to the RawHash tree.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <errno.h>
#include <dirent.h>
#include <dlfcn.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <math.h>
#include <zlib.h>
#include <algorithm>
#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <immintrin.h>
#include <cuda_runtime_api.h>

#include "rh_host.h"
#include "rh_inflate.h"

/* ================================================================================================================
 * `.ind` writer
 * ================================================================================================================ */
namespace {

struct file_out {
	FILE *f; bool ok = true;
	explicit file_out(FILE *fp) : f(fp) {}
	void put(const void *p, size_t bytes) { if (ok && bytes && fwrite(p, 1, bytes, f) != bytes) ok = false; }
	template <class T> void val(T v) { put(&v, sizeof(T)); }
};

uint32_t revcomp_code(uint32_t x, int k)
{ /* rev_complement of a 2-bit packed k-mer (src/rutils.c:71-82) */
	uint32_t y = 0;
	for (int i = 0; i < k; ++i) { y = (y << 2) | (3u - (x & 3u)); x >>= 2; }
	return y;
}

const int IND_BUCKET_BITS = 14; /* ri_idx_load hard-codes b = 14 (src/rindex.c:669) */

/* The reference writes a bucket's (key, value) pairs in the slot order of its khash table (src/rindex.c:622-627), so
 * a byte-identical `.ind` needs that order: an open-addressing table of 2^m slots, slot = (key>>1) & mask with
 * triangular probing, created with room for n keys (kh_resize, src/khash.h:232-290), filled in ascending key order
 * (worker_post, src/rindex.c:334-352) and doubled in place — entries re-seated in slot order, displacing not yet
 * moved ones — whenever the load reaches 0.77 (kh_put, src/khash.h:291-300).  No deletions occur. */
struct slot_table {
	std::vector<uint64_t> key, val; std::vector<uint8_t> used;
	uint32_t n_slots = 0, size = 0, limit = 0;
	static uint32_t pow2_at_least(uint32_t x) { uint32_t p = 4; while (p < x) p <<= 1; return p; }
	uint32_t probe(const std::vector<uint8_t> &occ, uint64_t k, uint32_t mask) const
	{
		uint32_t i = (uint32_t)(k >> 1) & mask, step = 0;
		while (occ[i]) i = (i + (++step)) & mask;
		return i;
	}
	void reserve(uint32_t want)
	{
		const uint32_t nn = pow2_at_least(want);
		if (size >= (uint32_t)(nn * 0.77 + 0.5)) return;
		std::vector<uint8_t> occ(nn, 0);
		key.resize(std::max<size_t>(key.size(), nn)); val.resize(key.size());
		for (uint32_t j = 0; j < n_slots; ++j) {
			if (!used[j]) continue;
			uint64_t k = key[j], v = val[j];
			used[j] = 0;
			for (;;) {
				const uint32_t i = probe(occ, k, nn - 1);
				occ[i] = 1;
				if (i < n_slots && used[i]) { std::swap(k, key[i]); std::swap(v, val[i]); used[i] = 0; }
				else { key[i] = k; val[i] = v; break; }
			}
		}
		used.swap(occ); n_slots = nn; limit = (uint32_t)(nn * 0.77 + 0.5);
	}
	void insert(uint64_t k, uint64_t v)
	{
		if (size >= limit) reserve(n_slots + 1);
		const uint32_t i = probe(used, k, n_slots - 1);
		key[i] = k; val[i] = v; used[i] = 1; ++size;
	}
};

} // namespace

extern "C" int rh_index_dump(const rh_index_t *idx, const char *path, const float *pore_vals, uint32_t n_pore_vals)
{
	if (!idx || !path) { rh_set_error("rh_index_dump: bad arguments"); return RH_ERR_ARG; }
	if (rh_index_sync_host(idx) != RH_OK) return RH_ERR_CUDA; /* a device-resident index is downloaded first */
	for (const std::string &nm : idx->names)
		if (nm.size() > 255) { rh_set_error("rh_index_dump: sequence name longer than 255 bytes: %s", nm.c_str()); return RH_ERR_ARG; }
	FILE *fp = fopen(path, "wb");
	if (!fp) { rh_set_error("cannot create %s: %s", path, strerror(errno)); return RH_ERR_IO; }
	file_out o(fp);
	/* header: magic, w e n q k n_seq flag, diff fine_min fine_max fine_range (src/rindex.c:547-558) */
	o.put("RI", 2);
	const uint32_t pars[7] = {(uint32_t)idx->w, (uint32_t)idx->e, (uint32_t)idx->n, (uint32_t)idx->q, (uint32_t)idx->k, (uint32_t)idx->names.size(), (uint32_t)idx->flag};
	o.put(pars, sizeof(pars));
	o.val(idx->diff); o.val(idx->fine_min); o.val(idx->fine_max); o.val(idx->fine_range);
	/* ri_pore_t written raw (src/rutils.h:26): two pointers (meaningless in a file; zero here), n_pore_vals, k,
	 * padding, max_val, min_val; then pore_vals[] and the value-sorted pore_inds[] (src/rutils.c:84-114) */
	if (!pore_vals) n_pore_vals = 0;
	{
		unsigned char raw[32]; memset(raw, 0, sizeof(raw));
		const int16_t k16 = (int16_t)idx->k; const float mx = -5000.0f, mn = 5000.0f; /* as main() initialises them (src/main.cpp:573-576) */
		memcpy(raw + 16, &n_pore_vals, 4); memcpy(raw + 20, &k16, 2); memcpy(raw + 24, &mx, 4); memcpy(raw + 28, &mn, 4);
		o.put(raw, 32);
		o.put(pore_vals, (size_t)n_pore_vals * 4);
		struct porei { float v; uint32_t ind, rev; };
		std::vector<porei> pi(n_pore_vals);
		/* create_sorted_pairs normalises the already normalised levels once more (src/rutils.c:89-110; variance as
		 * fma(-mean, mean, E[x^2]), the contraction the reference is built with) */
		double sum = 0, sum2 = 0;
		for (uint32_t i = 0; i < n_pore_vals; ++i) { sum += pore_vals[i]; sum2 += pore_vals[i] * pore_vals[i]; }
		const double mean = n_pore_vals ? sum / n_pore_vals : 0, sd = n_pore_vals ? sqrt(fma(-mean, mean, sum2 / n_pore_vals)) : 1;
		for (uint32_t i = 0; i < n_pore_vals; ++i) pi[i] = {(float)((pore_vals[i] - mean) / sd), i, revcomp_code(i, idx->k)};
		std::stable_sort(pi.begin(), pi.end(), [](const porei &a, const porei &b) { return a.v < b.v; });
		o.put(pi.data(), pi.size() * sizeof(porei));
	}
	for (size_t i = 0; i < idx->names.size(); ++i) {
		const uint8_t l = (uint8_t)idx->names[i].size();
		o.val(l); o.put(idx->names[i].data(), l); o.val(idx->lens[i]);
	}
	/* buckets: key h lives in bucket h & (2^14-1) under the khash key (h>>14)<<1 | singleton; a singleton's value
	 * is the position itself, otherwise offset<<32 | count into the bucket's position array (src/rindex.c:334-352) */
	const uint32_t nb = 1u << IND_BUCKET_BITS, bmask = nb - 1;
	const size_t nk = idx->keys.size();
	std::vector<uint32_t> cnt(nb + 1, 0);
	for (size_t i = 0; i < nk; ++i) ++cnt[(idx->keys[i] & bmask) + 1];
	for (uint32_t b = 0; b < nb; ++b) cnt[b + 1] += cnt[b];
	std::vector<uint32_t> order(nk);
	{
		std::vector<uint32_t> cur(cnt.begin(), cnt.end() - 1);
		for (size_t i = 0; i < nk; ++i) order[cur[idx->keys[i] & bmask]++] = (uint32_t)i; /* keys ascending inside a bucket */
	}
	std::vector<uint64_t> kv;
	slot_table tab;
	for (uint32_t b = 0; b < nb && o.ok; ++b) {
		uint64_t np = 0;
		for (uint32_t j = cnt[b]; j < cnt[b + 1]; ++j) { const uint64_t c = idx->off[order[j] + 1] - idx->off[order[j]]; if (c > 1) np += c; }
		if (np > 0x7fffffffULL) { rh_set_error("rh_index_dump: bucket %u holds %llu positions (> 2^31-1)", b, (unsigned long long)np); fclose(fp); return RH_ERR_FORMAT; }
		o.val((int32_t)np);
		const uint32_t n_keys = cnt[b + 1] - cnt[b];
		tab = slot_table();
		if (n_keys) tab.reserve(n_keys);
		uint64_t at = 0;
		for (uint32_t j = cnt[b]; j < cnt[b + 1]; ++j) {
			const uint32_t ki = order[j];
			const uint64_t c = idx->off[ki + 1] - idx->off[ki], key = (uint64_t)(idx->keys[ki] >> IND_BUCKET_BITS) << 1;
			if (c == 1) tab.insert(key | 1, idx->pos[idx->off[ki]]);
			else { o.put(&idx->pos[idx->off[ki]], c * 8); tab.insert(key, at << 32 | c); at += c; }
		}
		kv.clear();
		for (uint32_t i = 0; i < tab.n_slots; ++i) if (tab.used[i]) { kv.push_back(tab.key[i]); kv.push_back(tab.val[i]); }
		o.val(n_keys);
		o.put(kv.data(), kv.size() * 8);
	}
	const bool ok = o.ok && fflush(fp) == 0;
	fclose(fp);
	if (!ok) { rh_set_error("write to %s failed", path); return RH_ERR_IO; }
	return RH_OK;
}

/* ================================================================================================================
 * FASTA
 * ================================================================================================================ */
extern "C" int rh_fasta_load(const char *path, uint32_t *n_seq, char ***names, char ***seqs, uint32_t **lens)
{
	if (!path || !n_seq || !names || !seqs || !lens) { rh_set_error("rh_fasta_load: bad arguments"); return RH_ERR_ARG; }
	/* transparent for plain files; "-" is standard input, like the reference's gzdopen(0) (src/bseq.c:38-50) */
	gzFile g = strcmp(path, "-") == 0 ? gzdopen(dup(0), "rb") : gzopen(path, "rb");
	if (!g) { rh_set_error("cannot open %s: %s", path, strerror(errno)); return RH_ERR_IO; }
	gzbuffer(g, 1 << 20);
	std::vector<std::string> nm, sq;
	std::vector<char> buf(1 << 20);
	/* kseq semantics (src/kseq.h): '>' or '@' opens a record, the name ends at the first white space, sequence lines
	 * run until a line starting with '>', '@' or '+'; after '+' as many quality characters as bases are skipped */
	enum { HEADER, SEQ, PLUS, QUAL } mode = SEQ;
	bool line_start = true, name_done = false;
	size_t qual_left = 0;
	int n;
	while ((n = gzread(g, buf.data(), (unsigned)buf.size())) > 0) {
		for (int i = 0; i < n; ++i) {
			const char c = buf[i];
			const bool nl = c == '\n';
			switch (mode) {
			case HEADER:
				if (nl) mode = SEQ;
				else if (!name_done) { if (c == ' ' || c == '\t' || c == '\r') name_done = true; else nm.back().push_back(c); }
				break;
			case SEQ:
				if (line_start && (c == '>' || c == '@')) { nm.emplace_back(); sq.emplace_back(); mode = HEADER; name_done = false; }
				else if (line_start && c == '+' && !sq.empty()) mode = PLUS;
				else if (!nl) { /* the rest of this line (as far as it is buffered) in one append */
					const char *e = (const char *)memchr(buf.data() + i, '\n', (size_t)(n - i));
					const int end = e ? (int)(e - buf.data()) : n;
					if (!sq.empty()) {
						std::string &dst = sq.back();
						const size_t before = dst.size();
						dst.append(buf.data() + i, (size_t)(end - i));
						if (memchr(buf.data() + i, '\r', (size_t)(end - i)) || memchr(buf.data() + i, ' ', (size_t)(end - i)) || memchr(buf.data() + i, '\t', (size_t)(end - i))) {
							size_t w = before;
							for (size_t r = before; r < dst.size(); ++r) if (dst[r] != '\r' && dst[r] != ' ' && dst[r] != '\t') dst[w++] = dst[r];
							dst.resize(w);
						}
					}
					i = end - 1; /* the newline (if buffered) is seen by the next iteration */
				}
				break;
			case PLUS:
				if (nl) { qual_left = sq.back().size(); mode = qual_left ? QUAL : SEQ; }
				break;
			case QUAL:
				if (!nl && c != '\r' && --qual_left == 0) mode = SEQ;
				break;
			}
			line_start = nl;
		}
	}
	const bool bad = n < 0;
	gzclose(g);
	if (bad) { rh_set_error("%s: read error", path); return RH_ERR_IO; }
	for (const std::string &s : sq) if (s.size() > 0xffffffffULL) { rh_set_error("%s: sequence longer than 2^32-1", path); return RH_ERR_FORMAT; }
	const uint32_t N = (uint32_t)nm.size();
	*names = (char **)malloc(sizeof(char *) * std::max(1u, N)); *seqs = (char **)malloc(sizeof(char *) * std::max(1u, N));
	*lens = (uint32_t *)malloc(sizeof(uint32_t) * std::max(1u, N));
	for (uint32_t i = 0; i < N; ++i) {
		(*names)[i] = strdup(nm[i].c_str());
		(*seqs)[i] = (char *)malloc(sq[i].size() + 1); memcpy((*seqs)[i], sq[i].c_str(), sq[i].size() + 1);
		(*lens)[i] = (uint32_t)sq[i].size();
		std::string().swap(sq[i]);
	}
	*n_seq = N;
	return RH_OK;
}

extern "C" void rh_fasta_free(uint32_t n_seq, char **names, char **seqs, uint32_t *lens)
{
	for (uint32_t i = 0; i < n_seq; ++i) { if (names) free(names[i]); if (seqs) free(seqs[i]); }
	free(names); free(seqs); free(lens);
}

/* ================================================================================================================
 * SLOW5 / BLOW5
 * ================================================================================================================ */
namespace {

const char BLOW5_MAGIC[6] = {'B', 'L', 'O', 'W', '5', '\1'};  /* extern/slow5lib/include/slow5/slow5_defs.h:132 */
const char BLOW5_EOF[5] = {'5', 'W', 'O', 'L', 'B'};           /* :133 */
const long BLOW5_HDR_SIZE_AT = 64;                              /* :134 */
enum { PRESS_NONE = 0, PRESS_ZLIB = 1, PRESS_ZSTD = 2, SIG_NONE = 0, SIG_SVB_ZD = 1 }; /* slow5_press.c:51-145 */

/* ---- streamvbyte + zig-zag delta (extern/slow5lib/thirdparty/streamvbyte; slow5_press.c:1054-1135) ------------- */
/* SSSE3 path: one control byte = four values.  A 16-byte shuffle spreads their 4..16 data bytes over four 32-bit
 * lanes; zig-zag, the running sum of the deltas and the narrowing to int16 stay in registers.  Selected at run time. */
struct svb_tables_t {
	alignas(16) uint8_t shuffle[256][16];
	uint8_t length[256];
	bool ssse3;
	svb_tables_t()
	{
		for (unsigned c = 0; c < 256; ++c) {
			unsigned at = 0;
			for (unsigned j = 0; j < 4; ++j) {
				const unsigned n = ((c >> (2 * j)) & 3u) + 1;
				for (unsigned b = 0; b < 4; ++b) shuffle[c][4 * j + b] = b < n ? (uint8_t)(at + b) : 0x80; /* 0x80: zero the byte */
				at += n;
			}
			length[c] = (uint8_t)at;
		}
		const char *e = getenv("RH_NO_SIMD");
		ssse3 = __builtin_cpu_supports("ssse3") && !(e && e[0] == '1');
	}
};
const svb_tables_t &svb_tables() { static const svb_tables_t t; return t; }

__attribute__((target("ssse3")))
void svbzd_decode_ssse3(const uint8_t *ctl, const uint8_t *&dat, const uint8_t *end, int16_t *out, uint32_t count, uint32_t &i, int32_t &prev)
{
	const svb_tables_t &T = svb_tables();
	const __m128i one = _mm_set1_epi32(1), zero = _mm_setzero_si128();
	const __m128i narrow = _mm_setr_epi8(0, 1, 4, 5, 8, 9, 12, 13, (char)0x80, (char)0x80, (char)0x80, (char)0x80, (char)0x80, (char)0x80, (char)0x80, (char)0x80);
	__m128i run = _mm_set1_epi32(prev); /* the last decoded sample in every lane */
	const uint8_t *d = dat;
	uint32_t k = i;
	for (; k + 4 <= count && d + 16 <= end; k += 4) {
		const unsigned c = ctl[k >> 2];
		__m128i v = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i *)d), _mm_load_si128((const __m128i *)T.shuffle[c]));
		d += T.length[c];
		v = _mm_xor_si128(_mm_srli_epi32(v, 1), _mm_sub_epi32(zero, _mm_and_si128(v, one))); /* zig-zag */
		v = _mm_add_epi32(v, _mm_slli_si128(v, 4));                                             /* prefix sum over the four deltas */
		v = _mm_add_epi32(v, _mm_slli_si128(v, 8));
		v = _mm_add_epi32(v, run);
		run = _mm_shuffle_epi32(v, 0xff);
		_mm_storel_epi64((__m128i *)(out + k), _mm_shuffle_epi8(v, narrow));                    /* low 16 bits of each lane, like (int16_t) */
	}
	prev = _mm_cvtsi128_si32(run);
	dat = d; i = k;
}

/* layout: u32 count | ceil(count/4) control bytes (2 bits per value: 1..4 data bytes) | little-endian data bytes */
bool svbzd_decode(const uint8_t *in, size_t in_bytes, int16_t *out, uint32_t expect)
{
	if (in_bytes < 4) return false;
	uint32_t count; memcpy(&count, in, 4);
	if (count != expect) return false;
	const size_t nctl = ((size_t)count + 3) / 4;
	if (in_bytes < 4 + nctl) return false;
	const uint8_t *ctl = in + 4, *dat = ctl + nctl, *end = in + in_bytes;
	int32_t prev = 0;
	uint32_t i = 0;
	if (svb_tables().ssse3) svbzd_decode_ssse3(ctl, dat, end, out, count, i, prev); /* whole control bytes while 16 input bytes remain */
	static const uint32_t KEEP[4] = {0xffu, 0xffffu, 0xffffffu, 0xffffffffu};
	/* four values per control byte; a 4-byte load per value is safe while 16 bytes of input remain */
	for (; i + 4 <= count && dat + 16 <= end; i += 4) {
		const unsigned c = ctl[i >> 2];
		for (int j = 0; j < 4; ++j) {
			const unsigned code = (c >> (2 * j)) & 3u;
			uint32_t v; memcpy(&v, dat, 4);
			v &= KEEP[code];
			dat += code + 1;
			prev += (int32_t)(v >> 1) ^ -(int32_t)(v & 1);
			out[i + j] = (int16_t)prev;
		}
	}
	for (; i < count; ++i) {
		const unsigned code = (ctl[i >> 2] >> ((i & 3) * 2)) & 3u;
		if (dat + code + 1 > end) return false;
		uint32_t v = 0;
		for (unsigned b = 0; b <= code; ++b) v |= (uint32_t)dat[b] << (8 * b);
		dat += code + 1;
		prev += (int32_t)(v >> 1) ^ -(int32_t)(v & 1);
		out[i] = (int16_t)prev;
	}
	return dat == end;
}

uint32_t svbzd_count(const uint8_t *in, size_t in_bytes) { uint32_t c = 0; if (in_bytes >= 4) memcpy(&c, in, 4); return c; }

void svbzd_encode(const int16_t *in, uint32_t count, std::vector<uint8_t> &out)
{
	const size_t nctl = ((size_t)count + 3) / 4;
	out.assign(4 + nctl + (size_t)count * 4, 0);
	memcpy(out.data(), &count, 4);
	uint8_t *ctl = out.data() + 4, *dat = ctl + nctl;
	int32_t prev = 0;
	for (uint32_t i = 0; i < count; ++i) {
		const int32_t d = (int32_t)in[i] - prev; prev = in[i];
		const uint32_t v = ((uint32_t)d << 1) ^ (uint32_t)(d >> 31);
		const unsigned code = v < (1u << 8) ? 0 : v < (1u << 16) ? 1 : v < (1u << 24) ? 2 : 3;
		ctl[i >> 2] |= (uint8_t)(code << ((i & 3) * 2));
		memcpy(dat, &v, 4); /* little endian; the buffer has room for 4 bytes per value */
		dat += code + 1;
	}
	out.resize((size_t)(dat - out.data()));
}

bool zlib_inflate(const uint8_t *in, size_t in_bytes, std::vector<uint8_t> &out)
{ /* one zlib stream per record (ptr_depress_zlib_solo, slow5_press.c:945-982) */
	z_stream s; memset(&s, 0, sizeof(s));
	if (inflateInit2(&s, MAX_WBITS) != Z_OK) return false;
	out.resize(std::max<size_t>(in_bytes * 4, 1 << 12));
	s.next_in = (Bytef *)in; s.avail_in = (uInt)in_bytes;
	size_t done = 0; int rc;
	do {
		if (done == out.size()) out.resize(out.size() * 2);
		s.next_out = out.data() + done; s.avail_out = (uInt)std::min<size_t>(out.size() - done, 1u << 30);
		const size_t before = s.avail_out;
		rc = inflate(&s, Z_NO_FLUSH);
		done += before - s.avail_out;
	} while (rc == Z_OK || (rc == Z_BUF_ERROR && s.avail_in > 0 && done == out.size()));
	inflateEnd(&s);
	out.resize(done);
	return rc == Z_STREAM_END;
}

/* zstd records (slow5lib built with zstd=1: ptr_compress_zstd / ptr_depress_zstd, slow5_press.c:1146-1200).  This image
 * ships libzstd.so.1 without headers, so the four stable entry points are bound at run time; without the library zstd
 * files are refused by name. */
struct zstd_api {
	size_t (*decompress)(void *, size_t, const void *, size_t) = nullptr;
	unsigned long long (*frame_content_size)(const void *, size_t) = nullptr;
	unsigned (*is_error)(size_t) = nullptr;
	size_t (*compress)(void *, size_t, const void *, size_t, int) = nullptr;
	size_t (*compress_bound)(size_t) = nullptr;
	bool ok = false;
};
const zstd_api &zstd()
{
	static const zstd_api api = []() {
		zstd_api a;
		void *h = dlopen("libzstd.so.1", RTLD_NOW | RTLD_LOCAL);
		if (!h) h = dlopen("libzstd.so", RTLD_NOW | RTLD_LOCAL);
		if (!h) return a;
		a.decompress = (size_t (*)(void *, size_t, const void *, size_t))dlsym(h, "ZSTD_decompress");
		a.frame_content_size = (unsigned long long (*)(const void *, size_t))dlsym(h, "ZSTD_getFrameContentSize");
		a.is_error = (unsigned (*)(size_t))dlsym(h, "ZSTD_isError");
		a.compress = (size_t (*)(void *, size_t, const void *, size_t, int))dlsym(h, "ZSTD_compress");
		a.compress_bound = (size_t (*)(size_t))dlsym(h, "ZSTD_compressBound");
		a.ok = a.decompress && a.frame_content_size && a.is_error && a.compress && a.compress_bound;
		return a;
	}();
	return api;
}

bool zstd_inflate(const uint8_t *in, size_t in_bytes, std::vector<uint8_t> &out)
{
	const zstd_api &z = zstd();
	if (!z.ok) return false;
	const unsigned long long n = z.frame_content_size(in, in_bytes);
	if (n >= 0xfffffffffffffffeULL || n > ((unsigned long long)1 << 34)) return false; /* unknown / error / implausible */
	out.resize((size_t)n);
	const size_t got = z.decompress(out.data(), out.size(), in, in_bytes);
	return !z.is_error(got) && got == n;
}

/* records go through the table-driven decoder of rh_inflate.h (it needs readable bytes after the input: the file
 * mapping provides them except for the last records of a file, which are copied); RH_ZLIB=1 selects zlib's inflate */
bool inflate_record(const uint8_t *in, size_t in_bytes, size_t readable_after, std::vector<uint8_t> &out)
{
	static const bool use_zlib = []() { const char *e = getenv("RH_ZLIB"); return e && e[0] == '1'; }();
	if (use_zlib) return zlib_inflate(in, in_bytes, out);
	bool ok;
	if (readable_after >= rhz::RH_INFLATE_SLACK) ok = rhz::inflate_zlib(in, in_bytes, out);
	else {
		std::vector<uint8_t> padded(in_bytes + rhz::RH_INFLATE_SLACK, 0);
		memcpy(padded.data(), in, in_bytes);
		ok = rhz::inflate_zlib(padded.data(), in_bytes, out);
	}
	return ok || zlib_inflate(in, in_bytes, out); /* second opinion: zlib decides what is malformed */
}

bool zlib_deflate(const uint8_t *in, size_t in_bytes, std::vector<uint8_t> &out)
{
	uLongf cap = compressBound((uLong)in_bytes);
	out.resize(cap);
	if (compress2(out.data(), &cap, in, (uLong)in_bytes, Z_DEFAULT_COMPRESSION) != Z_OK) return false;
	out.resize(cap);
	return true;
}

/* ---- one record between the file and the arena ------------------------------------------------------------------ */
struct rec_t {
	const uint8_t *src = nullptr; size_t src_bytes = 0, src_slack = 0; /* binary: the stored record inside the file mapping (+ readable bytes after it) */
	std::vector<uint8_t> mem;       /* binary: the inflated record when records are compressed; ASCII: the line */
	const uint8_t *body = nullptr; size_t body_bytes = 0; /* what the fields are parsed from (src or mem) */
	std::string name;
	double digitisation = 0, offset = 0, range = 0, sampling_rate = 0;
	uint64_t n_samples = 0;         /* decoded sample count */
	size_t sig_at = 0, sig_bytes = 0; /* where the signal payload sits inside mem */
	bool ok = false;
};

bool parse_binary_record(rec_t &r, int rec_method, int sig_method)
{ /* slow5_rec_parse, binary branch (slow5.c:2806-2925): u16 id_len, id, u32 read_group, 4 doubles, u64 length, signal */
	r.body = r.src; r.body_bytes = r.src_bytes;
	if (rec_method == PRESS_ZLIB || rec_method == PRESS_ZSTD) {
		if (!(rec_method == PRESS_ZLIB ? inflate_record(r.src, r.src_bytes, r.src_slack, r.mem) : zstd_inflate(r.src, r.src_bytes, r.mem))) return false;
		r.body = r.mem.data(); r.body_bytes = r.mem.size();
	}
	const uint8_t *p = r.body; const size_t n = r.body_bytes; size_t at = 0;
	uint16_t idl;
	if (n < 2) return false;
	memcpy(&idl, p, 2); at = 2;
	if (at + idl + 4 + 32 + 8 > n) return false;
	r.name.assign((const char *)p + at, idl); at += idl;
	at += 4; /* read_group */
	memcpy(&r.digitisation, p + at, 8); at += 8;
	memcpy(&r.offset, p + at, 8); at += 8;
	memcpy(&r.range, p + at, 8); at += 8;
	memcpy(&r.sampling_rate, p + at, 8); at += 8;
	uint64_t len; memcpy(&len, p + at, 8); at += 8;
	r.sig_at = at;
	if (len > n - at) return false; /* before any arithmetic on it: a crafted length must not wrap (2 bytes per sample, or the compressed byte count) */
	if (sig_method == SIG_NONE) { r.sig_bytes = len * 2; r.n_samples = len; }
	else { r.sig_bytes = len; r.n_samples = svbzd_count(p + at, std::min<size_t>(len, n - at)); } /* length field = compressed bytes */
	if (r.sig_bytes > n - r.sig_at) return false; /* auxiliary fields after the signal are not needed */
	return sig_method == SIG_NONE || r.n_samples + (r.n_samples + 3) / 4 + 4 <= r.sig_bytes; /* every value has >= 1 data byte */
}

bool parse_ascii_record(rec_t &r)
{ /* slow5_rec_parse, ASCII branch (slow5.c:2648-2770): read_id read_group digitisation offset range sampling_rate len signal[,..] */
	char *s = (char *)r.mem.data(); /* NUL-terminated by the reader */
	r.body = r.mem.data(); r.body_bytes = r.mem.size();
	char *f[8]; int nf = 0;
	f[nf++] = s;
	for (char *c = s; *c && nf < 8; ++c) if (*c == '\t') { *c = 0; f[nf++] = c + 1; }
	if (nf < 8) return false;
	r.name = f[0];
	char *e;
	r.digitisation = strtod(f[2], &e); if (e == f[2]) return false;
	r.offset = strtod(f[3], &e); if (e == f[3]) return false;
	r.range = strtod(f[4], &e); if (e == f[4]) return false;
	r.sampling_rate = strtod(f[5], &e); if (e == f[5]) return false;
	r.n_samples = strtoull(f[6], &e, 10); if (e == f[6]) return false;
	r.sig_at = (size_t)(f[7] - s);
	char *t = strchr(f[7], '\t'); /* auxiliary columns follow */
	r.sig_bytes = t ? (size_t)(t - f[7]) : strlen(f[7]);
	return r.n_samples <= (r.sig_bytes + 1) / 2; /* "d,d,...,d" */
}

bool decode_signal(const rec_t &r, bool binary, int sig_method, int16_t *out)
{
	const uint8_t *p = r.body + r.sig_at;
	if (binary) {
		if (sig_method == SIG_NONE) { memcpy(out, p, r.sig_bytes); return true; }
		return svbzd_decode(p, r.sig_bytes, out, (uint32_t)r.n_samples);
	}
	if (r.n_samples == 0) return true; /* an empty signal is written as "" or "." */
	const char *c = (const char *)p, *end = c + r.sig_bytes;
	uint64_t i = 0;
	while (c < end && i < r.n_samples) {
		bool neg = false; int v = 0;
		if (*c == '-') { neg = true; ++c; }
		if (c >= end || *c < '0' || *c > '9') return false;
		while (c < end && *c >= '0' && *c <= '9') { v = v * 10 + (*c - '0'); if (v > 40000) return false; ++c; }
		v = neg ? -v : v;
		if (v < -32768 || v > 32767) return false;
		out[i++] = (int16_t)v;
		if (c < end) { if (*c != ',') return false; ++c; }
	}
	return i == r.n_samples && c >= end;
}

template <class F> void parallel_for(size_t n, int n_threads, F fn)
{
	if (n_threads <= 1 || n < 2) { for (size_t i = 0; i < n; ++i) fn(i); return; }
	std::atomic<size_t> next(0);
	auto work = [&]() { for (size_t i; (i = next.fetch_add(16)) < n;) for (size_t j = i; j < std::min(n, i + 16); ++j) fn(j); };
	std::vector<std::thread> th;
	const int nt = (int)std::min<size_t>((size_t)n_threads, (n + 15) / 16);
	for (int t = 1; t < nt; ++t) th.emplace_back(work);
	work();
	for (auto &t : th) t.join();
}

struct arena_t { int16_t *p = nullptr; size_t cap = 0; bool pinned = false; };

arena_t arena_alloc(size_t samples)
{
	arena_t a; a.cap = std::max<size_t>(samples, 1);
	static std::atomic<int> pin_state(-1); /* -1 unknown, 0 no usable device, 1 page-lock */
	const char *env = getenv("RH_PIN");
	if (pin_state.load() != 0 && !(env && env[0] == '0')) {
		void *p = nullptr;
		if (cudaHostAlloc(&p, a.cap * 2, cudaHostAllocPortable) == cudaSuccess) { a.p = (int16_t *)p; a.pinned = true; pin_state = 1; return a; }
		(void)cudaGetLastError();
		pin_state = 0;
	}
	a.p = (int16_t *)malloc(a.cap * 2);
	return a;
}

void arena_free(arena_t &a) { if (a.p) { if (a.pinned) cudaFreeHost(a.p); else free(a.p); } a = arena_t(); }

/* Page-locking a 1 GB arena costs a few hundred milliseconds, so arenas are recycled: a process-wide pool of at most
 * three (one being decoded into, one queued, one being mapped), shared by all signal files that are read one after
 * the other.  What is still pooled at exit is left to the operating system. */
std::mutex g_pool_mu;
std::vector<arena_t> g_pool;

arena_t arena_take(size_t samples)
{
	arena_t drop;
	{
		std::lock_guard<std::mutex> g(g_pool_mu);
		for (size_t i = 0; i < g_pool.size(); ++i)
			if (g_pool[i].cap >= samples) { arena_t a = g_pool[i]; g_pool.erase(g_pool.begin() + i); return a; }
		if (!g_pool.empty()) { drop = g_pool.back(); g_pool.pop_back(); } /* too small for this batch: replace it */
	}
	arena_free(drop);
	return arena_alloc(samples + samples / 8);
}

void arena_give_back(arena_t &a)
{
	if (!a.p) return;
	{
		std::lock_guard<std::mutex> g(g_pool_mu);
		if (g_pool.size() < 3) { g_pool.push_back(a); a = arena_t(); return; }
	}
	arena_free(a);
}

} // namespace

struct rh_sigfile_s {
	FILE *fp = nullptr; std::string path;
	bool binary = false; int rec_method = PRESS_NONE, sig_method = SIG_NONE;
	int n_threads = 1; bool eof = false;
	std::mutex mu;
	std::atomic<int> live_batches{0}; bool closed = false;
	char *line = nullptr; size_t line_cap = 0;
	const uint8_t *map = nullptr; size_t map_bytes = 0, map_at = 0; /* binary files are read through a mapping: no copies */
	std::vector<rec_t> pending; size_t pending_at = 0; /* parsed records not yet handed out */
};

namespace {
struct batch_impl {
	rh_sigbatch_t pub;
	rh_sigfile_s *owner = nullptr;
	arena_t arena;
	std::vector<const int16_t *> raw; std::vector<uint64_t> len;
	std::vector<double> offset, range, digitisation, sampling_rate;
	std::vector<std::string> name_store; std::vector<const char *> names;
};

void sigfile_destroy(rh_sigfile_s *f)
{
	if (f->map) munmap((void *)f->map, f->map_bytes);
	if (f->fp) fclose(f->fp);
	free(f->line);
	delete f;
}

bool has_suffix(const std::string &s, const char *suf) { const size_t n = strlen(suf); return s.size() >= n && s.compare(s.size() - n, n, suf) == 0; }
} // namespace

extern "C" rh_sigfile_t *rh_sigfile_open(const char *path, int n_threads)
{ /* slow5_open + slow5_hdr_init (slow5.c:636-900): the format follows the extension */
	if (!path) { rh_set_error("rh_sigfile_open: no path"); return NULL; }
	const std::string p(path);
	const bool bin = has_suffix(p, ".blow5");
	if (!bin && !has_suffix(p, ".slow5")) { rh_set_error("%s: only .slow5 and .blow5 signal files are supported (FAST5/POD5 need HDF5/Arrow)", path); return NULL; }
	FILE *fp = fopen(path, "rb");
	if (!fp) { rh_set_error("cannot open %s: %s", path, strerror(errno)); return NULL; }
	rh_sigfile_s *f = new rh_sigfile_s();
	f->fp = fp; f->path = p; f->binary = bin; f->n_threads = std::max(1, n_threads);
	setvbuf(fp, NULL, _IOFBF, 4 << 20);
	bool ok = true;
	if (bin) {
		struct stat st;
		ok = fstat(fileno(fp), &st) == 0 && (size_t)st.st_size >= BLOW5_HDR_SIZE_AT + 4;
		if (ok) {
			void *m = mmap(NULL, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fileno(fp), 0);
			ok = m != MAP_FAILED;
			if (ok) { f->map = (const uint8_t *)m; f->map_bytes = (size_t)st.st_size; madvise(m, f->map_bytes, MADV_SEQUENTIAL); }
		}
		const uint8_t *h = f->map;
		ok = ok && memcmp(h, BLOW5_MAGIC, 6) == 0;
		if (!ok) rh_set_error("%s: not a BLOW5 file (bad magic number)", path);
		else {
			const unsigned major = h[6], minor = h[7];
			f->rec_method = h[9];
			f->sig_method = (major > 0 || minor >= 2) ? (int)h[14] : (int)SIG_NONE; /* signal compression byte exists from 0.2.0 (slow5.c:824) */
			if (major > 0 || minor > 2) { rh_set_error("%s: BLOW5 version %u.%u.%u is newer than 0.2.0", path, major, minor, (unsigned)h[8]); ok = false; }
			else if (f->rec_method == PRESS_ZSTD && !zstd().ok) { rh_set_error("%s: zstd record compression needs libzstd.so.1 at run time, which is not installed (zlib and none are built in)", path); ok = false; }
			else if (f->rec_method > PRESS_ZSTD || f->sig_method > SIG_SVB_ZD) { rh_set_error("%s: unknown compression method (record %d, signal %d)", path, f->rec_method, f->sig_method); ok = false; }
			uint32_t hdr_bytes; memcpy(&hdr_bytes, h + BLOW5_HDR_SIZE_AT, 4);
			if (ok) { /* the ASCII header block must end with the column-name line */
				const size_t body_at = BLOW5_HDR_SIZE_AT + 4;
				ok = hdr_bytes <= f->map_bytes - body_at;
				if (ok) { const std::string hb((const char *)h + body_at, hdr_bytes); ok = hb.find("#read_id\t") != std::string::npos; }
				if (!ok) rh_set_error("%s: malformed BLOW5 header", path);
				f->map_at = body_at + hdr_bytes;
			}
		}
	} else {
		bool seen_cols = false;
		for (;;) {
			const int c = fgetc(fp);
			if (c == EOF) break;
			if (c != '#' && c != '@') { ungetc(c, fp); break; }
			ungetc(c, fp);
			if (getline(&f->line, &f->line_cap, fp) < 0) break;
			if (strncmp(f->line, "#read_id\t", 9) == 0) { seen_cols = true; break; }
		}
		ok = seen_cols;
		if (!ok) rh_set_error("%s: malformed SLOW5 header (no #read_id column line)", path);
	}
	if (!ok) { sigfile_destroy(f); return NULL; }
	return f;
}

extern "C" void rh_sigfile_close(rh_sigfile_t *f)
{
	if (!f) return;
	bool destroy;
	{ std::lock_guard<std::mutex> g(f->mu); f->closed = true; destroy = f->live_batches.load() == 0; }
	if (destroy) sigfile_destroy(f); /* otherwise the last rh_sigbatch_free does it */
}

/* next stored record -> r.src (binary) or r.mem (ASCII); returns 1 record, 0 end of file, <0 error */
static int read_stored_record(rh_sigfile_s *f, rec_t &r)
{
	if (f->binary) { /* slow5_get_next_mem (slow5.c:3191-3270): u64 size, then the bytes; "5WOLB" closes the file */
		const size_t left = f->map_bytes - f->map_at;
		if (left < 8) {
			if (left == 5 && memcmp(f->map + f->map_at, BLOW5_EOF, 5) == 0) return 0;
			rh_set_error("%s: truncated BLOW5 (no end-of-file marker)", f->path.c_str());
			return RH_ERR_FORMAT;
		}
		uint64_t sz; memcpy(&sz, f->map + f->map_at, 8);
		if (sz > left - 8) { rh_set_error("%s: truncated BLOW5 record", f->path.c_str()); return RH_ERR_FORMAT; }
		r.src = f->map + f->map_at + 8; r.src_bytes = sz; r.src_slack = left - 8 - sz;
		f->map_at += 8 + sz;
		return 1;
	}
	const ssize_t n = getline(&f->line, &f->line_cap, f->fp);
	if (n < 0) return 0;
	size_t l = (size_t)n;
	while (l && (f->line[l - 1] == '\n' || f->line[l - 1] == '\r')) --l;
	r.mem.assign((const uint8_t *)f->line, (const uint8_t *)f->line + l);
	r.mem.push_back(0);
	return 1;
}

extern "C" int rh_sigfile_next_batch(rh_sigfile_t *f, uint64_t max_samples, uint32_t max_reads, rh_sigbatch_t **out)
{
	if (!f || !out) { rh_set_error("rh_sigfile_next_batch: bad arguments"); return RH_ERR_ARG; }
	*out = NULL;
	if (max_samples == 0) max_samples = 500000000ULL; /* mini_batch_size default (src/roptions.c:88) */
	std::vector<rec_t> recs;
	uint64_t total = 0;
	const size_t group = 4096; /* the reference pulls 4096 records at a time (src/rsig.c:186) */
	/* ri_sig_read_frag (src/rmap.cpp:600-660): reads are appended until the running sample count reaches the limit */
	while (total < max_samples && (max_reads == 0 || recs.size() < max_reads)) {
		if (f->pending_at == f->pending.size()) {
			f->pending.clear(); f->pending_at = 0;
			if (f->eof) break;
			for (size_t i = 0; i < group; ++i) {
				rec_t r;
				const int rc = read_stored_record(f, r);
				if (rc < 0) { f->pending.clear(); f->eof = true; return rc; } /* a damaged file ends here: later calls report end of file */
				if (rc == 0) { f->eof = true; break; }
				if (!f->binary && r.mem.size() <= 1) continue; /* blank line */
				f->pending.emplace_back(std::move(r));
			}
			const bool binary = f->binary; const int rm = f->rec_method, sm = f->sig_method;
			std::vector<rec_t> &P = f->pending;
			parallel_for(P.size(), f->n_threads, [&](size_t i) { P[i].ok = binary ? parse_binary_record(P[i], rm, sm) : parse_ascii_record(P[i]); });
			for (size_t i = 0; i < P.size(); ++i)
				if (!P[i].ok) { rh_set_error("%s: a record cannot be parsed (record %zu of the current group)", f->path.c_str(), i); f->pending.clear(); f->eof = true; return RH_ERR_FORMAT; }
			if (P.empty()) break;
		}
		total += f->pending[f->pending_at].n_samples;
		recs.emplace_back(std::move(f->pending[f->pending_at++]));
	}
	if (recs.empty()) return RH_OK;
	batch_impl *b = new batch_impl();
	b->owner = f;
	b->arena = arena_take(total);
	++f->live_batches;
	if (!b->arena.p) { rh_set_error("out of host memory for a %llu-sample batch", (unsigned long long)total); --f->live_batches; delete b; return RH_ERR_NOMEM; }
	const size_t n = recs.size();
	b->raw.resize(n); b->len.resize(n); b->offset.resize(n); b->range.resize(n); b->digitisation.resize(n); b->sampling_rate.resize(n);
	b->name_store.resize(n); b->names.resize(n);
	uint64_t at = 0;
	for (size_t i = 0; i < n; ++i) {
		b->raw[i] = b->arena.p + at; b->len[i] = recs[i].n_samples; at += recs[i].n_samples;
		b->offset[i] = recs[i].offset; b->range[i] = recs[i].range; b->digitisation[i] = recs[i].digitisation; b->sampling_rate[i] = recs[i].sampling_rate;
		b->name_store[i].swap(recs[i].name);
	}
	for (size_t i = 0; i < n; ++i) b->names[i] = b->name_store[i].c_str();
	std::atomic<bool> bad(false);
	{
		const bool binary = f->binary; const int sm = f->sig_method;
		parallel_for(n, f->n_threads, [&](size_t i) {
			if (!decode_signal(recs[i], binary, sm, const_cast<int16_t *>(b->raw[i]))) bad = true;
			std::vector<uint8_t>().swap(recs[i].mem);
		});
	}
	b->pub.n = (uint32_t)n; b->pub.n_samples = total;
	b->pub.raw = b->raw.data(); b->pub.raw_len = b->len.data();
	b->pub.offset = b->offset.data(); b->pub.range = b->range.data(); b->pub.digitisation = b->digitisation.data(); b->pub.sampling_rate = b->sampling_rate.data();
	b->pub.names = b->names.data(); b->pub.arena_pinned = b->arena.pinned ? 1 : 0; b->pub.priv = b;
	if (bad) { rh_set_error("%s: a raw signal cannot be decoded", f->path.c_str()); rh_sigbatch_free(&b->pub); return RH_ERR_FORMAT; }
	*out = &b->pub;
	return RH_OK;
}

extern "C" void rh_sigbatch_free(rh_sigbatch_t *pub)
{
	if (!pub) return;
	batch_impl *b = (batch_impl *)pub->priv;
	rh_sigfile_s *f = b->owner;
	arena_give_back(b->arena);
	bool destroy = false;
	{
		std::lock_guard<std::mutex> g(f->mu);
		destroy = --f->live_batches == 0 && f->closed;
	}
	delete b;
	if (destroy) sigfile_destroy(f);
}

extern "C" int rh_zlib_inflate(const void *in, size_t in_bytes, void **out, size_t *out_bytes)
{
	if (!in || !out || !out_bytes) { rh_set_error("rh_zlib_inflate: bad arguments"); return RH_ERR_ARG; }
	std::vector<uint8_t> padded(in_bytes + rhz::RH_INFLATE_SLACK, 0), plain;
	memcpy(padded.data(), in, in_bytes);
	if (!rhz::inflate_zlib(padded.data(), in_bytes, plain)) { rh_set_error("rh_zlib_inflate: malformed zlib stream"); return RH_ERR_FORMAT; }
	*out = malloc(plain.size() ? plain.size() : 1);
	if (!*out) return RH_ERR_NOMEM;
	memcpy(*out, plain.data(), plain.size());
	*out_bytes = plain.size();
	return RH_OK;
}

/* ---- directory scan -------------------------------------------------------------------------------------------- */
static bool is_sigfile_name(const char *s) { return strstr(s, ".slow5") || strstr(s, ".blow5"); }
static bool is_directory(const std::string &p) { struct stat st; return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode); }
static void scan_sigfiles(const std::string &p, std::vector<std::string> &out)
{
	if (!is_directory(p)) { if (is_sigfile_name(p.c_str())) out.push_back(p); return; }
	DIR *d = opendir(p.c_str());
	if (!d) return;
	std::vector<std::string> here;
	for (struct dirent *e; (e = readdir(d)) != NULL;) {
		if (!strcmp(e->d_name, ".") || !strcmp(e->d_name, "..")) continue;
		here.push_back(p + "/" + e->d_name);
	}
	closedir(d);
	std::sort(here.begin(), here.end()); /* readdir order is file-system dependent; sorted = reproducible output order */
	for (const std::string &c : here) { if (is_directory(c)) scan_sigfiles(c, out); else if (is_sigfile_name(c.c_str())) out.push_back(c); }
}

extern "C" int rh_find_sigfiles(const char *path, char ***files, uint32_t *n_files)
{
	if (!path || !files || !n_files) { rh_set_error("rh_find_sigfiles: bad arguments"); return RH_ERR_ARG; }
	std::vector<std::string> v;
	scan_sigfiles(path, v);
	*files = (char **)malloc(sizeof(char *) * std::max<size_t>(1, v.size()));
	for (size_t i = 0; i < v.size(); ++i) (*files)[i] = strdup(v[i].c_str());
	*n_files = (uint32_t)v.size();
	return RH_OK;
}

/* ---- writer ------------------------------------------------------------------------------------------------------ */
extern "C" int rh_slow5_write(const char *path, uint32_t n, const char *const *names,
                              const int16_t *const *raw, const uint64_t *raw_len,
                              const double *offset, const double *range, const double *digitisation, double sampling_rate,
                              int record_press, int signal_press)
{ /* slow5_hdr_to_mem + slow5_rec_to_mem (slow5.c:940-1150, 3900-4050) */
	if (!path || (n && (!names || !raw || !raw_len || !offset || !range || !digitisation))) { rh_set_error("rh_slow5_write: bad arguments"); return RH_ERR_ARG; }
	const std::string p(path);
	const bool bin = has_suffix(p, ".blow5");
	if (!bin && !has_suffix(p, ".slow5")) { rh_set_error("rh_slow5_write: %s must end in .slow5 or .blow5", path); return RH_ERR_ARG; }
	if (record_press < 0 || record_press > PRESS_ZSTD || signal_press < 0 || signal_press > SIG_SVB_ZD) { rh_set_error("rh_slow5_write: unsupported compression"); return RH_ERR_ARG; }
	if (record_press == PRESS_ZSTD && !zstd().ok) { rh_set_error("rh_slow5_write: libzstd.so.1 is not installed"); return RH_ERR_ARG; }
	FILE *fp = fopen(path, "wb");
	if (!fp) { rh_set_error("cannot create %s: %s", path, strerror(errno)); return RH_ERR_IO; }
	file_out o(fp);
	char num[64];
	snprintf(num, sizeof(num), "%.0f", sampling_rate);
	const std::string attrs = std::string("@asic_id\t0\n@exp_start_time\t0\n@flow_cell_id\tsynthetic\n@run_id\trawhash_b200\n@sample_frequency\t") + num + "\n";
	const std::string cols = "#char*\tuint32_t\tdouble\tdouble\tdouble\tdouble\tuint64_t\tint16_t*\n"
	                         "#read_id\tread_group\tdigitisation\toffset\trange\tsampling_rate\tlen_raw_signal\traw_signal\n";
	if (bin) {
		unsigned char h[BLOW5_HDR_SIZE_AT]; memset(h, 0, sizeof(h));
		memcpy(h, BLOW5_MAGIC, 6); h[6] = 0; h[7] = 2; h[8] = 0; h[9] = (unsigned char)record_press;
		const uint32_t groups = 1; memcpy(h + 10, &groups, 4); h[14] = (unsigned char)signal_press;
		o.put(h, sizeof(h));
		const std::string body = attrs + cols;
		o.val((uint32_t)body.size()); o.put(body.data(), body.size());
	} else {
		const std::string top = "#slow5_version\t0.2.0\n#num_read_groups\t1\n";
		o.put(top.data(), top.size()); o.put(attrs.data(), attrs.size()); o.put(cols.data(), cols.size());
	}
	std::vector<uint8_t> rec, sig, packed;
	std::string line;
	for (uint32_t i = 0; i < n && o.ok; ++i) {
		if (bin) {
			const size_t idl = strlen(names[i]);
			if (idl > 0xffff) { fclose(fp); rh_set_error("rh_slow5_write: read id too long"); return RH_ERR_ARG; }
			rec.clear();
			auto push = [&](const void *q, size_t b) { const uint8_t *c = (const uint8_t *)q; rec.insert(rec.end(), c, c + b); };
			const uint16_t l16 = (uint16_t)idl; const uint32_t rg = 0;
			push(&l16, 2); push(names[i], idl); push(&rg, 4);
			push(&digitisation[i], 8); push(&offset[i], 8); push(&range[i], 8); push(&sampling_rate, 8);
			if (signal_press == SIG_SVB_ZD) {
				svbzd_encode(raw[i], (uint32_t)raw_len[i], sig);
				const uint64_t sb = sig.size(); push(&sb, 8); push(sig.data(), sig.size());
			} else { const uint64_t sl = raw_len[i]; push(&sl, 8); push(raw[i], raw_len[i] * 2); }
			const std::vector<uint8_t> *body = &rec;
			if (record_press == PRESS_ZLIB) { if (!zlib_deflate(rec.data(), rec.size(), packed)) { o.ok = false; break; } body = &packed; }
			else if (record_press == PRESS_ZSTD) {
				const zstd_api &z = zstd();
				packed.resize(z.compress_bound(rec.size()));
				const size_t got = z.compress(packed.data(), packed.size(), rec.data(), rec.size(), 1); /* SLOW5_ZSTD_COMPRESS_LEVEL */
				if (z.is_error(got)) { o.ok = false; break; }
				packed.resize(got); body = &packed;
			}
			o.val((uint64_t)body->size()); o.put(body->data(), body->size());
		} else {
			line.clear();
			char head[512];
			snprintf(head, sizeof(head), "%s\t0\t%.17g\t%.17g\t%.17g\t%.17g\t%llu\t", names[i], digitisation[i], offset[i], range[i], sampling_rate, (unsigned long long)raw_len[i]);
			line = head;
			for (uint64_t j = 0; j < raw_len[i]; ++j) { snprintf(num, sizeof(num), j ? ",%d" : "%d", (int)raw[i][j]); line += num; }
			line += '\n';
			o.put(line.data(), line.size());
		}
	}
	if (bin) o.put(BLOW5_EOF, 5);
	const bool ok = o.ok && fflush(fp) == 0;
	fclose(fp);
	if (!ok) { rh_set_error("write to %s failed", path); return RH_ERR_IO; }
	return RH_OK;
}
