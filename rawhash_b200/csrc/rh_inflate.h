/*
 * rh_inflate.h — a DEFLATE / zlib-stream decoder for BLOW5 records (host code, header only).
 *
 * slow5tools writes BLOW5 with zlib-compressed records by default, one zlib stream per record
 * (ptr_compress_zlib / ptr_depress_zlib_solo, extern/slow5lib/src/slow5_press.c:809-982); the payload is mostly
 * streamvbyte bytes, i.e. high-entropy literals, on which zlib's one-symbol-per-iteration inflate delivers
 * ≈80 MB/s per core and becomes the bound of step 0 once mapping runs on the GPU.  This decoder
 * follows RFC 1950/1951 with the usual fast-path structure: a 64-bit bit buffer refilled with one unaligned load,
 * an 11-bit primary table (+ sub-tables) for literal/length codes and an 8-bit one for distances, up to three
 * literals per refill, word-wise match copies.  The Adler-32 trailer is verified.
 *
 * The caller guarantees RH_INFLATE_SLACK readable bytes after the input (rh_io.cpp: the file mapping, or a padded
 * copy); the output vector grows on demand.
 */
#ifndef RH_INFLATE_H
#define RH_INFLATE_H

#include <stdint.h>
#include <string.h>
#include <vector>
#include <zlib.h> /* adler32() only */

namespace rhz {

enum { RH_INFLATE_SLACK = 64 };

namespace detail {

enum { K_LITERAL = 0, K_BASE = 1, K_END = 2, K_SUBTABLE = 3, K_INVALID = 4 };
struct entry_t { uint16_t value; uint8_t bits; uint8_t op; }; /* op = kind << 5 | extra bits (or sub-table index bits) */
static inline entry_t mk(unsigned value, unsigned bits, unsigned kind, unsigned extra) { entry_t e; e.value = (uint16_t)value; e.bits = (uint8_t)bits; e.op = (uint8_t)(kind << 5 | extra); return e; }
static inline unsigned kind_of(entry_t e) { return e.op >> 5; }
static inline unsigned extra_of(entry_t e) { return e.op & 31u; }

enum { LIT_TBITS = 11, DIST_TBITS = 8, MAX_CODE_BITS = 15, N_LITLEN = 288, N_DIST = 32 };

static const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

static inline unsigned reverse_bits(unsigned code, unsigned len)
{
	unsigned r = 0;
	for (unsigned i = 0; i < len; ++i) { r = (r << 1) | (code & 1u); code >>= 1; }
	return r;
}

/* what a symbol decodes to */
static inline entry_t symbol_entry(bool litlen, unsigned sym, unsigned bits)
{
	if (litlen) {
		if (sym < 256) return mk(sym, bits, K_LITERAL, 0);
		if (sym == 256) return mk(0, bits, K_END, 0);
		if (sym - 257 < 29) return mk(LEN_BASE[sym - 257], bits, K_BASE, LEN_EXTRA[sym - 257]);
		return mk(0, bits, K_INVALID, 0);
	}
	if (sym < 30) return mk(DIST_BASE[sym], bits, K_BASE, DIST_EXTRA[sym]);
	return mk(0, bits, K_INVALID, 0);
}

/* Canonical Huffman table over bit-reversed codes (deflate packs codes starting at the least significant bit).
 * Returns false for an over-subscribed code.  Unused slots stay K_INVALID (incomplete codes are legal for a
 * distance tree with a single code, RFC 1951 §3.2.7). */
static bool build_table(const uint8_t *lens, unsigned n_syms, bool litlen, unsigned tbits, std::vector<entry_t> &tab)
{
	unsigned count[MAX_CODE_BITS + 1] = {0}, first[MAX_CODE_BITS + 2] = {0};
	for (unsigned s = 0; s < n_syms; ++s) ++count[lens[s]];
	count[0] = 0;
	unsigned code = 0; int64_t space = 1;
	for (unsigned l = 1; l <= MAX_CODE_BITS; ++l) {
		code = (code + count[l - 1]) << 1; first[l] = code;
		space = (space << 1) - count[l];
		if (space < 0) return false;
	}
	tab.assign((size_t)1 << tbits, mk(0, 0, K_INVALID, 0));
	unsigned next[MAX_CODE_BITS + 1];
	for (unsigned l = 0; l <= MAX_CODE_BITS; ++l) next[l] = first[l];
	/* pass 1: longest code behind every primary slot that needs a sub-table */
	std::vector<uint8_t> sub_bits;
	bool any_long = false;
	for (unsigned l = tbits + 1; l <= MAX_CODE_BITS; ++l) any_long = any_long || count[l];
	if (any_long) {
		sub_bits.assign((size_t)1 << tbits, 0);
		unsigned nx[MAX_CODE_BITS + 1];
		for (unsigned l = 0; l <= MAX_CODE_BITS; ++l) nx[l] = first[l];
		for (unsigned s = 0; s < n_syms; ++s) {
			const unsigned l = lens[s];
			if (!l) continue;
			const unsigned c = nx[l]++;
			if (l > tbits) { const unsigned slot = reverse_bits(c, l) & ((1u << tbits) - 1); if (l - tbits > sub_bits[slot]) sub_bits[slot] = (uint8_t)(l - tbits); }
		}
		for (size_t slot = 0; slot < sub_bits.size(); ++slot)
			if (sub_bits[slot]) {
				tab[slot] = mk((unsigned)tab.size(), tbits, K_SUBTABLE, sub_bits[slot]);
				tab.resize(tab.size() + ((size_t)1 << sub_bits[slot]), mk(0, 0, K_INVALID, 0));
			}
		if (tab.size() > 0xffff) return false;
	}
	/* pass 2: fill */
	for (unsigned s = 0; s < n_syms; ++s) {
		const unsigned l = lens[s];
		if (!l) continue;
		const unsigned r = reverse_bits(next[l]++, l);
		if (l <= tbits) {
			const entry_t e = symbol_entry(litlen, s, l);
			for (unsigned i = r; i < (1u << tbits); i += 1u << l) tab[i] = e;
		} else {
			const entry_t p = tab[r & ((1u << tbits) - 1)];
			const unsigned sb = extra_of(p), rest = l - tbits;
			const entry_t e = symbol_entry(litlen, s, rest);
			for (unsigned i = r >> tbits; i < (1u << sb); i += 1u << rest) tab[p.value + i] = e;
		}
	}
	return true;
}

static inline uint64_t load64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; } /* little endian host */

struct stream_t {
	const uint8_t *in, *in_end;
	uint64_t bitbuf = 0; unsigned bitcnt = 0;
	/* one unaligned load tops the buffer up to >= 56 valid bits; bytes are re-read at the same bit positions */
	inline void refill() { bitbuf |= load64(in) << bitcnt; in += (63 - bitcnt) >> 3; bitcnt |= 56; }
	inline unsigned peek(unsigned n) const { return (unsigned)(bitbuf & ((1ull << n) - 1)); }
	inline void drop(unsigned n) { bitbuf >>= n; bitcnt -= n; }
	inline unsigned take(unsigned n) { const unsigned v = peek(n); drop(n); return v; }
	inline const uint8_t *byte_pos() const { return in - (bitcnt >> 3); } /* first byte not yet consumed (after aligning) */
	inline bool overrun() const { return byte_pos() > in_end; }
};

static inline entry_t lookup(const std::vector<entry_t> &tab, unsigned tbits, stream_t &s)
{
	entry_t e = tab[s.peek(tbits)];
	if (kind_of(e) == K_SUBTABLE) { s.drop(tbits); e = tab[e.value + s.peek(extra_of(e))]; }
	return e;
}

struct tables_t { std::vector<entry_t> lit, dist; };

static bool fixed_tables(tables_t &t)
{
	uint8_t l[N_LITLEN], d[N_DIST];
	for (unsigned i = 0; i < N_LITLEN; ++i) l[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
	for (unsigned i = 0; i < N_DIST; ++i) d[i] = 5;
	return build_table(l, N_LITLEN, true, LIT_TBITS, t.lit) && build_table(d, N_DIST, false, DIST_TBITS, t.dist);
}

static bool dynamic_tables(stream_t &s, tables_t &t)
{
	static const uint8_t ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
	s.refill();
	const unsigned hlit = s.take(5) + 257, hdist = s.take(5) + 1, hclen = s.take(4) + 4;
	if (hlit > 286 || hdist > 30) return false;
	uint8_t pre[19] = {0};
	for (unsigned i = 0; i < hclen; ++i) { if (s.bitcnt < 3) { s.refill(); if (s.overrun()) return false; } pre[ORDER[i]] = (uint8_t)s.take(3); }
	std::vector<entry_t> ptab;
	{ /* the code-length code: at most 7 bits, decoded through the generic builder as "distance-like" symbols */
		unsigned count[8] = {0}, first[9] = {0}, code = 0; int space = 1;
		for (unsigned i = 0; i < 19; ++i) ++count[pre[i]];
		count[0] = 0;
		for (unsigned l = 1; l <= 7; ++l) { code = (code + count[l - 1]) << 1; first[l] = code; space = (space << 1) - (int)count[l]; if (space < 0) return false; }
		ptab.assign(128, mk(0, 0, K_INVALID, 0));
		for (unsigned sym = 0; sym < 19; ++sym) {
			const unsigned l = pre[sym];
			if (!l) continue;
			const unsigned r = reverse_bits(first[l]++, l);
			for (unsigned i = r; i < 128; i += 1u << l) ptab[i] = mk(sym, l, K_LITERAL, 0);
		}
	}
	uint8_t lens[N_LITLEN + N_DIST + 138];
	unsigned n = 0;
	while (n < hlit + hdist) {
		s.refill();
		if (s.overrun()) return false;
		const entry_t e = ptab[s.peek(7)];
		if (kind_of(e) != K_LITERAL) return false;
		s.drop(e.bits);
		const unsigned sym = e.value;
		if (sym < 16) { lens[n++] = (uint8_t)sym; continue; }
		unsigned rep; uint8_t v = 0;
		if (sym == 16) { if (!n) return false; v = lens[n - 1]; rep = 3 + s.take(2); }
		else if (sym == 17) rep = 3 + s.take(3);
		else rep = 11 + s.take(7);
		if (n + rep > hlit + hdist) return false;
		memset(lens + n, v, rep); n += rep;
	}
	if (lens[256] == 0) return false; /* no end-of-block code */
	return build_table(lens, hlit, true, LIT_TBITS, t.lit) && build_table(lens + hlit, hdist, false, DIST_TBITS, t.dist);
}

/* one Huffman-coded block; `out` may be re-allocated (positions are kept as offsets).  The bit buffer, the input
 * cursor and the output cursor live in locals: byte stores may alias anything reachable through a pointer, and
 * state kept in the stream object would be reloaded after every literal. */
static bool huffman_block(stream_t &s, const tables_t &t, std::vector<uint8_t> &out, size_t &n_out)
{
	const entry_t *const lit = t.lit.data(), *const dst_tab = t.dist.data();
	uint64_t bitbuf = s.bitbuf; unsigned bitcnt = s.bitcnt;
	const uint8_t *in = s.in; const uint8_t *const in_limit = s.in_end + RH_INFLATE_SLACK - 16; /* an iteration refills twice: 8-byte loads up to in + 15 */
	uint8_t *base = out.data(), *o = base + n_out, *o_limit = base + out.size() - (3 + 258 + 16);
	bool ok = false;
#define RHZ_REFILL() do { bitbuf |= load64(in) << bitcnt; in += (63 - bitcnt) >> 3; bitcnt |= 56; } while (0)
#define RHZ_DROP(n) do { bitbuf >>= (n); bitcnt -= (n); } while (0)
#define RHZ_LOOKUP(e, tab, tbits) do { e = tab[bitbuf & ((1u << (tbits)) - 1)]; \
		if (kind_of(e) == K_SUBTABLE) { RHZ_DROP(tbits); e = tab[e.value + (bitbuf & ((1u << extra_of(e)) - 1))]; } } while (0)
	for (;;) {
		if (o > o_limit) {
			const size_t at = (size_t)(o - base);
			out.resize(out.size() * 2 + 4096);
			base = out.data(); o = base + at; o_limit = base + out.size() - (3 + 258 + 16);
		}
		if (in > in_limit) break; /* ran past the input (and its readable slack) */
		RHZ_REFILL();
		entry_t e;
		RHZ_LOOKUP(e, lit, LIT_TBITS);
		if (kind_of(e) == K_LITERAL) { /* up to three literals on one refill: 3 x 15 bits <= 56 */
			*o++ = (uint8_t)e.value; RHZ_DROP(e.bits);
			RHZ_LOOKUP(e, lit, LIT_TBITS);
			if (kind_of(e) == K_LITERAL) {
				*o++ = (uint8_t)e.value; RHZ_DROP(e.bits);
				RHZ_LOOKUP(e, lit, LIT_TBITS);
				if (kind_of(e) == K_LITERAL) { *o++ = (uint8_t)e.value; RHZ_DROP(e.bits); continue; }
			}
			/* e is a length / end / invalid entry decoded from >= 26 valid bits (the lookup may already have dropped the
			 * primary bits of a sub-table code); top up before its extra bits and the distance */
			RHZ_REFILL();
		}
		const unsigned k = kind_of(e);
		if (k == K_END) { RHZ_DROP(e.bits); ok = true; break; }
		if (k != K_BASE) break;
		RHZ_DROP(e.bits);
		const unsigned xl = extra_of(e);
		const unsigned len = e.value + (unsigned)(bitbuf & ((1u << xl) - 1)); RHZ_DROP(xl);   /* <= 15 + 5 bits used: >= 36 left */
		entry_t d;
		RHZ_LOOKUP(d, dst_tab, DIST_TBITS);
		if (kind_of(d) != K_BASE) break;
		RHZ_DROP(d.bits);
		const unsigned xd = extra_of(d);
		const size_t dist = (size_t)d.value + (size_t)(bitbuf & ((1u << xd) - 1)); RHZ_DROP(xd); /* <= 15 + 13 bits */
		if (dist > (size_t)(o - base)) break;
		const uint8_t *src = o - dist;
		if (dist >= 8) { /* word copies, may write up to 7 bytes past the match (room is reserved above) */
			uint8_t *w = o, *const end = o + len;
			do { memcpy(w, src, 8); w += 8; src += 8; } while (w < end);
		} else if (dist == 1) {
			memset(o, *src, len);
		} else {
			for (unsigned i = 0; i < len; ++i) o[i] = src[i];
		}
		o += len;
	}
#undef RHZ_REFILL
#undef RHZ_DROP
#undef RHZ_LOOKUP
	s.bitbuf = bitbuf; s.bitcnt = bitcnt; s.in = in;
	n_out = (size_t)(o - base);
	return ok;
}

} // namespace detail

/* Inflates one zlib stream.  `in` must be readable for in_bytes + RH_INFLATE_SLACK bytes.  Returns false on malformed
 * input or an Adler-32 mismatch. */
static inline bool inflate_zlib(const uint8_t *in, size_t in_bytes, std::vector<uint8_t> &out)
{
	using namespace detail;
	if (in_bytes < 6) return false;
	const unsigned cmf = in[0], flg = in[1];
	if ((cmf & 15) != 8 || (cmf >> 4) > 7 || ((cmf << 8) | flg) % 31 != 0 || (flg & 0x20)) return false; /* deflate, <= 32 KiB window, no preset dictionary */
	stream_t s; s.in = in + 2; s.in_end = in + in_bytes;
	if (out.size() < in_bytes * 3 + 4096) out.resize(in_bytes * 3 + 4096);
	size_t n_out = 0;
	tables_t dyn;
	static const tables_t fixed = []() { tables_t t; fixed_tables(t); return t; }();
	for (bool last = false; !last;) {
		s.refill();
		if (s.overrun()) return false;
		last = s.take(1) != 0;
		const unsigned type = s.take(2);
		if (type == 0) { /* stored: skip to the byte boundary, LEN, ~LEN, bytes */
			s.drop(s.bitcnt & 7);
			const uint8_t *p = s.byte_pos();
			if (p + 4 > s.in_end) return false;
			const unsigned len = p[0] | p[1] << 8, nlen = p[2] | p[3] << 8;
			if ((len ^ nlen) != 0xffff || p + 4 + len > s.in_end) return false;
			if (out.size() - n_out < len) out.resize(n_out + len + out.size());
			memcpy(out.data() + n_out, p + 4, len); n_out += len;
			s.in = p + 4 + len; s.bitbuf = 0; s.bitcnt = 0;
		} else if (type == 1) {
			if (!huffman_block(s, fixed, out, n_out)) return false;
		} else if (type == 2) {
			if (!dynamic_tables(s, dyn) || !huffman_block(s, dyn, out, n_out)) return false;
		} else return false;
		if (s.overrun()) return false;
	}
	s.drop(s.bitcnt & 7);
	const uint8_t *p = s.byte_pos();
	if (p + 4 > s.in_end) return false;
	const uint32_t want = (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3];
	uint32_t have = (uint32_t)adler32(0L, Z_NULL, 0);
	for (size_t at = 0; at < n_out;) { const size_t c = n_out - at < (1u << 30) ? n_out - at : (1u << 30); have = (uint32_t)adler32(have, out.data() + at, (uInt)c); at += c; }
	if (have != want) return false;
	out.resize(n_out);
	return true;
}

} // namespace rhz

#endif /* RH_INFLATE_H */
