/*
 * slow5_tap.c — TEST INFRASTRUCTURE ONLY.  A thin C harness over the reference's vendored slow5lib
 * (extern/slow5lib, compiled from where it lies by oracle/Makefile into oracle/_ref/libslow5_tap.so) so that
 * tests can check rawhash_b200's own SLOW5/BLOW5 reader and writer (csrc/rh_io.cpp) against the library the
 * reference reads its signals with (src/rsig.c:170-207,478-533).  Never linked into the product.
 */
#include <stdint.h>
#include <stdlib.h>
#include <slow5/slow5.h>

typedef struct { slow5_file_t *sp; slow5_rec_t *rec; } s5tap_t;

void *s5tap_open(const char *path)
{
	slow5_file_t *sp = slow5_open(path, "r");
	if (!sp) return 0;
	s5tap_t *t = (s5tap_t *)calloc(1, sizeof(s5tap_t));
	t->sp = sp;
	return t;
}

/* 1 = a record was returned, 0 = end of file, <0 = slow5lib error */
int s5tap_next(void *h, const char **id, const int16_t **raw, uint64_t *len, double *digitisation, double *offset, double *range, double *sampling_rate)
{
	s5tap_t *t = (s5tap_t *)h;
	int ret = slow5_get_next(&t->rec, t->sp);
	if (ret < 0) return ret == SLOW5_ERR_EOF ? 0 : ret;
	*id = t->rec->read_id; *raw = t->rec->raw_signal; *len = t->rec->len_raw_signal;
	*digitisation = t->rec->digitisation; *offset = t->rec->offset; *range = t->rec->range; *sampling_rate = t->rec->sampling_rate;
	return 1;
}

void s5tap_close(void *h)
{
	s5tap_t *t = (s5tap_t *)h;
	if (!t) return;
	if (t->rec) slow5_rec_free(t->rec);
	slow5_close(t->sp);
	free(t);
}

/* Re-writes `in` as BLOW5 with the given compression (record: 0 none, 1 zlib, 2 zstd; signal: 0 none, 1 svb-zd)
 * through slow5lib's own writer.  Returns the number of records, -100 when zstd was asked for but not compiled in. */
int s5tap_convert(const char *in, const char *out, int record_press, int signal_press)
{
#ifndef SLOW5_USE_ZSTD
	if (record_press == 2) return -100;
#endif
	slow5_file_t *src = slow5_open(in, "r");
	if (!src) return -1;
	slow5_file_t *dst = slow5_open(out, "w");
	if (!dst) { slow5_close(src); return -2; }
	/* same header: attributes of every read group, auxiliary field table */
	slow5_hdr_t *tmp = dst->header;
	dst->header = src->header;
	enum slow5_press_method rp = record_press == 2 ? SLOW5_COMPRESS_ZSTD : record_press == 1 ? SLOW5_COMPRESS_ZLIB : SLOW5_COMPRESS_NONE;
	enum slow5_press_method sp = signal_press == 1 ? SLOW5_COMPRESS_SVB_ZD : SLOW5_COMPRESS_NONE;
	int n = -3;
	if (slow5_set_press(dst, rp, sp) == 0 && slow5_hdr_write(dst) >= 0) {
		slow5_rec_t *rec = NULL;
		n = 0;
		while (slow5_get_next(&rec, src) >= 0) { if (slow5_write(rec, dst) < 0) { n = -4; break; } ++n; }
		slow5_rec_free(rec);
	}
	dst->header = tmp;
	slow5_close(dst);
	slow5_close(src);
	return n;
}
