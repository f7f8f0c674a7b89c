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
