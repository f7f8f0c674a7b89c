/*
 * zstd.h — TEST INFRASTRUCTURE ONLY.  This image ships libzstd.so.1 without its development header; these are the
 * five stable public entry points (zstd 1.x ABI) the reference's slow5lib uses when built with zstd=1
 * (extern/slow5lib/src/slow5_press.c:1146-1200), so that oracle/Makefile can compile that code path of the
 * reference and the checkers can read and write zstd-compressed BLOW5.
 */
#ifndef RH_ORACLE_ZSTD_SHIM_H
#define RH_ORACLE_ZSTD_SHIM_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
size_t ZSTD_compress(void *dst, size_t dstCapacity, const void *src, size_t srcSize, int compressionLevel);
size_t ZSTD_decompress(void *dst, size_t dstCapacity, const void *src, size_t compressedSize);
size_t ZSTD_compressBound(size_t srcSize);
unsigned ZSTD_isError(size_t code);
unsigned long long ZSTD_getFrameContentSize(const void *src, size_t srcSize);
#define ZSTD_CONTENTSIZE_UNKNOWN (0ULL - 1)
#define ZSTD_CONTENTSIZE_ERROR (0ULL - 2)
#ifdef __cplusplus
}
#endif
#endif
