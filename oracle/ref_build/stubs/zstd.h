/* Minimal prototypes of the zstd functions the reference's mdoc circuit cache uses
 * (lib/circuits/mdoc/mdoc_decompress.cc, mdoc_generate_circuit.cc); the image has
 * libzstd.so.1 but no development header.  Test infrastructure only. */
#ifndef LF_STUB_ZSTD_H_
#define LF_STUB_ZSTD_H_
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
size_t ZSTD_compress(void* dst, size_t dstCapacity, const void* src, size_t srcSize, int compressionLevel);
size_t ZSTD_decompress(void* dst, size_t dstCapacity, const void* src, size_t compressedSize);
unsigned ZSTD_isError(size_t code);
const char* ZSTD_getErrorName(size_t code);
size_t ZSTD_compressBound(size_t srcSize);
unsigned long long ZSTD_getFrameContentSize(const void* src, size_t srcSize);
#ifdef __cplusplus
}
#endif
#endif
