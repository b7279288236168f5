/* On-disk cache of built BLAS data: the reference rebuilds every BLAS at start-up
 * (BVH.BlasesBuild, Source/Bvh/BVH.cs:300-470; Readme.md:515-522 reports 30 ms - 1.1 s per model); this file format lets
 * the host skip the SweepSAH build when the source geometry and build settings are unchanged.
 *
 * Layout: 64-byte header, array directory, then the arrays at 64-byte aligned offsets (so a mapped file can be handed to
 * idkpt_set_scene without a copy). Every array carries an FNV-1a-64 checksum; `SourceKey` is the caller's hash of the
 * inputs (positions, indices, build settings, builder version). */
#ifndef IDKHOST_CACHE_H
#define IDKHOST_CACHE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define IDKHOST_CACHE_MAGIC 0x3148564249444b49ull /* "IKDIBVH1" little endian */
#define IDKHOST_CACHE_VERSION 1u
#define IDKHOST_CACHE_MAX_ARRAYS 16

typedef enum IdkHostCacheArrayId {
    IDKHOST_CACHE_BLAS_NODES = 1,      /* GpuBlasNode, 32 B */
    IDKHOST_CACHE_BLAS_TRIANGLES = 2,  /* GpuBlasTriangle, 16 B */
    IDKHOST_CACHE_BLAS_DESCS = 3,      /* GpuBlasDesc, 40 B */
    IDKHOST_CACHE_USER = 100           /* ids >= 100 are free for the host (e.g. build statistics) */
} IdkHostCacheArrayId;

typedef struct IdkHostCacheArray {
    uint32_t    Id;
    uint32_t    ElemSize;
    uint64_t    Count;
    const void* Data;
} IdkHostCacheArray;

typedef struct IdkHostCacheView IdkHostCacheView;

enum { IDKHOST_CACHE_OK = 0, IDKHOST_CACHE_ERR_IO = -1, IDKHOST_CACHE_ERR_FORMAT = -2, IDKHOST_CACHE_ERR_KEY = -3, IDKHOST_CACHE_ERR_CHECKSUM = -4,
       IDKHOST_CACHE_ERR_ARGUMENT = -5 };

uint64_t idkhost_hash64(const void* data, uint64_t bytes, uint64_t seed);   /* FNV-1a-64, chainable through seed (0 = offset basis) */
/* Writes to `path` atomically (temp file + rename). */
int idkhost_cache_save(const char* path, uint64_t source_key, const IdkHostCacheArray* arrays, uint32_t array_count);
/* Maps the file read-only, checks magic/version/key and every checksum. */
int idkhost_cache_open(const char* path, uint64_t expect_source_key, IdkHostCacheView** out);
/* Pointer into the mapping (valid until close), or NULL if the id is absent. */
const void* idkhost_cache_array(const IdkHostCacheView* view, uint32_t id, uint32_t* elem_size, uint64_t* count);
void idkhost_cache_close(IdkHostCacheView* view);

#ifdef __cplusplus
}
#endif
#endif
