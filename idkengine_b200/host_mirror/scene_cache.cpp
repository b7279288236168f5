// On-disk BLAS cache (include/idkhost_cache.h). Plain POSIX I/O + mmap; no dependency on the builder.
#include "idkhost_cache.h"

#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <string>
#include <vector>

#define IDKHOST_API extern "C" __attribute__((visibility("default")))

namespace {

struct FileHeader {          // 64 bytes
    uint64_t magic;
    uint32_t version;
    uint32_t arrayCount;
    uint64_t sourceKey;
    uint64_t fileBytes;
    uint64_t directoryChecksum;
    uint8_t  pad[24];
};
static_assert(sizeof(FileHeader) == 64, "header must be 64 bytes");

struct DirEntry {            // 40 bytes
    uint32_t id, elemSize;
    uint64_t count, offset, checksum;
    uint64_t reserved;
};
static_assert(sizeof(DirEntry) == 40, "directory entry must be 40 bytes");

inline uint64_t align64(uint64_t v) { return (v + 63) & ~(uint64_t)63; }

bool write_all(int fd, const void* p, size_t n) {
    const char* c = (const char*)p;
    while (n) {
        const ssize_t w = write(fd, c, n);
        if (w <= 0) return false;
        c += w; n -= (size_t)w;
    }
    return true;
}

} // namespace

struct IdkHostCacheView {
    void* base = nullptr;
    size_t bytes = 0;
    std::vector<DirEntry> dir;
};

IDKHOST_API uint64_t idkhost_hash64(const void* data, uint64_t bytes, uint64_t seed) {
    uint64_t h = seed ? seed : 0xcbf29ce484222325ull;
    const uint8_t* p = (const uint8_t*)data;
    // 8 bytes per step (word-wise FNV-1a variant): the cache hashes hundreds of MB of geometry
    uint64_t i = 0;
    for (; i + 8 <= bytes; i += 8) {
        uint64_t w;
        memcpy(&w, p + i, 8);
        h = (h ^ w) * 0x100000001b3ull;
    }
    for (; i < bytes; i++) h = (h ^ p[i]) * 0x100000001b3ull;
    return h;
}

IDKHOST_API int idkhost_cache_save(const char* path, uint64_t sourceKey, const IdkHostCacheArray* arrays, uint32_t arrayCount) {
    if (!path || (!arrays && arrayCount) || arrayCount > IDKHOST_CACHE_MAX_ARRAYS) return IDKHOST_CACHE_ERR_ARGUMENT;
    std::vector<DirEntry> dir(arrayCount);
    uint64_t off = align64(sizeof(FileHeader) + (uint64_t)arrayCount * sizeof(DirEntry));
    for (uint32_t i = 0; i < arrayCount; i++) {
        if (arrays[i].Count && !arrays[i].Data) return IDKHOST_CACHE_ERR_ARGUMENT;
        for (uint32_t j = 0; j < i; j++) if (arrays[j].Id == arrays[i].Id) return IDKHOST_CACHE_ERR_ARGUMENT;
        dir[i] = DirEntry{arrays[i].Id, arrays[i].ElemSize, arrays[i].Count, off,
                          idkhost_hash64(arrays[i].Data, arrays[i].Count * arrays[i].ElemSize, 0), 0};
        off = align64(off + arrays[i].Count * arrays[i].ElemSize);
    }
    FileHeader h;
    memset(&h, 0, sizeof(h));
    h.magic = IDKHOST_CACHE_MAGIC; h.version = IDKHOST_CACHE_VERSION; h.arrayCount = arrayCount; h.sourceKey = sourceKey; h.fileBytes = off;
    h.directoryChecksum = idkhost_hash64(dir.data(), dir.size() * sizeof(DirEntry), 0);
    const std::string tmp = std::string(path) + ".tmp." + std::to_string((long)getpid());
    const int fd = open(tmp.c_str(), O_CREAT | O_TRUNC | O_WRONLY, 0644);
    if (fd < 0) return IDKHOST_CACHE_ERR_IO;
    bool ok = write_all(fd, &h, sizeof(h)) && write_all(fd, dir.data(), dir.size() * sizeof(DirEntry));
    uint64_t pos = sizeof(h) + dir.size() * sizeof(DirEntry);
    static const char zeros[64] = {0};
    for (uint32_t i = 0; ok && i < arrayCount; i++) {
        ok = write_all(fd, zeros, (size_t)(dir[i].offset - pos));
        ok = ok && write_all(fd, arrays[i].Data, (size_t)(arrays[i].Count * arrays[i].ElemSize));
        pos = dir[i].offset + arrays[i].Count * arrays[i].ElemSize;
    }
    ok = ok && write_all(fd, zeros, (size_t)(off - pos));
    ok = ok && fsync(fd) == 0;
    close(fd);
    if (!ok || rename(tmp.c_str(), path) != 0) { unlink(tmp.c_str()); return IDKHOST_CACHE_ERR_IO; }
    return IDKHOST_CACHE_OK;
}

IDKHOST_API int idkhost_cache_open(const char* path, uint64_t expectKey, IdkHostCacheView** out) {
    if (!path || !out) return IDKHOST_CACHE_ERR_ARGUMENT;
    *out = nullptr;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return IDKHOST_CACHE_ERR_IO;
    struct stat st;
    if (fstat(fd, &st) != 0 || (size_t)st.st_size < sizeof(FileHeader)) { close(fd); return st.st_size >= 0 ? IDKHOST_CACHE_ERR_FORMAT : IDKHOST_CACHE_ERR_IO; }
    void* base = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (base == MAP_FAILED) return IDKHOST_CACHE_ERR_IO;
    int rc = IDKHOST_CACHE_OK;
    IdkHostCacheView* v = new IdkHostCacheView();
    v->base = base; v->bytes = (size_t)st.st_size;
    do {
        FileHeader h;
        memcpy(&h, base, sizeof(h));
        if (h.magic != IDKHOST_CACHE_MAGIC || h.version != IDKHOST_CACHE_VERSION || h.arrayCount > IDKHOST_CACHE_MAX_ARRAYS || h.fileBytes != (uint64_t)st.st_size ||
            sizeof(FileHeader) + (uint64_t)h.arrayCount * sizeof(DirEntry) > (uint64_t)st.st_size) { rc = IDKHOST_CACHE_ERR_FORMAT; break; }
        if (h.sourceKey != expectKey) { rc = IDKHOST_CACHE_ERR_KEY; break; }
        v->dir.resize(h.arrayCount);
        memcpy(v->dir.data(), (const char*)base + sizeof(FileHeader), h.arrayCount * sizeof(DirEntry));
        if (idkhost_hash64(v->dir.data(), v->dir.size() * sizeof(DirEntry), 0) != h.directoryChecksum) { rc = IDKHOST_CACHE_ERR_CHECKSUM; break; }
        for (const DirEntry& e : v->dir) {
            const uint64_t bytes = e.count * e.elemSize;
            if ((e.offset & 63) || e.offset > (uint64_t)st.st_size || bytes > (uint64_t)st.st_size - e.offset || (e.elemSize && bytes / e.elemSize != e.count)) { rc = IDKHOST_CACHE_ERR_FORMAT; break; }
            if (idkhost_hash64((const char*)base + e.offset, bytes, 0) != e.checksum) { rc = IDKHOST_CACHE_ERR_CHECKSUM; break; }
        }
    } while (0);
    if (rc != IDKHOST_CACHE_OK) { idkhost_cache_close(v); return rc; }
    *out = v;
    return IDKHOST_CACHE_OK;
}

IDKHOST_API const void* idkhost_cache_array(const IdkHostCacheView* v, uint32_t id, uint32_t* elemSize, uint64_t* count) {
    if (!v) return nullptr;
    for (const DirEntry& e : v->dir)
        if (e.id == id) {
            if (elemSize) *elemSize = e.elemSize;
            if (count) *count = e.count;
            return (const char*)v->base + e.offset;
        }
    return nullptr;
}

IDKHOST_API void idkhost_cache_close(IdkHostCacheView* v) {
    if (!v) return;
    if (v->base) munmap(v->base, v->bytes);
    delete v;
}
