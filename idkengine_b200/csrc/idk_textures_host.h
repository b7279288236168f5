// Host side of the material texture table (include/idkpt.h IdkPtTextureDesc), shared by the path tracer and the voxeliser:
// validation, packing layout (all decoded level-0 images in one allocation, 256-byte aligned), BCn decode at upload and the
// sRGB decode table.
#pragma once
#include <math.h>
#include <stdint.h>
#include <vector>

#include "../../include/idkpt.h"
#include "idk_device.cuh"
#include "idk_bcn.cuh"

// storage kind of the decoded texels: 0 = RGBA8 (4 B), 1 = RG32F (8 B), 2 = R32F (4 B), 3 = RGBA32F (16 B); -1 = unknown format
static inline int idk_tex_kind(int fmt) {
    switch (fmt) {
        case IDKPT_TEX_RGBA8_UNORM: case IDKPT_TEX_RGBA8_SRGB: case IDKPT_TEX_BC7_UNORM: case IDKPT_TEX_BC7_SRGB: return 0;
        case IDKPT_TEX_BC5_RG_UNORM: case IDKPT_TEX_RG32F: return 1;
        case IDKPT_TEX_BC4_R_UNORM: case IDKPT_TEX_R32F: return 2;
        case IDKPT_TEX_RGBA32F: return 3;
        default: return -1;
    }
}
static inline size_t idk_tex_texel_bytes(int kind) { return kind == 0 ? 4 : (kind == 1 ? 8 : (kind == 2 ? 4 : 16)); }
static inline int idk_tex_block_bytes(int fmt) {   // 0 = not block compressed
    return (fmt == IDKPT_TEX_BC7_UNORM || fmt == IDKPT_TEX_BC7_SRGB || fmt == IDKPT_TEX_BC5_RG_UNORM) ? 16 : (fmt == IDKPT_TEX_BC4_R_UNORM ? 8 : 0);
}
// bytes the host provides at IdkPtTextureDesc.Pixels
static inline size_t idk_tex_source_bytes(const IdkPtTextureDesc& t) {
    const int bb = idk_tex_block_bytes(t.Format);
    if (bb) return (size_t)((t.Width + 3) / 4) * ((t.Height + 3) / 4) * bb;
    return (size_t)t.Width * t.Height * idk_tex_texel_bytes(idk_tex_kind(t.Format));
}

static inline const char* idk_validate_textures(const IdkPtSceneDesc* s) {
    if (s->TextureCount && !s->Textures) return "TextureCount without Textures";
    for (uint64_t i = 0; i < s->TextureCount; i++) {
        const IdkPtTextureDesc& t = s->Textures[i];
        if (!t.Pixels || t.Width < 1 || t.Height < 1 || t.Width > 16384 || t.Height > 16384) return "texture without pixels or with an invalid size";
        if (idk_tex_kind(t.Format) < 0) return "texture format not supported (RGBA8 unorm / sRGB, BC7, BC5, BC4, R/RG/RGBA32F)";
        if (t.Flags & ~(IDKPT_TEX_FLAG_R_FROM_B | IDKPT_TEX_FLAG_MAG_NEAREST)) return "unknown texture flag";
        for (int k = 0; k < 2; k++) {
            const int wm = k ? t.WrapT : t.WrapS;
            if (wm != 10497 && wm != 33071 && wm != 33648) return "texture wrap mode must be REPEAT, CLAMP_TO_EDGE or MIRRORED_REPEAT";
        }
    }
    for (uint64_t i = 0; i < s->MaterialCount; i++) {
        const GpuMaterial& m = s->Materials[i];
        const uint64_t h[5] = {m.BaseColorTexture, m.MetallicRoughnessTexture, m.NormalTexture, m.EmissiveTexture, m.TransmissionTexture};
        for (int k = 0; k < 5; k++) if (h[k] > s->TextureCount) return "material texture handle outside the texture table (0 = white, k = Textures[k-1])";
    }
    return nullptr;
}

// Byte offset of every decoded texture inside the packed texel allocation; off[TextureCount] = total bytes.
static inline std::vector<size_t> idk_texture_offsets(const IdkPtSceneDesc* s) {
    std::vector<size_t> off(s->TextureCount + 1, 0);
    for (uint64_t i = 0; i < s->TextureCount; i++) {
        const size_t bytes = (size_t)s->Textures[i].Width * s->Textures[i].Height * idk_tex_texel_bytes(idk_tex_kind(s->Textures[i].Format));
        off[i + 1] = off[i] + ((bytes + 255) & ~(size_t)255);
    }
    return off;
}

// Fill the packed texel allocation `dPixels` (off[count] bytes, device) and the records: uncompressed images are copied,
// block-compressed ones are staged and decoded by k_bcn_decode. Synchronises the stream before returning (host arrays are
// borrowed for the duration of the call only).
static inline cudaError_t idk_upload_texture_table(const IdkPtTextureDesc* textures, uint64_t count, const std::vector<size_t>& off, void* dPixels,
                                                   cudaStream_t stream, std::vector<TexRec>& recs) {
    recs.assign(std::max<uint64_t>(count, 1), TexRec{});
    size_t stagingBytes = 0;
    for (uint64_t i = 0; i < count; i++) if (idk_tex_block_bytes(textures[i].Format)) stagingBytes = std::max(stagingBytes, idk_tex_source_bytes(textures[i]));
    void* staging = nullptr;
    cudaError_t e = cudaSuccess;
    if (stagingBytes && (e = cudaMalloc(&staging, stagingBytes)) != cudaSuccess) return e;
    for (uint64_t i = 0; i < count && e == cudaSuccess; i++) {
        const IdkPtTextureDesc& t = textures[i];
        const int kind = idk_tex_kind(t.Format);
        char* dst = (char*)dPixels + off[i];
        if (idk_tex_block_bytes(t.Format)) {
            e = cudaMemcpyAsync(staging, t.Pixels, idk_tex_source_bytes(t), cudaMemcpyHostToDevice, stream);
            if (e != cudaSuccess) break;
            BcnDecodeArgs a;
            a.blocks = (const uint8_t*)staging; a.texels = dst; a.width = t.Width; a.height = t.Height;
            a.kind = t.Format == IDKPT_TEX_BC5_RG_UNORM ? 1 : (t.Format == IDKPT_TEX_BC4_R_UNORM ? 2 : 0);
            const int nBlocks = ((t.Width + 3) / 4) * ((t.Height + 3) / 4);
            k_bcn_decode<<<(nBlocks + 127) / 128, 128, 0, stream>>>(a);
            e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaStreamSynchronize(stream);     // the staging buffer is reused by the next texture
        } else {
            e = cudaMemcpyAsync(dst, t.Pixels, idk_tex_source_bytes(t), cudaMemcpyHostToDevice, stream);
        }
        recs[i].px = dst;
        recs[i].w = t.Width; recs[i].h = t.Height; recs[i].wrapS = t.WrapS; recs[i].wrapT = t.WrapT;
        recs[i].srgb = (t.Format == IDKPT_TEX_RGBA8_SRGB || t.Format == IDKPT_TEX_BC7_SRGB) ? 1 : 0;
        recs[i].kind = kind | ((t.Flags & IDKPT_TEX_FLAG_R_FROM_B) ? 256 : 0) | ((t.Flags & IDKPT_TEX_FLAG_MAG_NEAREST) ? 512 : 0);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (staging) cudaFree(staging);
    return e;
}

// GL_SRGB8 decode (OpenGL 4.6 spec 8.24), evaluated in double and rounded once.
static inline void idk_srgb_lut(float lut[256]) {
    for (int i = 0; i < 256; i++) {
        const double cs = i / 255.0;
        lut[i] = (float)(cs <= 0.04045 ? cs / 12.92 : pow((cs + 0.055) / 1.055, 2.4));
    }
}
