// Host side of the material texture table (include/idkpt.h IdkPtTextureDesc), shared by the path tracer and the voxeliser:
// validation, packing layout (all base levels in one allocation, 256-byte aligned) and the sRGB decode table.
#pragma once
#include <math.h>
#include <stdint.h>
#include <vector>

#include "../../include/idkpt.h"

static inline const char* idk_validate_textures(const IdkPtSceneDesc* s) {
    if (s->TextureCount && !s->Textures) return "TextureCount without Textures";
    for (uint64_t i = 0; i < s->TextureCount; i++) {
        const IdkPtTextureDesc& t = s->Textures[i];
        if (!t.Pixels || t.Width < 1 || t.Height < 1 || t.Width > 16384 || t.Height > 16384) return "texture without pixels or with an invalid size";
        if (t.Format != IDKPT_TEX_RGBA8_UNORM && t.Format != IDKPT_TEX_RGBA8_SRGB) return "texture format not supported (RGBA8 unorm / sRGB only; transcode BCn on the host)";
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

// Byte offset of every texture inside the packed pixel allocation; off[TextureCount] = total bytes.
static inline std::vector<size_t> idk_texture_offsets(const IdkPtSceneDesc* s) {
    std::vector<size_t> off(s->TextureCount + 1, 0);
    for (uint64_t i = 0; i < s->TextureCount; i++) off[i + 1] = off[i] + ((((size_t)s->Textures[i].Width * s->Textures[i].Height * 4) + 255) & ~(size_t)255);
    return off;
}

// GL_SRGB8 decode (OpenGL 4.6 spec 8.24), evaluated in double and rounded once.
static inline void idk_srgb_lut(float lut[256]) {
    for (int i = 0; i < 256; i++) {
        const double cs = i / 255.0;
        lut[i] = (float)(cs <= 0.04045 ? cs / 12.92 : pow((cs + 0.055) / 1.055, 2.4));
    }
}
