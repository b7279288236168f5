// Block-compressed material textures: the engine's loader uploads BC7 (base colour / emissive / packed maps),
// BC5 (normal, metallic-roughness with IDK_BC5_normal_metallicRoughness) and BC4 (transmission) KTX2 levels straight to GL
// (ModelLoader.cs:954-968) and the texture unit decodes them. Here the blocks are decoded ONCE at upload into the texel
// arrays the explicit fp32 sampler reads (idk_device.cuh tex_sample_raw): one thread per 4x4 block.
//
//   BC7 (BPTC)  OpenGL 4.6 core spec "BPTC compressed texture image formats" / Direct3D 11 BC7: 8 modes, partition and
//               anchor tables in include/idk_bc7_tables.h, integer interpolation ((64-w)*e0 + w*e1 + 32) >> 6 -> exact RGBA8;
//               a reserved-mode block (first byte 0) decodes to (0,0,0,0).
//   BC4 / BC5   RGTC1 / RGTC2 (EXT_texture_compression_rgtc): palette on normalised floats, (a*R0 + b*R1) / 7 (or / 5),
//               evaluated in fp32 left to right -> R32F / RG32F texels (the hardware keeps more than 8 bits too).
//
// Checked against tests/bcn_ref.py (independent Python decoders, themselves pinned to Pillow's C decoder by
// tests/golden/bcn_blocks.npz) in tests/test_textures.py.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define IDK_BC7_TABLE_QUALIFIER static __device__ const
#include "../../include/idk_bc7_tables.h"

struct Bc7Bits {
    uint64_t lo, hi;
    uint32_t pos;
    __device__ __forceinline__ uint32_t get(uint32_t n) {
        if (n == 0) return 0u;
        uint64_t v;
        if (pos >= 64) v = hi >> (pos - 64);
        else if (pos + n <= 64) v = lo >> pos;
        else v = (lo >> pos) | (hi << (64 - pos));
        pos += n;
        return (uint32_t)(v & ((1ull << n) - 1ull));
    }
};

__device__ __forceinline__ uint32_t bc7_weight(uint32_t bits, uint32_t i) {
    // 2-bit {0,21,43,64}, 3-bit {0,9,18,27,37,46,55,64}, 4-bit {0,4,9,13,17,21,26,30,34,38,43,47,51,55,60,64}
    const unsigned char w2[4] = {0, 21, 43, 64};
    const unsigned char w3[8] = {0, 9, 18, 27, 37, 46, 55, 64};
    const unsigned char w4[16] = {0, 4, 9, 13, 17, 21, 26, 30, 34, 38, 43, 47, 51, 55, 60, 64};
    return bits == 2 ? w2[i] : (bits == 3 ? w3[i] : w4[i]);
}

// out: 16 texels, index = y * 4 + x
static __device__ void bc7_decode_block(const uint8_t* block, uchar4* out) {
    Bc7Bits b;
    b.lo = 0; b.hi = 0; b.pos = 0;
    for (int i = 0; i < 8; i++) { b.lo |= (uint64_t)block[i] << (8 * i); b.hi |= (uint64_t)block[8 + i] << (8 * i); }
    uint32_t mode = 0;
    while (mode < 8 && !((b.lo >> mode) & 1ull)) mode++;
    if (mode == 8) {
        for (int i = 0; i < 16; i++) out[i] = make_uchar4(0, 0, 0, 0);
        return;
    }
    // subsets, partition bits, rotation bits, index-selection bits, colour bits, alpha bits, endpoint p-bits, shared p-bits, index bits, second index bits
    const unsigned char M[8][10] = {{3, 4, 0, 0, 4, 0, 1, 0, 3, 0}, {2, 6, 0, 0, 6, 0, 0, 1, 3, 0}, {3, 6, 0, 0, 5, 0, 0, 0, 2, 0}, {2, 6, 0, 0, 7, 0, 1, 0, 2, 0},
                                    {1, 0, 2, 1, 5, 6, 0, 0, 2, 3}, {1, 0, 2, 0, 7, 8, 0, 0, 2, 2}, {1, 0, 0, 0, 7, 7, 1, 0, 4, 0}, {2, 6, 0, 0, 5, 5, 1, 0, 2, 0}};
    const uint32_t ns = M[mode][0], pb = M[mode][1], rb = M[mode][2], isb = M[mode][3], cb = M[mode][4], ab = M[mode][5];
    const uint32_t epb = M[mode][6], spb = M[mode][7], ib = M[mode][8], ib2 = M[mode][9];
    b.pos = mode + 1;
    const uint32_t part = b.get(pb), rot = b.get(rb), isel = b.get(isb);
    uint32_t ep[6][4];
    for (uint32_t c = 0; c < 3; c++)
        for (uint32_t e = 0; e < 2 * ns; e++) ep[e][c] = b.get(cb);
    for (uint32_t e = 0; e < 2 * ns; e++) ep[e][3] = ab ? b.get(ab) : 255u;
    uint32_t cbits = cb, abits = ab;
    if (epb) {
        for (uint32_t e = 0; e < 2 * ns; e++) {
            const uint32_t p = b.get(1);
            for (uint32_t c = 0; c < 3; c++) ep[e][c] = (ep[e][c] << 1) | p;
            if (ab) ep[e][3] = (ep[e][3] << 1) | p;
        }
        cbits++;
        if (ab) abits++;
    } else if (spb) {
        for (uint32_t s = 0; s < ns; s++) {
            const uint32_t p = b.get(1);
            for (uint32_t e = 2 * s; e < 2 * s + 2; e++)
                for (uint32_t c = 0; c < 3; c++) ep[e][c] = (ep[e][c] << 1) | p;
        }
        cbits++;
    }
    for (uint32_t e = 0; e < 2 * ns; e++) {
        for (uint32_t c = 0; c < 3; c++) { const uint32_t v = ep[e][c] << (8 - cbits); ep[e][c] = v | (v >> cbits); }
        if (ab) { const uint32_t v = ep[e][3] << (8 - abits); ep[e][3] = v | (v >> abits); }
    }
    const uint32_t anchor1 = ns == 2 ? IDK_BC7_ANCHOR2[part] : (ns == 3 ? IDK_BC7_ANCHOR3A[part] : 0u);
    const uint32_t anchor2 = ns == 3 ? IDK_BC7_ANCHOR3B[part] : 0u;
    uint32_t idx1[16], idx2[16];
    for (uint32_t i = 0; i < 16; i++) {
        const bool isAnchor = i == 0 || (ns >= 2 && i == anchor1) || (ns == 3 && i == anchor2);
        idx1[i] = b.get(isAnchor ? ib - 1 : ib);
    }
    for (uint32_t i = 0; i < 16; i++) idx2[i] = ib2 ? b.get(i == 0 ? ib2 - 1 : ib2) : 0u;
    for (uint32_t i = 0; i < 16; i++) {
        const uint32_t s = ns == 1 ? 0u : (ns == 2 ? IDK_BC7_PARTITION2[part][i] : IDK_BC7_PARTITION3[part][i]);
        const uint32_t* e0 = ep[2 * s];
        const uint32_t* e1 = ep[2 * s + 1];
        uint32_t ci, cw, ai, aw;
        if (ib2) {
            if (!isel) { ci = idx1[i]; cw = ib; ai = idx2[i]; aw = ib2; }
            else { ci = idx2[i]; cw = ib2; ai = idx1[i]; aw = ib; }
        } else { ci = idx1[i]; cw = ib; ai = idx1[i]; aw = ib; }
        const uint32_t wc = bc7_weight(cw, ci), wa = bc7_weight(aw, ai);
        uint32_t r = ((64 - wc) * e0[0] + wc * e1[0] + 32) >> 6;
        uint32_t g = ((64 - wc) * e0[1] + wc * e1[1] + 32) >> 6;
        uint32_t bl = ((64 - wc) * e0[2] + wc * e1[2] + 32) >> 6;
        uint32_t a = ((64 - wa) * e0[3] + wa * e1[3] + 32) >> 6;
        if (rot == 1) { const uint32_t t = r; r = a; a = t; }
        else if (rot == 2) { const uint32_t t = g; g = a; a = t; }
        else if (rot == 3) { const uint32_t t = bl; bl = a; a = t; }
        out[i] = make_uchar4((unsigned char)r, (unsigned char)g, (unsigned char)bl, (unsigned char)a);
    }
}

// RGTC1 block -> 16 floats
static __device__ void bc4_decode_block(const uint8_t* block, float* out) {
    const uint32_t r0 = block[0], r1 = block[1];
    const float R0 = (float)r0 / 255.0f, R1 = (float)r1 / 255.0f;
    float pal[8];
    pal[0] = R0; pal[1] = R1;
    if (r0 > r1) {
        for (int k = 1; k < 7; k++) pal[1 + k] = ((float)(7 - k) * R0 + (float)k * R1) / 7.0f;
    } else {
        for (int k = 1; k < 5; k++) pal[1 + k] = ((float)(5 - k) * R0 + (float)k * R1) / 5.0f;
        pal[6] = 0.0f; pal[7] = 1.0f;
    }
    uint64_t bits = 0;
    for (int i = 0; i < 6; i++) bits |= (uint64_t)block[2 + i] << (8 * i);
    for (int i = 0; i < 16; i++) out[i] = pal[(bits >> (3 * i)) & 7ull];
}

// kind: 0 = BC7 -> uchar4 texels, 1 = BC5 -> float2 texels, 2 = BC4 -> float texels
struct BcnDecodeArgs {
    const uint8_t* blocks;
    void* texels;
    int width, height;
    int kind;
};

static __global__ void __launch_bounds__(128) k_bcn_decode(BcnDecodeArgs a) {
    const int bw = (a.width + 3) / 4, bh = (a.height + 3) / 4;
    const int bIdx = blockIdx.x * blockDim.x + threadIdx.x;
    if (bIdx >= bw * bh) return;
    const int bx = bIdx % bw, by = bIdx / bw;
    if (a.kind == 0) {
        uchar4 t[16];
        bc7_decode_block(a.blocks + (size_t)bIdx * 16, t);
        for (int i = 0; i < 16; i++) {
            const int x = 4 * bx + (i & 3), y = 4 * by + (i >> 2);
            if (x < a.width && y < a.height) ((uchar4*)a.texels)[(size_t)y * a.width + x] = t[i];
        }
    } else if (a.kind == 1) {
        float r[16], g[16];
        bc4_decode_block(a.blocks + (size_t)bIdx * 16, r);
        bc4_decode_block(a.blocks + (size_t)bIdx * 16 + 8, g);
        for (int i = 0; i < 16; i++) {
            const int x = 4 * bx + (i & 3), y = 4 * by + (i >> 2);
            if (x < a.width && y < a.height) ((float2*)a.texels)[(size_t)y * a.width + x] = make_float2(r[i], g[i]);
        }
    } else {
        float r[16];
        bc4_decode_block(a.blocks + (size_t)bIdx * 8, r);
        for (int i = 0; i < 16; i++) {
            const int x = 4 * bx + (i & 3), y = 4 * by + (i >> 2);
            if (x < a.width && y < a.height) ((float*)a.texels)[(size_t)y * a.width + x] = r[i];
        }
    }
}
