"""Golden vectors for the block-compressed texture decoders (BC7 / BC5 / BC4), generated from an INDEPENDENT implementation:
Pillow's C decoder (PIL DdsImagePlugin -> BcnDecode.c). Run here (Pillow is in this image, not a runtime dependency):

    python tests/golden/make_bcn_golden.py

writes
  include/idk_bc7_tables.h        BC7 partition tables and anchor (fix-up) indices, recovered by probing Pillow's decoder with
                                  crafted blocks (the tables are data of the BC7 / BPTC specification)
  tests/golden/bcn_blocks.npz     random + structured blocks of every BC7 mode, BC5 and BC4 with Pillow's decoded texels
"""
import io
import os
import struct

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
DXGI = {"bc7": 98, "bc5": 83, "bc4": 80}


def dds(fmt, width, height, payload):
    """Minimal DDS container with a DX10 header."""
    block = 16 if fmt in ("bc7", "bc5") else 8
    assert len(payload) == (width // 4) * (height // 4) * block
    DDSD = 0x1 | 0x2 | 0x4 | 0x1000 | 0x80000
    pf = struct.pack("<II4sIIIII", 32, 0x4, b"DX10", 0, 0, 0, 0, 0)
    hdr = struct.pack("<4sIIIIIII44s", b"DDS ", 124, DDSD, height, width, len(payload), 0, 1, b"\0" * 44) + pf + struct.pack("<IIIII", 0x1000, 0, 0, 0, 0)
    dx10 = struct.pack("<IIIII", DXGI[fmt], 3, 0, 1, 0)
    return hdr + dx10 + payload


def pillow_decode(fmt, blocks):
    """blocks: [N, block_bytes] uint8 -> [N, 4, 4, C] uint8 (C = 4 for bc7, 2 for bc5, 1 for bc4), texel (y, x) order."""
    n = len(blocks)
    w = 4 * n
    img = Image.open(io.BytesIO(dds(fmt, w, 4, blocks.tobytes())))
    img.load()
    a = np.asarray(img)
    if fmt == "bc7":
        a = np.asarray(img.convert("RGBA"))
        return a.reshape(4, n, 4, 4).transpose(1, 0, 2, 3).copy()
    if fmt == "bc5":
        a = np.asarray(img.convert("RGB"))[..., :2]
        return a.reshape(4, n, 4, 2).transpose(1, 0, 2, 3).copy()
    a = np.asarray(img.convert("L"))
    return a.reshape(4, n, 4, 1).transpose(1, 0, 2, 3).copy()


class Bits:
    def __init__(self):
        self.v, self.n = 0, 0

    def put(self, value, bits):
        self.v |= (value & ((1 << bits) - 1)) << self.n
        self.n += bits

    def bytes16(self):
        assert self.n <= 128
        return np.frombuffer(self.v.to_bytes(16, "little"), np.uint8)


def mode1_block(partition, ep, index_bits_all_ones=False):
    """Mode 1 (2 subsets, 6-bit endpoints, shared p-bit, 3-bit indices). ep[subset][end] = 6-bit grey value."""
    b = Bits()
    b.put(0b10, 2)
    b.put(partition, 6)
    for _c in range(3):
        for s in range(2):
            for e in range(2):
                b.put(ep[s][e], 6)
    b.put(0, 1); b.put(0, 1)
    b.put((1 << 46) - 1 if index_bits_all_ones else 0, 46)
    return b.bytes16()


def mode2_block(partition, ep, index_bits_all_ones=False):
    """Mode 2 (3 subsets, 5-bit endpoints, no p-bits, 2-bit indices)."""
    b = Bits()
    b.put(0b100, 3)
    b.put(partition, 6)
    for _c in range(3):
        for s in range(3):
            for e in range(2):
                b.put(ep[s][e], 5)
    b.put((1 << 29) - 1 if index_bits_all_ones else 0, 29)
    return b.bytes16()


def derive_tables():
    # --- partitions: subset k gets the constant grey level k * step, all indices 0 -> the texel colour names its subset
    p2 = np.zeros((64, 16), np.uint8)
    blocks = np.stack([mode1_block(p, [[0, 0], [63, 63]]) for p in range(64)])
    dec = pillow_decode("bc7", blocks)
    p2[:] = (dec[..., 0] > 127).reshape(64, 16)
    p3 = np.zeros((64, 16), np.uint8)
    blocks = np.stack([mode2_block(p, [[0, 0], [15, 15], [31, 31]]) for p in range(64)])
    dec = pillow_decode("bc7", blocks)[..., 0].reshape(64, 16).astype(np.int32)
    p3[:] = np.where(dec < 60, 0, np.where(dec < 190, 1, 2))
    # --- anchors: with every stored index bit = 1 the anchor texel of a subset (one bit fewer, implicit leading 0) is the only
    # texel of the subset that does not reach the second endpoint
    a2 = np.zeros(64, np.uint8)
    blocks = np.stack([mode1_block(p, [[0, 63], [0, 63]], True) for p in range(64)])
    dec = pillow_decode("bc7", blocks)[..., 0].reshape(64, 16)
    for p in range(64):
        low = [i for i in range(16) if dec[p, i] < 250]
        assert low[0] == 0 and p2[p, 0] == 0, (p, low)
        second = [i for i in low if p2[p, i] == 1]
        assert len(second) == 1 and len(low) == 2, (p, low)
        a2[p] = second[0]
    a3a, a3b = np.zeros(64, np.uint8), np.zeros(64, np.uint8)
    blocks = np.stack([mode2_block(p, [[0, 31]] * 3, True) for p in range(64)])
    dec = pillow_decode("bc7", blocks)[..., 0].reshape(64, 16)
    for p in range(64):
        low = [i for i in range(16) if dec[p, i] < 250]
        s1 = [i for i in low if p3[p, i] == 1]
        s2 = [i for i in low if p3[p, i] == 2]
        assert low[0] == 0 and len(low) == 3 and len(s1) == 1 and len(s2) == 1, (p, low)
        a3a[p], a3b[p] = s1[0], s2[0]
    return p2, p3, a2, a3a, a3b


def write_header(p2, p3, a2, a3a, a3b):
    def rows(t):
        return ",\n".join("    {" + ",".join(str(int(v)) for v in r) + "}" for r in t)

    def row(t):
        return ",".join(str(int(v)) for v in t)
    text = f"""/* BC7 (BPTC) partition tables and anchor indices: data of the BC7 specification (OpenGL 4.6 core, section 8.? "BPTC";
 * Direct3D 11 BC7 format), recovered by probing an independent decoder -- tests/golden/make_bcn_golden.py. GENERATED FILE. */
#ifndef IDK_BC7_TABLES_H
#define IDK_BC7_TABLES_H
#ifndef IDK_BC7_TABLE_QUALIFIER
#define IDK_BC7_TABLE_QUALIFIER static const
#endif
IDK_BC7_TABLE_QUALIFIER unsigned char IDK_BC7_PARTITION2[64][16] = {{
{rows(p2)}
}};
IDK_BC7_TABLE_QUALIFIER unsigned char IDK_BC7_PARTITION3[64][16] = {{
{rows(p3)}
}};
IDK_BC7_TABLE_QUALIFIER unsigned char IDK_BC7_ANCHOR2[64] = {{{row(a2)}}};        /* 2 subsets: anchor of subset 1 */
IDK_BC7_TABLE_QUALIFIER unsigned char IDK_BC7_ANCHOR3A[64] = {{{row(a3a)}}};      /* 3 subsets: anchor of subset 1 */
IDK_BC7_TABLE_QUALIFIER unsigned char IDK_BC7_ANCHOR3B[64] = {{{row(a3b)}}};      /* 3 subsets: anchor of subset 2 */
#endif
"""
    open(os.path.join(REPO, "include", "idk_bc7_tables.h"), "w").write(text)


def golden_blocks():
    rng = np.random.default_rng(0xBC7)
    blocks = rng.integers(0, 256, (6000, 16), dtype=np.uint8)
    # force every mode to appear often: mode m = m zero bits then a one in byte 0
    for i in range(len(blocks)):
        m = i % 10
        if m < 8:
            blocks[i, 0] = (int(blocks[i, 0]) & (0xFF ^ ((1 << (m + 1)) - 1))) | (1 << m)
        elif m == 8:
            blocks[i, 0] = 0          # reserved mode: decodes to zeros
    bc7 = pillow_decode("bc7", blocks)
    b5 = rng.integers(0, 256, (2000, 16), dtype=np.uint8)
    b5[::7, 1] = b5[::7, 0]           # r0 == r1 edge case
    b5[1::7, 9] = b5[1::7, 8]
    bc5 = pillow_decode("bc5", b5)
    b4 = rng.integers(0, 256, (2000, 8), dtype=np.uint8)
    b4[::5, 1] = b4[::5, 0]
    bc4 = pillow_decode("bc4", b4)
    np.savez_compressed(os.path.join(HERE, "bcn_blocks.npz"), bc7_blocks=blocks, bc7_texels=bc7, bc5_blocks=b5, bc5_texels=bc5,
                        bc4_blocks=b4, bc4_texels=bc4)
    return blocks, bc7


if __name__ == "__main__":
    t = derive_tables()
    write_header(*t)
    golden_blocks()
    print("partition 2-subset row 0:", t[0][0].tolist(), "anchors", t[2][:8].tolist())
