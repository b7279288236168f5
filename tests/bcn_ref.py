"""Reference decoders (plain Python / numpy, test infrastructure) for the block-compressed texture formats the engine's loader
produces (ModelLoader.cs:954-968: BC7 base colour / emissive, BC5 normal / metallic-roughness, BC4 transmission), written
from the format definitions (OpenGL 4.6 core spec, appendix "BPTC" / "RGTC"; Direct3D 11 BC7) and pinned to Pillow's
independent C decoder by tests/golden/bcn_blocks.npz. They decode the textures handed to the CPU oracle, so that the CUDA
decoders in idkengine_b200/csrc/idk_bcn.cuh are checked by an implementation that shares no code with them.

Also here: two small ENCODERS used to build test textures (BC7 mode 6 and BC5 / BC4 range fit) -- any valid block stream
would do, the parity tests additionally use random blocks, which exercise every mode.
"""
import os
import re

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tables():
    text = open(os.path.join(REPO, "include", "idk_bc7_tables.h")).read()
    out = {}
    for name in ("PARTITION2", "PARTITION3", "ANCHOR2", "ANCHOR3A", "ANCHOR3B"):
        body = re.search(r"IDK_BC7_%s(?:\[\d+\])+ = \{(.*?)\};" % name, text, re.S).group(1)
        vals = np.array([int(v) for v in re.findall(r"\d+", body)], np.int64)
        out[name] = vals.reshape(64, 16) if name.startswith("PART") else vals
    return out


_T = None
# mode: subsets, partition bits, rotation bits, index-selection bits, colour bits, alpha bits, endpoint p-bits, shared p-bits, index bits, second index bits
_MODES = [(3, 4, 0, 0, 4, 0, 1, 0, 3, 0), (2, 6, 0, 0, 6, 0, 0, 1, 3, 0), (3, 6, 0, 0, 5, 0, 0, 0, 2, 0), (2, 6, 0, 0, 7, 0, 1, 0, 2, 0),
          (1, 0, 2, 1, 5, 6, 0, 0, 2, 3), (1, 0, 2, 0, 7, 8, 0, 0, 2, 2), (1, 0, 0, 0, 7, 7, 1, 0, 4, 0), (2, 6, 0, 0, 5, 5, 1, 0, 2, 0)]
_WEIGHTS = {2: [0, 21, 43, 64], 3: [0, 9, 18, 27, 37, 46, 55, 64], 4: [0, 4, 9, 13, 17, 21, 26, 30, 34, 38, 43, 47, 51, 55, 60, 64]}


def decode_bc7_block(block):
    """16 bytes -> [4, 4, 4] uint8 (texel [y, x] = RGBA)."""
    global _T
    if _T is None:
        _T = _tables()
    v = int.from_bytes(bytes(block), "little")
    pos = 0

    def get(n):
        nonlocal pos
        r = (v >> pos) & ((1 << n) - 1)
        pos += n
        return r
    mode = 0
    while mode < 8 and not (v >> mode) & 1:
        mode += 1
    if mode == 8:
        return np.zeros((4, 4, 4), np.uint8)
    pos = mode + 1
    ns, pb, rb, isb, cb, ab, epb, spb, ib, ib2 = _MODES[mode]
    part = get(pb)
    rot = get(rb)
    isel = get(isb)
    ep = np.zeros((ns * 2, 4), np.int64)
    for c in range(3):
        for e in range(ns * 2):
            ep[e, c] = get(cb)
    if ab:
        for e in range(ns * 2):
            ep[e, 3] = get(ab)
    cbits, abits = cb, ab
    if epb:
        for e in range(ns * 2):
            p = get(1)
            ep[e, :3] = (ep[e, :3] << 1) | p
            if ab:
                ep[e, 3] = (ep[e, 3] << 1) | p
        cbits += 1
        abits += 1 if ab else 0
    elif spb:
        for s in range(ns):
            p = get(1)
            for e in (2 * s, 2 * s + 1):
                ep[e, :3] = (ep[e, :3] << 1) | p
        cbits += 1
    ep[:, :3] = (ep[:, :3] << (8 - cbits)) | ((ep[:, :3] << (8 - cbits)) >> cbits)
    if ab:
        ep[:, 3] = (ep[:, 3] << (8 - abits)) | ((ep[:, 3] << (8 - abits)) >> abits)
    else:
        ep[:, 3] = 255
    subset = np.zeros(16, np.int64) if ns == 1 else (_T["PARTITION2"][part] if ns == 2 else _T["PARTITION3"][part])
    anchors = [0] if ns == 1 else ([0, int(_T["ANCHOR2"][part])] if ns == 2 else [0, int(_T["ANCHOR3A"][part]), int(_T["ANCHOR3B"][part])])

    def read_indices(bits, use_anchors):
        idx = np.zeros(16, np.int64)
        for i in range(16):
            idx[i] = get(bits - 1 if i in use_anchors else bits)
        return idx
    idx1 = read_indices(ib, anchors)
    idx2 = read_indices(ib2, [0]) if ib2 else None
    out = np.zeros((16, 4), np.int64)
    for i in range(16):
        e0, e1 = ep[2 * subset[i]], ep[2 * subset[i] + 1]
        if ib2:
            ci, cbw, ai, abw = (idx1[i], ib, idx2[i], ib2) if not isel else (idx2[i], ib2, idx1[i], ib)
        else:
            ci, cbw, ai, abw = idx1[i], ib, idx1[i], ib
        wc, wa = _WEIGHTS[cbw][ci], _WEIGHTS[abw][ai]
        out[i, :3] = ((64 - wc) * e0[:3] + wc * e1[:3] + 32) >> 6
        out[i, 3] = ((64 - wa) * e0[3] + wa * e1[3] + 32) >> 6
        if rot:
            c = rot - 1
            out[i, c], out[i, 3] = out[i, 3], out[i, c]
    return out.reshape(4, 4, 4).astype(np.uint8)


def _rgtc_palette(r0, r1):
    """RGTC1 palette (EXT_texture_compression_rgtc), evaluated on normalised floats in fp32, left to right."""
    f = np.float32
    R0, R1 = f(r0) / f(255.0), f(r1) / f(255.0)
    if r0 > r1:
        return [R0, R1] + [(f(7 - k) * R0 + f(k) * R1) / f(7.0) for k in range(1, 7)]
    return [R0, R1] + [(f(5 - k) * R0 + f(k) * R1) / f(5.0) for k in range(1, 5)] + [f(0.0), f(1.0)]


def decode_bc4_block(block):
    """8 bytes -> [4, 4] float32."""
    pal = _rgtc_palette(int(block[0]), int(block[1]))
    bits = int.from_bytes(bytes(block[2:8]), "little")
    return np.array([pal[(bits >> (3 * i)) & 7] for i in range(16)], np.float32).reshape(4, 4)


def decode_texture(fmt, blocks, width, height):
    """blocks: uint8 [ceil(h/4) * ceil(w/4), block bytes] in row-major block order -> texels [height, width, C]
    (bc7: uint8 x 4; bc5: float32 x 2; bc4: float32 x 1)."""
    bw, bh = (width + 3) // 4, (height + 3) // 4
    blocks = np.asarray(blocks, np.uint8).reshape(bh * bw, -1)
    if fmt == "bc7":
        full = np.zeros((bh * 4, bw * 4, 4), np.uint8)
    else:
        full = np.zeros((bh * 4, bw * 4, 2 if fmt == "bc5" else 1), np.float32)
    for b in range(bh * bw):
        y, x = 4 * (b // bw), 4 * (b % bw)
        if fmt == "bc7":
            full[y:y + 4, x:x + 4] = decode_bc7_block(blocks[b])
        elif fmt == "bc5":
            full[y:y + 4, x:x + 4, 0] = decode_bc4_block(blocks[b][:8])
            full[y:y + 4, x:x + 4, 1] = decode_bc4_block(blocks[b][8:])
        else:
            full[y:y + 4, x:x + 4, 0] = decode_bc4_block(blocks[b])
    return full[:height, :width]


# ---------------------------------------------------------------------------------------------------------------- encoders (test data)
def encode_bc4_channel(t):
    """[4, 4] values in [0, 255] -> 8 bytes (8-value mode, nearest palette entry)."""
    t = np.asarray(t, np.float64).reshape(16)
    r0, r1 = int(np.clip(np.round(t.max()), 0, 255)), int(np.clip(np.round(t.min()), 0, 255))
    if r0 == r1:
        r0, r1 = (r0, r0 - 1) if r0 > 0 else (1, 0)
    pal = [r0, r1] + [((7 - k) * r0 + k * r1) / 7.0 for k in range(1, 7)]
    idx = [int(np.argmin([abs(v - p) for p in pal])) for v in t]
    bits = 0
    for i, k in enumerate(idx):
        bits |= k << (3 * i)
    return np.frombuffer(bytes([r0, r1]) + bits.to_bytes(6, "little"), np.uint8)


def encode_bc7_mode6(t):
    """[4, 4, 4] uint8 RGBA -> 16 bytes: mode 6 (one subset, 7-bit endpoints + p-bit, 4-bit indices), endpoints = the two
    texels that are farthest apart along the block's principal channel sum."""
    px = np.asarray(t, np.int64).reshape(16, 4)
    key = px.sum(1)
    lo, hi = px[int(np.argmin(key))], px[int(np.argmax(key))]
    e0, e1 = lo >> 1, hi >> 1
    p0, p1 = 0, 1

    def full(e, p):
        return (e << 1) | p
    f0, f1 = full(e0, p0), full(e1, p1)
    d = (f1 - f0).astype(np.float64)
    dd = float((d * d).sum())
    idx = []
    for i in range(16):
        tpos = 0.0 if dd == 0 else float(((px[i] - f0) * d).sum()) / dd
        idx.append(int(np.clip(np.round(tpos * 15), 0, 15)))
    if idx[0] >= 8:                     # anchor index must have its top bit clear: swap the endpoints
        e0, e1, p0, p1 = e1, e0, p1, p0
        idx = [15 - k for k in idx]
    v, pos = 0, 0

    def put(val, n):
        nonlocal v, pos
        v |= (int(val) & ((1 << n) - 1)) << pos
        pos += n
    put(1 << 6, 7)
    for c in range(4):
        put(e0[c], 7)
        put(e1[c], 7)
    put(p0, 1)
    put(p1, 1)
    put(idx[0], 3)
    for k in idx[1:]:
        put(k, 4)
    assert pos == 128
    return np.frombuffer(v.to_bytes(16, "little"), np.uint8)


def encode_texture(fmt, img):
    """img [H, W, C] (uint8-range values) -> block array [ceil(H/4) * ceil(W/4), block bytes]; edge blocks replicate the border."""
    img = np.asarray(img)
    h, w = img.shape[:2]
    bh, bw = (h + 3) // 4, (w + 3) // 4
    pad = np.pad(img, ((0, bh * 4 - h), (0, bw * 4 - w), (0, 0)), mode="edge")
    out = []
    for by in range(bh):
        for bx in range(bw):
            t = pad[4 * by:4 * by + 4, 4 * bx:4 * bx + 4]
            if fmt == "bc7":
                out.append(encode_bc7_mode6(t))
            elif fmt == "bc5":
                out.append(np.concatenate([encode_bc4_channel(t[..., 0]), encode_bc4_channel(t[..., 1])]))
            else:
                out.append(encode_bc4_channel(t[..., 0]))
    return np.stack(out)
