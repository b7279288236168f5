#!/usr/bin/env python
"""Strict multi-GPU parity (SURVEY 8e option ii) checked inside ONE process: `world` tile contexts created with
IDKPT_CREATE_GLOBAL_SLOTS and wired to each other with idkpt_gather_connect (on one GPU, or on GPUs r % device_count) must
produce, bit for bit, the image ONE untiled context produces -- in each tile's own rows and in every tile's gathered frame.
The control run without the flag must differ (tile-local slot numbers draw other random numbers) -- otherwise the scene does
not exercise the exchange.

Run it in a fresh process with enough hardware queues for all the streams (several contexts on one GPU spin-wait on each other):

    CUDA_DEVICE_MAX_CONNECTIONS=32 python scripts/check_global_slots.py --world 3

tests/test_async.py runs it that way under -m gpu. Multi-process (one rank per GPU): bench.py --global-slots.
"""
import argparse
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=3)
    ap.add_argument("--width", type=int, default=200)
    ap.add_argument("--height", type=int, default=132)      # 16.5 stripes: the last stripe is partial
    ap.add_argument("--samples", type=int, default=5)
    ap.add_argument("--tris", type=int, default=20000)
    ap.add_argument("--lanes", type=int, default=2)
    args = ap.parse_args()

    import torch
    from idkengine_b200 import capi, scenes
    from idkengine_b200.pathtracer import PathTracer

    ndev = torch.cuda.device_count()
    if ndev == 0:
        raise SystemExit("check_global_slots.py: no CUDA device")
    scene, cam = scenes.atrium(args.tris)
    frame = scenes.camera_frame(cam, args.width, args.height)
    s = capi.default_settings()
    s.RayDepth = 7
    sky = (0.6, 0.7, 0.9)
    w, h, W = args.width, args.height, args.world

    def prepare(pt):
        pt.SetScene(scene)
        pt.SetSky(sky)
        pt.SetFrame(frame)

    with PathTracer(w, h, s, device=0, lanes=args.lanes) as one:
        prepare(one)
        for _ in range(args.samples):
            one.ComputeAsync()
        one.Sync()
        ref = one.Result.copy()

    def run_tiles(global_slots):
        tiles = [PathTracer(w, h, s, device=r % ndev, tile=(8, r, W), lanes=args.lanes, global_slots=global_slots) for r in range(W)]
        try:
            PathTracer.ConnectPeers(tiles)
            for t in tiles:
                prepare(t)
            # every context only ever queues work (stats == NULL): one host thread can drive all of them
            for _ in range(args.samples):
                for t in tiles:
                    t.ComputeAsync()
            gathered = [np.zeros((h, w, 4), np.float32) for _ in tiles]
            for t, g in zip(tiles, gathered):
                t.PresentAsync(g.ctypes.data, g.nbytes, capi.IDKPT_IMAGE_GATHERED)
            for t in tiles:
                t.PresentWait()
                t.Sync()
            img = np.zeros((h, w, 4), np.float32)
            for t in tiles:
                rows = t.TileRows()
                img[rows] = t.Result[rows]
            return img, gathered
        finally:
            for t in tiles:
                t.Dispose()

    strict, strict_gathered = run_tiles(True)
    local, _ = run_tiles(False)
    eq = lambda a, b: bool(np.array_equal(a.view(np.uint32), b.view(np.uint32)))
    out = {
        "world": W, "devices": ndev, "size": [w, h], "samples": args.samples,
        "tiles_eq_one_gpu": eq(strict, ref),
        "gathered_eq_one_gpu": [eq(g, ref) for g in strict_gathered],
        "control_local_slots_differ": not eq(local, ref),
        "control_max_abs_diff": float(np.abs(local - ref).max()),
    }
    out["ok"] = out["tiles_eq_one_gpu"] and all(out["gathered_eq_one_gpu"]) and out["control_local_slots_differ"]
    print(json.dumps(out))
    sys.exit(0 if out["ok"] else 1)


if __name__ == "__main__":
    main()
