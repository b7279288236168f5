"""Regenerate tests/golden/*.json from the CPU oracle (python tests/golden/make_golden.py).

The reference ships no golden vectors and cannot run here (no .NET / GL), so these fixtures are produced by the oracle
itself: they pin the oracle and the CUDA path against *regressions*; they do not pin the oracle to the reference
(DESIGN.md "parity unpinned")."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib as ol  # noqa: E402
from idkengine_b200 import capi, scenes  # noqa: E402

CASES = {
    "cornell_96x64_depth7_spp2_aov": dict(scene="cornell_1k", w=96, h=64, depth=7, spp=2, aov=1, sort=0, lights=0, rr=1),
    "cornell_80x80_depth6_sort": dict(scene="cornell_1k", w=80, h=80, depth=6, spp=1, aov=0, sort=1, lights=0, rr=1),
    "multi_blas_64x48_lights_norr": dict(scene="multi_blas", w=64, h=48, depth=5, spp=1, aov=1, sort=0, lights=1, rr=0),
}


def settings_for(c):
    s = capi.default_settings()
    s.RayDepth, s.SamplesPerPixel, s.OutputAOVs, s.DoRaySorting = c["depth"], c["spp"], c["aov"], c["sort"]
    s.Gpu.DoTraceLights, s.Gpu.DoRussianRoulette = c["lights"], c["rr"]
    return s


def digest(a):
    a = np.ascontiguousarray(a)
    a = np.where(a == 0, np.zeros_like(a), a) if a.dtype.kind == "f" else a   # -0 -> +0
    return hashlib.sha256(a.tobytes()).hexdigest()


def run_case(c):
    scene, cam = getattr(scenes, c["scene"])(threads=1)
    frame = scenes.camera_frame(cam, c["w"], c["h"])
    o = ol.path_trace(scene, frame, settings_for(c), c["w"], c["h"])
    return dict(result=digest(o.result), albedo=digest(o.albedo), normal=digest(o.normal),
                rays=int(o.stats.Rays), bounce_rays=[int(v) for v in o.stats.BounceRays][: c["depth"]],
                node_pair_fetches=int(o.stats.NodePairFetches), triangle_tests=int(o.stats.TriangleTests),
                mean_rgb=[float(v) for v in o.result[..., :3].mean(axis=(0, 1))])


if __name__ == "__main__":
    out = {name: dict(case=c, expect=run_case(c)) for name, c in CASES.items()}
    json.dump(out, open(os.path.join(HERE, "path_trace_golden.json"), "w"), indent=1)
    print(json.dumps(out, indent=1)[:600])
