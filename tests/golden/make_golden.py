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
    "textured_room_96x72_aov_lights": dict(scene="textured_room", w=96, h=72, depth=6, spp=1, aov=1, sort=0, lights=1, rr=1),
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


# ---- the widening rows: shadows / any-hit, skin + refit, present chain ------------------------------------------
def next_rows_inputs():
    """Seeded inputs shared by the oracle run here and the GPU run in tests/test_golden.py."""
    import copy
    from idkengine_b200 import gpu_types as gt
    scene, cam = scenes.multi_blas(threads=1)
    w, h = 96, 64
    frame = scenes.camera_frame(cam, w, h)
    depth, nrg, _ = ol.synth_gbuffer(scene, frame, w, h)
    rng = np.random.default_rng(77)
    rays = np.zeros(5000, gt.IdkPtRay)
    rays["Origin"] = rng.uniform(-2.5, 2.5, (5000, 3)).astype(np.float32)
    d = rng.normal(size=(5000, 3))
    rays["Direction"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays["TMax"] = np.float32(3.4028235e38)
    rays["TMax"][::3] = 1.25
    hdr = rng.uniform(0.0, 1.5, (54, 80, 4)).astype(np.float32)
    hdr[20:24, 30:36, :3] += 40.0
    # skinning of BLAS 2 (the refittable crate): 5 joints, fixed matrices
    dsc = scene.blas_descs[2]
    tris = scene.blas_triangles[dsc["TriangleOffset"]:dsc["TriangleOffset"] + dsc["TriangleCount"]]
    idx = np.concatenate([tris["X"], tris["Y"], tris["Z"]])
    v0, v1 = int(idx.min()), int(idx.max()) + 1
    u = np.zeros(v1 - v0, gt.GpuUnskinnedVertex)
    u["JointIndices"] = rng.integers(0, 5, (v1 - v0, 4))
    wts = rng.uniform(0.0, 1.0, (v1 - v0, 4)).astype(np.float32)
    u["JointWeights"] = wts / wts.sum(1, keepdims=True)
    for k, c in enumerate("xyz"):
        u["Position"][:, k] = scene.positions[c][v0:v1]
    u["Normal"], u["Tangent"] = scene.vertices["Normal"][v0:v1], scene.vertices["Tangent"][v0:v1]
    jm = np.zeros((5, 3, 4), np.float32)
    for j in range(5):
        a = 0.1 * (j - 2)
        jm[j, :, :3] = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]], np.float32) * np.float32(1.0 + 0.05 * j)
        jm[j, :, 3] = np.array([0.03 * j, -0.02 * j, 0.01 * j], np.float32)
    cmd = np.zeros(1, gt.IdkPtSkinningCmd)
    cmd["OutputVertexOffset"], cmd["VertexCount"] = v0, v1 - v0
    return dict(scene=scene, frame=frame, depth=depth, nrg=nrg, rays=rays, hdr=hdr, unskinned=u, joints=jm, cmd=cmd, copy=copy)


def next_rows_expect():
    x = next_rows_inputs()
    scene = x["scene"]
    vis = ol.shadows_ray_traced(scene, x["frame"], x["depth"], x["nrg"], 0, samples=3, noise_index=6)
    anyh = ol.trace_rays_any(scene, x["rays"], trace_lights=True)
    ldr = ol.post_process(x["hdr"])
    moved = x["copy"].deepcopy(scene)
    ol.skin_vertices(x["unskinned"], x["joints"], moved.positions, moved.vertices, x["cmd"][0])
    ol.blas_refit(moved, 2)
    return dict(shadows=digest(vis), any_hit_flags=digest(anyh["NodePairFetches"]), any_hit_t=digest(anyh["T"]), ldr=digest(ldr),
                skinned_positions=digest(moved.positions.view(np.float32)), skinned_vertices=digest(moved.vertices.view(np.uint32)),
                refit_nodes=digest(moved.blas_nodes.view(np.uint32)), occluded=int((anyh["NodePairFetches"] == 1).sum()),
                lit_pixels=int((vis == 1.0).sum()), ldr_mean=float(ldr[..., :3].mean()))


if __name__ == "__main__":
    json.dump(next_rows_expect(), open(os.path.join(HERE, "next_rows_golden.json"), "w"), indent=1)
    out = {name: dict(case=c, expect=run_case(c)) for name, c in CASES.items()}
    json.dump(out, open(os.path.join(HERE, "path_trace_golden.json"), "w"), indent=1)
    print(json.dumps(out, indent=1)[:600])
