"""Timings of the 8f rows on a large scene: any-hit vs closest-hit batch, ray-traced shadows, skin + refit."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import oracle_lib as ol
from idkengine_b200 import scenes, capi, gpu_types as gt
from idkengine_b200.pathtracer import PathTracer

out = {}
scene, cam = scenes.atrium(262144, seed=1)
scene.add_light((0.0, 6.0, 0.0), (40.0, 40.0, 40.0), 0.3)
w, h = 1920, 1080
frame = scenes.camera_frame(cam, w, h)
with PathTracer(64, 64) as pt:
    pt.SetScene(scene)
    rays = ol.gui_test_rays(frame, w, h)
    for _ in range(2):
        hc, ms_c = pt.TraceRays(rays)
        ha, ms_a = pt.TraceRaysAny(rays)
    out["primary_closest_ms"], out["primary_any_ms"] = ms_c, ms_a
    # G-buffer from the closest hits (GPU), same construction as tests/oracle_lib.synth_gbuffer
    hit = hc["TriangleId"] != 0xFFFFFFFF
    o = rays["Origin"].astype(np.float64); d = rays["Direction"].astype(np.float64)
    pos = o + d * hc["T"][:, None].astype(np.float64)
    pv = frame["ProjView"][0].astype(np.float64).reshape(4, 4)
    clip = np.concatenate([pos, np.ones((len(pos), 1))], 1) @ pv
    depth = np.where(hit, clip[:, 2] / clip[:, 3], 1.0).astype(np.float32)
    depth = np.where(hit & (depth >= 1.0), np.float32(0.999999), depth).reshape(h, w)
    nrg = np.zeros((h, w, 2), np.float32); nrg[..., 0] = 0.5; nrg[..., 1] = 0.5   # +Y normals: floor-like, fine for timing
    for samples in (1, 4):
        for _ in range(2):
            vis, ms = pt.ShadowsRayTraced(frame, depth, nrg, 0, samples=samples)
        out[f"shadows_{samples}spp_ms"] = ms
    out["shadow_lit_fraction"] = float((vis[depth < 1.0] > 0).mean())
    # skin + refit of the whole BLAS
    n = len(scene.positions)
    u = np.zeros(n, gt.GpuUnskinnedVertex)
    rng = np.random.default_rng(0)
    u["JointIndices"] = rng.integers(0, 32, (n, 4))
    wts = rng.uniform(0, 1, (n, 4)).astype(np.float32); u["JointWeights"] = wts / wts.sum(1, keepdims=True)
    u["Position"][:, 0], u["Position"][:, 1], u["Position"][:, 2] = scene.positions["x"], scene.positions["y"], scene.positions["z"]
    u["Normal"], u["Tangent"] = scene.vertices["Normal"], scene.vertices["Tangent"]
    jm = np.zeros((32, 3, 4), np.float32); jm[:, 0, 0] = jm[:, 1, 1] = jm[:, 2, 2] = 1; jm[:, :, 3] = rng.uniform(-0.01, 0.01, (32, 3))
    cmd = np.zeros(1, gt.IdkPtSkinningCmd); cmd["VertexCount"] = n
    pt.SetSkinningData(u)
    for _ in range(3):
        ms_s = pt.SkinVertices(jm, cmd)
        ms_r = pt.BlasRefit(0, 1)
    out["skin_ms"], out["refit_ms"] = ms_s, ms_r
    out["vertices"], out["triangles"], out["nodes"] = n, len(scene.blas_triangles), len(scene.blas_nodes)
print("NEXT", json.dumps(out))
open("gpurun_out/next_rows.json", "w").write(json.dumps(out, indent=1))
