"""Bench workload with every material textured (8 x 512^2 sRGB base colour + 128^2 metallic-roughness): serial and pipelined
Mrays/s next to the factor-only scene, plus parity against the oracle on a small frame."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import torch
import bench
import oracle_lib as ol
from idkengine_b200 import capi, scenes
from idkengine_b200.pathtracer import PathTracer

sys.argv = sys.argv[:1]
args = bench.parse_args()
out = {}
for name in ("factor_only", "textured"):
    scene, cam, frame = bench.build_scene(args)
    if name == "textured":
        scenes.texturize(scene)
    s = capi.default_settings(); s.RayDepth = args.ray_depth
    with PathTracer(args.width, args.height, s) as pt:
        pt.SetScene(scene); pt.SetSky(bench.SKY); pt.SetFrame(frame)
        K = 24
        dev = rays = shade = 0.0
        for _ in range(3): pt.Compute()
        pt.ResetAccumulation()
        for _ in range(K):
            st = pt.Compute(); dev += st.TotalMs; rays += st.Rays; shade += st.ShadeMs
        ext = torch.cuda.ExternalStream(pt.StreamHandle())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for rep in range(2):
            pt.ResetAccumulation(); torch.cuda.synchronize(); e0.record(ext)
            for _ in range(K): pt.ComputeAsync()
            e1.record(ext); pt.Sync()
        out[name] = {"serial_mrays_s": rays / dev / 1e3, "shade_ms": shade / K, "pipelined_mrays_s": rays / e0.elapsed_time(e1) / 1e3, "rays_per_sample": rays / K}
    if name == "textured":
        w, h = 160, 90
        f2 = scenes.camera_frame(cam, w, h)
        with PathTracer(w, h, s) as pt:
            pt.SetScene(scene); pt.SetSky(bench.SKY); pt.SetFrame(f2); pt.Compute()
            out["textured_parity"] = bool(np.array_equal(pt.Result, ol.path_trace(scene, f2, s, w, h, sky=bench.SKY).result))
    print(name, out[name], flush=True)
print("TEXBENCH", json.dumps(out))
json.dump(out, open("gpurun_out/textured_bench.json", "w"), indent=1)
