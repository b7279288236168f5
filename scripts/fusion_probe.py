"""What would fusing the launches of S samples buy at N = 8? Emulated with bigger tiles: S samples of a 1/8 tile per fused launch
carry the rays of one S/8 tile. Pipelined ms per sample of tiles 1/8, 1/4, 1/2 with 8, 4, 2 samples in flight, per 1/8-tile equivalent."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bench
from idkengine_b200 import capi
from idkengine_b200.pathtracer import PathTracer

sys.argv = sys.argv[:1]
args = bench.parse_args()
scene, cam, frame = bench.build_scene(args)
out = {}
for tcount, lanes_list in ((8, (8,)), (4, (8, 4, 3)), (2, (8, 4, 2)), (1, (8, 2))):
    for lanes in lanes_list:
        s = capi.default_settings(); s.RayDepth = args.ray_depth
        pt = PathTracer(args.width, args.height, s, device=0, tile=(8, 0, tcount), lanes=lanes)
        pt.SetScene(scene); pt.SetSky(bench.SKY); pt.SetFrame(frame)
        rays = pt.Compute().Rays
        serial = min(pt.Compute().TotalMs for _ in range(3))
        ext = torch.cuda.ExternalStream(pt.StreamHandle())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 40
        for rep in range(2):
            pt.ResetAccumulation(); torch.cuda.synchronize()
            e0.record(ext)
            for _ in range(K if rep else 8):
                pt.ComputeAsync()
            e1.record(ext); pt.Sync()
        ms = e0.elapsed_time(e1) / K
        key = f"tile1/{tcount}_lanes{lanes}"
        out[key] = {"ms_per_sample": round(ms, 4), "ms_per_eighth_equivalent": round(ms * tcount / 8, 4), "serial_ms": round(serial, 4), "mrays_s": round(rays / ms / 1e3, 1)}
        print(key, out[key], flush=True)
        pt.Dispose()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/fusion_probe.json", "w"), indent=1)
