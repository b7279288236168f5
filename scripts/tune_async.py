"""Phase thresholds / lane grids of the pipelined path (8 lanes): device ms per sample, full frame and 1/8 tile."""
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
K = 48
variants = [dict(), dict(IDKPT_TUNE_SETUP="8"), dict(IDKPT_TUNE_SETUP="16"), dict(IDKPT_TUNE_SETUP="20", IDKPT_TUNE_LEAF="6"),
            dict(IDKPT_TUNE_LEAF="2"), dict(IDKPT_TUNE_LEAF="8"), dict(IDKPT_TUNE_SETUP="6", IDKPT_TUNE_LEAF="2"),
            dict(IDKPT_TRAVERSE_VARIANT="2"), dict(IDKPT_LANES="12"), dict(IDKPT_LANES="15"), dict(IDKPT_L2_PERSIST="0")]
for tname, tile in (("full", (8, 0, 1)), ("eighth", (8, 0, 8))):
    for env in variants:
        os.environ.update(env)
        s = capi.default_settings(); s.RayDepth = args.ray_depth
        pt = PathTracer(args.width, args.height, s, device=0, tile=tile)
        pt.SetScene(scene); pt.SetSky(bench.SKY); pt.SetFrame(frame)
        for k in env: os.environ.pop(k)
        rays = pt.Compute().Rays
        ext = torch.cuda.ExternalStream(pt.StreamHandle())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for rep in range(3):
            pt.ResetAccumulation()
            torch.cuda.synchronize()
            e0.record(ext)
            for _ in range(K if rep else 16):
                pt.ComputeAsync()
            e1.record(ext)
            pt.Sync()
            if rep: best = min(best, e0.elapsed_time(e1) / K)
        key = tname + " " + (" ".join(f"{k[6:]}={v}" for k, v in env.items()) or "default")
        out[key] = {"ms_per_sample": best, "mrays_s": rays / best / 1e3}
        print(key, out[key], flush=True)
        pt.Dispose()
json.dump(out, open("gpurun_out/tune_async.json", "w"), indent=1)
