"""Where does the e2e gap go? Same workload as bench.py at N=1; wall-clock per step of four loops."""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import bench
from idkengine_b200 import capi
from idkengine_b200.pathtracer import PathTracer

sys.argv = sys.argv[:1]
args = bench.parse_args()
scene, cam, frame = bench.build_scene(args)
s = capi.default_settings()
s.RayDepth = args.ray_depth
pt = PathTracer(args.width, args.height, s, device=0)
pt.SetScene(scene); pt.SetSky(bench.SKY); pt.SetFrame(frame)
pinned = [torch.empty((args.height, args.width, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
ldr = [torch.empty((args.height, args.width, 4), dtype=torch.uint8).pin_memory() for _ in range(2)]
K = 30
out = {}

def run(name, body, drain=None):
    pt.ResetAccumulation()
    for k in range(3):
        body(k)
    if drain: drain()
    torch.cuda.synchronize()
    pt.ResetAccumulation()
    t0 = time.perf_counter()
    dev = 0.0
    for k in range(K):
        st = body(k)
        dev += st.TotalMs
    if drain: drain()
    torch.cuda.synchronize()
    out[name] = {"wall_ms": (time.perf_counter() - t0) * 1e3 / K, "device_ms": dev / K}

def a(k): return pt.Compute()
def b(k):
    st = pt.Compute(); pt.PresentAsync(pinned[k & 1].data_ptr(), pinned[k & 1].numel() * 4); return st
post = capi.default_post_settings()
import ctypes
def c(k):
    st = pt.Compute()
    ms = ctypes.c_float()
    pt._check(pt._lib.idkpt_post_process(pt._ctx, ctypes.byref(post), 0, ldr[k & 1].data_ptr(), ctypes.byref(ms)), "post")
    return st
def d(k):
    st = pt.Compute(); pt.PresentAsync(pinned[k & 1].data_ptr(), pinned[k & 1].numel() * 4); pt.PresentWait(); return st
run("compute_only", a)
run("compute_present_async_f32", b, pt.PresentWait)
run("compute_post_rgba8_sync", c)
run("compute_present_sync_f32", d)
print("E2EPROBE", json.dumps(out))
open("gpurun_out/e2e_probe.json", "w").write(json.dumps(out, indent=1))
