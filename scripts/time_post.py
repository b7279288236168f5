"""Timing of the present chain (8f.3) at 1080p and 4K on a path-traced-like HDR image."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from idkengine_b200 import capi
from idkengine_b200.pathtracer import PathTracer

out = {}
for name, (w, h) in {"1080p": (1920, 1080), "2160p": (3840, 2160)}.items():
    rng = np.random.default_rng(1)
    img = rng.uniform(0, 2.5, (h, w, 4)).astype(np.float32)
    with PathTracer(w, h) as pt:
        pt.WriteResult(img)
        for bloom in (1, 0):
            st = capi.default_post_settings()
            st.IsBloom = bloom
            ms = [pt.PostProcess(st, download=False)[1] for _ in range(6)][2:]
            out[f"{name}_bloom{bloom}_kernel_ms"] = float(np.median(ms))
        import time
        st = capi.default_post_settings()
        t0 = time.perf_counter()
        for _ in range(5):
            pt.PostProcess(st, download=True)
        out[f"{name}_with_download_wall_ms"] = (time.perf_counter() - t0) / 5 * 1e3
print("POST", json.dumps(out))
open("gpurun_out/post_chain.json", "w").write(json.dumps(out, indent=1))
