"""Per-bounce ray counts and kernel times (CUDA events) for the bench workload, full frame and a 1/8 stripe tile."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from idkengine_b200 import capi, scenes
from idkengine_b200.pathtracer import PathTracer
scene, cam = scenes.atrium(262144)
w, h = 1920, 1080
frame = scenes.camera_frame(cam, w, h)
import os
for tile in (((8, 0, 1), (8, 0, 8)) if not os.environ.get('ONLY_EIGHTH') else ((8, 0, 8),)):
    s = capi.default_settings(); s.RayDepth = 9
    with PathTracer(w, h, s, tile=tile) as pt:
        pt.SetScene(scene); pt.SetSky((0.6, 0.7, 0.9)); pt.SetFrame(frame)
        pt.CollectStats = 1
        mx = pt.Compute().as_dict()["BounceMaxSteps"]
        pt.CollectStats = 0
        print("  max steps per bounce:", mx[:9])
        for _ in range(3): pt.Compute()
        pt.ResetAccumulation()
        acc = None
        K = 10
        for _ in range(K):
            st = pt.Compute().as_dict()
            if acc is None: acc = st
            else:
                for k in ("BounceRays", "BounceTraverseMs", "BounceShadeMs"): acc[k] = [a + b for a, b in zip(acc[k], st[k])]
                for k in ("TotalMs", "TraverseMs", "ShadeMs", "OtherMs"): acc[k] += st[k]
        print("TILE", tile, "total %.3f trav %.3f shade %.3f other %.3f ms" % tuple(acc[k] / K for k in ("TotalMs", "TraverseMs", "ShadeMs", "OtherMs")))
        for j in range(9):
            print("  bounce %d rays %8d traverse %7.1f us shade+compact %6.1f us" % (j, acc["BounceRays"][j] / K, acc["BounceTraverseMs"][j] / K * 1e3, acc["BounceShadeMs"][j] / K * 1e3))
