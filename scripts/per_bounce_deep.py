import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from idkengine_b200 import capi, scenes
from idkengine_b200.pathtracer import PathTracer
scene, cam = scenes.atrium(262144)
w, h = 1920, 1080
frame = scenes.camera_frame(cam, w, h)
D = 40
s = capi.default_settings(); s.RayDepth = D
with PathTracer(w, h, s, tile=(8, 0, 8)) as pt:
    pt.SetScene(scene); pt.SetSky((0.6, 0.7, 0.9)); pt.SetFrame(frame)
    pt.CollectStats = 1
    mx = pt.Compute().as_dict()["BounceMaxSteps"]
    pt.CollectStats = 0
    for _ in range(3): pt.Compute()
    pt.ResetAccumulation()
    K = 10
    acc = None
    for _ in range(K):
        st = pt.Compute().as_dict()
        if acc is None: acc = st
        else:
            for k in ("BounceRays", "BounceTraverseMs", "BounceShadeMs"): acc[k] = [a + b for a, b in zip(acc[k], st[k])]
    for j in range(D):
        print("bounce %2d rays %8d maxS %4d traverse %7.1f us shade+compact %6.1f us" % (j, acc["BounceRays"][j] / K, mx[j], acc["BounceTraverseMs"][j] / K * 1e3, acc["BounceShadeMs"][j] / K * 1e3))
