"""ctypes wrapper of oracle/liboracle.so -- imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs only."""
import ctypes
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "oracle"))

from idkengine_b200 import capi, gpu_types as gt  # noqa: E402

_lib = None


def lib():
    global _lib
    if _lib is None:
        import importlib.util
        spec = importlib.util.spec_from_file_location("oracle_build", os.path.join(REPO, "oracle", "build.py"))
        ob = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ob)
        path = ob.LIBORACLE if os.path.exists(ob.LIBORACLE) else ob.build()
        try:
            path = ob.build()
        except Exception:
            pass
        L = ctypes.CDLL(path)
        P = ctypes.POINTER
        vp, u64, i32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int32
        L.oracle_trace_rays.restype = i32
        L.oracle_trace_rays.argtypes = [P(capi.IdkPtSceneDesc), vp, u64, i32, vp, i32]
        L.oracle_brute_force.restype = i32
        L.oracle_brute_force.argtypes = [P(capi.IdkPtSceneDesc), vp, u64, vp, i32]
        L.oracle_cpu_intersect.restype = ctypes.c_double
        L.oracle_cpu_intersect.argtypes = [P(capi.IdkPtSceneDesc), vp, u64, vp, i32]
        L.oracle_gui_test_rays.restype = None
        L.oracle_gui_test_rays.argtypes = [vp, i32, i32, i32, i32, vp]
        L.oracle_path_trace.restype = i32
        L.oracle_path_trace.argtypes = [P(capi.IdkPtSceneDesc), P(capi.IdkPtSkyDesc), vp, P(capi.IdkPtSettings),
                                        i32, i32, i32, i32, i32, P(ctypes.c_uint32), vp, vp, vp, vp,
                                        P(capi.IdkPtStats), i32]
        L.oracle_det_sincos.argtypes = [vp, u64, vp, vp]
        L.oracle_det_exp.argtypes = [vp, u64, vp]
        L.oracle_encode_decode.argtypes = [vp, u64, vp, vp]
        L.oracle_pcg.restype = ctypes.c_uint32
        L.oracle_pcg.argtypes = [ctypes.c_uint32, P(ctypes.c_uint32)]
        _lib = L
    return _lib


def default_threads():
    return max(1, min(os.cpu_count() or 1, 64))


def make_rays(origins, directions, tmax=3.4028235e+38):
    r = np.zeros(len(origins), gt.IdkPtRay)
    r["Origin"] = origins
    r["Direction"] = directions
    r["TMax"] = tmax
    return r


def trace_rays(scene, rays, trace_lights=False, threads=None):
    d, keep = capi.scene_desc(scene)
    out = np.zeros(len(rays), gt.IdkPtHit)
    lib().oracle_trace_rays(ctypes.byref(d), rays.ctypes.data, len(rays), int(trace_lights), out.ctypes.data,
                            threads or default_threads())
    return out


def brute_force(scene, rays, threads=None):
    d, keep = capi.scene_desc(scene)
    out = np.zeros(len(rays), gt.IdkPtHit)
    lib().oracle_brute_force(ctypes.byref(d), rays.ctypes.data, len(rays), out.ctypes.data, threads or default_threads())
    return out


def cpu_intersect(scene, rays, threads=None, want_hits=True):
    d, keep = capi.scene_desc(scene)
    out = np.zeros(len(rays), gt.IdkPtHit) if want_hits else None
    secs = lib().oracle_cpu_intersect(ctypes.byref(d), rays.ctypes.data, len(rays),
                                      out.ctypes.data if want_hits else None, threads or default_threads())
    return out, secs


def gui_test_rays(frame, width, height, y0=0, y1=None):
    y1 = height if y1 is None else y1
    out = np.zeros((y1 - y0) * width, gt.IdkPtRay)
    lib().oracle_gui_test_rays(frame.ctypes.data, width, height, y0, y1, out.ctypes.data)
    return out


def primary_rays(frame, width, height, stride=1):
    """Pinhole rays through pixel centres (test inputs for the stand-alone traversal; not a reference code path)."""
    rays = gui_test_rays(frame, width, height)
    return rays[::stride].copy()


class PathTraceResult:
    pass


def path_trace(scene, frame, settings, width, height, sky=(0.6, 0.7, 0.9), tile=(8, 0, 1), accumulated=0,
               result=None, albedo=None, normal=None, want_rays=True, threads=None):
    d, keep = capi.scene_desc(scene)
    sk = capi.sky_desc(sky)
    res = np.zeros((height, width, 4), np.float32) if result is None else result
    alb = np.zeros((height, width, 4), np.float32) if albedo is None else albedo
    nrm = np.zeros((height, width, 4), np.float32) if normal is None else normal
    rays = np.zeros(width * height, gt.GpuWavefrontRay) if want_rays else None
    acc = ctypes.c_uint32(accumulated)
    stats = capi.IdkPtStats()
    rc = lib().oracle_path_trace(ctypes.byref(d), ctypes.byref(sk), frame.ctypes.data, ctypes.byref(settings),
                                 width, height, tile[0], tile[1], tile[2], ctypes.byref(acc),
                                 res.ctypes.data, alb.ctypes.data, nrm.ctypes.data,
                                 rays.ctypes.data if want_rays else None, ctypes.byref(stats),
                                 threads or default_threads())
    assert rc == 0, rc
    out = PathTraceResult()
    out.result, out.albedo, out.normal, out.rays, out.accumulated, out.stats = res, alb, nrm, rays, acc.value, stats
    return out
