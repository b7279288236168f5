"""Asynchronous idkpt_compute: several samples in flight on separate lanes must produce exactly the images of the
synchronous path (and hence of the oracle), whatever is interleaved between the calls."""
import numpy as np
import pytest

import oracle_lib as ol
from idkengine_b200 import capi, scenes

pytestmark = pytest.mark.gpu


def render_sync(scene, cam, w, h, s, steps, tile=(8, 0, 1), frames=None):
    from idkengine_b200.pathtracer import PathTracer
    with PathTracer(w, h, s, tile=tile) as pt:
        pt.SetScene(scene); pt.SetSky((0.6, 0.7, 0.9))
        for k in range(steps):
            pt.SetFrame(frames[k] if frames else scenes.camera_frame(cam, w, h))
            pt.Compute()
        return pt.Result.copy(), pt.AlbedoTexture.copy(), pt.NormalTexture.copy()


@pytest.mark.parametrize("lanes", [2, 3, 4, 8])
@pytest.mark.parametrize("sorting,aov", [(0, 1), (1, 0)])
def test_async_equals_sync(multi_blas, lanes, sorting, aov):
    from idkengine_b200.pathtracer import PathTracer
    scene, cam = multi_blas
    w, h, steps = 144, 96, 11
    s = capi.default_settings()
    s.DoRaySorting, s.OutputAOVs = sorting, aov
    s.Gpu.DoTraceLights = 1
    want = render_sync(scene, cam, w, h, s, steps)
    with PathTracer(w, h, s, lanes=lanes) as pt:
        pt.SetScene(scene); pt.SetSky((0.6, 0.7, 0.9)); pt.SetFrame(scenes.camera_frame(cam, w, h))
        for _ in range(steps):
            pt.ComputeAsync()
        assert pt.AccumulatedSamples == steps
        pt.Sync()
        got = (pt.Result.copy(), pt.AlbedoTexture.copy(), pt.NormalTexture.copy())
    for a, b in zip(got, want):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_async_matches_oracle_with_tiles_and_spp(cornell):
    from idkengine_b200.pathtracer import PathTracer
    scene, cam = cornell
    w, h = 128, 96
    s = capi.default_settings()
    s.SamplesPerPixel = 3
    tile = (8, 1, 3)
    frame = scenes.camera_frame(cam, w, h)
    with PathTracer(w, h, s, tile=tile, lanes=3) as pt:
        pt.SetScene(scene); pt.SetSky((0.6, 0.7, 0.9)); pt.SetFrame(frame)
        pt.ComputeAsync(); pt.ComputeAsync()
        got = pt.Result.copy()                         # read-back is ordered after the queued samples
        assert pt.AccumulatedSamples == 6
    res = np.zeros((h, w, 4), np.float32)
    o = ol.path_trace(scene, frame, s, w, h, tile=tile, result=res)
    o = ol.path_trace(scene, frame, s, w, h, tile=tile, accumulated=o.accumulated, result=res)
    assert np.array_equal(got.view(np.uint32), res.view(np.uint32))


def test_async_interleaved_with_everything(multi_blas):
    """Camera changes, accumulation resets, presents, post-process, a material update and a synchronous Compute between
    asynchronous ones: same images as the fully synchronous sequence."""
    import torch
    from idkengine_b200.pathtracer import PathTracer
    scene, cam = multi_blas
    w, h = 128, 80
    s = capi.default_settings()
    cam2 = dict(cam, position=(0.6, 1.4, 4.6))
    fa, fb = scenes.camera_frame(cam, w, h), scenes.camera_frame(cam2, w, h)
    mats = scene.materials[:1].copy()
    mats["RoughnessFactor"] = 0.1

    def sequence(pt, compute):
        snaps = []
        pt.SetFrame(fa)
        compute(); compute(); compute()
        snaps.append(pt.Result.copy())
        pt.ResetAccumulation(); pt.SetFrame(fb)        # the engine's camera-moved path
        compute(); compute()
        ldr, _ = pt.PostProcess()
        snaps.append(ldr.copy())
        compute()
        pt.UpdateRange(capi.IDKPT_ARRAY_MATERIALS, 0, mats)     # drains, resets the accumulation
        compute(); compute(); compute(); compute(); compute()
        pt.Compute()                                   # synchronous call in the middle of the stream
        compute()
        snaps.append(pt.Result.copy())
        return snaps

    with PathTracer(w, h, s, lanes=1) as pt:
        pt.SetScene(scene); pt.SetSky((0.6, 0.7, 0.9))
        want = sequence(pt, pt.Compute)
    with PathTracer(w, h, s, lanes=4) as pt:
        pt.SetScene(scene); pt.SetSky((0.6, 0.7, 0.9))
        got = sequence(pt, pt.ComputeAsync)
        # presents of consecutive asynchronous samples see consecutive images
        pinned = [torch.empty((h, w, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
        pt.ResetAccumulation()
        pt.ComputeAsync(); pt.PresentAsync(pinned[0].data_ptr(), pinned[0].numel() * 4)
        pt.PresentWait()
        first = pinned[0].numpy().copy()
        pt.ComputeAsync(); pt.PresentAsync(pinned[1].data_ptr(), pinned[1].numel() * 4)
        pt.ComputeAsync(); pt.ComputeAsync()
        pt.PresentWait()
        second = pinned[1].numpy().copy()
        pt.Sync()
        final = pt.Result.copy()
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    with PathTracer(w, h, s) as pt:
        pt.SetScene(scene); pt.SetSky((0.6, 0.7, 0.9)); pt.UpdateRange(capi.IDKPT_ARRAY_MATERIALS, 0, mats); pt.SetFrame(fb)
        pt.Compute(); r1 = pt.Result.copy()
        pt.Compute(); r2 = pt.Result.copy()
        pt.Compute(); pt.Compute(); r4 = pt.Result.copy()
    assert np.array_equal(first, r1) and np.array_equal(second, r2) and np.array_equal(final, r4)


def test_async_resize_and_stats_fallback(cornell):
    from idkengine_b200.pathtracer import PathTracer
    scene, cam = cornell
    s = capi.default_settings()
    with PathTracer(96, 64, s, lanes=3) as pt:
        pt.SetScene(scene); pt.SetSky((0.6, 0.7, 0.9)); pt.SetFrame(scenes.camera_frame(cam, 96, 64))
        for _ in range(5):
            pt.ComputeAsync()
        pt.SetSize(64, 48)                              # drains, reallocates every lane lazily
        pt.SetFrame(scenes.camera_frame(cam, 64, 48))
        for _ in range(4):
            pt.ComputeAsync()
        pt.CollectStats = 1                              # counters need the synchronous path: still correct
        pt.ComputeAsync()
        got = pt.Result.copy()
    with PathTracer(64, 48, s) as pt:
        pt.SetScene(scene); pt.SetSky((0.6, 0.7, 0.9)); pt.SetFrame(scenes.camera_frame(cam, 64, 48))
        for _ in range(5):
            pt.Compute()
        assert np.array_equal(got, pt.Result)


def test_tiled_present_into_a_registered_host_frame(cornell):
    """Multi-GPU presentation path on one device: a tiled context delivers ONLY its own stripes, at their final rows, into a
    full-frame host buffer page-locked through idkpt_register_host_buffer (the strided copy of idkpt_present_async)."""
    from idkengine_b200.pathtracer import PathTracer
    scene, cam = cornell
    w, h = 96, 84                       # 10.5 stripes of 8 rows: the last stripe of the image is partial
    s = capi.default_settings()
    frame = scenes.camera_frame(cam, w, h)
    host = np.full((h, w, 4), -7.0, np.float32)
    for tile in ((8, 0, 3), (8, 1, 3), (8, 2, 3)):
        with PathTracer(w, h, s, tile=tile, lanes=2) as pt:
            pt.SetScene(scene); pt.SetSky((0.6, 0.7, 0.9)); pt.SetFrame(frame)
            pt.RegisterHostBuffer(host.ctypes.data, host.nbytes)
            before = host.copy()
            pt.ComputeAsync(); pt.ComputeAsync()
            pt.PresentAsync(host.ctypes.data, host.nbytes)
            pt.PresentWait()
            rows = pt.TileRows()
            other = np.setdiff1d(np.arange(h), rows)
            assert np.array_equal(host[other], before[other])             # nobody else's rows were touched
            assert np.array_equal(host[rows], pt.Result[rows])
            assert len(rows) % 8 != 0 or tile[1] != 1                      # tile 1 owns the partial last stripe (rows 80..83)
            pt.UnregisterHostBuffer(host.ctypes.data)
    assert (host != -7.0).all()                                            # the three tiles together filled the frame
