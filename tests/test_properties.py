"""Size-independent properties of the path (oracle level, CPU): they hold for the CUDA path too because it is bit-identical
to the oracle (tests/test_gpu_parity.py, test_async.py)."""
import copy

import numpy as np

import oracle_lib as ol
from idkengine_b200 import capi, scenes


def test_sample_depends_only_on_its_index(cornell):
    """Sample k's radiance depends on AccumulatedSamples = k alone, not on the image it is folded into: sample k rendered
    onto a black image is r_k / (k + 1); folding those r_k by hand reproduces the sequential accumulation."""
    scene, cam = cornell
    w, h = 64, 48
    frame = scenes.camera_frame(cam, w, h)
    s = capi.default_settings()
    seq = np.zeros((h, w, 4), np.float32)
    acc = 0
    for _ in range(4):
        acc = ol.path_trace(scene, frame, s, w, h, accumulated=acc, result=seq, want_rays=False).accumulated
    assert acc == 4
    manual = np.zeros((h, w, 3), np.float64)
    for k in range(4):
        single = np.zeros((h, w, 4), np.float32)
        ol.path_trace(scene, frame, s, w, h, accumulated=k, result=single, want_rays=False)     # mix(0, r_k, 1/(k+1))
        r_k = single[..., :3].astype(np.float64) * (k + 1)
        manual = manual + (r_k - manual) / (k + 1)
    assert np.allclose(seq[..., :3], manual, rtol=2e-5, atol=1e-6)
    assert seq[..., 3].min() == 1.0 and np.isfinite(seq).all()


def test_primary_bounce_is_tile_independent(multi_blas):
    """NHit seeds are slot ids, so deeper bounces depend on the tile map; the first hit does not (seed = pixel, sample)."""
    scene, cam = multi_blas
    w, h = 96, 64
    frame = scenes.camera_frame(cam, w, h)
    s = capi.default_settings()
    s.RayDepth = 1
    full = ol.path_trace(scene, frame, s, w, h, want_rays=False).result
    stitched = np.zeros_like(full)
    for t in range(3):
        part = ol.path_trace(scene, frame, s, w, h, tile=(8, t, 3), want_rays=False).result
        rows = [y for y in range(h) if (y // 8) % 3 == t]
        stitched[rows] = part[rows]
    assert np.array_equal(full, stitched)


def test_refit_is_idempotent_and_a_no_op_on_unmoved_refittable_geometry(multi_blas):
    """The builder's boxes of a non-presplit (refittable) BLAS are the exact bounds of its triangles, so BLAS.Refit leaves
    them untouched; on a presplit BLAS refit grows the clipped fragment boxes once and is idempotent afterwards."""
    scene = copy.deepcopy(multi_blas[0])
    before = scene.blas_nodes.copy()
    ol.blas_refit(scene, 2)                                             # crate: refittable, no presplit
    assert scene.blas_nodes.tobytes() == before.tobytes()
    ol.blas_refit(scene, 0)                                             # room: presplit fragments
    once = scene.blas_nodes.copy()
    ol.blas_refit(scene, 0)
    assert scene.blas_nodes.tobytes() == once.tobytes()
    d = scene.blas_descs[0]
    a, b = before[d["NodeOffset"] + 1:d["NodeOffset"] + d["NodeCount"]], once[d["NodeOffset"] + 1:d["NodeOffset"] + d["NodeCount"]]
    assert (b["Min"] <= a["Min"] + 1e-6).all() and (b["Max"] >= a["Max"] - 1e-6).all()      # refit only grows presplit boxes


def test_any_hit_is_monotone_in_tmax(multi_blas):
    scene, _ = multi_blas
    rng = np.random.default_rng(9)
    from idkengine_b200 import gpu_types as gt
    rays = np.zeros(4000, gt.IdkPtRay)
    rays["Origin"] = rng.uniform(-2.5, 2.5, (4000, 3)).astype(np.float32)
    d = rng.normal(size=(4000, 3))
    rays["Direction"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    prev = np.zeros(4000, bool)
    for tmax in (0.25, 1.0, 4.0, 3.4028235e38):
        rays["TMax"] = np.float32(tmax)
        occ = ol.trace_rays_any(scene, rays)["NodePairFetches"] == 1
        assert (occ | ~prev).all()                                      # once occluded, stays occluded for a longer ray
        prev = occ
    assert prev.mean() > 0.3


def _cube_face_dirs(size):
    """Direction through every texel centre of the six faces, GL layout (+X,-X,+Y,-Y,+Z,-Z; spec table 8.19)."""
    c = (np.arange(size, dtype=np.float64) + 0.5) / size * 2.0 - 1.0
    sc, tc = np.meshgrid(c, c)                      # [t, s]
    one = np.ones_like(sc)
    return np.stack([np.stack([one, -tc, -sc], -1), np.stack([-one, -tc, sc], -1), np.stack([sc, one, tc], -1),
                     np.stack([sc, -one, -tc], -1), np.stack([sc, -tc, one], -1), np.stack([-sc, -tc, -one], -1)])


def test_cube_map_filtering_is_seamless():
    """GL_TEXTURE_CUBE_MAP_SEAMLESS (the engine enables it, SkyBoxManager.cs:74): a cube map that stores a smooth function of
    the direction must be reproduced smoothly ACROSS face edges and corners, not only inside faces. With per-face clamping the
    error at an edge is half a texel of gradient; with seamless filtering it stays at the curvature level everywhere."""
    import oracle_lib as ol
    size = 16
    dirs = _cube_face_dirs(size)
    unit = dirs / np.linalg.norm(dirs, axis=-1, keepdims=True)
    faces = np.zeros((6, size, size, 4), np.float32)
    faces[..., :3] = (0.5 + 0.5 * unit).astype(np.float32)          # f(d) = 0.5 + 0.5 d
    faces[..., 3] = 1.0
    # texel centres return the texel exactly
    flat = dirs.reshape(-1, 3)
    got = ol.sample_sky(faces, flat)
    assert np.abs(got - faces[..., :3].reshape(-1, 3)).max() < 1e-6
    # random directions, with a bias towards edges and corners of the cube
    rng = np.random.default_rng(12)
    d = rng.normal(size=(60000, 3))
    d[:20000] = np.sign(d[:20000]) * (1.0 - rng.uniform(0, 0.08, (20000, 3)) * rng.integers(0, 2, (20000, 3)))   # near edges / corners
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    got = ol.sample_sky(faces, d)
    want = 0.5 + 0.5 * d
    err = np.abs(got - want).max(axis=1)
    # bilinear interpolation of a smooth function on this grid: a few 1e-3; a clamped edge would be ~0.5 * (2/size) * 0.5 = 0.03
    assert err.max() < 0.012, err.max()
    m = np.abs(d).max(axis=1, keepdims=True)
    on_cube = d / m
    near_edge = (np.sort(np.abs(on_cube), axis=1)[:, 1] > 1.0 - 1.0 / size)
    assert near_edge.sum() > 5000 and err[near_edge].max() < 0.012
    # continuity: two directions a hair apart on either side of an edge give (almost) the same colour
    a = np.array([[1.0, 0.3, 0.999999], [1.0, 0.3, 1.000001]])
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    s2 = ol.sample_sky(faces, a)
    assert np.abs(s2[0] - s2[1]).max() < 1e-4
