"""Scope table 8f.3: Bloom + TonemapAndGammaCorrect present chain (Application.cs:217-223)."""
import numpy as np
import pytest

import oracle_lib as ol
from idkengine_b200 import capi, scenes


def synthetic_hdr(w, h, seed=0):
    rng = np.random.default_rng(seed)
    img = rng.uniform(0.0, 1.2, (h, w, 4)).astype(np.float32)
    yy, xx = np.mgrid[0:h, 0:w]
    blob = 30.0 * np.exp(-(((xx - w * 0.3) ** 2 + (yy - h * 0.6) ** 2) / (0.002 * w * h)))
    img[..., :3] += blob[..., None].astype(np.float32)
    img[..., 3] = 1.0
    return img


def test_oracle_tonemap_properties():
    """Oracle sanity: black stays (dithered) black, the curve is monotonic in luminance and saturates at white, the sRGB
    transfer of mid-grey lands where the closed-form AgX-less formula says when tonemapping is off."""
    st = capi.default_post_settings()
    st.IsBloom = 0
    ramp = np.zeros((8, 256, 4), np.float32)
    ramp[..., :3] = (np.arange(256, dtype=np.float32) / 32.0)[None, :, None] ** 2
    out = ol.post_process(ramp, st)
    assert out[..., 3].min() == 255
    lum = out[0, :, 0].astype(int)
    assert lum[0] <= 1 and lum[-1] >= 250
    assert (np.diff(out[0, ::8, 0].astype(int)) >= 0).all()
    st.DoTonemapAndSrgbTransform = 0
    flat = np.full((8, 8, 4), 0.5, np.float32)
    out = ol.post_process(flat, st)
    vals = np.unique(out[..., :3]).astype(int)
    assert vals.min() >= 125 and vals.max() <= 130 and len(vals) >= 3      # 0.5 * 255 = 127.5 +- Bayer dither of +-2 LSB


def test_oracle_bloom_properties():
    st = capi.default_post_settings()
    img = np.full((96, 128, 4), 0.3, np.float32)
    _, bloom = ol.post_process(img, st, want_bloom=True)
    assert bloom.shape == (48, 64, 3) and np.abs(bloom).max() == 0.0            # nothing above the threshold: no bloom
    img[40:44, 60:64, :3] = 50.0
    out, bloom = ol.post_process(img, st, want_bloom=True)
    assert bloom.max() > 0.5 and bloom[22, 31].sum() > bloom[2, 2].sum()         # energy around the hot spot, falling off
    st.IsBloom = 0
    plain = ol.post_process(img, st)
    assert (out.astype(int) >= plain.astype(int) - 1).all() and (out.astype(int) > plain.astype(int) + 3).any()


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,bloom,tonemap,minus", [(256, 144, 1, 1, 3), (250, 131, 1, 1, 1), (64, 48, 0, 1, 3), (97, 33, 1, 0, 0), (2, 2, 1, 1, 3)])
def test_post_process_bit_exact(w, h, bloom, tonemap, minus):
    from idkengine_b200.pathtracer import PathTracer
    st = capi.default_post_settings()
    st.IsBloom, st.DoTonemapAndSrgbTransform, st.BloomMinusLods = bloom, tonemap, minus
    st.Exposure = 0.3
    img = synthetic_hdr(w, h, seed=w)
    with PathTracer(w, h) as pt:
        pt.WriteResult(img)
        g, ms = pt.PostProcess(st)
        g2, _ = pt.PostProcess(st)
    o = ol.post_process(img, st)
    assert np.array_equal(g, o), int(np.abs(g.astype(int) - o.astype(int)).max())
    assert np.array_equal(g, g2) and ms > 0


@pytest.mark.gpu
def test_post_process_of_path_traced_frame(cornell):
    from idkengine_b200.pathtracer import PathTracer
    scene, cam = cornell
    w, h = 160, 120
    with PathTracer(w, h) as pt:
        pt.SetScene(scene)
        pt.SetSky((0.6, 0.7, 0.9))
        pt.SetFrame(scenes.camera_frame(cam, w, h))
        for _ in range(3):
            pt.Compute()
        hdr = pt.Result.copy()
        ldr, _ = pt.PostProcess()
        assert pt.AccumulatedSamples == 3                    # presenting does not disturb the accumulation
        none, _ = pt.PostProcess(download=False)
        assert none is None
    assert np.array_equal(ldr, ol.post_process(hdr))
    assert ldr[..., :3].std() > 10


@pytest.mark.gpu
def test_post_process_errors():
    from idkengine_b200.pathtracer import PathTracer
    with PathTracer(64, 64, tile=(8, 0, 2)) as pt:
        with pytest.raises(RuntimeError):
            pt.PostProcess()                                 # a tile holds only its own rows
    with PathTracer(1, 1) as pt:
        with pytest.raises(RuntimeError):
            pt.PostProcess()                                 # bloom needs 2x2
