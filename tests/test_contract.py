"""The data contract: numpy dtypes == C header == reference struct sizes."""
import os
import re

import numpy as np

from idkengine_b200 import gpu_types as gt

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sizes_match_header_static_asserts():
    hdr = open(os.path.join(REPO, "include", "idk_gpu_types.h")).read() + open(os.path.join(REPO, "include", "idkpt.h")).read()
    asserted = dict(re.findall(r"IDK_STATIC_ASSERT\(sizeof\((\w+)\) == (\d+)", hdr))
    assert len(asserted) >= 17
    for name, size in asserted.items():
        assert getattr(gt, name).itemsize == int(size), name


def test_offsets():
    assert gt.GpuPerFrameData.fields["InvView"][1] == 128
    assert gt.GpuPerFrameData.fields["ViewPos"][1] == 256
    assert gt.GpuPerFrameData.fields["InvProjection"][1] == 336
    assert gt.GpuMaterial.fields["BaseColorTexture"][1] == 48
    assert gt.GpuMaterial.fields["IsVolumetric"][1] == 88
    assert gt.GpuMesh.fields["TintOnTransmissive"][1] == 92
    assert gt.GpuBlasNode.fields["TriStartOrChild"][1] == 12
    assert gt.GpuBlasNode.fields["TriCount"][1] == 28


def test_header_compiles_as_c(tmp_path):
    import subprocess
    src = tmp_path / "t.c"
    src.write_text('#include "idkpt.h"\nint main(void){return (int)sizeof(IdkPtStats) == 0;}\n')
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(REPO, "include"), "-c", str(src),
                    "-o", str(tmp_path / "t.o")], check=True)


def test_compression_roundtrip():
    n = np.array([[0.0, 1.0, 0.0], [0.6, -0.8, 0.0], [-1.0, 0.0, 0.0]], np.float32)
    packed = gt.compress_sr11g11b10(n)
    r = (packed & 2047).astype(np.float32) / 2047 * 2 - 1
    assert np.allclose(r, n[:, 0], atol=1e-3)
    assert gt.pack_unorm4x8(np.array([1.0, 0.0, 0.5, 1.0])) == (255 | (0 << 8) | (128 << 16) | (255 << 24))
