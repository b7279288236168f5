"""CPU-side tests of the boundary: the C-ABI library loads, exports every declared symbol, and refuses to run
without a GPU (no CPU fallback); the oracle is deterministic; multi-GPU tiling logic (gloo, world_size 2)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as ol
from idkengine_b200 import build, capi, scenes, multigpu
from idkengine_b200 import gpu_types as gt

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libidkpt():
    if not os.path.exists(build.LIBIDKPT):
        build.build_cuda()
    return capi.load()


def test_library_exports_every_declared_symbol(libidkpt):
    hdr = open(os.path.join(REPO, "include", "idkpt.h")).read()
    declared = set(re.findall(r"IDKPT_API\s+[\w\s\*]+?\b(idkpt_\w+)\s*\(", hdr))
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)
    for name in declared:
        assert hasattr(libidkpt, name), name
    assert libidkpt.idkpt_abi_version() == 4


def test_host_library_exports_cache_symbols():
    from idkengine_b200 import host
    hdr = open(os.path.join(REPO, "include", "idkhost_cache.h")).read()
    declared = set(re.findall(r"\b(idkhost_\w+)\s*\(", hdr))
    assert {"idkhost_hash64", "idkhost_cache_save", "idkhost_cache_open", "idkhost_cache_array", "idkhost_cache_close"} == declared
    L = host.lib()
    for name in declared:
        assert hasattr(L, name), name


def test_integration_doc_names_every_entry_point():
    doc = open(os.path.join(REPO, "INTEGRATION.md")).read()
    vx = set(re.findall(r"\b(idkvx_\w+)\s*\(", open(os.path.join(REPO, "include", "idkvx.h")).read()))
    missing = [n for n in list(capi.EXPORTS) + sorted(vx) if n not in doc]
    assert not missing, missing


# ---- INTEGRATION.md's C# structs vs include/idkpt.h (field names, order, offsets, sizes) -------------------------------------
_C_SIZES = {"int32_t": 4, "uint32_t": 4, "float": 4, "uint64_t": 8, "int64_t": 8, "uint8_t": 1, "IdkPtGpuSettings": 20}
_CS_SIZES = {"int": 4, "uint": 4, "float": 4, "ulong": 8, "long": 8, "nint": 8, "byte": 1, "PathTracer.GpuSettings": 20}


def _layout(fields):
    """fields: [(name, elem_size, count, align)] -> [(name, offset, bytes)] with natural (C / LayoutKind.Sequential) alignment."""
    out, off, amax = [], 0, 1
    for name, size, count, align in fields:
        off = (off + align - 1) // align * align
        out.append((name, off, size * count))
        off += size * count
        amax = max(amax, align)
    return out, (off + amax - 1) // amax * amax


def _c_struct_fields(hdr, name):
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), hdr, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        m = re.match(r"(?:const )?([\w ]+?)\s*(\*?)\s*((?:\w+(?:\[\w+\])?(?:\s*,\s*)?)+)$", decl)
        assert m, decl
        ctype, ptr, names = m.group(1).strip(), m.group(2), m.group(3)
        for n in names.split(","):
            n = n.strip()
            count = 1
            am = re.match(r"(\w+)\[(\w+)\]", n)
            if am:
                n, count = am.group(1), {"IDKPT_MAX_RAY_DEPTH": 64}.get(am.group(2)) or int(am.group(2))
            size = 8 if ptr else _C_SIZES[ctype]
            fields.append((n, size, count, min(size, 8) if ctype != "IdkPtGpuSettings" or ptr else 4))
    return fields


def _cs_struct_fields(doc, name):
    body = re.search(r"public struct %s\s*\{(.*?)\}" % name, doc, re.S).group(1)
    body = re.sub(r"//[^\n]*|/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        m = re.match(r"public (fixed )?([\w\.]+?)(\*?) (.+)$", decl)
        assert m, decl
        fixed, ctype, ptr, names = m.groups()
        for n in names.split(","):
            n = n.strip()
            count = 1
            am = re.match(r"(\w+)\[(\d+)\]", n)
            if am:
                assert fixed, decl
                n, count = am.group(1), int(am.group(2))
            size = 8 if ptr else _CS_SIZES[ctype]
            fields.append((n, size, count, 4 if ctype == "PathTracer.GpuSettings" and not ptr else min(size, 8)))
    return fields


def test_integration_doc_structs_match_the_c_header():
    """Every [StructLayout] struct of INTEGRATION.md section 1 must be byte-compatible with its C twin in include/idkpt.h:
    same field names in the same order at the same offsets, same total size (the round-1 doc missed Textures/TextureCount)."""
    doc = open(os.path.join(REPO, "INTEGRATION.md")).read()
    hdr = open(os.path.join(REPO, "include", "idkpt.h")).read()
    cs_names = re.findall(r"\[StructLayout\(LayoutKind\.Sequential\)\]\s*public struct (\w+)", doc)
    c_names = [n for n in re.findall(r"typedef struct (IdkPt\w+) \{", hdr)]
    assert sorted("IdkPt" + n for n in cs_names) == sorted(c_names), (cs_names, c_names)
    ctypes_twins = {"IdkPtCreateInfo": capi.IdkPtCreateInfo, "IdkPtSceneDesc": capi.IdkPtSceneDesc, "IdkPtSettings": capi.IdkPtSettings,
                    "IdkPtStats": capi.IdkPtStats}
    for n in cs_names:
        cl, csize = _layout(_c_struct_fields(hdr, "IdkPt" + n))
        sl, ssize = _layout(_cs_struct_fields(doc, n))
        assert cl == sl, (n, cl, sl)
        assert csize == ssize, (n, csize, ssize)
        if "IdkPt" + n in ctypes_twins:      # and the ctypes twin the tests call through agrees with both
            assert ctypes.sizeof(ctypes_twins["IdkPt" + n]) == csize, n


def test_no_cpu_fallback(libidkpt):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ctx = ctypes.c_void_p()
    ci = capi.IdkPtCreateInfo(0, 64, 64, 8, 0, 1, 0)
    rc = libidkpt.idkpt_create(ctypes.byref(ci), ctypes.byref(ctx))
    assert rc == -2 and not ctx.value            # IDKPT_ERR_NO_DEVICE
    assert b"no CUDA device" in libidkpt.idkpt_last_error(None)


def test_product_never_imports_oracle():
    # the strict check: no file of the package opens, imports or links anything under oracle/
    pat = re.compile(r"(import\s+oracle|from\s+oracle|liboracle|oracle/\w+\.(so|cpp))")
    for root, _, files in os.walk(os.path.join(REPO, "idkengine_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(root, f)).read()
                assert not pat.search(text.replace("oracle/build.py", "")), os.path.join(root, f)


def test_oracle_deterministic_and_thread_independent(cornell):
    scene, cam = cornell
    frame = scenes.camera_frame(cam, 64, 64)
    s = capi.default_settings()
    s.OutputAOVs = 1
    a = ol.path_trace(scene, frame, s, 64, 64, threads=1)
    b = ol.path_trace(scene, frame, s, 64, 64, threads=4)
    assert np.array_equal(a.result, b.result) and np.array_equal(a.albedo, b.albedo)
    assert a.stats.Rays == b.stats.Rays and a.stats.NodePairFetches == b.stats.NodePairFetches
    assert a.accumulated == 1
    assert np.all(a.result[..., 3] == 1.0) and np.isfinite(a.result).all()
    assert list(a.stats.BounceRays)[:7] == sorted(list(a.stats.BounceRays)[:7], reverse=True)


def test_oracle_accumulation_is_running_mean(cornell):
    scene, cam = cornell
    frame = scenes.camera_frame(cam, 32, 32)
    s = capi.default_settings()
    one = ol.path_trace(scene, frame, s, 32, 32)
    two = ol.path_trace(scene, frame, s, 32, 32, accumulated=1, result=one.result.copy())
    s2 = capi.default_settings()
    s2.SamplesPerPixel = 2
    both = ol.path_trace(scene, frame, s2, 32, 32)
    assert np.array_equal(two.result, both.result) and both.accumulated == 2


def test_det_math_against_libm():
    x = np.linspace(0, 2 * np.pi, 20001).astype(np.float32)
    s, c = np.zeros_like(x), np.zeros_like(x)
    ol.lib().oracle_det_sincos(x.ctypes.data, len(x), s.ctypes.data, c.ctypes.data)
    assert np.abs(s - np.sin(x.astype(np.float64))).max() < 3e-7
    assert np.abs(c - np.cos(x.astype(np.float64))).max() < 3e-7
    e = np.linspace(-80, 0, 8001).astype(np.float32)
    y = np.zeros_like(e)
    ol.lib().oracle_det_exp(e.ctypes.data, len(e), y.ctypes.data)
    ref = np.exp(e.astype(np.float64))
    assert np.all(np.abs(y - ref) <= 3e-7 * ref + 1e-44)


def test_pcg_known_values():
    """Random.glsl:16-23 PCG hash: seed 0 -> state 2891336453; reference values computed by hand from the formula."""
    nxt = ctypes.c_uint32()
    r = ol.lib().oracle_pcg(0, ctypes.byref(nxt))
    assert nxt.value == 2891336453
    state = 2891336453
    word = (((state >> ((state >> 28) + 4)) ^ state) * 277803737) & 0xFFFFFFFF
    assert r == ((word >> 22) ^ word)


def test_octahedral_roundtrip():
    rng = np.random.RandomState(3)
    d = rng.normal(size=(5000, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    enc, dec = np.zeros((5000, 2), np.float32), np.zeros((5000, 3), np.float32)
    ol.lib().oracle_encode_decode(d.ctypes.data, 5000, enc.ctypes.data, dec.ctypes.data)
    assert enc.min() >= 0 and enc.max() <= 1
    assert np.abs(dec - d).max() < 1e-5


def test_tile_rows_partition():
    for h, stripe, world in [(1080, 8, 8), (1080, 8, 2), (2160, 16, 4), (67, 8, 3)]:
        rows = [multigpu.tile_rows(h, stripe, i, world) for i in range(world)]
        assert sorted(np.concatenate(rows).tolist()) == list(range(h))


def test_oracle_tiles_first_hit_is_tile_independent(cornell):
    """FirstHit seeds depend only on the pixel, so with RayDepth 1 the union of per-tile images equals the full image."""
    scene, cam = cornell
    frame = scenes.camera_frame(cam, 64, 48)
    s = capi.default_settings()
    s.RayDepth = 1
    full = ol.path_trace(scene, frame, s, 64, 48)
    img = np.zeros_like(full.result)
    for t in range(3):
        ol.path_trace(scene, frame, s, 64, 48, tile=(8, t, 3), result=img)
    assert np.array_equal(img, full.result)


GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, torch, torch.distributed as dist
import oracle_lib as ol
from idkengine_b200 import scenes, capi, multigpu
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
scene, cam = scenes.cornell_1k(threads=1)
W, H, stripe = 64, 40, 8
frame = scenes.camera_frame(cam, W, H)
s = capi.default_settings(); s.RayDepth = 4
img = np.zeros((H, W, 4), np.float32)
ol.path_trace(scene, frame, s, W, H, tile=(stripe, rank, world), result=img, threads=1)
rows = multigpu.tile_rows(H, stripe, rank, world)
full = multigpu.all_gather_tiles(torch.from_numpy(img[rows].copy()), H, stripe, world).numpy()
ref = np.zeros((H, W, 4), np.float32)
for t in range(world):
    ol.path_trace(scene, frame, s, W, H, tile=(stripe, t, world), result=ref, threads=1)
assert np.array_equal(full, ref), "gathered image differs"
if rank == 0: print("GLOO_OK")
'''


def test_two_rank_gloo_tile_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script), REPO],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "GLOO_OK" in out.stdout


def test_slab_partition_properties():
    """multigpu.slab_range (z-slabs of the voxel grid / row bands of the cone trace): a partition of [0, n) into `world`
    contiguous ranges whose sizes differ by at most one, equal when world divides n (the in-place all-gather case)."""
    for n in (1, 7, 30, 31, 384, 1080):
        for world in (1, 2, 3, 4, 8):
            rs = [multigpu.slab_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1 and (n % world != 0 or len(set(sizes)) == 1)


GLOO_SLAB_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, torch, torch.distributed as dist
import oracle_lib as ol
from idkengine_b200 import scenes, vxgi, multigpu
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
scene, cam = scenes.cornell_1k(threads=1)
scene.add_light((0.0, 1.6, 0.3), (6.0, 5.5, 5.0), 0.2)
ci = vxgi.create_info((20, 24, 18), (-1.2, -0.2, -1.2), (1.2, 2.2, 1.2))
levels, raw, frags = ol.vx_voxelize(scene, ci)
level0 = levels[0].view(np.uint16).reshape(18, -1).astype(np.int32)       # [z, y*x*4]
z0, z1 = multigpu.slab_range(18, rank, world)
mine = np.zeros_like(level0); mine[z0:z1] = level0[z0:z1]                 # what this rank's slab voxelisation leaves in its grid
parts = [torch.zeros((multigpu.slab_range(18, r, world)[1] - multigpu.slab_range(18, r, world)[0], level0.shape[1]), dtype=torch.int32) for r in range(world)]
dist.all_gather(parts, torch.from_numpy(mine[z0:z1].copy()))
merged = torch.cat(parts).numpy()
assert np.array_equal(merged, level0), "slab gather differs from the full grid"
if rank == 0: print("GLOO_SLAB_OK")
'''


def test_two_rank_gloo_slab_gather(tmp_path):
    """Host logic of the multi-GPU VXGI split on CPU (gloo, world_size 2): contiguous z-slabs of the oracle's level 0, gathered in
    rank order, are the full level."""
    script = tmp_path / "slab_worker.py"
    script.write_text(GLOO_SLAB_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29537")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29537", str(script), REPO],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "GLOO_SLAB_OK" in out.stdout


def test_oracle_any_hit_and_shadows(multi_blas):
    """Oracle-only sanity for the 8f.1 rows: any-hit occlusion == closest-hit "found something"; the shadow pass is dark
    behind occluders, lit in the open, and leaves sky pixels untouched."""
    scene, cam = multi_blas
    rng = np.random.default_rng(4)
    rays = np.zeros(6000, gt.IdkPtRay)
    rays["Origin"] = rng.uniform(-2.5, 2.5, (6000, 3)).astype(np.float32)
    d = rng.normal(size=(6000, 3))
    rays["Direction"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays["TMax"] = np.float32(3.4028235e38)
    rays["TMax"][::3] = 1.0
    a = ol.trace_rays_any(scene, rays, trace_lights=True)
    c = ol.trace_rays(scene, rays, trace_lights=True)
    assert np.array_equal(a["NodePairFetches"] == 1, c["T"] != rays["TMax"])
    assert (a["T"][a["NodePairFetches"] == 1] >= c["T"][a["NodePairFetches"] == 1]).all()   # first accepted hit is never closer than the closest
    w, h = 96, 64
    frame = scenes.camera_frame(cam, w, h)
    depth, nrg, _ = ol.synth_gbuffer(scene, frame, w, h)
    vis0 = np.full((h, w), 7.0, np.float32)
    vis = ol.shadows_ray_traced(scene, frame, depth, nrg, 0, samples=2, visibility=vis0.copy())
    assert (vis[depth == 1.0] == 7.0).all()
    inside = vis[depth < 1.0]
    assert ((inside >= 0.0) & (inside <= 1.0)).all() and (inside == 0.0).any() and (inside == 1.0).any()


def test_cpp_host_mirror_compiles_links_and_fails_loudly_without_gpu(tmp_path):
    """include/idkpt.hpp (the C++ stand-in for the C# PathTracerNative) against the built library."""
    exe = str(tmp_path / "hpp_smoke")
    libdir = os.path.dirname(build.LIBIDKPT)
    cmd = ["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-Wno-comment", "-I", os.path.join(REPO, "include"), os.path.join(REPO, "tests", "cpp", "hpp_smoke.cpp"),
           "-L", libdir, "-lidkpt", "-Wl,-rpath," + libdir, "-o", exe]
    subprocess.run(cmd, check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("OK")
