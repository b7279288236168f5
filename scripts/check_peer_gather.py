"""torchrun --nproc-per-node N scripts/check_peer_gather.py : the peer-memory gather fused into Compute() must produce
the same full image on every rank as the NCCL all_gather + de-interleave path, and as the CPU oracle per tile."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np, torch, torch.distributed as dist
from idkengine_b200 import capi, scenes, multigpu
from idkengine_b200.pathtracer import PathTracer
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
scene, cam = scenes.cornell_1k(threads=1)
W, H, stripe = 320, 200, 8
frame = scenes.camera_frame(cam, W, H)
s = capi.default_settings(); s.RayDepth = 5; s.OutputAOVs = 1
pt = PathTracer(W, H, s, device=lr, tile=(stripe, rank, world))
pt.SetScene(scene); pt.SetSky((0.6, 0.7, 0.9)); pt.SetFrame(frame)
def exchange(blob):
    out = [None] * world
    dist.all_gather_object(out, blob)
    return out
pt.EnablePeerGather(rank, world, exchange)
rows = pt.TileRows()
ptr, _ = pt.ResultDevicePtr()
local = torch.as_tensor(multigpu.DeviceArray(ptr, (len(rows), W, 4)), device=f"cuda:{lr}")
g = multigpu.TileGatherer(H, W, 4, stripe, world, torch.device("cuda", lr))
pinned = torch.zeros((H, W, 4), dtype=torch.float32).pin_memory()
ok = True
for k in range(4):
    pt.Compute()
    p, n = pt.GatheredDevicePtr()
    peer = torch.as_tensor(multigpu.DeviceArray(p, (H, W, 4)), device=f"cuda:{lr}").clone()
    ref = g.gather(local).clone()
    ok &= bool(torch.equal(peer, ref))
    pt.PresentAsync(pinned.data_ptr(), pinned.numel() * 4, which=capi.IDKPT_IMAGE_GATHERED); pt.PresentWait()
    ok &= bool(np.array_equal(pinned.numpy(), ref.cpu().numpy()))
if rank == 0:
    import oracle_lib as ol
    img = np.zeros((H, W, 4), np.float32)
    for t in range(world):
        acc = 0
        for k in range(4):
            o = ol.path_trace(scene, frame, s, W, H, tile=(stripe, t, world), accumulated=acc, result=img, want_rays=False)
            acc = o.accumulated
    ok &= bool(np.array_equal(img, ref.cpu().numpy()))
t = torch.tensor([1.0 if ok else 0.0], device=f"cuda:{lr}")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0: print("PEER_GATHER_OK" if t.item() == 1.0 else "PEER_GATHER_MISMATCH")
pt.Dispose(); dist.destroy_process_group()
