"""Multi-GPU screen tiling: one process per GPU, the scene replicated, image rows cut into stripes
dealt round-robin to the ranks, and ONE all-gather of tile radiance per frame (torch.distributed; NCCL over NVLink on
the GPU box, gloo in the CPU tests). The reference has no multi-GPU path; this is new work behind the same surface."""
import numpy as np


def tile_rows(height, stripe, index, count):
    """Rows owned by tile `index`: stripes of `stripe` rows dealt round-robin (must match idkpt_create)."""
    y = np.arange(height)
    if count <= 1:
        return y
    return y[(y // stripe) % count == index]


def max_tile_rows(height, stripe, count):
    return max(len(tile_rows(height, stripe, i, count)) for i in range(max(count, 1)))


class TileGatherer:
    """Pre-allocated single-collective gather of tile rows into the full image (one all_gather + one index_select)."""

    def __init__(self, height, width, channels, stripe, world, device, dtype=None, group=None):
        import torch
        self.world, self.group, self.height = world, group, height
        dtype = dtype or torch.float32
        self.pad_rows = max_tile_rows(height, stripe, world)
        self.send = torch.zeros((self.pad_rows, width, channels), dtype=dtype, device=device)
        self.recv = torch.empty((world, self.pad_rows, width, channels), dtype=dtype, device=device)
        # source row (in the flattened [world*pad_rows] receive buffer) of every image row
        src = np.zeros(height, np.int64)
        for r in range(world):
            rows = tile_rows(height, stripe, r, world)
            src[rows] = r * self.pad_rows + np.arange(len(rows))
        self.src = torch.as_tensor(src, device=device)
        self.full = torch.empty((height, width, channels), dtype=dtype, device=device)

    def gather(self, local_rows):
        import torch
        import torch.distributed as dist
        if self.world <= 1:
            return local_rows
        n = local_rows.shape[0]
        send = local_rows if n == self.pad_rows and local_rows.is_contiguous() else self.send
        if send is self.send:
            self.send[:n].copy_(local_rows)
        if local_rows.is_cuda:
            dist.all_gather_into_tensor(self.recv.view(-1), send.view(-1), group=self.group)
        else:
            dist.all_gather(list(self.recv.unbind(0)), send, group=self.group)
        torch.index_select(self.recv.view(self.world * self.pad_rows, *self.recv.shape[2:]), 0, self.src, out=self.full)
        return self.full


def all_gather_tiles(local_rows, height, stripe, world, group=None):
    """Convenience wrapper (allocates): full [height, W, C] image on every rank after a single all_gather."""
    if world <= 1:
        return local_rows
    g = TileGatherer(height, local_rows.shape[1], local_rows.shape[2], stripe, world, local_rows.device, local_rows.dtype, group)
    return g.gather(local_rows).clone()


class DeviceArray:
    """Zero-copy view of a libidkpt device buffer for torch (torch.as_tensor(DeviceArray(...), device='cuda'))."""

    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3}


# ---- VXGI over N GPUs: z-slab voxelisation + ONE all-gather of the slabs, mip chain and cone tracing replicated / tiled
def slab_range(depth, rank, world):
    """z range of rank's slab: equal slabs when world divides depth, otherwise the last ranks get one layer less."""
    base, extra = divmod(depth, world)
    z0 = rank * base + min(rank, extra)
    return z0, z0 + base + (1 if rank < extra else 0)


def voxelize_multi_gpu(vx, rank, world, device, group=None):
    """Voxelizer.Render() across `world` ranks: every rank voxelises its z-slab of level 0, the slabs are all-gathered straight
    into every rank's grid (a slab of the linear x-fastest level is one contiguous range; NCCL all_gather_into_tensor, in place
    when the slabs are equal), then every rank builds the mip chain. The merged grid equals the single-GPU grid bit for bit
    (the voxel merge is a per-channel max). Returns (voxelise stats of this rank, mip stats)."""
    import torch
    import torch.distributed as dist
    w, h, d = vx.sizes[0]
    z0, z1 = slab_range(d, rank, world)
    vx.SetSlab(z0, z1)
    st = vx.Render()
    ptr, nbytes = vx.LevelDevicePtr(0)
    level0 = torch.as_tensor(DeviceArray(ptr, (d, h * w * 2), "<i4"), device=device)     # 8 bytes per texel as 2 x int32 (NCCL / torch ops have no uint32)
    if world > 1:
        if d % world == 0:
            dist.all_gather_into_tensor(level0.view(-1), level0[z0:z1].reshape(-1), group=group)
        else:
            parts = [torch.empty((slab_range(d, r, world)[1] - slab_range(d, r, world)[0], h * w * 2), dtype=level0.dtype, device=device) for r in range(world)]
            dist.all_gather(parts, level0[z0:z1].contiguous(), group=group)
            for r, part in enumerate(parts):
                a, b = slab_range(d, r, world)
                level0[a:b].copy_(part)
        torch.cuda.synchronize(device)
    mst = vx.Mipmap()
    vx.SetSlab(0, d)
    return st, mst
