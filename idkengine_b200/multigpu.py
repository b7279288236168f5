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


def all_gather_tiles(local_rows, height, stripe, world, group=None):
    """local_rows: torch tensor [rows_local, W, C] on this rank's device (compact tile rows).
    Returns the full [height, W, C] image on every rank after a single all_gather."""
    import torch
    import torch.distributed as dist
    if world <= 1:
        return local_rows
    pad_rows = max_tile_rows(height, stripe, world)
    w, c = local_rows.shape[1], local_rows.shape[2]
    send = torch.zeros((pad_rows, w, c), dtype=local_rows.dtype, device=local_rows.device)
    send[: local_rows.shape[0]] = local_rows
    recv = torch.empty((world, pad_rows, w, c), dtype=local_rows.dtype, device=local_rows.device)
    dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=group) if local_rows.is_cuda else \
        dist.all_gather(list(recv.unbind(0)), send, group=group)
    full = torch.empty((height, w, c), dtype=local_rows.dtype, device=local_rows.device)
    for r in range(world):
        rows = torch.as_tensor(tile_rows(height, stripe, r, world), device=local_rows.device, dtype=torch.long)
        full[rows] = recv[r, : len(rows)]
    return full


class DeviceArray:
    """Zero-copy view of a libidkpt device buffer for torch (torch.as_tensor(DeviceArray(...), device='cuda'))."""

    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3}
