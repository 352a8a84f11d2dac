"""View sharding across the GPUs of one box.

Reference views are independent units (rank r takes view r as its reference at
models/trainer.py:101 of the reference and all-gathers the depth maps at :246-247), so the path shards
with no data-path collective: each rank processes a contiguous block of samples and ONE all-gather of the
per-view depth maps at the end makes every rank hold all of them.
"""
import torch
import torch.distributed as dist


def block_partition(n_items, world_size, rank):
    """Contiguous, balanced [start, stop) block of `n_items` for `rank` (first ranks take the remainder)."""
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_depth_maps(local_maps, n_items=None):
    """local_maps [n_local, ...] on every rank -> [n_items, ...] on every rank, ordered by global sample index.

    Equal block sizes use a single all_gather_into_tensor (one NCCL all-gather over NVLink/NVSwitch);
    ragged blocks are padded to the largest block, gathered once, and trimmed."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_maps
    world = dist.get_world_size()
    n_local = local_maps.shape[0]
    if n_items is None:
        n_items = n_local * world
    sizes = [block_partition(n_items, world, r) for r in range(world)]
    n_max = max(b - a for a, b in sizes)
    send = local_maps
    if n_local != n_max:
        send = torch.zeros((n_max,) + tuple(local_maps.shape[1:]), dtype=local_maps.dtype, device=local_maps.device)
        send[:n_local] = local_maps
    out = torch.empty((world * n_max,) + tuple(local_maps.shape[1:]), dtype=local_maps.dtype, device=local_maps.device)
    dist.all_gather_into_tensor(out, send.contiguous())
    if all(b - a == n_max for a, b in sizes):
        return out
    return torch.cat([out[r * n_max: r * n_max + (b - a)] for r, (a, b) in enumerate(sizes)], 0)
