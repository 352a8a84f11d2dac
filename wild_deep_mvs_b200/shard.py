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


class DepthGather:
    """The ONE collective of the path with its buffers owned up front: a preallocated
    [world * n_local, ...] buffer per rank; the regression kernel (K3) writes this rank's maps straight into
    `local` -- the rank's slice of that buffer -- and `all_gather()` is an IN-PLACE ncclAllGather (send buffer = the
    slice, no staging copy, no allocation), which is legal inside a CUDA-graph capture: a captured step is then
    K1 ... K3 -> all-gather as one graph launch with nothing issued from Python in between.

    Blocks must be equal (n_items % world == 0): that is how BASELINE cfg5 (64 reference views over 1/2/4/8 GPUs) and
    the reference's trainer (one view per rank, models/trainer.py:101,246-247) shard; ragged blocks go through
    gather_depth_maps()."""

    def __init__(self, n_local, map_shape, device, dtype=torch.float32, group=None):
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.group = group
        self.n_local = n_local
        self.all = torch.zeros((self.world * n_local,) + tuple(map_shape), device=device, dtype=dtype)
        self.local = self.all[self.rank * n_local:(self.rank + 1) * n_local]

    def all_gather(self):
        """Every rank's `local` slice -> `all` on every rank (ordered by global sample index).  Stream-ordered on the
        current stream; capturable.  Returns `all`."""
        if self.world > 1:
            dist.all_gather_into_tensor(self.all, self.local, group=self.group)
        return self.all
