"""View sharding across the GPUs of one box.

Reference views are independent units (rank r takes view r as its reference at
models/trainer.py:101 of the reference and all-gathers the depth maps at :246-247), so the path shards
with no data-path collective: each rank processes a contiguous block of samples and ONE all-gather of the
per-view depth maps at the end makes every rank hold all of them.
"""
import ctypes
import os

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

    backend "lib" (default on CUDA; MVSB200_GATHER=torch selects the other): the library's own NCCL communicator and
    `mvsb200_allgather_depth` (include/mvsb200.h) -- torch.distributed only carries the 128-byte ncclUniqueId from rank 0
    to the other ranks once, at construction;  backend "torch": `dist.all_gather_into_tensor` on the process group
    (CPU / gloo tests, or a box without NCCL).

    Blocks must be equal (n_items % world == 0): that is how BASELINE cfg5 (64 reference views over 1/2/4/8 GPUs) and
    the reference's trainer (one view per rank, models/trainer.py:101,246-247) shard; ragged blocks go through
    gather_depth_maps()."""

    _comm = {}   # (device index, world, rank) -> library communicator, shared by every DepthGather of the process

    def __init__(self, n_local, map_shape, device, dtype=torch.float32, group=None, backend=None):
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.group = group
        self.n_local = n_local
        self.all = torch.zeros((self.world * n_local,) + tuple(map_shape), device=device, dtype=dtype)
        self.local = self.all[self.rank * n_local:(self.rank + 1) * n_local]
        if backend is None:
            backend = os.environ.get("MVSB200_GATHER", "lib" if self.all.is_cuda else "torch")
        if backend not in ("lib", "torch"):
            raise ValueError("DepthGather: backend %r (expected 'lib' or 'torch')" % (backend,))
        if backend == "lib" and not (self.all.is_cuda and dtype == torch.float32 and group is None):
            backend = "torch"
        self.backend = backend
        self.comm = self._communicator() if (backend == "lib" and self.world > 1) else None

    def _communicator(self):
        from . import _lib as L
        key = (self.all.device.index, self.world, self.rank)
        if key not in DepthGather._comm:
            lib = L.load()
            ident = ctypes.create_string_buffer(128)
            if self.rank == 0:
                L.check(lib.mvsb200_gather_unique_id(ident), "mvsb200_gather_unique_id")
            box = [ident.raw]
            dist.broadcast_object_list(box, src=0)          # the one use of torch.distributed: carry the id to every rank
            ident = ctypes.create_string_buffer(box[0], 128)
            comm = ctypes.c_void_p()
            with torch.cuda.device(self.all.device):
                L.check(lib.mvsb200_gather_init(ident, self.world, self.rank, ctypes.byref(comm)), "mvsb200_gather_init")
            DepthGather._comm[key] = comm
        return DepthGather._comm[key]

    def all_gather(self):
        """Every rank's `local` slice -> `all` on every rank (ordered by global sample index).  Stream-ordered on the
        current stream; capturable.  Returns `all`."""
        if self.world > 1:
            if self.comm is not None:
                from . import _lib as L
                with torch.cuda.device(self.all.device):
                    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
                    L.check(L.load().mvsb200_allgather_depth(self.comm, ctypes.c_void_p(self.all.data_ptr()), self.local.numel(),
                                                              self.rank, stream), "mvsb200_allgather_depth")
            else:
                dist.all_gather_into_tensor(self.all, self.local, group=self.group)
        return self.all

    @classmethod
    def destroy_communicators(cls):
        """Destroy the library communicators of this process (after every CUDA graph that captured a gather is gone)."""
        from . import _lib as L
        for comm in cls._comm.values():
            L.load().mvsb200_gather_destroy(comm)
        cls._comm.clear()
