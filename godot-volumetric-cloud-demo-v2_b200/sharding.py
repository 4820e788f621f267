"""Multi-GPU sharding of the cloud march.

Every pixel depends only on the push constants, the three textures, the sky LUT and its own coordinates
(clouds.glsl:258-266); the reference already renders the texture as independent tiles addressed by
update_position (cloud_sky.gd:156-161).  So ranks need no data-path exchange while rendering; the only
collective is ONE all-gather of the finished RGBA16F texture(s):

  * one frame   : contiguous row bands, H/N rows per rank, rendered straight into the rank's slice of the
                  full-size buffer, then an in-place all-gather (each band is a contiguous chunk).
  * sun sweep   : n/N whole frames per rank (each rank builds its own sky LUTs), gathered [n, H, W, 4].

One process per GPU, torch.distributed for the plumbing (NCCL on GPUs; gloo in the CPU tests, where the
"device" buffers are host tensors and the compute backend is the oracle library).

Two gathers: gather="nccl" is the library collective (ncclAllGather, whose copy CTAs share the SMs with the march);
gather="peer" is the fused one (PeerBuffers): every rank maps its peers' copies of the gathered buffer (CUDA IPC over
NVLink/NVSwitch peer access), the march kernel itself stores each finished pixel into all copies, and what is left of the
collective is a flag barrier (cs_peer_barrier) — no collective kernel, no copy pass."""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np


def row_bands(height: int, world: int) -> List[Tuple[int, int]]:
    """Equal contiguous row bands [r0, r1) per rank (height must divide evenly so that the in-place
    all-gather sees equal chunks)."""
    if world < 1 or height % world != 0:
        raise ValueError(f"height {height} is not divisible by world size {world}")
    rows = height // world
    return [(r * rows, (r + 1) * rows) for r in range(world)]


def interleaved_bands(height: int, world: int, bands_per_rank: int) -> List[List[Tuple[int, int]]]:
    """Row bands for a frame whose cost varies with the row (the lit fraction does): the image is cut into
    world * bands_per_rank equal bands and rank r takes bands r, r + world, r + 2 world, ...  Band group g (bands g*world ..
    g*world + world - 1) is a contiguous block of `world` equal chunks in rank order, so it is gathered by one in-place
    all-gather.  Returns, per rank, its bands_per_rank (r0, r1) ranges in group order."""
    n = world * bands_per_rank
    if world < 1 or bands_per_rank < 1 or height % n != 0:
        raise ValueError(f"height {height} is not divisible by {world} ranks x {bands_per_rank} bands")
    rows = height // n
    return [[((g * world + r) * rows, (g * world + r + 1) * rows) for g in range(bands_per_rank)] for r in range(world)]


def sun_sweep(n: int) -> np.ndarray:
    """BASELINE config 4 / SURVEY §8(d) C4: dir_k = (cos th_k, sin th_k, 0), th_k = pi (k + 0.5) / n."""
    th = math.pi * (np.arange(n, dtype=np.float64) + 0.5) / n
    return np.stack([np.cos(th), np.sin(th), np.zeros(n)], -1).astype(np.float32)


def sun_shard(n: int, world: int, rank: int) -> Tuple[int, int]:
    if n % world != 0:
        raise ValueError(f"{n} suns are not divisible by world size {world}")
    per = n // world
    return rank * per, (rank + 1) * per


def _all_gather_inplace(full_flat, chunk, group=None):
    """In-place all-gather of equal chunks.  The collective is chosen from the backend up front (never by catching an
    error: a rank that retried with a different collective than its peers would hang the job)."""
    import torch.distributed as dist
    if dist.get_backend(group) == "gloo":  # CPU tests: gloo has no flat in-place variant
        world = dist.get_world_size(group)
        parts = list(full_flat.chunk(world))
        dist.all_gather(parts, chunk.clone(), group=group)
    else:
        dist.all_gather_into_tensor(full_flat, chunk, group=group)


class _DevicePtr:
    """Minimal __cuda_array_interface__ holder so torch can view library-owned device memory without copying."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3, "strides": None}


class PeerBuffers:
    """`slots` copies-in-time of a gathered buffer of `slot_bytes`, replicated on every rank and mapped into every rank.

    After activate(), any march dispatch that writes into this rank's copy also writes the same pixels, at the same offset,
    into every peer's copy (cs_set_output_mirrors); barrier() then makes the copy complete on this rank's stream.
    Two slots used alternately make back-to-back steps safe: a peer starts overwriting slot s (step k+2) only after this rank
    published step k+1, which is stream-ordered after whatever this rank queued to consume step k.
    One registered range per context: activating a PeerBuffers deactivates the one that was active before.  Tensors returned
    by tensor() alias library memory and die with close()."""

    def __init__(self, ctx, slot_bytes: int, slots: int = 2, group=None):
        import torch.distributed as dist
        self.ctx, self.group = ctx, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if self.world > 8:
            raise ValueError("peer gather is a single-node mechanism (<= 8 GPUs)")
        self.slot_bytes, self.slots = int(slot_bytes), int(slots)
        self.nbytes = self.slot_bytes * self.slots
        self.base, handle = ctx.peer_alloc(self.nbytes)
        self.flags, fhandle = ctx.peer_alloc(256)
        handles = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(handles, (handle, fhandle), group=group)
        else:
            handles[0] = (handle, fhandle)
        self.bases, self.flag_ptrs = [], []
        for k, (h, fh) in enumerate(handles):
            self.bases.append(self.base if k == self.rank else ctx.peer_open(h))
            self.flag_ptrs.append(self.flags if k == self.rank else ctx.peer_open(fh))
        self.epoch = 0
        self.active = False
        self._views = {}

    def activate(self):
        prev = getattr(self.ctx, "_active_peer_buffers", None)
        if prev is not None and prev is not self:
            prev.active = False  # cs_set_output_mirrors replaces the registered range
        self.ctx.set_output_mirrors(self.base, self.nbytes, [b for k, b in enumerate(self.bases) if k != self.rank])
        self.ctx._active_peer_buffers = self
        self.active = True

    def deactivate(self):
        if self.active:
            self.ctx.set_output_mirrors(0, 0, [])
            self.ctx._active_peer_buffers = None
        self.active = False

    def slot_ptr(self, slot: int) -> int:
        return self.base + (slot % self.slots) * self.slot_bytes

    def barrier(self):
        """Stream-ordered: after it, this rank's copy holds every rank's pixels of the step."""
        self.epoch += 1
        self.ctx.peer_barrier(self.rank, self.world, self.flag_ptrs, self.epoch)

    def tensor(self, slot: int, shape, dtype="float16"):
        """torch view of this rank's copy of `slot` (cached: the view is created once per slot and shape)."""
        import torch
        key = (slot % self.slots, tuple(shape), dtype)
        t = self._views.get(key)
        if t is None:
            typestr = {"float16": "<f2", "uint8": "|u1"}[dtype]
            t = self._views[key] = torch.as_tensor(_DevicePtr(self.slot_ptr(slot), shape, typestr), device="cuda")
        return t

    def close(self):
        import torch.distributed as dist
        if self.base is None:
            return
        self._views = {}
        if self.active:
            self.deactivate()
        self.ctx.sync()
        if self.world > 1:
            dist.barrier(group=self.group)  # nobody unmaps or frees while a peer may still store into it
        for k in range(self.world):
            if k != self.rank:
                self.ctx.peer_close(self.bases[k])
                self.ctx.peer_close(self.flag_ptrs[k])
        if self.world > 1:
            dist.barrier(group=self.group)
        self.ctx.peer_free(self.base)
        self.ctx.peer_free(self.flags)
        self.base = None


class ShardedRenderer:
    """Renders with `ctx` (any implementation of the C-ABI) into torch tensors on `device` and gathers."""

    def __init__(self, ctx, width: int, height: int, device="cuda", group=None, gather: str = "nccl"):
        import torch
        import torch.distributed as dist
        if gather not in ("nccl", "peer"):
            raise ValueError("gather must be 'nccl' or 'peer'")
        if gather == "peer" and device == "cpu":
            raise ValueError("the peer gather needs CUDA devices")
        self.torch, self.dist = torch, dist
        self.gather = gather
        self._peer = {}   # slot_bytes -> PeerBuffers (allocated on first use, reused across calls)
        self._step = 0
        self.ctx, self.W, self.H, self.device, self.group = ctx, width, height, device, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if ctx.width != width or ctx.height != height:
            ctx.resize(width, height)
        if device != "cpu":
            # kernels, the output allocation and the NCCL all-gather must be ordered on ONE stream: the context's own stream is
            # cudaStreamNonBlocking, so it would race torch's allocator/fill work and NCCL otherwise
            self.use_torch_stream()

    def render_frame_rows(self, params, bands_per_rank=1):
        """One frame split into row bands; returns the gathered [H, W, 4] fp16 tensor (identical on all ranks).
        bands_per_rank > 1 interleaves the bands over the ranks (interleaved_bands) to even out the lit fraction, "max" takes the
        finest interleave (bands of 8 rows, one CTA row each); the result is the same texture.  A rank's bands are ONE launch
        (cs_render_row_bands_to) whenever the band height is a multiple of 8.  With the NCCL gather every band group costs one
        all-gather, so keep bands_per_rank small there; the peer gather has no such cost."""
        torch = self.torch
        if bands_per_rank == "max":
            bands_per_rank = max(1, self.H // (8 * self.world))
        mine = interleaved_bands(self.H, self.world, bands_per_rank)[self.rank]
        rows = mine[0][1] - mine[0][0]

        def render_into(ptr):
            if bands_per_rank > 1 and rows % 8 == 0:
                self.ctx.render_row_bands_to(params, mine[0][0], rows, rows * self.world, bands_per_rank, ptr)
            else:
                for r0, r1 in mine:
                    self.ctx.render_rows_to(params, r0, r1, ptr)

        if self.gather == "peer":
            pb, slot = self._peer_slot(self.H * self.W * 8)
            render_into(pb.slot_ptr(slot))  # the kernel stores into every rank's copy
            pb.barrier()
            return pb.tensor(slot, (self.H, self.W, 4))
        full = torch.empty((self.H, self.W, 4), dtype=torch.float16, device=self.device)  # every row is written by a render or the gather
        render_into(full.data_ptr())
        if self.world > 1:
            flat = full.view(-1)
            per = rows * self.W * 4          # elements of one band
            for g, (r0, _) in enumerate(mine):
                group_flat = flat[g * self.world * per:(g + 1) * self.world * per]
                _all_gather_inplace(group_flat, group_flat[self.rank * per:(self.rank + 1) * per], self.group)
        return full

    def render_sun_sweep(self, params, suns: Sequence[Sequence[float]]):
        """n sun angles sharded n/N per rank; returns the gathered [n, H, W, 4] fp16 tensor."""
        torch = self.torch
        suns = np.asarray(suns, np.float32).reshape(-1, 3)
        n = suns.shape[0]
        k0, k1 = sun_shard(n, self.world, self.rank)
        if self.gather == "peer":
            pb, slot = self._peer_slot(n * self.H * self.W * 8)
            self.ctx.render_sun_batch_to(params, suns[k0:k1], pb.slot_ptr(slot) + k0 * self.H * self.W * 8)
            pb.barrier()
            return pb.tensor(slot, (n, self.H, self.W, 4))
        full = torch.empty((n, self.H, self.W, 4), dtype=torch.float16, device=self.device)
        self.ctx.render_sun_batch_to(params, suns[k0:k1], full[k0].data_ptr())
        if self.world > 1:
            flat = full.view(-1)
            per = (k1 - k0) * self.H * self.W * 4
            _all_gather_inplace(flat, flat[self.rank * per:(self.rank + 1) * per], self.group)
        return full

    def _peer_slot(self, slot_bytes: int):
        """The PeerBuffers for results of this size (created collectively on first use) and the slot of this call.  The returned
        tensors alias library memory: a result stays valid until the call after the next one of the same size."""
        pb = self._peer.get(slot_bytes)
        if pb is None:
            pb = self._peer[slot_bytes] = PeerBuffers(self.ctx, slot_bytes, 2, self.group)
        if not pb.active:
            pb.activate()
        self._step += 1
        return pb, self._step & 1

    def close(self):
        for pb in self._peer.values():
            pb.close()
        self._peer = {}

    def use_torch_stream(self):
        """Make the context launch on torch's current stream so kernels and the NCCL all-gather are stream-ordered."""
        s = self.torch.cuda.current_stream()
        if s.cuda_stream == 0:
            s = self.torch.cuda.Stream()
            self.torch.cuda.set_stream(s)
        self.ctx.set_stream(s.cuda_stream)
        self.ctx._shares_torch_stream = True
        return s
