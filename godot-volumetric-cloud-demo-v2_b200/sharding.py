"""Multi-GPU sharding of the cloud march.

Every pixel depends only on the push constants, the three textures, the sky LUT and its own coordinates
(clouds.glsl:258-266); the reference already renders the texture as independent tiles addressed by
update_position (cloud_sky.gd:156-161).  So ranks need no data-path exchange while rendering; the only
collective is ONE all-gather of the finished RGBA16F texture(s):

  * one frame   : contiguous row bands, H/N rows per rank, rendered straight into the rank's slice of the
                  full-size buffer, then an in-place all-gather (each band is a contiguous chunk).
  * sun sweep   : n/N whole frames per rank (each rank builds its own sky LUTs), gathered [n, H, W, 4].

One process per GPU, torch.distributed for the plumbing (NCCL on GPUs; gloo in the CPU tests, where the
"device" buffers are host tensors and the compute backend is the oracle library)."""
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


class ShardedRenderer:
    """Renders with `ctx` (any implementation of the C-ABI) into torch tensors on `device` and gathers."""

    def __init__(self, ctx, width: int, height: int, device="cuda", group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.ctx, self.W, self.H, self.device, self.group = ctx, width, height, device, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if ctx.width != width or ctx.height != height:
            ctx.resize(width, height)
        if device != "cpu":
            # kernels, the output allocation and the NCCL all-gather must be ordered on ONE stream: the context's own stream is
            # cudaStreamNonBlocking, so it would race torch's allocator/fill work and NCCL otherwise
            self.use_torch_stream()

    def render_frame_rows(self, params, bands_per_rank: int = 1):
        """One frame split into row bands; returns the gathered [H, W, 4] fp16 tensor (identical on all ranks).
        bands_per_rank > 1 interleaves the bands over the ranks (interleaved_bands) to even out the lit fraction; the result is
        the same texture, gathered with one all-gather per band group."""
        torch = self.torch
        full = torch.empty((self.H, self.W, 4), dtype=torch.float16, device=self.device)  # every row is written by a render or the gather
        mine = interleaved_bands(self.H, self.world, bands_per_rank)[self.rank]
        for r0, r1 in mine:
            self.ctx.render_rows_to(params, r0, r1, full.data_ptr())
        if self.world > 1:
            flat = full.view(-1)
            per = (mine[0][1] - mine[0][0]) * self.W * 4          # elements of one band
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
        full = torch.empty((n, self.H, self.W, 4), dtype=torch.float16, device=self.device)
        self.ctx.render_sun_batch_to(params, suns[k0:k1], full[k0].data_ptr())
        if self.world > 1:
            flat = full.view(-1)
            per = (k1 - k0) * self.H * self.W * 4
            _all_gather_inplace(flat, flat[self.rank * per:(self.rank + 1) * per], self.group)
        return full

    def use_torch_stream(self):
        """Make the context launch on torch's current stream so kernels and the NCCL all-gather are stream-ordered."""
        s = self.torch.cuda.current_stream()
        if s.cuda_stream == 0:
            s = self.torch.cuda.Stream()
            self.torch.cuda.set_stream(s)
        self.ctx.set_stream(s.cuda_stream)
        self.ctx._shares_torch_stream = True
        return s
