"""world_size-2 gloo test of the N>1 path on CPU: row-band and sun-angle sharding + the single all-gather
reassemble exactly what one rank renders alone.  The compute backend here is the oracle library (CPU), the
plumbing (sharding.py + torch.distributed) is the same code the GPU ranks run with NCCL."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_helpers(cs):
    from cloudsky_b200 import sharding
    assert sharding.row_bands(1024, 8) == [(128 * r, 128 * (r + 1)) for r in range(8)]
    assert sharding.row_bands(64, 1) == [(0, 64)]
    with pytest.raises(ValueError):
        sharding.row_bands(100, 8)
    assert sharding.interleaved_bands(64, 2, 1) == [[(0, 32)], [(32, 64)]]
    assert sharding.interleaved_bands(64, 2, 2) == [[(0, 16), (32, 48)], [(16, 32), (48, 64)]]
    bands = sharding.interleaved_bands(1024, 8, 4)
    assert sorted(b for r in bands for b in r) == [(32 * i, 32 * (i + 1)) for i in range(32)]   # a partition of the rows
    with pytest.raises(ValueError):
        sharding.interleaved_bands(100, 2, 3)
    s = sharding.sun_sweep(64)
    assert s.shape == (64, 3) and np.allclose(np.linalg.norm(s, axis=1), 1.0, atol=1e-6)
    assert s[0, 0] > 0.99 and s[-1, 0] < -0.99 and (s[:, 1] > 0).all() and (s[:, 2] == 0).all()
    assert [sharding.sun_shard(64, 8, r) for r in (0, 7)] == [(0, 8), (56, 64)]
    with pytest.raises(ValueError):
        sharding.sun_shard(10, 4, 0)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import cloudsky_b200 as cs
    from cloudsky_b200 import assets, sharding
    from conftest import make_params, prepared_context
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    torch.set_num_threads(1)
    lib = cs.Library(os.path.join(ROOT, "oracle", "libcloudsky_oracle.so"))
    tex = assets.synthetic_textures(seed=5, large_n=16, small_n=8, weather_n=32)
    W, H = 32, 16
    ctx = prepared_context(lib, tex, W, H, threads=2)
    ctx.set_march_config(32, 3)
    p = make_params(lib, W, H, time=3.0, coverage=0.6)
    r = sharding.ShardedRenderer(ctx, W, H, device="cpu")
    assert (r.world, r.rank) == (world, rank)
    frame = r.render_frame_rows(p).numpy()
    inter = r.render_frame_rows(p, bands_per_rank=4).numpy()  # interleaved bands, one all-gather per band group
    assert (inter.view(np.uint16) == frame.view(np.uint16)).all()
    sweep = r.render_sun_sweep(p, sharding.sun_sweep(4)).numpy()
    # bands of 8 rows: the rank's interleaved bands are ONE cs_render_row_bands_to call (the GPU path's single launch)
    W2, H2 = 24, 64
    ctx.resize(W2, H2)
    ctx.build_sky_lut((0.0, 1.0, 0.0))  # the sweep left every rank with the LUT of its own last sun
    r2 = sharding.ShardedRenderer(ctx, W2, H2, device="cpu")
    p2 = make_params(lib, W2, H2, time=3.0, coverage=0.6)
    tall = r2.render_frame_rows(p2, bands_per_rank=1).numpy()
    for bpr in (2, 4, "max"):
        again = r2.render_frame_rows(p2, bands_per_rank=bpr).numpy()
        assert (again.view(np.uint16) == tall.view(np.uint16)).all(), bpr
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), frame=frame, sweep=sweep, tall=tall)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_rank(cs, oracle_lib, helpers, tmp_path):
    import torch.multiprocessing as mp
    from cloudsky_b200 import assets, sharding
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    # single-rank reference of the same work
    tex = assets.synthetic_textures(seed=5, large_n=16, small_n=8, weather_n=32)
    W, H = 32, 16
    ctx = helpers.prepared_context(oracle_lib, tex, W, H, threads=2)
    ctx.set_march_config(32, 3)
    p = helpers.make_params(oracle_lib, W, H, time=3.0, coverage=0.6)
    ctx.render_frame(p)
    single = ctx.read_image()
    suns = sharding.sun_sweep(4)
    singles = []
    for k in range(4):
        q = p.copy(); q.light_direction[:] = suns[k].tolist()
        ctx.build_sky_lut(tuple(q.light_direction)); ctx.render_frame(q)
        singles.append(ctx.read_image().copy())
    for rank in range(2):
        d = np.load(tmp_path / f"rank{rank}.npz")
        assert (d["frame"].view(np.uint16) == single.view(np.uint16)).all()          # bit-identical to 1 rank
        for k in range(4):
            assert (d["sweep"][k].view(np.uint16) == singles[k].view(np.uint16)).all()
    assert single.astype(np.float32)[..., 3].max() > 0
    # the 24x64 frame rendered with interleaved 8-row bands (one cs_render_row_bands_to call per rank) == one dispatch
    ctx.resize(24, 64)
    ctx.build_sky_lut((0.0, 1.0, 0.0))
    p2 = helpers.make_params(oracle_lib, 24, 64, time=3.0, coverage=0.6)
    ctx.render_frame(p2)
    for rank in range(2):
        assert (np.load(tmp_path / f"rank{rank}.npz")["tall"].view(np.uint16) == ctx.read_image().view(np.uint16)).all()
