"""Two ranks on two GPUs (skipped with fewer): row-band sharding + one all-gather reproduces the single-GPU texture bit for
bit, and so does the sun-angle sweep — with the NCCL all-gather and with the fused peer gather (the march
kernel stores into every rank's copy over NVLink, cs_peer_barrier completes it)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import cloudsky_b200 as cs
    from cloudsky_b200 import assets, sharding
    from conftest import make_params, prepared_context
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    lib = cs.load_product()
    tex = assets.load_fixture()
    W, H = 512, 256
    ctx = prepared_context(lib, tex, W, H, device=rank)
    ctx.set_march_config(128, 6, cs.MODE_FAST)
    p = make_params(lib, W, H, time=3.0)
    r = sharding.ShardedRenderer(ctx, W, H, device="cuda")
    r.use_torch_stream()
    frame = r.render_frame_rows(p)
    sweep = r.render_sun_sweep(p, sharding.sun_sweep(4))
    torch.cuda.synchronize()
    # fused peer gather: same results, several back-to-back calls (slot rotation), interleaved bands, then the sweep
    rp = sharding.ShardedRenderer(ctx, W, H, device="cuda", gather="peer")
    peer_frames = []
    for k in range(5):
        pk = make_params(lib, W, H, time=3.0 + k)
        peer_frames.append(rp.render_frame_rows(pk, bands_per_rank=1 + (k % 2) * 3).clone())
    peer_sweep = rp.render_sun_sweep(p, sharding.sun_sweep(4)).clone()
    torch.cuda.synchronize()
    ctx.peer_check()
    rp.close()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), frame=frame.cpu().numpy(), sweep=sweep.cpu().numpy(),
             peer_frames=torch.stack(peer_frames).cpu().numpy(), peer_sweep=peer_sweep.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_nccl_gather_is_bit_identical(cs, product_lib, textures, helpers, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from cloudsky_b200 import sharding
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    W, H = 512, 256
    ctx = helpers.prepared_context(product_lib, textures, W, H)
    ctx.set_march_config(128, 6, cs.MODE_FAST)
    p = helpers.make_params(product_lib, W, H, time=3.0)
    single = ctx.render_frame_host(p)
    suns = sharding.sun_sweep(4)
    for rank in range(2):
        d = np.load(tmp_path / f"rank{rank}.npz")
        assert (d["frame"].view(np.uint16) == single.view(np.uint16)).all()
        assert (d["peer_frames"][0].view(np.uint16) == single.view(np.uint16)).all()
        for k in range(1, 5):
            pk = helpers.make_params(product_lib, W, H, time=3.0 + k)
            assert (d["peer_frames"][k].view(np.uint16) == ctx.render_frame_host(pk).view(np.uint16)).all(), k
        assert (d["peer_sweep"].view(np.uint16) == d["sweep"].view(np.uint16)).all()
        for k in range(4):
            q = p.copy(); q.light_direction[:] = suns[k].tolist()
            assert (d["sweep"][k].view(np.uint16) == ctx.render_frame_host(q).view(np.uint16)).all()
    ctx.close()


def test_output_mirrors_and_barrier_on_one_gpu(cs, product_lib, textures, helpers):
    """The fused-gather mechanism without peers: a second buffer on the same device stands in for a peer's copy.  Every
    march kernel (fast, texture-unit, strict, sun batch) must store the same pixels at the same offset into the mirror;
    the barrier kernel must complete for world = 1 and trip nothing."""
    import torch
    W, H = 200, 96
    ctx = helpers.prepared_context(product_lib, textures, W, H)
    frame_bytes = W * H * 8
    own, _ = ctx.peer_alloc(4 * frame_bytes)
    mirror, _ = ctx.peer_alloc(4 * frame_bytes)
    flags, _ = ctx.peer_alloc(256)
    from cloudsky_b200.sharding import _DevicePtr
    view = lambda ptr, n: torch.as_tensor(_DevicePtr(ptr, (n, H, W, 4), "<f2"), device="cuda")
    ctx.set_output_mirrors(own, 4 * frame_bytes, [mirror])
    p = helpers.make_params(product_lib, W, H, time=2.0, sun=(0.4, 0.7, 0.1))
    ctx.build_sky_lut(tuple(p.light_direction))
    for i, mode in enumerate((cs.MODE_FAST, cs.MODE_FAST | cs.MODE_TEX, cs.MODE_STRICT, cs.MODE_FAST | cs.MODE_HALF)):
        ctx.set_march_config(64, 6, mode)
        ctx.render_rows_to(p, 0, H // 2, own + i * frame_bytes)       # two bands -> offsets inside a frame
        ctx.render_rows_to(p, H // 2, H, own + i * frame_bytes)
    ctx.peer_barrier(0, 1, [flags], 1)
    ctx.sync()
    a, b = view(own, 4).cpu().numpy(), view(mirror, 4).cpu().numpy()
    assert (a.view(np.uint16) == b.view(np.uint16)).all() and a.astype(np.float32).max() > 0.1
    ctx.set_march_config(64, 6, cs.MODE_FAST)
    ctx.render_frame(p)                                               # a dispatch outside the registered range is not mirrored
    assert (ctx.read_image().view(np.uint16) == a[0].view(np.uint16)).all()
    suns = np.array([[0.0, 1.0, 0.0], [0.6, 0.8, 0.0], [-0.5, 0.5, 0.3]], np.float32)
    ctx.render_sun_batch_to(p, suns, own + frame_bytes)               # sun-batch kernel into frames 1..3
    ctx.peer_barrier(0, 1, [flags], 2)
    ctx.sync()
    ctx.peer_check()
    a, b = view(own, 4).cpu().numpy(), view(mirror, 4).cpu().numpy()
    assert (a.view(np.uint16) == b.view(np.uint16)).all()
    assert (a[1].view(np.uint16) != a[2].view(np.uint16)).any()
    ctx.set_output_mirrors(0, 0, [])
    ctx.render_rows_to(p, 0, H, own)                                  # mirrors off: only the own copy changes
    ctx.sync()
    assert (view(mirror, 4).cpu().numpy().view(np.uint16) == b.view(np.uint16)).all()
    for ptr in (own, mirror, flags):
        ctx.peer_free(ptr)
    ctx.close()


def test_peer_gather_world_one(cs, product_lib, textures, helpers):
    """ShardedRenderer(gather='peer') without a process group: one rank, no mirrors, the barrier still runs."""
    from cloudsky_b200 import sharding
    W, H = 128, 64
    ctx = helpers.prepared_context(product_lib, textures, W, H)
    ctx.set_march_config(128, 6, cs.MODE_FAST)
    p = helpers.make_params(product_lib, W, H, time=1.5)
    want = ctx.render_frame_host(p)
    r = sharding.ShardedRenderer(ctx, W, H, device="cuda", gather="peer")
    for _ in range(3):
        got = r.render_frame_rows(p, bands_per_rank=2).cpu().numpy()
        assert (got.view(np.uint16) == want.view(np.uint16)).all()
    sweep = r.render_sun_sweep(p, sharding.sun_sweep(3)).cpu().numpy()
    assert sweep.shape == (3, H, W, 4) and np.isfinite(sweep.astype(np.float32)).all()
    ctx.peer_check()
    r.close()
    ctx.close()


def test_peer_barrier_watchdog_reports_a_missing_peer(cs, product_lib, monkeypatch):
    """A rank that never arrives must not hang the GPU: the barrier kernel gives up after the watchdog time and
    cs_peer_check turns that into an error (here: a 'world' of 2 whose second flag array nobody ever writes)."""
    monkeypatch.setenv("CLOUDSKY_PEER_WATCHDOG_MS", "150")
    ctx = product_lib.context(0)
    mine, _ = ctx.peer_alloc(256)
    absent, _ = ctx.peer_alloc(256)
    ctx.peer_barrier(0, 2, [mine, absent], 1)
    with pytest.raises(cs.CloudSkyError) as e:
        ctx.peer_check()
    assert e.value.code == 5 and "did not arrive" in str(e.value)
    ctx.peer_check()  # the error is reported once
    ctx.peer_barrier(0, 1, [mine], 2)  # and the context keeps working
    ctx.sync()
    ctx.peer_check()
    ctx.peer_free(mine); ctx.peer_free(absent)
    ctx.close()
