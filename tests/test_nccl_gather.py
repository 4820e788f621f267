"""Two ranks over NCCL on two GPUs (skipped with fewer): row-band sharding + one all-gather reproduces the
single-GPU texture bit for bit, and so does the sun-angle sweep."""
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
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), frame=frame.cpu().numpy(), sweep=sweep.cpu().numpy())
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
        for k in range(4):
            q = p.copy(); q.light_direction[:] = suns[k].tolist()
            assert (d["sweep"][k].view(np.uint16) == ctx.render_frame_host(q).view(np.uint16)).all()
    ctx.close()
