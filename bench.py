#!/usr/bin/env python
"""bench.py — the measurement contract for the cloud-march hot path.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's algorithm on host cores

Workload (BASELINE.json configs[2], the one `metric` is quoted on): 2048x1024 hemisphere,
128 primary x 8 light steps (7 cone + 1 distant density evaluations per lit step),
animated wind: step k renders frame k of the wind animation at absolute time t_k = 1 + k*64/60 s,
noon sun, demo settings (coverage 0.2, density 0.05).  One "step" = _update_per_frame_data +
_render_process of the reference (cloud_sky.gd:165-187,234-248): advance the wind offsets on the
host, rebuild the sky LUT, run the march prologue and the march kernel for the whole image.

Metric: Mray-steps/s = marched pixels (dir.y > 0) x nominal primary steps / seconds / 1e6.

N > 1 (launched under torchrun, one rank per GPU): weak scaling over sun-angle batches
(BASELINE configs[3]): every rank renders one full frame for its own sun angle straight into its
slice of the gathered [N, H, W, 4] fp16 tensor, followed by ONE NCCL all-gather of the finished
textures over NVLink; value = N frames' ray-steps / max-over-ranks device time.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W, H, PRIMARY, LIGHT = 2048, 1024, 128, 8
CONE = LIGHT - 1
METRIC = "Mray-steps/sec, 2048x1024 hemisphere @128x8 steps; achieved HBM GB/s vs peak"
UNIT = "Mray-steps/s"
WORKLOAD = "C3: 2048x1024 hemisphere, 128 primary / 8 light (7 cone + 1 distant) steps, animated wind, noon sun"
L2_FLUSH_BYTES = 256 << 20


def frame_time(k):
    return 1.0 + k * (64.0 / 60.0)


def sun_for_rank(rank, world):
    if world == 1:
        return (0.0, 1.0, 0.0)
    th = math.pi * (rank + 0.5) / world  # dir_k = (cos th_k, sin th_k, 0)
    return (math.cos(th), math.sin(th), 0.0)


def frame_params(lib, k, sun):
    """Settings -> FrameData -> push constants for animation frame k (host logic of cloud_sky.gd)."""
    s = lib.settings_demo()
    st = lib.frame_state_init()
    st.light_direction[:] = list(sun)
    for j in range(k + 1):  # integrate the wind offsets frame by frame like the running demo does
        lib.frame_advance(st, s, frame_time(j))
    return lib.fill_cloud_params(s, st, W, H)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm = sorted(int(r[1]) for r in self.rows if len(r) > 2 and r[1].isdigit())
        mx = max([int(r[2]) for r in self.rows if len(r) > 2 and r[2].isdigit()], default=None)
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        # samples taken while the SMs were busy have the highest clocks; report the median of the upper half
        busy = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def oracle_library(cs):
    """The CPU oracle (test infrastructure): only used for the cpu_baseline leg and --impl reference."""
    path = os.path.join(ROOT, "oracle", "libcloudsky_oracle.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
    return cs.Library(path)


class OracleSampler:
    """The CPU oracle timed on evenly spaced row bands of a frame (the whole frame when it fits the time budget).
    One context (textures + transmittance LUT) is reused across calls."""

    def __init__(self, cs, textures, threads):
        import numpy as np
        self.np = np
        self.threads = threads
        self.ctx = oracle_library(cs).context(0)
        self.ctx.set_threads(threads)
        self.ctx.upload_textures(*textures)
        self.ctx.build_transmittance_lut()
        self.ctx.resize(W, H)
        self.ctx.set_march_config(PRIMARY, CONE)
        self.buf = np.zeros((H, W, 4), np.float16)
        self.band = min(H, max(threads, 8))  # rows per call: the oracle parallelises over the rows of one call
        self.t_band = None

    def sample(self, params, sun, target_seconds):
        """Returns (Mray-steps/s, sample description, seconds)."""
        ctx, band = self.ctx, self.band
        ctx.build_sky_lut(sun)
        if self.t_band is None:  # calibration band (second of two runs: the first one warms caches and thread pool), not part of any sample
            for _ in range(2):
                t0 = time.perf_counter()
                ctx.render_rows_to(params, H // 3, H // 3 + band, self.buf.ctypes.data)
                self.t_band = time.perf_counter() - t0
        max_bands = H // band
        n_bands = max(1, min(max_bands, int(target_seconds / max(self.t_band, 1e-6))))
        steps = 0
        t0 = time.perf_counter()
        if n_bands == max_bands:
            ctx.render_rows_to(params, 0, H, self.buf.ctypes.data)
            steps = ctx.get_counters().primary_steps
            desc = "the whole 2048x1024 frame"
        else:
            pitch = H / n_bands
            for i in range(n_bands):
                r0 = min(H - band, int(i * pitch + (pitch - band) / 2))
                ctx.render_rows_to(params, r0, r0 + band, self.buf.ctypes.data)
                steps += ctx.get_counters().primary_steps
            desc = f"{n_bands} evenly spaced bands of {band} rows ({n_bands * band} of {H} rows) of the same 2048x1024 frame"
        sec = time.perf_counter() - t0
        desc += f", {PRIMARY}/{LIGHT} steps, scalar fp32 C++ oracle (-O2 -ffp-contract=off), {self.threads} threads"
        return steps / sec / 1e6, desc, sec


def cpu_rows_sample(cs, textures, params, sun, target_seconds, threads):
    return OracleSampler(cs, textures, threads).sample(params, sun, target_seconds)


def run_reference(args):
    """--impl reference: the reference's own algorithm (the oracle port: nothing in this container or on
    the GPU box can execute Godot/GLSL) on all host cores, same config/metric, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import cloudsky_b200 as cs
    from cloudsky_b200 import assets
    large, small, weather, desc = assets.load_default_textures()
    threads = os.cpu_count() or 1
    ora = oracle_library(cs)
    sampler = OracleSampler(cs, (large, small, weather), threads)
    # bounded sample per step: the whole run (warm-up + steps) stays near two minutes whatever K the driver asks for
    per_step = min(4.0, max(0.25, 110.0 / max(1, args.steps + args.warmup)))
    vals, secs = [], []
    sample = ""
    for k in range(args.warmup + args.steps):
        p = frame_params(ora, k % 16, (0.0, 1.0, 0.0))
        v, sample, sec = sampler.sample(p, (0.0, 1.0, 0.0), per_step)
        if k >= args.warmup:
            vals.append(v); secs.append(sec)
    value = sum(vals) / len(vals)
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(1e3 * sum(secs) / len(secs), 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": desc,
            "config": {"workload": WORKLOAD, "width": W, "height": H, "primary_steps": PRIMARY, "light_steps": LIGHT},
            "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy bandwidth)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(name="roofline_latest.json"):
    """dram bytes per launch of the march kernel from the committed `ncu --set full` capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            d = json.load(f)
        return d.get("dram_bytes_per_launch"), d.get("source"), {k: d.get(k) for k in ("l2_hit_pct", "l1_hit_pct", "issue_active_pct")}
    except Exception:
        return None, None, {}


def run_ours(args):
    import numpy as np
    import torch
    import cloudsky_b200 as cs
    from cloudsky_b200 import assets

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    lib = cs.load_product()
    large, small, weather, desc = assets.load_default_textures()
    ctx = lib.context(local)
    stream = torch.cuda.Stream()  # an explicit stream: torch's default stream is handle 0, which cs_set_stream reads as "own stream"
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.upload_textures(large, small, weather)
    ctx.build_transmittance_lut()
    ctx.resize(W, H)
    base_mode = cs.MODE_FAST | (cs.MODE_TEX if args.sampler == "texture" else 0)
    ctx.set_march_config(PRIMARY, CONE, base_mode)
    sun = sun_for_rank(rank, world)
    # two gathered buffers: the all-gather of step k (on a side stream) overlaps the kernels of step k+1
    gathered = [torch.zeros((world, H, W, 4), dtype=torch.float16, device="cuda") for _ in range(2 if world > 1 else 1)]
    gather_stream = torch.cuda.Stream() if world > 1 else None
    rendered = [torch.cuda.Event() for _ in gathered]
    gather_done = [torch.cuda.Event() for _ in gathered]
    gather_used = [False for _ in gathered]
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")
    params = [frame_params(lib, k % 16, sun) for k in range(16)]

    # work counters of frame 0 (deterministic; measured outside the timed region with the instrumented kernel)
    ctx.build_sky_lut(sun)
    ctx.set_counters_enabled(True)
    ctx.render_frame(params[0])
    counters = ctx.get_counters().as_dict()
    ctx.set_counters_enabled(False)

    def step(k):
        p = params[k % 16]
        b = k % len(gathered)
        buf = gathered[b]
        if gather_used[b]:
            stream.wait_event(gather_done[b])         # this buffer's previous all-gather (step k-2) has finished
        ctx.build_sky_lut(sun)                        # sky_lut.update_lut (cloud_sky.gd:187)
        ctx.render_rows_to(p, 0, H, buf[rank].data_ptr())  # prologue + march into this rank's slice
        if dist is not None:
            rendered[b].record(stream)
            gather_stream.wait_event(rendered[b])
            with torch.cuda.stream(gather_stream):   # ONE NCCL all-gather of the finished textures, overlapping the next step
                dist.all_gather_into_tensor(buf.view(-1), buf[rank].reshape(-1))
                gather_done[b].record(gather_stream)
            gather_used[b] = True

    def drain():
        if gather_stream is not None:
            stream.wait_stream(gather_stream)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for k in range(args.warmup):
        step(k)
    drain()
    torch.cuda.synchronize()
    if rank == 0:
        time.sleep(0.3)   # let nvidia-smi come up, then drop what it saw during warm-up / idle
        sampler.rows.clear()
    ctx.set_kernel_timing(True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()  # evict L2 between timed steps (outside the per-step event pair)
        ev[k][0].record(stream)
        step(args.warmup + k)
        if k == args.steps - 1:
            drain()  # the last step's all-gather is inside the timed region
        ev[k][1].record(stream)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    kt = ctx.read_kernel_timings()
    ctx.set_kernel_timing(False)
    per_rank_kernel_ms = [round((kt["march_ms"] + kt["sky_ms"]) / max(1, kt["march_launches"]), 4)]
    if dist is not None:
        t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
        mine_ms = torch.tensor([per_rank_kernel_ms[0]], dtype=torch.float64, device="cuda")
        all_ms = torch.zeros(world, dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(all_ms, mine_ms)
        per_rank_kernel_ms = [round(float(v), 4) for v in all_ms.tolist()]

    ray_steps_per_frame = counters["marched_pixels"] * PRIMARY
    value = world * ray_steps_per_frame * args.steps / (dev_ms * 1e-3) / 1e6
    ms_per_step = dev_ms / args.steps

    # end-to-end through the C-ABI with HOST buffers: params from host memory, result into pinned host memory
    e2e = None
    if True:
        host = torch.empty((H, W, 4), dtype=torch.float16).pin_memory()
        for k in range(min(3, args.warmup)):
            ctx.render_frame_host(params[k % 16], out_ptr=host.data_ptr())
        if dist is not None:
            dist.barrier()
        hosts = [host, torch.empty((H, W, 4), dtype=torch.float16).pin_memory()]
        for k in range(max(3, min(args.warmup, 10))):  # warm-up of the streaming path (second image, copy stream, both host buffers)
            ctx.render_frame_host_async(params[k % 16], hosts[k & 1].data_ptr())
        ctx.wait_host()
        t0 = time.perf_counter()
        for k in range(args.steps):  # streaming: frame k+1's kernels overlap frame k's device->host copy
            ctx.render_frame_host_async(params[(args.warmup + k) % 16], hosts[k & 1].data_ptr())
        ctx.wait_host()  # every result is in host memory when the clock stops
        e2e_s = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        assert torch.isfinite(host.float()).all()
        e2e = {"value": round(world * ray_steps_per_frame * args.steps / e2e_s / 1e6, 1), "unit": UNIT,
               "h2d_bytes_per_step": 112 + 12, "d2h_bytes_per_step": W * H * 8,
               "note": "cs_render_frame_host_async + cs_wait_host: push constants from host, sky LUT + prologue + march per step, every 16 MiB RGBA16F result copied to pinned host memory (copy of frame k overlaps the kernels of frame k+1); wall clock around all steps incl. the final wait"}

    # Extra, NOT the headline: the opt-in CS_MODE_EARLY_OUT flag (rays stop once T < 2^-12; results within 2 fp16 ulps,
    # SURVEY 7.3-5 asks for nominal AND executed steps to be reported).  Measured after the timed region.
    early = None
    if world == 1:
        ctx.set_march_config(PRIMARY, CONE, base_mode | cs.MODE_EARLY_OUT)
        ctx.set_counters_enabled(True)
        ctx.render_frame(params[0])
        k_early = ctx.get_counters().as_dict()
        ctx.set_counters_enabled(False)
        ms_early = ctx.time_render_frame(params[0], 3, 10)
        early = {"march_ms": round(ms_early, 4), "value_on_nominal_steps": round(ray_steps_per_frame / ms_early / 1e3, 1),
                 "value_on_executed_steps": round(k_early["primary_steps"] / ms_early / 1e3, 1),
                 "executed_step_fraction": round(k_early["primary_steps"] / (counters["marched_pixels"] * PRIMARY), 4),
                 "note": "opt-in mode flag, not reference behaviour (clouds.glsl:172 runs every step); not used for value / e2e"}
        ctx.set_march_config(PRIMARY, CONE, base_mode)

    # Extra, NOT the headline: the opt-in CS_MODE_TEX flag (the texture unit filters the three input textures with its 8-bit
    # fixed-point weights, like the reference's own sampler bindings; the headline keeps the in-kernel fp32 filter).
    tex_extra = None
    if world == 1 and args.sampler == "kernel":
        ctx.set_march_config(PRIMARY, CONE, cs.MODE_FAST | cs.MODE_TEX)
        ms_tex = ctx.time_render_frame(params[0], 3, 10)
        tex_extra = {"march_ms": round(ms_tex, 4), "value_march_only": round(ray_steps_per_frame / ms_tex / 1e3, 1),
                     "note": "opt-in mode flag (bench.py --sampler texture runs the whole bench in it); same parity tolerance as FAST, tests/test_gpu_parity.py"}
        ctx.set_march_config(PRIMARY, CONE, base_mode)

    # Extra, NOT the headline: BASELINE config 4's shape on this GPU — 8 sun angles of the same cloud field through
    # cs_render_sun_batch_to, whose kernel marches 4 suns per launch and shares the sun-independent primary march between them
    # (every image bit-identical to a single-sun dispatch, tests/test_gpu_parity.py).  Includes the 8 sky-LUT builds.
    sun_batch = None
    if world == 1 and args.sampler == "kernel":
        n_b = 8
        th = np.pi * (np.arange(n_b) + 0.5) / n_b
        suns_b = np.stack([np.cos(th), np.sin(th), np.zeros(n_b)], 1).astype(np.float32)  # dir_k = (cos, sin, 0)
        out_b = torch.empty((n_b, H, W, 4), dtype=torch.float16, device="cuda")
        best = None
        for it in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); ctx.render_sun_batch_to(params[0], suns_b, out_b.data_ptr()); e1.record(stream)
            torch.cuda.synchronize()
            if it:
                best = e0.elapsed_time(e1) if best is None else min(best, e0.elapsed_time(e1))
        sun_batch = {"suns": n_b, "ms_per_frame": round(best / n_b, 4), "value": round(ray_steps_per_frame * n_b / best / 1e3, 1), "unit": UNIT,
                     "note": "opt-in call, not the headline workload: 8 sun angles of one cloud field, 4 suns per launch sharing the primary march"}
        del out_b

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peak()
    alg_bytes = 80 * counters["density_evals"] + 8 * W * H  # 80 B per density evaluation + 8 B per pixel
    march_ms = kt["march_ms"] / max(1, kt["march_launches"])
    achieved = alg_bytes / (march_ms * 1e-3) / 1e9
    traffic, traffic_src, ncu_extra = ncu_traffic("roofline_tex.json" if args.sampler == "texture" else "roofline_latest.json")
    roofline = {"bound": "hbm", "kernel": "clouds_fast_kernel", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "kernel_ms": round(march_ms, 4), "algorithmic_bytes_per_launch": alg_bytes, "ncu": ncu_extra,
                "note": "algorithmic bytes = 80 B x executed density evaluations + 8 B x pixels (SURVEY 8(d)); the texel working set is L1/L2-resident, "
                        "so DRAM traffic is far below this figure and frac can exceed 1"}
    cpu_threads = os.cpu_count() or 1
    cpu_v, cpu_sample, _ = cpu_rows_sample(cs, (large, small, weather), params[0], sun, 12.0, cpu_threads) if world == 1 else (None, None, None)
    line = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": desc,
            "config": {"workload": WORKLOAD, "width": W, "height": H, "primary_steps": PRIMARY, "light_steps": LIGHT, "cone_samples": CONE,
                       "sampler": "in-kernel fp32 trilinear (coefficient records)" if args.sampler == "kernel" else "texture unit (CS_MODE_TEX, 8-bit filter weights)",
                       "l2": f"flushed between timed steps ({L2_FLUSH_BYTES >> 20} MiB memset outside the per-step event pairs)",
                       "parallelism": "1 GPU" if world == 1 else f"{world} GPUs: one sun-angle frame per rank + one NCCL all-gather of the finished textures per step (side stream, overlaps the next step)",
                       "lit_fraction": round(counters["lit_steps"] / counters["primary_steps"], 4),
                       "density_evals_per_frame": counters["density_evals"], "marched_pixels": counters["marched_pixels"]},
            "roofline": roofline, "e2e": e2e, "gpu_launches": 3 * args.steps,
            "kernels": {"march_ms_avg": round(march_ms, 4), "sky_lut_ms_avg": round(kt["sky_ms"] / max(1, kt["sky_launches"]), 4)},
            "gevals_per_s": round(world * counters["density_evals"] * args.steps / (dev_ms * 1e-3) / 1e9, 2),
            "wall_ms_per_step_incl_flush": round(1e3 * t_wall / args.steps, 3), "clocks": clocks,
            "early_out_mode": early, "texture_unit_mode": tex_extra, "sun_batch_mode": sun_batch,
            "per_rank_kernel_ms": per_rank_kernel_ms,  # sky LUT + march per step on every rank (load balance)
            "value_without_gather": round(world * ray_steps_per_frame / (max(per_rank_kernel_ms) * 1e-3) / 1e6, 1)}
    if cpu_v is not None:
        line["cpu_baseline"] = {"value": round(cpu_v, 3), "unit": UNIT, "cores": cpu_threads, "kind": "port", "sample": cpu_sample}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sampler", default="kernel", choices=["kernel", "texture"],
                    help="kernel: in-kernel fp32 trilinear filter (default, the headline); texture: CS_MODE_TEX, the GPU texture unit filters")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
