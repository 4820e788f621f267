#!/usr/bin/env python
"""bench.py — the measurement contract for the cloud-march hot path.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's own shaders on host cores (oracle/_ref)

Workload (BASELINE.json configs[2], the one `metric` is quoted on): 2048x1024 hemisphere,
128 primary x 8 light steps (7 cone + 1 distant density evaluations per lit step),
animated wind: step k renders frame k of the wind animation at absolute time t_k = 1 + k*64/60 s,
noon sun, demo settings (coverage 0.2, density 0.05).  One "step" = _update_per_frame_data +
_render_process of the reference (cloud_sky.gd:165-187,234-248): advance the wind offsets on the
host, rebuild the sky LUT, run the march prologue and the march kernel for the whole image.

Metric: Mray-steps/s = marched pixels (dir.y > 0) x nominal primary steps / seconds / 1e6.

N > 1 (launched under torchrun, one rank per GPU): weak scaling of the same workload — step k renders
the N consecutive wind frames k*N .. k*N+N-1, one per rank, every rank straight into its slice of the
gathered [N, H, W, 4] fp16 buffer; the all-gather is FUSED into the march kernel (each
finished pixel is stored into every rank's copy over NVLink, godot-volumetric-cloud-demo-v2_b200/csrc/peer.cu)
and completed by one flag barrier per step; value = N frames' ray-steps / max-over-ranks device time.
`--gather nccl` uses one ncclAllGather per step on a side stream instead (the round-1 path).
After the timed region (never part of `value`): BASELINE configs[3] and [4] as configured and the
single-frame row-band strong scaling, reported under "extra".
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
N_SM = 148


def bench_config(world):
    """The workload description — identical in both arms (the driver compares the two `config` objects)."""
    return {"workload": WORKLOAD, "width": W, "height": H, "primary_steps": PRIMARY, "light_steps": LIGHT, "cone_samples": CONE,
            "animation": "wind frame j has t_j = 1 + j*64/60 s (cloud_sky.gd:165-187), 16 frames cycled; step k renders frame k (N = 1) / frames k*N .. k*N+N-1, one per rank (N > 1)",
            "sun": "noon (0,1,0)",
            "l2": f"GPU arm: flushed between timed steps ({L2_FLUSH_BYTES >> 20} MiB memset outside the per-step event pairs); CPU arm: not applicable",
            "parallelism": "1 GPU" if world == 1 else f"{world} GPUs, weak scaling: {world} consecutive wind frames per step, one per rank, gathered on every rank"}


def frame_time(k):
    return 1.0 + k * (64.0 / 60.0)


NOON = (0.0, 1.0, 0.0)


def frame_params(lib, k, sun, width=W, height=H, coverage=None):
    """Settings -> FrameData -> push constants for animation frame k (host logic of cloud_sky.gd)."""
    s = lib.settings_demo()
    if coverage is not None:
        s.cloud_coverage = coverage
    st = lib.frame_state_init()
    st.light_direction[:] = list(sun)
    for j in range(k + 1):  # integrate the wind offsets frame by frame like the running demo does
        lib.frame_advance(st, s, frame_time(j))
    return lib.fill_cloud_params(s, st, width, height)


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
        pw = [float(r[3]) for r in self.rows if len(r) > 3 and r[3].replace(".", "", 1).isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        # samples taken while the SMs were busy have the highest clocks; report the median of the upper half
        busy = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(pw) if pw else None}


# ---------------------------------------------------------------------------------------------------------
# CPU arms (test infrastructure, timed as the baseline — never on the product path)
# ---------------------------------------------------------------------------------------------------------
def oracle_library(cs):
    """The hand-written CPU oracle: host logic (settings -> push constants) for the CPU arm, and the port fallback."""
    path = os.path.join(ROOT, "oracle", "libcloudsky_oracle.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
    return cs.Library(path)


def reference_binding():
    """tests/refbind.py over oracle/_ref (the reference's own GLSL compiled by g++), or None when it was not built."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refbind
    return refbind if refbind.available() else None


class RefSampler:
    """The compiled reference shaders (oracle/_ref) timed on evenly spaced row bands of a frame (the whole frame when it fits
    the time budget).  The reference hard-codes 128 primary steps and 6 cone + 1 distant light samples (clouds.glsl:186-199,228):
    7 of the configured 8 density evaluations per lit step — the metric counts primary steps, so this favours the reference."""
    kind = "reference"

    def __init__(self, refbind, textures, threads):
        import numpy as np
        self.threads = threads
        self.ref = refbind.Reference(threads=threads)
        self.ref.upload_textures(*textures)
        self.ref.build_transmittance_lut()
        self.buf = np.zeros((H, W, 4), np.float16)
        self.band = min(H, max(threads // 2, 8))
        self.t_band = None
        self.sun = None

    def _steps(self, r0, r1):
        return (W - 1) * (r1 - max(r0, 1)) * PRIMARY  # marched pixels (dir.y > 0: all but row 0 / column 0) x 128

    def sample(self, params, sun, target_seconds):
        band = self.band
        if self.sun != tuple(sun):
            self.ref.build_sky_lut(sun)
            self.sun = tuple(sun)
        if self.t_band is None:  # calibration band (second of two runs), not part of any sample
            for _ in range(2):
                t0 = time.perf_counter()
                self.ref.render(params, W, H, rows=range(H // 3, H // 3 + band), out=self.buf)
                self.t_band = time.perf_counter() - t0
        max_bands = H // band
        n_bands = max(1, min(max_bands, int(target_seconds / max(self.t_band, 1e-6))))
        steps = 0
        t0 = time.perf_counter()
        if n_bands == max_bands:
            self.ref.render(params, W, H, out=self.buf)
            steps = self._steps(0, H)
            desc = "the whole 2048x1024 frame"
        else:
            pitch = H / n_bands
            for i in range(n_bands):
                r0 = min(H - band, int(i * pitch + (pitch - band) / 2))
                self.ref.render(params, W, H, rows=range(r0, r0 + band), out=self.buf)
                steps += self._steps(r0, r0 + band)
            desc = f"{n_bands} evenly spaced bands of {band} rows ({n_bands * band} of {H} rows) of the same 2048x1024 frame"
        sec = time.perf_counter() - t0
        desc += (f"; the reference's own clouds.glsl compiled by g++ -O2 -ffp-contract=off (oracle/_ref), its hard-coded 128 primary / 6 cone + 1 distant "
                 f"light samples, {self.threads} threads")
        return steps / sec / 1e6, desc, sec


class OracleSampler:
    """Fallback when oracle/_ref is absent: the hand-written port at exactly the configured 128 / 7 + 1 steps."""
    kind = "port"

    def __init__(self, cs, textures, threads):
        import numpy as np
        self.threads = threads
        self.ctx = oracle_library(cs).context(0)
        self.ctx.set_threads(threads)
        self.ctx.upload_textures(*textures)
        self.ctx.build_transmittance_lut()
        self.ctx.resize(W, H)
        self.ctx.set_march_config(PRIMARY, CONE)
        self.buf = np.zeros((H, W, 4), np.float16)
        self.band = min(H, max(threads, 8))
        self.t_band = None

    def sample(self, params, sun, target_seconds):
        ctx, band = self.ctx, self.band
        ctx.build_sky_lut(sun)
        if self.t_band is None:
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
        desc += f", {PRIMARY}/{LIGHT} steps, scalar fp32 C++ port (-O2 -ffp-contract=off), {self.threads} threads"
        return steps / sec / 1e6, desc, sec


def make_cpu_sampler(cs, textures, threads):
    rb = reference_binding()
    return RefSampler(rb, textures, threads) if rb is not None else OracleSampler(cs, textures, threads)


def run_reference(args):
    """--impl reference: the reference's own CPU-executable implementation of the path — its three GLSL shaders compiled by g++
    (oracle/_ref; the hand port only if that library is missing) — on all host cores, same config/metric, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import cloudsky_b200 as cs
    from cloudsky_b200 import assets
    large, small, weather, desc = assets.load_default_textures()
    threads = os.cpu_count() or 1
    ora = oracle_library(cs)  # host logic only: settings -> FrameData -> push constants (cloud_sky.gd:165-187,251-289)
    sampler = make_cpu_sampler(cs, (large, small, weather), threads)
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
            "vs_baseline": None, "dtype": "f32", "data": desc, "config": bench_config(args.gpus),
            "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": threads, "kind": sampler.kind, "sample": sample},
            "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------------
# roofline inputs
# ---------------------------------------------------------------------------------------------------------
def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy bandwidth)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_capture(name="roofline_latest.json"):
    """Per-launch figures of the march kernel from the committed `ncu --set full` capture of frame 0 of this workload."""
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            return json.load(f)
    except Exception:
        return {}


def build_roofline(kernel_name, march_ms_avg, frame0_ms, counters_avg, counters0, clocks, capture, peak, peak_src):
    """`roofline`: what binds the march kernel is the SM issue rate (the texel working set is L1/L2 resident, DRAM ~0.1 % of
    peak), so the top-level fraction is issued warp-instructions over issue slots.  SURVEY 8(d)'s algorithmic-byte figure
    against the HBM peak is kept beside it under `algorithmic` (it exceeds 1 by construction), with the bytes the kernel really
    fetches and the DRAM traffic ncu measured."""
    alg_bytes = 80 * counters_avg["density_evals"] + 8 * W * H  # 80 B per density evaluation + 8 B per pixel
    alg_gbs = alg_bytes / (march_ms_avg * 1e-3) / 1e9
    fetched = 16 * counters_avg["density_evals"] + 32 * counters_avg["large_fetches"] + 16 * counters_avg["small_fetches"] + 8 * W * H
    sm_mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965
    issue_peak = N_SM * 4 * sm_mhz * 1e6  # warp instructions per second: 4 schedulers per SM, one instruction per clock each
    inst = capture.get("warp_instructions_per_launch")
    traffic = capture.get("dram_bytes_per_launch")
    r = {"bound": "issue", "kernel": kernel_name, "unit": "Gwarp-inst/s", "peak": round(issue_peak / 1e9, 1),
         "peak_source": f"{N_SM} SMs x 4 schedulers x {sm_mhz} MHz (SM clock sampled during the timed region)",
         "achieved": None, "frac": None, "traffic": traffic, "traffic_source": capture.get("source"),
         "kernel_ms": round(march_ms_avg, 4), "kernel_ms_frame0": round(frame0_ms, 4),
         "warp_instructions_per_launch": inst,
         "algorithmic": {"bound": "hbm", "achieved": round(alg_gbs, 1), "peak": peak, "unit": "GB/s", "frac": round(alg_gbs / peak, 4),
                         "bytes_per_launch": int(alg_bytes), "peak_source": peak_src,
                         "note": "SURVEY 8(d): 80 B x executed density evaluations + 8 B x pixels over the kernel's average CUDA-event time in the timed region; "
                                 "not a physical bound here — the texels are cache-resident and early exits skip fetches, so it exceeds 1"},
         "fetched": {"bytes_per_launch": int(fetched), "GBps": round(fetched / (march_ms_avg * 1e-3) / 1e9, 1),
                     "note": "bytes the kernel's loads really request: 16 B weather record per evaluation + 32 B per large-volume fetch + 16 B per small-volume fetch "
                             "(device counters) + 8 B per pixel stored; served by L1/L2"},
         "dram_frac": round(traffic / (frame0_ms * 1e-3) / 1e9 / peak, 5) if traffic else None,
         "l1_hit_pct": capture.get("l1_hit_pct"), "l2_hit_pct": capture.get("l2_hit_pct"), "ncu_issue_active_pct": capture.get("issue_active_pct")}
    if inst:
        ach = inst / (frame0_ms * 1e-3)
        r["achieved"] = round(ach / 1e9, 1)
        r["frac"] = round(ach / issue_peak, 4)
        r["note"] = ("issue-slot fraction = warp instructions of the frame-0 launch (ncu smsp__inst_executed.sum of the committed capture; "
                     f"{counters0['density_evals']} density evaluations, same workload) / (live CUDA-event time of that launch x issue slots)")
    return r


# ---------------------------------------------------------------------------------------------------------
# the CUDA arm
# ---------------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import cloudsky_b200 as cs
    from cloudsky_b200 import assets, sharding

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

    def allmax(x):
        if dist is None:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allgather_floats(x):
        if dist is None:
            return [float(x)]
        mine = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        out = torch.zeros(world, dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(out, mine)
        return [float(v) for v in out.tolist()]

    def allmin_flag(ok):
        if dist is None:
            return bool(ok)
        t = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

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
    sun = NOON
    frame_bytes = W * H * 8

    # ---- gathered output: [world, H, W, 4] fp16 on every rank, two slots -----------------------------------------------
    gather = "none"
    peer = None
    if world > 1:
        gather = args.gather
        if gather in ("auto", "peer"):
            ok = True
            try:
                peer = sharding.PeerBuffers(ctx, world * frame_bytes, slots=2)
                peer.activate()
                peer.barrier()
                ctx.sync()
                ctx.peer_check()
            except Exception as e:  # no CUDA IPC / peer access between these devices
                ok = False
                sys.stderr.write(f"[bench rank {rank}] peer gather unavailable: {e}\n")
            if not allmin_flag(ok):
                if gather == "peer":
                    raise SystemExit("bench.py: --gather peer requested but peer mapping failed")
                peer = None
            gather = "peer" if peer is not None else "nccl"
    if peer is not None:
        gathered = [peer.tensor(b, (world, H, W, 4)) for b in range(2)]
    else:
        gathered = [torch.zeros((world, H, W, 4), dtype=torch.float16, device="cuda") for _ in range(2 if world > 1 else 1)]
    gather_stream = torch.cuda.Stream() if gather == "nccl" else None
    rendered = [torch.cuda.Event() for _ in gathered]
    gather_done = [torch.cuda.Event() for _ in gathered]
    gather_used = [False for _ in gathered]
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")
    params = [frame_params(lib, k % 16, sun) for k in range(16)]

    # work counters of all 16 frames (deterministic; measured outside the timed region with the instrumented kernel)
    ctx.build_sky_lut(sun)
    ctx.set_counters_enabled(True)
    counters_all = []
    for k in range(16):
        ctx.render_frame(params[k])
        counters_all.append(ctx.get_counters().as_dict())
    ctx.set_counters_enabled(False)
    counters = counters_all[0]
    counters_avg = {key: sum(c[key] for c in counters_all) / 16.0 for key in counters}

    def step(k):
        p = params[(k * world + rank) % 16]  # N = 1: frame k; N ranks: frames k*N .. k*N+N-1, this rank's is k*N + rank
        b = k % len(gathered)
        buf = gathered[b]
        if gather == "nccl" and gather_used[b]:
            stream.wait_event(gather_done[b])         # this buffer's previous all-gather (step k-2) has finished
        ctx.build_sky_lut(sun)                        # sky_lut.update_lut (cloud_sky.gd:187)
        ctx.render_rows_to(p, 0, H, buf[rank].data_ptr())  # prologue + march into this rank's slice (peer: and into every peer's copy)
        if gather == "peer":
            peer.barrier()                            # the fused all-gather's completion: every rank's frame k is in this copy
        elif gather == "nccl":
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
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    kt = ctx.read_kernel_timings()
    # Soak: when the driver asks for few steps the timed region is far shorter than nvidia-smi's sampling period.  Keep stepping
    # (untimed for `value`, same loop) until >= 1.2 s of device work has passed so the clock record covers real load.
    soak = None
    if dev_ms < 1200.0:
        n_soak = int(math.ceil((1200.0 - dev_ms) / max(dev_ms / args.steps, 1e-3)))
        if dist is not None:
            t = torch.tensor([n_soak], dtype=torch.int64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            n_soak = int(t.item())
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        for k in range(n_soak):
            flush.zero_()
            step(args.warmup + args.steps + k)
        drain()
        s1.record(stream)
        torch.cuda.synchronize()
        soak = {"steps": n_soak, "ms_per_step_incl_flush": round(s0.elapsed_time(s1) / max(1, n_soak), 4),
                "note": "same step loop continued after the timed region so that the nvidia-smi clock record spans > 1 s of load; not part of value"}
        ctx.read_kernel_timings()
    clocks = sampler.stop() if rank == 0 else None
    ctx.set_kernel_timing(False)
    if peer is not None:
        ctx.peer_check()
    my_kernel_ms = (kt["march_ms"] + kt["sky_ms"]) / max(1, kt["march_launches"])
    dev_ms = allmax(dev_ms)
    per_rank_kernel_ms = [round(v, 4) for v in allgather_floats(my_kernel_ms)]
    per_rank_march_ms = [round(v, 4) for v in allgather_floats(kt["march_ms"] / max(1, kt["march_launches"]))]

    ray_steps_per_frame = counters["marched_pixels"] * PRIMARY
    value = world * ray_steps_per_frame * args.steps / (dev_ms * 1e-3) / 1e6
    ms_per_step = dev_ms / args.steps

    # ---- gather bit-identity (the driver's 1-GPU test box skips the multi-GPU tests): the last two steps' gathered buffers must
    # hold, in the slice of rank (r+1) % N, exactly what this rank renders locally for that rank's frame -----------------------
    gather_bit_identical = None
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        nb = (rank + 1) % world
        last = args.warmup + args.steps + (soak["steps"] if soak else 0) - 1
        okb = True
        check = torch.empty((H, W, 4), dtype=torch.float16, device="cuda")
        for k in (last - 1, last):
            ctx.render_rows_to(params[(k * world + nb) % 16], 0, H, check.data_ptr())
            ctx.sync()
            okb = okb and bool(torch.equal(check.view(torch.int16), gathered[k % len(gathered)][nb].view(torch.int16)))
        gather_bit_identical = allmin_flag(okb)
        dist.barrier()

    # ---- end-to-end through the C-ABI with HOST buffers ------------------------------------------------------------------------
    e2e = None
    hosts = [torch.empty((H, W, 4), dtype=torch.float16).pin_memory() for _ in range(2)]
    if world == 1:
        for k in range(max(3, min(args.warmup, 10))):  # warm-up of the streaming path (second image, copy stream, both host buffers)
            ctx.render_frame_host_async(params[k % 16], hosts[k & 1].data_ptr())
        ctx.wait_host()
        t0 = time.perf_counter()
        for k in range(args.steps):  # streaming: frame k+1's kernels overlap frame k's device->host copy
            ctx.render_frame_host_async(params[(args.warmup + k) % 16], hosts[k & 1].data_ptr())
        ctx.wait_host()  # every result is in host memory when the clock stops
        e2e_s = time.perf_counter() - t0
        e2e_note = ("cs_render_frame_host_async + cs_wait_host: push constants from host, sky LUT + prologue + march per step, every 16 MiB RGBA16F result copied to "
                    "pinned host memory (copy of frame k overlaps the kernels of frame k+1); wall clock around all steps incl. the final wait")
    else:
        # N ranks: the same loop as the timed region INCLUDING the gather (fused stores + barrier, or NCCL), then each rank copies
        # its own finished frame to pinned host memory on a copy stream — the union over ranks is the job's complete result, once.
        copy_stream = torch.cuda.Stream()
        copied = [torch.cuda.Event() for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]
        used = [False, False]

        def e2e_step(k):
            b = k % 2
            if used[b]:
                stream.wait_event(copied[b])  # this rank's slice of slot b was read out (step k-2)
            step(k)
            if gather == "nccl":
                stream.wait_event(gather_done[b])
            done[b].record(stream)
            copy_stream.wait_event(done[b])
            with torch.cuda.stream(copy_stream):
                hosts[b].copy_(gathered[b][rank], non_blocking=True)
                copied[b].record(copy_stream)
            used[b] = True

        for k in range(4):
            e2e_step(k)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for k in range(args.steps):
            e2e_step(args.warmup + k)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        e2e_note = (f"per step and rank: push constants from host, sky LUT + prologue + march into the gathered buffer, the gather ({gather}), then this rank's own "
                    "16 MiB frame copied to pinned host memory on a copy stream (overlaps the next step); wall clock, max over ranks")
    e2e_s = allmax(e2e_s)
    assert torch.isfinite(hosts[0].float()).all()
    e2e = {"value": round(world * ray_steps_per_frame * args.steps / e2e_s / 1e6, 1), "unit": UNIT,
           "h2d_bytes_per_step": 112 + 12, "d2h_bytes_per_step": W * H * 8, "includes_gather": world > 1, "note": e2e_note}

    # ---- frame-0 kernel time (the launch the committed ncu capture profiles) -----------------------------------------------------
    ctx.build_sky_lut(sun)
    frame0_ms = ctx.time_render_frame(params[0], 3, 10)

    extra = {}
    # Extra, NOT the headline: the opt-in CS_MODE_EARLY_OUT flag (rays stop once T < 2^-12; results within 2 fp16 ulps,
    # SURVEY 7.3-5 asks for nominal AND executed steps to be reported).
    if world == 1:
        ctx.set_march_config(PRIMARY, CONE, base_mode | cs.MODE_EARLY_OUT)
        ctx.set_counters_enabled(True)
        ctx.render_frame(params[0])
        k_early = ctx.get_counters().as_dict()
        ctx.set_counters_enabled(False)
        ms_early = ctx.time_render_frame(params[0], 3, 10)
        extra["early_out_mode"] = {"march_ms": round(ms_early, 4), "value_on_nominal_steps": round(ray_steps_per_frame / ms_early / 1e3, 1),
                                   "value_on_executed_steps": round(k_early["primary_steps"] / ms_early / 1e3, 1),
                                   "executed_step_fraction": round(k_early["primary_steps"] / (counters["marched_pixels"] * PRIMARY), 4),
                                   "note": "opt-in mode flag, not reference behaviour (clouds.glsl:172 runs every step); not used for value / e2e"}
        ctx.set_march_config(PRIMARY, CONE, base_mode)
    # Extra: the opt-in CS_MODE_TEX flag (the texture unit filters the three input textures with its 8-bit fixed-point weights,
    # like the reference's own sampler bindings; the headline keeps the in-kernel fp32 filter).
    if world == 1 and args.sampler == "kernel":
        ctx.set_march_config(PRIMARY, CONE, cs.MODE_FAST | cs.MODE_TEX)
        ms_tex = ctx.time_render_frame(params[0], 3, 10)
        extra["texture_unit_mode"] = {"march_ms": round(ms_tex, 4), "value_march_only": round(ray_steps_per_frame / ms_tex / 1e3, 1),
                                      "note": "opt-in mode flag (bench.py --sampler texture runs the whole bench in it); same parity tolerance as FAST, tests/test_gpu_parity.py"}
        # Extra: the opt-in CS_MODE_HALF flag (the in-kernel filter in packed fp16 on the exact-integer records: 11-bit weights)
        ctx.set_march_config(PRIMARY, CONE, cs.MODE_FAST | cs.MODE_HALF)
        ms_half = ctx.time_render_frame(params[0], 3, 10)
        extra["half_filter_mode"] = {"march_ms": round(ms_half, 4), "value_march_only": round(ray_steps_per_frame / ms_half / 1e3, 1),
                                     "note": "opt-in mode flag: HFMA2 filter, no half->float conversions; same parity tolerance as FAST, tests/test_gpu_parity.py"}
        ctx.set_march_config(PRIMARY, CONE, base_mode)

    if not args.no_extra:
        extra.update(run_extras(cs, sharding, lib, ctx, torch, dist, world, rank, allmax, allgather_floats, allmin_flag, base_mode, peer, gather))

    if peer is not None:
        peer.close()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peak()
    march_ms = kt["march_ms"] / max(1, kt["march_launches"])
    capture = ncu_capture("roofline_tex.json" if args.sampler == "texture" else "roofline_latest.json")
    roofline = build_roofline("clouds_fast_kernel", march_ms, frame0_ms if frame0_ms else march_ms, counters_avg, counters, clocks, capture, peak, peak_src)
    line = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": desc, "config": bench_config(world),
            "sampler": "in-kernel fp32 trilinear (coefficient records)" if args.sampler == "kernel" else "texture unit (CS_MODE_TEX, 8-bit filter weights)",
            "gather": {"none": "single GPU", "peer": "fused into the march kernel: every pixel stored into all ranks' copies over NVLink (peer-mapped, CUDA IPC) + one flag barrier per step; no collective kernel",
                       "nccl": "one ncclAllGather of the finished textures per step on a side stream (overlaps the next step)"}[gather],
            "gather_bit_identical": gather_bit_identical,
            "nvlink_bytes_per_step_per_gpu": (world - 1) * frame_bytes if world > 1 else 0,
            "workload_stats": {"lit_fraction": round(counters["lit_steps"] / counters["primary_steps"], 4), "density_evals_frame0": counters["density_evals"],
                               "density_evals_avg": counters_avg["density_evals"], "marched_pixels": counters["marched_pixels"]},
            "roofline": roofline, "e2e": e2e, "gpu_launches": (3 + (1 if gather == "peer" else 0)) * args.steps,
            "kernels": {"march_ms_avg": round(march_ms, 4), "sky_lut_ms_avg": round(kt["sky_ms"] / max(1, kt["sky_launches"]), 4)},
            "gevals_per_s": round(world * counters_avg["density_evals"] * args.steps / (dev_ms * 1e-3) / 1e9, 2),
            "wall_ms_per_step_incl_flush": round(1e3 * t_wall / args.steps, 3), "clocks": clocks, "soak": soak,
            "per_rank_kernel_ms": per_rank_kernel_ms,  # sky LUT + march per step on every rank (load balance)
            "per_rank_march_ms": per_rank_march_ms,
            "value_without_gather": round(world * ray_steps_per_frame / (max(per_rank_kernel_ms) * 1e-3) / 1e6, 1),
            "extra": extra}
    if world == 1 and not args.no_cpu:
        cpu_threads = os.cpu_count() or 1
        sampler_cpu = make_cpu_sampler(cs, (large, small, weather), cpu_threads)
        host_lib = oracle_library(cs)
        p0 = frame_params(host_lib, 0, sun)
        cpu_v, cpu_sample, _ = sampler_cpu.sample(p0, sun, 12.0)
        one = make_cpu_sampler(cs, (large, small, weather), 1)
        one_v, one_sample, _ = one.sample(p0, sun, 6.0)
        line["cpu_baseline"] = {"value": round(cpu_v, 3), "unit": UNIT, "cores": cpu_threads, "kind": sampler_cpu.kind, "sample": cpu_sample,
                                "single_thread": {"value": round(one_v, 3), "unit": UNIT, "cores": 1, "sample": one_sample}}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def run_extras(cs, sharding, lib, ctx, torch, dist, world, rank, allmax, allgather_floats, allmin_flag, base_mode, peer, gather):
    """BASELINE configs[3] and [4] as configured, and the row-band strong scaling of single frames.  Outside the
    timed region; every number is device time (CUDA events), max over ranks."""
    import numpy as np
    out = {}
    stream = torch.cuda.current_stream()
    use_peer = world == 1 or gather == "peer"
    R = sharding.ShardedRenderer(ctx, W, H, device="cuda", gather="peer" if use_peer else "nccl")

    def timed(fn, iters, warm=2, kernel_timing=True):
        """-> (ms per call, max over ranks; this rank's kernel ms per call; last result).  kernel_timing=False: the sun-batch kernel
        only runs with per-kernel event timing off (cs_render_sun_batch_to), so this rank's own elapsed time is returned instead."""
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if kernel_timing:
            ctx.set_kernel_timing(True)
        e0.record(stream)
        for _ in range(iters):
            res = fn()
        e1.record(stream)
        torch.cuda.synchronize()
        mine = e0.elapsed_time(e1) / iters
        if kernel_timing:
            kt = ctx.read_kernel_timings()
            ctx.set_kernel_timing(False)
            return allmax(mine), (kt["march_ms"] + kt["sky_ms"]) / iters, res
        return allmax(mine), mine, res

    noon = (0.0, 1.0, 0.0)
    # ---- C4: 2048x1024 x 64 sun angles of ONE cloud field, 64/N suns per rank through the sun-batch kernel, gathered on every rank
    if base_mode == cs.MODE_FAST:
        n_suns = 64
        suns = sharding.sun_sweep(n_suns)
        p0 = frame_params(lib, 0, noon)
        ctx.set_march_config(PRIMARY, CONE, base_mode)
        ctx.set_counters_enabled(True); ctx.build_sky_lut(noon); ctx.render_frame(p0); marched = ctx.get_counters().as_dict()["marched_pixels"]; ctx.set_counters_enabled(False)
        ms, kern_ms, sweep = timed(lambda: R.render_sun_sweep(p0, suns), 3, warm=1, kernel_timing=False)
        k = (rank * (n_suns // world) + 3) % n_suns  # one of this rank's ... and one of the neighbour's frames against a local single dispatch
        okc = True
        for kk in (k, (k + n_suns // world) % n_suns):
            q = p0.copy(); q.light_direction[:] = suns[kk].tolist()
            single = torch.from_numpy(ctx.render_frame_host(q)).cuda()
            okc = okc and bool(torch.equal(single.view(torch.int16), sweep[kk].view(torch.int16)))
        out["c4_sun_sweep"] = {"config": "BASELINE configs[3]: 2048x1024 x 64 sun angles of one cloud field, 128 primary / 8 light steps, strong scaling: 64/N suns per rank, gathered on every rank",
                               "ms_per_sweep": round(ms, 3), "ms_per_frame": round(ms / n_suns, 4), "value": round(marched * PRIMARY * n_suns / ms / 1e3, 1), "unit": UNIT,
                               "per_rank_ms_incl_barrier": [round(v, 3) for v in allgather_floats(kern_ms)],
                               "gathered_bytes_per_rank": n_suns * W * H * 8, "bit_identical_to_single_dispatch": allmin_flag(okc),
                               "kernel": "clouds_fast_sunbatch_kernel: 4 suns per launch share the sun-independent primary march; 64 sky LUT builds included",
                               "gather": "fused peer stores + flag barrier" if use_peer and world > 1 else ("nccl all-gather" if world > 1 else "single GPU")}
        del sweep
    # ---- single-frame row-band strong scaling (C3 and C5 shapes): bands_per_rank 1 (contiguous) and 4 (interleaved) -----------------
    for tag, (w, h, P, cone, cov, iters) in {"c3_row_bands": (W, H, PRIMARY, CONE, None, 10), "c5_row_bands": (8192, 4096, 256, 11, 1.0, 2)}.items():
        ctx.resize(w, h)
        Rb = sharding.ShardedRenderer(ctx, w, h, device="cuda", gather="peer" if use_peer else "nccl")
        ctx.set_march_config(P, cone, base_mode)
        p = frame_params(lib, 0, noon, w, h, coverage=cov)
        ctx.build_sky_lut(noon)
        ctx.set_counters_enabled(True); ctx.render_frame(p); marched = ctx.get_counters().as_dict()["marched_pixels"]; ctx.set_counters_enabled(False)
        want = torch.from_numpy(ctx.read_image()).cuda()
        res = {"config": f"{w}x{h}, {P} primary / {cone + 1} light steps, coverage {cov if cov is not None else 0.2}, ONE frame split into row bands over {world} GPU(s), gathered on every rank"}
        for bpr in (1, 4, "max"):
            if bpr == "max" and not use_peer:
                continue  # one NCCL all-gather per band group: only sensible for few bands
            if bpr != "max" and h % (world * bpr):
                continue
            ms, kern_ms, img = timed(lambda: Rb.render_frame_rows(p, bands_per_rank=bpr), iters, warm=3)
            per_rank = allgather_floats(kern_ms)
            res[f"bands_per_rank_{bpr}"] = {"bands": "contiguous" if bpr == 1 else (f"interleaved, {h // (world * bpr)} rows each, one launch per rank" if bpr != "max" else "interleaved, 8 rows each (one CTA row), one launch per rank"),
                                            "ms_per_frame": round(ms, 4), "value": round(marched * P / ms / 1e3, 1), "unit": UNIT,
                                            "per_rank_kernel_ms": [round(v, 4) for v in per_rank],
                                            "imbalance_max_over_mean": round(max(per_rank) / (sum(per_rank) / len(per_rank)), 4),
                                            "bit_identical_to_one_gpu_dispatch": allmin_flag(bool(torch.equal(img.view(torch.int16), want.view(torch.int16))))}
        if tag == "c5_row_bands":  # the opt-in early-out flag on the same shape (exact to 2 fp16 ulps)
            ctx.set_march_config(P, cone, base_mode | cs.MODE_EARLY_OUT)
            ms, kern_ms, img = timed(lambda: Rb.render_frame_rows(p, bands_per_rank="max" if use_peer else 4), iters, warm=2)
            res["early_out_interleaved"] = {"ms_per_frame": round(ms, 4), "value_on_nominal_steps": round(marched * P / ms / 1e3, 1), "unit": UNIT,
                                                 "per_rank_kernel_ms": [round(v, 4) for v in allgather_floats(kern_ms)]}
        out[tag] = res
        Rb.close()
        del want
    R.close()
    ctx.resize(W, H)
    ctx.set_march_config(PRIMARY, CONE, base_mode)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sampler", default="kernel", choices=["kernel", "texture"],
                    help="kernel: in-kernel fp32 trilinear filter (default, the headline); texture: CS_MODE_TEX, the GPU texture unit filters")
    ap.add_argument("--gather", default="auto", choices=["auto", "peer", "nccl"],
                    help="N > 1: peer = all-gather fused into the march kernel over peer-mapped memory; nccl = ncclAllGather on a side stream; auto = peer if it maps")
    ap.add_argument("--no-extra", action="store_true", help="skip the configs[3]/[4]/row-band extras after the timed region")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (N = 1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
