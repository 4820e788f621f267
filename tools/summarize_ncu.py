#!/usr/bin/env python
"""Summarise an `ncu --set full` report of the march kernel into profiles/<name>.json and refresh
profiles/roofline_latest.json (read by bench.py for `roofline.traffic`).  Runs where ncu is installed (no GPU needed).

  python tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01_v4_clouds_fast_ncu_summary.json "<command that was profiled>"
"""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_static",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma_type_fp16.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_active.avg.per_cycle_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main():
    rep, out, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    roofline_out = sys.argv[4] if len(sys.argv) > 4 else None  # e.g. profiles/roofline_latest.json; not written unless named
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    kernels = []
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        m = {h: {"value": v, "unit": u} for h, u, v in zip(hdr, units, vals) if h in KEEP}
        dram = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            if k in m:
                dram += float(m[k]["value"]) * UNIT.get(m[k]["unit"], 1.0)
        kernels.append({"kernel": d.get("Kernel Name"), "dram_bytes": dram, "metrics": m})
    summary = {"report": rep, "command": cmd, "launches": kernels}
    json.dump(summary, open(out, "w"), indent=1)
    if kernels and roofline_out:
        k = kernels[-1]
        hit = lambda name: float(k["metrics"][name]["value"]) if name in k["metrics"] else None
        json.dump({"dram_bytes_per_launch": k["dram_bytes"], "kernel": k["kernel"], "warp_instructions_per_launch": hit("smsp__inst_executed.sum"),
                   "kernel_ms_under_ncu": (hit("gpu__time_duration.sum") or 0.0) * {"us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "ns": 1e-6, "nsecond": 1e-6, "s": 1e3, "second": 1e3}.get(k["metrics"].get("gpu__time_duration.sum", {}).get("unit"), 1e-3),
                   "l2_hit_pct": hit("lts__t_sector_hit_rate.pct"),
                   "l1_hit_pct": hit("l1tex__t_sector_hit_rate.pct"), "issue_active_pct": hit("smsp__issue_active.avg.pct_of_peak_sustained_active"), "source": f"{out} (ncu --set full --clock-control none: dram__bytes_read.sum + dram__bytes_write.sum)"},
                  open(roofline_out, "w"), indent=1)
    print(json.dumps({k["kernel"]: k["dram_bytes"] for k in kernels}))


if __name__ == "__main__":
    main()
