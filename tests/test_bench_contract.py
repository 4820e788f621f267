"""bench.py contract checks that need no GPU: the reference arm runs the compiled reference shaders (oracle/_ref) on host
cores and prints the contract keys; both arms describe the workload with the same `config` object."""
import json
import os
import subprocess
import sys

import bench
import refbind

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    refbind.build()
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == bench.UNIT and line["higher_is_better"] is True
    assert line["config"] == bench.bench_config(1)                      # the same object the CUDA arm prints
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    cb = line["cpu_baseline"]
    assert cb["value"] == line["value"] and cb["cores"] == (os.cpu_count() or 1)
    assert cb["kind"] == ("reference" if refbind.available() else "port")
    if refbind.available():
        assert "clouds.glsl compiled by g++" in cb["sample"]


def test_config_is_shared_and_names_the_workload():
    for n in (1, 2, 8):
        c = bench.bench_config(n)
        assert c["workload"].startswith("C3: 2048x1024") and c["primary_steps"] == 128 and c["light_steps"] == 8 and c["cone_samples"] == 7
        assert "flushed" in c["l2"]
    assert bench.bench_config(1)["parallelism"] == "1 GPU" and "8 consecutive wind frames" in bench.bench_config(8)["parallelism"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
