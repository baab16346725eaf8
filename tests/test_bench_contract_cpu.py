"""CPU (-m "not gpu"): the reference arm of bench.py (the reference's own TraceKernel on host
cores, or the oracle port when oracle/_ref is absent) prints the contract's JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    # OMP_NUM_THREADS=1 is what torch.distributed.run exports to its workers: the arm must
    # still run on -- and report -- all host cores
    env = dict(os.environ, OMP_NUM_THREADS="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--steps", "1", "--warmup", "0", "--slices", "20", "--cpu-seconds", "1"],
                         capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = res.stdout.strip().splitlines()
    assert len(lines) == 1, "stdout must carry the JSON line and nothing else"
    line = json.loads(lines[-1])
    assert line["impl"] == "reference" and line["metric"] == "rays/s" and line["unit"] == "rays/s"
    assert line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port")
    assert line["cpu_baseline"]["cores"] == os.cpu_count()
    assert ("%d host threads" % os.cpu_count()) in line["cpu_baseline"]["sample"]
    assert line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "rays/s", "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and line["gpu_launches"] == 0
