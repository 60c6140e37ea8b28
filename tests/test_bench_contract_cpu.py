"""CPU: the reference arm of bench.py prints the JSON contract (metric/unit/config, cpu_baseline, e2e zeros)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_contract():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                          "msda_encoder", "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["value"] > 0 and line["unit"] == "images/s" and "workload" in line["config"]


def test_workload_registry_and_defaults():
    sys.path.insert(0, ROOT)
    import bench_workloads as B
    assert B.DEFAULT_WORKLOAD == "pair_forward"
    assert set(B.WORKLOADS) >= {"pair_forward", "msda_encoder", "gdino_head"}
    assert set(B._CPU) >= set(B.WORKLOADS)
    p = B.measured_peaks()
    assert p["hbm_gbs"] > 1000 and p["bf16_tflops_sustained"] > 100
