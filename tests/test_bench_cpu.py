"""bench.py contract on CPU: the reference arm (CPU port of the path) must print ONE JSON line with the driver's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "c1-tiny", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "decode tokens/sec" and d["unit"] == "tokens/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["config"]["workload"].startswith("internlm2-1.8b fp16")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    e = d["e2e"]
    assert e == {"value": d["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_configs_name_the_baseline_workloads():
    sys.path.insert(0, ROOT)
    import bench
    from lmdeploy_b200.turbomind.specs import SPECS
    assert bench.CONFIGS["c2"] == ("internlm2-7b", 64, 2048, 8)          # BASELINE.json configs[1]: the quoted metric
    assert set(bench.CONFIGS) >= {"c1-tiny", "c2", "c3", "c4", "c5"}
    for name, (spec, batch, ctx, qp) in bench.CONFIGS.items():
        assert spec in SPECS and 1 <= batch <= 128 and qp in (0, 4, 8)
    assert SPECS["llama3-70b"].weight_format == "fp8" and SPECS["llama3-70b"].dtype == "bf16"
    assert bench.MIN_GPUS["c5"] >= 2
