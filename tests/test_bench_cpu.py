"""bench.py contract on CPU: the reference arm (CPU port of the path) must print ONE JSON line with the driver's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (module level is argparse-free: main() parses)


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
    # steps / ms_per_step describe what ran (one sampled layer + head per step); value comes from the extrapolated whole step
    assert d["steps_requested"] == 1 and d["ms_per_full_step_extrapolated"] >= d["ms_per_step"]
    assert abs(d["value"] - 1 / (d["ms_per_full_step_extrapolated"] * 1e-3)) < 1e-6 * d["value"]        # c1-tiny: bs = 1
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


class _FakeNvml:
    """The five NVML calls bench.ClockSampler makes, with a clock that dips and a power-cap flag mid-run."""
    NVML_CLOCK_SM = 1

    def __init__(self, by_uuid=True):
        self.n = 0
        self.by_uuid = by_uuid
        self.opened = None

    def nvmlInit(self):
        pass

    def nvmlDeviceGetHandleByUUID(self, u):
        if not self.by_uuid or isinstance(u, str):
            raise RuntimeError("not found")
        self.opened = ("uuid", u)
        return "h"

    def nvmlDeviceGetHandleByIndex(self, i):
        self.opened = ("index", i)
        return "h"

    def nvmlDeviceGetMaxClockInfo(self, h, which):
        return 1965

    def nvmlDeviceGetClockInfo(self, h, which):
        self.n += 1
        return 1965 if self.n % 4 else 1800

    def nvmlDeviceGetCurrentClocksEventReasons(self, h):
        return 0x4 if self.n > 5 else 0


def test_clock_sampler_counts_only_samples_inside_the_timed_region():
    import time
    nv = _FakeNvml()
    cs = bench.ClockSampler(3, uuid="GPU-abc", nvml=nv, period_s=0.001).start()
    assert cs.backend == "nvml" and nv.opened == ("uuid", b"GPU-abc")
    time.sleep(0.02)                                   # "warm-up": sampled, but outside the region
    before = len(cs.samples)
    assert before > 0
    with cs:
        time.sleep(0.05)
    time.sleep(0.01)
    s = cs.summary()
    assert s["source"] == "nvml" and s["sm_max_mhz"] == 1965 and s["sm_mhz"] == 1965
    assert 5 <= s["samples"] <= len(cs.samples) - before
    assert s["reasons"] == ["sw_power_cap"]
    n = len(cs.samples)
    time.sleep(0.01)
    assert len(cs.samples) == n                        # the thread stopped with the region


def test_clock_sampler_falls_back_to_the_device_index_and_never_raises():
    nv = _FakeNvml(by_uuid=False)
    cs = bench.ClockSampler(5, uuid="GPU-zzz", nvml=nv).start()
    assert nv.opened == ("index", 5)
    cs._stop.set()

    class Broken:
        def nvmlInit(self):
            raise RuntimeError("no driver")
    cs = bench.ClockSampler(0, nvml=Broken())
    with cs.wait_ready(limit_s=0.2):                   # no NVML, and nvidia-smi is absent or silent here
        pass
    s = cs.summary()
    assert s["samples"] == 0 and s["sm_mhz"] is None and s["source"] in ("nvidia-smi", "none")
