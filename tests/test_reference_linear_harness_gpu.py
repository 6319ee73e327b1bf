"""The reference's OWN linear test harness (tests/turbomind/linear/test_linear.py: `test_smoke_linear_correctness`,
bf16 x e4m3-128x128 on llama2_7b_o, batches 1 / 3 / 16) driving THIS repository's GEMM through the `_turbomind` shim
(lmdeploy_b200/shim/_turbomind.py), unchanged.  The harness is not part of this repository: it runs where a copy of the
reference's tests/turbomind/linear directory is reachable -- /root/reference in the build container (no GPU there), or the
directory B200_REF_LINEAR_TESTS names (a scratch copy shipped for one GPU call; log: profiles/r02_raw/ref_linear_harness.log)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _harness_dir():
    for d in (os.environ.get("B200_REF_LINEAR_TESTS"), "/root/reference/tests/turbomind/linear"):
        if d and os.path.exists(os.path.join(d, "test_linear.py")):
            return d
    return None


def test_reference_smoke_linear_suite_passes_against_the_shim():
    d = _harness_dir()
    if d is None:
        pytest.skip("the reference's tests/turbomind/linear directory is not reachable (set B200_REF_LINEAR_TESTS)")
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "lmdeploy_b200", "shim"), ROOT, env.get("PYTHONPATH", "")])
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(d, "test_linear.py"), "-q", "-p", "no:cacheprovider",
                        "-k", "smoke", "-rA"], env=env, capture_output=True, text=True, timeout=900)
    print(r.stdout[-4000:])
    print(r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-3000:]
    assert " passed" in r.stdout and "skipped" not in r.stdout.splitlines()[-1], r.stdout[-500:]
