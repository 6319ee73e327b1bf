"""The bodies of the GPU tests that have not run on a device yet (tests/test_zz_pending_gpu.py), executed on the CPU
against the dry-run stand-ins: the C library records calls and computes nothing, so every test is EXPECTED to end at its
first numeric assertion -- what must not happen is any other exception (a misspelt name, a wrong argument list, a tensor
of the wrong shape handed to the binding), which would otherwise cost a GPU call to discover."""
import inspect
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import dryrun  # noqa: E402

FACTORIES = ("tensor", "zeros", "ones", "empty", "arange", "randn", "rand", "randint", "full", "zeros_like", "empty_like")


@pytest.fixture
def cpu_as_cuda(monkeypatch, lib):
    rec = dryrun.install(monkeypatch.setattr, lib)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    for name in FACTORIES:
        orig = getattr(torch, name)

        def wrapped(*a, _orig=orig, **k):
            if "device" in k and str(k["device"]).startswith("cuda"):
                k["device"] = "cpu"
            return _orig(*a, **k)
        monkeypatch.setattr(torch, name, wrapped)
    to = torch.Tensor.to

    def to_cpu(self, *a, **k):
        a = tuple("cpu" if (isinstance(x, str) and x.startswith("cuda")) or (isinstance(x, torch.device) and x.type == "cuda")
                  else x for x in a)
        if "device" in k and str(k["device"]).startswith("cuda"):
            k["device"] = "cpu"
        return to(self, *a, **k)
    monkeypatch.setattr(torch.Tensor, "to", to_cpu)
    import lmdeploy_b200.turbomind.runtime as RT
    init = RT.DecodeModel.__init__

    def init_on_cpu(self, spec, **kw):
        kw["device"] = "cpu"
        init(self, spec, **kw)
    monkeypatch.setattr(RT.DecodeModel, "__init__", init_on_cpu)
    gen = torch.Generator
    monkeypatch.setattr(torch, "Generator", lambda device=None: gen())

    class Event:
        def __init__(self, *a, **k):
            pass

        def record(self, *a):
            pass

        def elapsed_time(self, other):
            return 1.0
    monkeypatch.setattr(torch.cuda, "Event", Event)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setenv("B200_RUN_UNVERIFIED", "1")

    # the oracle/_ref libraries (reference kernels) cannot be dlopen-ed without the CUDA driver: stand-ins that accept any
    # call, so the Python that drives them (oracle/ref_step.py, ref_gemm_lib.py, ref_lib.args_from_b200) runs too
    class AnyLib:
        def __init__(self):
            self.calls = []

        def __getattr__(self, name):
            if name.startswith("_"):
                raise AttributeError(name)

            def call(*args):
                self.calls.append(name)
                if name in ("ref_gemm_create", "ref_linear_create"):
                    args[0]._obj.value = 0x2000 + len(self.calls)
                return b"stand-in" if name.endswith("version") else 0
            return call
    from oracle import ref_gemm_lib, ref_lib
    ref = AnyLib()
    for mod, fn in ((ref_lib, "load"), (ref_lib, "load_norm"), (ref_lib, "load_prefill"), (ref_gemm_lib, "load")):
        monkeypatch.setattr(mod, fn, lambda _ref=ref: _ref)
    rec.ref = ref
    return rec


GPU_MODULES = ("test_zz_pending_gpu", "test_decode_attention_gpu", "test_linear_gpu", "test_norm_gpu", "test_ref_attention_gpu",
               "test_decode_step_gpu", "test_linear_fp8_gpu", "test_decode_step_variants_gpu", "test_prefill_attention_gpu",
               "test_engine_gpu", "test_sampling_gpu")          # test_tp_gpu spawns NCCL ranks: its host side is tests/test_tp_gloo.py


def _gpu_tests():
    out = {}
    for mod in GPU_MODULES:
        m = __import__(mod)
        for n, f in sorted(vars(m).items()):
            if n.startswith("test_") and inspect.isfunction(f) and f.__module__ == mod:
                out[f"{mod}::{n}"] = f
    return out


def _first_parameters(fn):
    """The first value of every @pytest.mark.parametrize of the test."""
    kw = {}
    for mark in getattr(fn, "pytestmark", []):
        if mark.name == "parametrize":
            names = [x.strip() for x in mark.args[0].split(",")] if isinstance(mark.args[0], str) else list(mark.args[0])
            first = mark.args[1][0]
            first = getattr(first, "values", first)           # pytest.param(...)
            kw.update(zip(names, first if len(names) > 1 else [first]))
    return kw


@pytest.mark.parametrize("name", sorted(_gpu_tests()))
def test_gpu_test_body_reaches_its_numeric_checks(cpu_as_cuda, lib, monkeypatch, name):
    """Pending AND verified GPU tests: the verified ones are what the round-end run executes with -x, so a host-side
    regression in code they share with the product (runtime, harness) must be caught here, without a GPU."""
    fn = _gpu_tests()[name]
    name = name.split("::")[1]
    kwargs = _first_parameters(fn)
    for p in inspect.signature(fn).parameters:
        if p in kwargs:
            continue
        if p == "lib":
            from lmdeploy_b200 import _lib as L
            kwargs[p] = L.load()                      # the recording stand-in
        elif p == "ref":
            kwargs[p] = cpu_as_cuda.ref               # fixture of test_ref_attention_gpu: the reference library
        elif p == "monkeypatch":
            kwargs[p] = monkeypatch
        else:
            pytest.fail(f"{name}: no dry-run value for parameter {p!r}")
    try:
        fn(**kwargs)
        if name == "test_greedy_tokens_match_a_step_assembled_from_reference_kernels":
            pytest.fail("the stand-ins compute nothing: the L2 test cannot pass")
    except AssertionError:
        pass                                          # the expected end: nothing was computed
    except pytest.skip.Exception:
        pass
    except torch.AcceleratorError:
        pass                                          # reached a genuine device read-back (DecodeModel.logits wraps raw memory)
    except OSError as e:                              # dlopen of an oracle/_ref library needs the CUDA driver
        assert "libcuda" in str(e) or "cannot open shared object" in str(e), e
    calls = cpu_as_cuda.ref.calls
    if name == "test_greedy_tokens_match_a_step_assembled_from_reference_kernels":
        # one full assembled step ran: 2 layers x (4 GEMMs, 1 attention, 2 norms) + the step's first norm
        assert calls.count("ref_linear_forward") >= 8 and calls.count("ref_decode_attention") >= 2
        assert calls.count("ref_bias_residual_rmsnorm") >= 5 and calls.count("ref_linear_create") == 8
    if "matches_reference_gemm" in name:
        assert "ref_linear_forward" in calls
    if name == "test_qk_rmsnorm_is_bit_identical_to_the_reference_kernel":
        assert "ref_qk_rmsnorm" in calls
    if name == "test_residual_rmsnorm_matches_reference_kernel":
        assert "ref_bias_residual_rmsnorm" in calls


def test_graft_entry_smoke_body(cpu_as_cuda, lib):
    """__graft_entry__.smoke() -- the first thing the round-end GPU run executes -- up to its first numeric check."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import __graft_entry__ as G
    with pytest.raises(AssertionError, match="KV cache bytes differ|differs from the oracle"):
        G.smoke()
    assert "b200_decode_attention" in cpu_as_cuda.names()


def test_reference_gemm_bench_tool_body(cpu_as_cuda, lib, capsys):
    """tools/ref_gemm_bench.py (a step of the first GPU call of a round, and bench.py's optional reference-GEMM leg)."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import ref_gemm_bench
    res = ref_gemm_bench.measure(4, 2, [("wo", 2048, 4096, 0), ("w1w3", 2048, 4096, 1)], verbose=True)
    assert set(res) == {"wo", "w1w3"} and set(res["wo"]) >= {"reference_us", "ours_us", "weight_bytes", "max_abs_diff"}
    assert res["wo"]["weight_bytes"] == 2048 * 4096 // 2 + 2048 // 128 * 4096 * 4
    assert "ours/ref speed-up" in capsys.readouterr().out
    names = cpu_as_cuda.names()
    assert "b200_pack_activations" in names and "b200_linear_forward_packed" in names
    assert cpu_as_cuda.ref.calls.count("ref_linear_forward") >= 2 * (1 + 3 + 2 + 1)
