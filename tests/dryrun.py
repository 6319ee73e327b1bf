"""CPU stand-ins for dry runs of the GPU-only host code (tests/test_runtime_dryrun_cpu.py, tests/test_tp_gloo.py)."""
import contextlib
import ctypes as C
import types

import numpy as np
import torch

from lmdeploy_b200 import _lib as L


class RecordingLib:
    """Stands in for libb200tm.so.  Every b200_* symbol the real library exports exists here; a call is checked against
    the declared argtypes (count, and that a pointer parameter never receives a Python float / a scalar parameter never a
    tensor) and recorded."""

    def __init__(self, real):
        self.calls = []
        self._real = real
        self._handles = 0

    def __getattr__(self, name):
        if not name.startswith("b200_"):
            raise AttributeError(name)
        real_fn = getattr(self._real, name)          # AttributeError if the library has no such symbol
        argtypes = real_fn.argtypes

        def call(*args):
            if argtypes is not None:
                assert len(args) == len(argtypes), (name, len(args), len(argtypes))
                for a, t in zip(args, argtypes):
                    if t in (C.c_int, C.c_int64, C.c_size_t):
                        assert isinstance(a, (int, np.integer)) and not isinstance(a, bool) or isinstance(a, bool), (name, a, t)
                    elif t is C.c_float:
                        assert isinstance(a, (int, float)), (name, a, t)
                    elif t is C.c_void_p:
                        assert a is None or isinstance(a, (int, C.c_void_p)) or hasattr(a, "_obj"), (name, type(a))
            self.calls.append((name, args))
            return self._result(name, args)
        return call

    def _result(self, name, args):
        if name == "b200_last_error":
            return b""
        if name == "b200_version":
            return b"recording stand-in sm_100a"
        if name == "b200_linear_prepared_bytes":
            d = args[0]._obj
            per = {0: 2.0, 1: 0.53125, 2: 1.001, 3: 1.03125}[d.weight_format]
            return int(d.input_dim * d.output_dim * per)
        if name.endswith("_bytes"):
            return 4096
        if name == "b200_model_create":
            self._handles += 1
            args[0]._obj.value = 0x1000 * self._handles
            cfg = args[1]._obj
            assert 1 <= cfg.max_batch <= 128 and cfg.rope.type in (0, 1, 2, 4, 5) and cfg.rope.dim > 0
            assert cfg.num_heads % cfg.num_kv_heads == 0 and cfg.head_dim == 128
            return 0
        if name == "b200_linear_load_plans":
            return 0
        return 0

    def names(self):
        return [c[0] for c in self.calls]


def install(setattr_fn, real_lib):
    """Replace the C library by a RecordingLib and the torch.cuda entry points the host code touches by CPU stand-ins.
    setattr_fn(obj, name, value): monkeypatch.setattr in a test, plain setattr in a spawned worker."""
    rec = RecordingLib(real_lib)
    setattr_fn(L, "load", lambda: rec)
    stream = types.SimpleNamespace(cuda_stream=0, synchronize=lambda: None)
    setattr_fn(torch.cuda, "current_stream", lambda *a, **k: stream)
    setattr_fn(torch.cuda, "current_device", lambda: 0)
    setattr_fn(torch.cuda, "set_device", lambda *a, **k: None)
    setattr_fn(torch.cuda, "device", lambda *a, **k: contextlib.nullcontext())
    setattr_fn(torch.cuda, "synchronize", lambda *a, **k: None)
    setattr_fn(torch.cuda, "mem_get_info", lambda *a, **k: (1 << 30, 2 << 30))
    setattr_fn(torch.Tensor, "pin_memory", lambda self, *a, **k: self)

    class Graph:
        def replay(self):
            rec.calls.append(("graph.replay", ()))
    setattr_fn(torch.cuda, "CUDAGraph", Graph)
    setattr_fn(torch.cuda, "graph", lambda g, *a, **k: contextlib.nullcontext())
    return rec
