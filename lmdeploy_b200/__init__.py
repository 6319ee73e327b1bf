"""lmdeploy_b200 -- B200-native (sm_100a) implementation of LMDeploy TurboMind's decode hot path.

All compute lives in lib/libb200tm.so (hand-written CUDA, C ABI in include/b200_turbomind.h); there is no
CPU or PyTorch fallback: importing a compute entry point without the built library raises."""
__version__ = "0.1.0"

from .messages import GenerationConfig, QuantPolicy, Response, TurbomindEngineConfig  # noqa: E402,F401


def pipeline(*args, **kw):
    from .api import pipeline as _p
    return _p(*args, **kw)
