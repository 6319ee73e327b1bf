"""Host-side mirror of lmdeploy.turbomind for the B200-native decode path."""
from .specs import ModelSpec, SPECS  # noqa: F401
from .runtime import DecodeModel, KvCachePool, kv_layer_bytes  # noqa: F401
