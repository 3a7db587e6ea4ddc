"""attentionalpoolingaction_b200 — the attentional-pooling head of
rohitgirdhar/AttentionalPoolingAction as hand-written sm_100a CUDA behind a C ABI.

Only what the hot path needs lives here:
  csrc/         CUDA kernels + the C-ABI (include/apa.h)
  custom_ops/   the built library + its ctypes wrapper (mirrors src/custom_ops/ of the reference)
  nets_factory  `get_network_fn(...) -> network_fn(images) -> (logits, end_points)` (nets_factory.py:94-118)
  config        the cfg.NET.* flags the head reads (src/config.py:182-231)
  dist          batch sharding + the single all-gather of the logits
Importing `custom_ops.attn_pool_factory` raises if the library is not built: no CPU fallback.
"""
__version__ = "0.1.0"
