"""Import alias for inference_examples/demo_streaming_inference.py:36-44 of the reference, which imports the
cache classes from a module called ``modeling_qwen2_5_vl`` (the file name the hub checkpoint ships its
modeling code under).  Put this directory on ``sys.path`` / ``PYTHONPATH`` and the demo's

    from modeling_qwen2_5_vl import StaticCachePrealloc, StaticSlidingWindowLayerPrealloc, StaticLinearLayerPrealloc

resolves to the B200 implementations (same attributes: ``_buf_keys``, ``_buf_values``, ``keys``, ``values``,
``size``, ``cumulative_length``, ``capacity``, ``recurrent_state``, ``conv_state_{q,k,v}``, ``seq_len``, ``start``,
``is_sliding`` -- what the demo's clone_inference_cache touches, demo:110-160)."""
from infinitevl_b200.cache import (StaticCachePrealloc, StaticLinearLayerPrealloc,  # noqa: F401
                                   StaticSlidingWindowLayerPrealloc)

__all__ = ["StaticCachePrealloc", "StaticSlidingWindowLayerPrealloc", "StaticLinearLayerPrealloc"]
