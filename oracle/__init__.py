"""CPU oracle for the InfiniteVL hybrid-attention hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker or as the timed CPU baseline -- never as the thing shipped.  The product
(``infinitevl_b200``) fails loudly when its CUDA library is missing; it never
falls back to this code.

What it is: a plain fp32 PyTorch/numpy restatement of the reference's algorithm
for every row of SURVEY.md section 8(a), each function citing the reference
file:line it follows (paths relative to /root/reference; ``fla/`` abbreviates
``src/llamafactory/model/fla/``; ``std`` abbreviates
``infinitevl/infinitevl_standard/modeling_infinitevl.py``).

Parity pinning: the reference ships NO tests and NO golden vectors for this
path.  The oracle is therefore pinned against outputs of
the reference's own Python code imported in the build container
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``): the in-tree
``delta_rule_recurrence`` / ``delta_rule_chunkwise`` (fla/ops/delta_rule/naive.py),
``apply_multimodal_rotary_pos_emb``, ``InfiniteVLRotaryEmbedding``,
``eager_attention_forward``, ``StaticSlidingWindowLayerPrealloc`` and
``StaticLinearLayerPrealloc`` (std), and -- for the gated recurrence whose
arithmetic lives in the un-vendored dependency flash-linear-attention
(requirements.txt:19-20 pins 0.4.0; 0.5.1 is what this image has) -- that
package's own ``naive_recurrent_gated_delta_rule`` /
``naive_chunk_gated_delta_rule``.
"""
from .gdn import (  # noqa: F401
    l2norm_ref, short_conv_ref, gdn_gate_ref, gdn_recurrent_ref, gdn_chunk_ref, gdn_chunk_segmented_ref,
    rmsnorm_gated_ref, rmsnorm_ref, err_ratio,
)
from .swa import (  # noqa: F401
    mrope_cos_sin_ref, mrope_apply_ref, swa_attention_ref, swa_visible_mask,
)
from .cache import (  # noqa: F401
    SlidingWindowCacheRef, LinearCacheRef, swa_mask_sizes_ref,
)
from .block import hybrid_decoder_ref, mlp_ref, swa_mixer_ref  # noqa: F401,E402
from .gdn import gdn_mixer_ref  # noqa: F401,E402
