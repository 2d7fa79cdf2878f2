"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the
header declares, and rejects bad arguments before touching the GPU."""
import ctypes
import os
import re

import pytest

from infinitevl_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    return _lib.load()


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "ivl_b200.h")) as f:
        text = f.read()
    return re.findall(r"IVL_API\s+[\w\s\*]+?\b(ivl_\w+)\s*\(", text)


def test_header_symbols_are_exported_and_bound(lib):
    names = _declared_symbols()
    assert len(names) >= 7 and "ivl_gdn_chunk_fwd" in names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ivl_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in infinitevl_b200/_lib.py"
    assert set(_lib.SIGNATURES) == set(names)


def test_version_and_strerror(lib):
    assert lib.ivl_abi_version() == 3
    assert lib.ivl_strerror(0) == b"ok"
    assert b"shape" in lib.ivl_strerror(-1)
    assert lib.ivl_strerror(-12345) == b"unknown error"


def test_workspace_size_formula(lib):
    # per (b, h, chunk): 57 KiB operand blob (56 KiB of images + gamma) + 8 x 4 KiB U slices + a 4 B ready flag;
    # per (b, h): 8 progress counters; flags + counters rounded to 1 KiB
    for B, T, H in ((1, 64, 16), (1, 65, 16), (2, 1000, 3), (1, 131072, 16)):
        nt = (T + 63) // 64
        n = B * H * nt
        want = n * (57344 + 1024 + 8 * 4096) + ((n + B * H * 8) * 4 + 1023) // 1024 * 1024
        assert lib.ivl_gdn_chunk_workspace_bytes(B, T, H) == want
    assert lib.ivl_gdn_chunk_workspace_bytes(0, 64, 16) == 0


def test_argument_validation_needs_no_gpu(lib):
    buf = ctypes.create_string_buffer(4096)
    p = ctypes.addressof(buf)
    # wrong head dims -> BAD_SHAPE before any CUDA call
    assert lib.ivl_gdn_chunk_fwd(p, p, p, p, p, None, 0, p, None, 0, 1, 64, 16, 64, 256, 0.0, 1, p, 4096, None) == -1
    assert lib.ivl_gdn_recurrent_fwd(p, p, p, p, p, None, 0, p, None, 0, 1, 1, 16, 128, 128, 0.0, 1, None) == -1
    assert lib.ivl_gdn_chunk_fwd(p, p, p, p, p, None, 0, p, None, 0, 1, 0, 16, 128, 256, 0.0, 1, p, 4096, None) == -1
    # NULL tensors
    assert lib.ivl_gdn_chunk_prep(None, p, p, p, p, 1, 64, 16, 0.0, 1, p, 4096, None) == -2
    # workspace too small
    aligned = (p + 1023) // 1024 * 1024
    assert lib.ivl_gdn_chunk_prep(p, p, p, p, p, 1, 64, 16, 0.0, 1, aligned, 1024, None) == -3
    # bad state dtype code
    assert lib.ivl_gdn_recurrent_fwd(p, p, p, p, p, p, 7, p, None, 0, 1, 1, 16, 128, 256, 0.0, 1, None) == -4


def test_ops_refuse_cpu_tensors():
    import torch

    from infinitevl_b200 import ops
    q = torch.zeros(1, 64, 1, 128, dtype=torch.bfloat16)
    v = torch.zeros(1, 64, 1, 256, dtype=torch.bfloat16)
    g = torch.zeros(1, 64, 1)
    with pytest.raises(_lib.IvlError):
        ops.chunk_gated_delta_rule(q, q, v, g, g.bfloat16())


def test_varlen_chunk_tables_host_logic():
    """Chunk geometry of a packed batch (the host side of ivl_gdn_chunk_fwd_varlen): 64-token chunks that never
    straddle a sequence boundary, none for empty sequences -- fla's prepare_chunk_indices
    (fla/ops/gated_delta_rule/chunk.py:211-214) in the layout the C ABI takes."""
    from infinitevl_b200.ops import varlen_chunk_tables
    bounds = [0, 0, 1, 64, 129, 129, 400]
    tok0, valid, begin, n = varlen_chunk_tables(bounds, "cpu")
    assert n == 9
    assert tok0.tolist() == [0, 1, 64, 128, 129, 193, 257, 321, 385]
    assert valid.tolist() == [1, 63, 64, 1, 64, 64, 64, 64, 15]
    assert begin.tolist() == [0, 0, 1, 2, 4, 4, 9]
    assert int(valid.sum()) == bounds[-1]
    assert all(t.dtype.is_floating_point is False and t.element_size() == 4 for t in (tok0, valid, begin))
