"""Sequence-sharded prefill over NCCL against the one-GPU run (BASELINE.json config 4):
`dist.sharded_prefill` on a HybridDecoder (8 layers, 3B mixer dims) at P = 2, 4, 8 ranks, one rank per GPU.
Rank-concatenated output and the last rank's cache must match the single-GPU run: error ratio <= 1e-3 on
activations and states, integers (size / cumulative_length / seq_len) exactly.  Needs >= 2 GPUs: skipped on a
one-GPU box (run by hand with `gpurun --gpus N`; results in profiles/)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-3


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(900)
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_prefill_matches_single_gpu(world):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "dist_gpu_worker.py"), "32768", "131072"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=850, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    lines = [json.loads(l.split(" ", 1)[1]) for l in p.stdout.splitlines() if l.startswith("DIST_PARITY ")]
    assert len(lines) == 2, p.stdout[-2000:]
    for r in lines:
        assert r["ints_equal"], r
        assert r["out"] <= TOL and r["state"] <= TOL and r["kv"] <= TOL and r["conv"] <= TOL, r
