"""GPU tests (-m gpu) over a REAL NCCL world: one process per GPU launched with torch.distributed.run, exactly as the
driver launches bench.py.  Skipped on a box with fewer than 2 GPUs (the gloo world-size-2 tests in test_sharded.py cover the
host logic on CPU; on one GPU test_cuda_shards_assemble_to_reference computes every shard with the CUDA engine)."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch

    return torch.cuda.device_count()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _torchrun(script, world, env=None, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, script)]
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=e, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert lines, r.stdout[-2000:]
    return json.loads(lines[-1])


@pytest.mark.parametrize("world", [2, 4, 8])
def test_nccl_sharded_commit_all_exchange_modes(world):
    """from_coeffs (all-gather one-shot / streaming, peer loads, all-to-all of LDE rows, blinding, more ranks than coset
    blocks), from_values and pinned-host streaming: cap, rows, paths, digests == CPU oracle on every rank"""
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    out = _torchrun("tests/harness/nccl_commit_check.py", world)
    assert out["world"] == world and out["all_ok"], out


@pytest.mark.parametrize("world", [2, 8])
def test_nccl_opening_proof_over_sharded_commitments(world):
    """prove_openings over 4 sharded oracles: the same proof on all ranks, equal to the CPU oracle's field by field, accepted by
    the restated verifier"""
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    out = _torchrun("tests/harness/sharded_opening_check.py", world, env={"CHECK_LG_D": "12"})
    assert out["same_proof_on_all_ranks"] and out["caps_equal"] and out["openings_equal"]
    assert out["proof_equal_to_cpu_oracle"] and out["verifier_accepts"]
