"""GPU suite, multi-GPU part: runs only when the box has at least two GPUs (the round-end 1-GPU run skips it; `gpurun --gpus N`
runs it).  tests/dist_parity.py under torchrun, one rank per GPU over NCCL: the sharded ProductPoly and sum-of-products provers
(strided sharding by the last-bound variables, exact all-reduce of the round polynomials, residual gather) bit-identical to the
single-GPU proof and to the CPU oracle, and zk_ntt_sharded (NCCL send/recv exchanges) against the oracle's fft, both ways."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _gpus():
    import torch

    return torch.cuda.device_count()


def _run(world, env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29520 + world), os.path.join(ROOT, "tests", "dist_parity.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=dict(os.environ, **(env or {})))
    lines = [l for l in r.stdout.splitlines() if l.startswith("{") and "dist_parity" in l]
    assert r.returncode == 0 and lines, r.stdout[-3000:] + r.stderr[-3000:]
    return json.loads(lines[-1])


@pytest.mark.parametrize("world", [2, 4, 8])
def test_dist_parity_under_torchrun(world):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    res = _run(world)
    assert res["dist_parity"] is True and res["world"] == world and res["checks"] >= 40, res


def test_dist_parity_with_the_nccl_all_reduce():
    """The fallback of the in-kernel mailbox all-reduce (ZK_B200_MAILBOX=0: ncclAllReduce + narrowing launch): same checks."""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    res = _run(2, {"ZK_B200_MAILBOX": "0"})
    assert res["dist_parity"] is True and res["uses_mailbox"] is False and res["checks"] >= 40, res
