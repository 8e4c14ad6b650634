"""Band-sharded multi-GPU parity (needs >= 2 GPUs; skipped on a single-GPU box): one process per GPU under torchrun,
replicas linked through CUDA IPC; the result must be bit-identical on every rank and identical to the single-GPU result."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        from texture_synthesis_b200 import capi
        return capi.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_gpus() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("args", [["384", "96"], ["384", "96", "tiling"]], ids=["plain", "tiling"])
def test_band_sharded_two_gpus_identical_to_single(args):
    env = dict(os.environ, TSB_MG_MIN_PHASE="1024")   # shard even the small phases of this small case
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "gpu_mg_check.py")] + args
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    text = out.stdout + out.stderr
    assert out.returncode == 0, text[-2000:]
    assert "replicas identical: True" in text and "multi-GPU result identical to single-GPU: True" in text, text[-2000:]
    assert "band-sharded phases 0" not in text
