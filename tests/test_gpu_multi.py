"""
Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): tools/mgpu_check.py under torchrun -- reads sharded
over ranks, canonical k-mers exchanged (fused NVLink peer-memory routing and the NCCL all-to-all fallback), the union
of the per-rank counted sets compared bit for bit with the oracle; all-pairs distance tiles sharded and added up.
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode,k", [("p2p", 25), ("p2p_reserve", 25), ("nccl", 25), ("p2p", 31)])
def test_two_rank_kmerize_and_allpairs(mode, k):
    """k = 31 is BASELINE.json config[4]'s k (a 1/1000-scale instance of it: same code path, 62-bit keys)"""
    from zotmer_b200 import _native
    if _native.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, ZB_EXCHANGE=mode, ZB_CHECK_K=str(k))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "mgpu_check.py")],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "mgpu_check ok" in r.stdout
