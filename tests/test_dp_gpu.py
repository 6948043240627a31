"""Data-parallel trainer on real GPUs (needs >= 2 B200s; skipped on a 1-GPU box): the gradient exchange through the
C ABI - NVLink peer memory (csrc/mt_p2p.cu) and ncclAllReduce (csrc/mt_comm.cu) - 2 ranks, post-SGD weights and
loss against the oracle."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("peer_memory", [True, False])
def test_two_gpu_dp_matches_oracle(peer_memory):
    from microtorch_b200 import _capi
    if _capi.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, MICROTORCH_B200_P2P="1" if peer_memory else "0")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533" if peer_memory else "29534",
                        os.path.join(REPO, "tests", "dp_parity_worker.py")],
                       capture_output=True, text=True, timeout=600, env=env)
    line = [l for l in r.stdout.splitlines() if l.startswith("DP_PARITY ")]
    assert line, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.loads(line[-1][10:])
    assert res["param_err"] < 1e-4 and res["loss_err"] < 1e-4, res
    assert res["exchanged_bytes"] == 3 * 3 * 256 * 256 * 4      # 3 weight matrices x 3 steps, never the biases
    assert res["peer_memory"] == peer_memory, res               # the path under test is the one that ran
    assert res["replicas_bit_identical"], res
    if peer_memory:
        assert res["p2p_fallbacks"] == 0, res
