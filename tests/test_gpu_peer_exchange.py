"""Multi-GPU (needs >= 2 GPUs on the box; skipped otherwise): the fused BRDF iteration with its exchange steps over PEER MEMORY
(in-kernel mailboxes for the three scalar sums, direct NVLink stores for the two halo exchanges; materialist_b200/parallel.py
PeerArena, csrc/mb200_optim.cu) — two ranks, 3 iterations each of mesh mode, G-buffer fused / direct and pos_mlp — equals the
single-GPU optimisation (tools/check_shard_equivalence.py, launched under torchrun as the bench is)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("peer", ["1", "0"])
def test_two_rank_optimisations_equal_single_gpu(peer):
    env = dict(os.environ, MB200_PEER=peer)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(ROOT, "tools", "check_shard_equivalence.py")],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "OK: 2-rank" in r.stdout
    assert ("exchanged over peer memory" in r.stdout) == (peer == "1")
