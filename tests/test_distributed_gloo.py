"""N > 1 host logic on CPU: world_size 2 over gloo (frame sharding, all-reduce of the dense
covariance table, all-gather + merge of the sparse label entries, decode to reference objects)."""
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_combine_over_gloo():
    env = dict(os.environ, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29531",
           os.path.join(REPO, "tests", "_gloo_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600, cwd=REPO)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "GLOO_OK" in out.stdout
