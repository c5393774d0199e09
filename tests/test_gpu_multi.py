"""A sharded run on hardware equals the single-GPU run bit for bit: two ranks launched with
torch.distributed.run -- NCCL on two GPUs when the box has them, otherwise both ranks on GPU 0 with gloo carrying the
(host-staged) all-gather -- against the same job in one process."""
import os
import subprocess
import sys

import pytest
import torch

import scene_util as su

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B", [6, 5])  # even shards (3 + 3) and ragged ones (3 + 2)
def test_two_rank_run_equals_single_gpu_bitwise(tmp_path, B):
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    worker = os.path.join(su.ROOT, "tests", "multi_gpu_worker.py")
    port = 29600 + (os.getpid() + 7 * B) % 300
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT")}
    one = subprocess.run([sys.executable, worker, str(tmp_path), str(B), backend], capture_output=True, text=True, env=env, timeout=600)
    assert one.returncode == 0, one.stderr[-3000:]
    two = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", str(port), worker, str(tmp_path), str(B), backend], capture_output=True, text=True, env=env, timeout=900)
    assert two.returncode == 0, two.stderr[-3000:]
    ref = torch.load(os.path.join(tmp_path, "rank0_of1.pt"))
    for r in range(2):
        got = torch.load(os.path.join(tmp_path, "rank%d_of2.pt" % r))
        assert torch.equal(got["lr"], ref["lr"]), "rank 0's multipliers are the job's on every rank"
        assert torch.equal(got["poses"], ref["poses"]), "pose history of every hypothesis, bitwise (%s)" % backend
        assert set(got["losses"]) == set(ref["losses"]) == {"rgb", "depth", "mask_selection"}
        for k in ref["losses"]:
            assert torch.equal(got["losses"][k], ref["losses"][k]), k
        assert got["argmin"] == ref["argmin"] and torch.equal(got["pose"], ref["pose"]) and torch.equal(got["final"], ref["final"])
