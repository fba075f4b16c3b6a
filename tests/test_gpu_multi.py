"""Row-sharded multi-GPU path on real devices (needs >= 2 GPUs: `gpurun --gpus 2`): the NCCL
all-reduce of the integer partials makes totals and the SWASA trajectory identical to 1 GPU."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_equals_single_gpu(world, tmp_path, hqlib):
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, have {torch.cuda.device_count()}")
    out = tmp_path / "res.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(REPO, "tests", "multi_gpu_worker.py"), str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, HQ_PEER_TIMEOUT_MS="20000"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    res = json.load(open(out))
    assert len(res) == world and all(x["same_on_all_ranks"] and x["pruned_equals_exhaustive"] and x["hook_equals_native"] for x in res)
    assert all(x["comm"]["size"] == world and x["comm"]["rank"] == x["rank"] and x["comm"]["nccl_version"] > 20000 for x in res)
    assert res[0]["totals_equal_single_gpu"] and res[0]["trajectory_equal_single_gpu"] and res[0]["iterations"] == 60
    assert res[0]["large_population_equal_single_gpu"]   # both host I/O paths of hq_eval_palettes go through the all-reduce
    assert res[0]["scielab_totals_equal_single_gpu"] and res[0]["scielab_trajectory_equal_single_gpu"]
    # the exchange over NVLink peer memory (CUDA IPC mailboxes between the ranks): open on every rank, the plugin's default
    # palette size through the one-launch evaluation, the device-pointer entry with HQ_EVAL_ALLREDUCE, and the same totals
    # once the mailboxes are closed again and everything runs on ncclAllReduce
    assert all(x["comm"]["peer_exchange"] for x in res), "hq_comm_open_peers did not open the peer path"
    assert res[0]["small_k_equal_single_gpu"]
    assert all(x["device_allreduce_equals_host_call"] and x["peer_equals_nccl"] and x["peers_closed"] for x in res)
