"""The exchange over NVLink peer memory under sustained load (needs >= 2 GPUs: `gpurun --gpus 2`): thousands of exchanges back to
back, with and without host synchronisation between them, against the single-GPU integers (tests/multi_gpu_soak_worker.py).
HQ_SOAK_ITERS / HQ_SOAK_LAUNCHES lengthen it (tools/peer_ab.sh soaks with 20,000 / 5,000)."""
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


@pytest.mark.parametrize("world", [2, 8])
def test_peer_exchange_soak(world, tmp_path, hqlib):
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, have {torch.cuda.device_count()}")
    out = tmp_path / "soak.json"
    iters, launches = os.environ.get("HQ_SOAK_ITERS", "2000"), os.environ.get("HQ_SOAK_LAUNCHES", "600")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(REPO, "tests", "multi_gpu_soak_worker.py"), str(out), iters, launches]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, HQ_PEER_TIMEOUT_MS="20000"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    res = json.load(open(out))
    assert len(res) == world and all(x["peer_exchange"] and x["same_every_launch"] for x in res), res
    assert res[0]["searches_equal_single_gpu"] and res[0]["device_launches_equal_single_gpu"], res
