"""Single-process multi-device context (hq_create_multi, native NCCL inside libhq_b200.so) driven by a plain C
program — no Python, no torch in the data path (needs >= 2 GPUs: `gpurun --gpus 2`).  Everything the N-device
context returns must be bit-identical to a 1-device context (tests/cpp/multi_c_test.c lists what is compared)."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "multi_c_test")
    libdir = os.path.join(REPO, "hybridquantization_b200")
    subprocess.run(["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(REPO, "include"), "-o", exe, os.path.join(REPO, "tests", "cpp", "multi_c_test.c"),
                    "-L", libdir, "-lhq_b200", f"-Wl,-rpath,{libdir}"], check=True)
    return exe


@pytest.mark.parametrize("ndev", [2, 4, 8])
def test_c_host_on_n_devices_equals_one_device(ndev, hqlib, tmp_path):
    import torch

    if torch.cuda.device_count() < ndev:
        pytest.skip(f"needs {ndev} GPUs, have {torch.cuda.device_count()}")
    exe = _build(tmp_path)
    # the exchange of small payloads over NVLink peer memory (default: inside the scoring kernel's last CTA for K <= 32, a
    # one-CTA launch otherwise) and, with HQ_PEER_EXCHANGE=0, everything on ncclAllReduce: the same integers either way
    for peers in ("1", "0"):
        env = dict(os.environ, HQ_PEER_EXCHANGE=peers, HQ_PEER_TIMEOUT_MS="20000")
        for args in (["1031", "517", "64", "5"],      # ragged rows, pruning on (AUTO) inside the search
                     ["640", "37", "16", "3"],        # fewer than 10 rows per device at 4/8 devices: halos overlap several neighbours
                     ["512", "512", "8", "4"]):       # the plugin's defaults: one launch per device and iteration, exchange included
            r = subprocess.run([exe, str(ndev)] + args, capture_output=True, text=True, timeout=600, env=env)
            assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
            assert "MULTI_C_TEST OK" in r.stdout and f"comm_size={ndev}" in r.stdout and f"peers={peers}" in r.stdout, r.stdout


def test_single_device_list_is_a_plain_context(hqlib, tmp_path):
    """hq_create_multi with one device needs no NCCL and behaves like hq_create (the same program, self-compared)."""
    exe = _build(tmp_path)
    r = subprocess.run([exe, "1", "320", "200", "16", "3"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "MULTI_C_TEST OK ndev=1" in r.stdout
