"""The N>1 host logic on CPU: world_size-2 gloo processes shard the rows, each computes its
partial integers (the oracle stands in for the kernel — it is only the checker's arithmetic
here), the int64 result words are all-reduced exactly as on NCCL, and every rank must hold the
whole-image result.  Also: both ranks take identical SWASA decisions from the reduced costs."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, h, w, K, B, out_q):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hybridquantization_b200 import synth
    from hybridquantization_b200.dist import allreduce_words, row_shard
    from oracle import hq_oracle as O

    img = synth.synth_image(w, h, 4242, smooth=True)
    pal = synth.synth_palettes(B, K)
    r0, r1 = row_shard(h, world, rank)
    part = O.assign_reduce(img[r0:r1], pal, threads=1)
    words = np.concatenate([part["err_fx"][:, None], part["counts"].astype(np.int64), part["sums_fx"].reshape(B, -1)], axis=1)
    t = torch.from_numpy(words.copy())
    allreduce_words(t)
    # identical accept/reject decisions on every rank: costs from the reduced integers
    costs = [O.cost(int(t[b, 0]), t[b, 1:1 + K].numpy().astype(np.uint64), h * w, 2.0) for b in range(B)]
    out_q.put((rank, t.numpy().copy(), costs))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("h", [32, 33])
def test_two_rank_row_sharding_allreduce(oracle, h):
    from hybridquantization_b200 import synth

    w, K, B = 40, 12, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, h, w, K, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    img = synth.synth_image(w, h, 4242, smooth=True)
    whole = oracle.assign_reduce(img, synth.synth_palettes(B, K), threads=1)
    want = np.concatenate([whole["err_fx"][:, None], whole["counts"].astype(np.int64), whole["sums_fx"].reshape(B, -1)], axis=1)
    for rank, words, costs in got:
        assert np.array_equal(words, want), rank
        assert costs == [oracle.cost(int(whole["err_fx"][b]), whole["counts"][b], h * w, 2.0) for b in range(B)]
    assert got[0][2] == got[1][2]


def test_row_shard_partition():
    from hybridquantization_b200.dist import row_shard

    for H in (1, 7, 2160, 8192):
        for G in (1, 2, 4, 8):
            cuts = [row_shard(H, G, r) for r in range(G)]
            assert cuts[0][0] == 0 and cuts[-1][1] == H
            assert all(a[1] == b[0] for a, b in zip(cuts[:-1], cuts[1:]))
    with pytest.raises(ValueError):
        row_shard(10, 2, 2)


class _FakeBackend:
    """Stands in for ImageManipulation in the handshake of dist.open_peer_exchange (the mailboxes themselves need GPUs:
    tests/test_gpu_multi*.py): records what the helper asks of it and fails where told to."""

    def __init__(self, rank, fail_handle=False, fail_open=False):
        self.rank, self.fail_handle, self.fail_open = rank, fail_handle, fail_open
        self.opened, self.closed = None, 0

    def commPeerHandle(self):
        from hybridquantization_b200._lib import HqError
        if self.fail_handle:
            raise HqError(-4, "no mailbox")
        return bytes([self.rank]) * 64

    def commOpenPeers(self, handles, rank):
        from hybridquantization_b200._lib import HqError
        if self.fail_open:
            raise HqError(-4, "cudaIpcOpenMemHandle failed")
        self.opened = (list(handles), rank)

    def commClosePeers(self):
        self.closed += 1


def _peer_worker(rank, world, port, scenario, out_q):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hybridquantization_b200.dist import open_peer_exchange
    be = _FakeBackend(rank, fail_handle=(scenario == "handle" and rank == 1), fail_open=(scenario == "open" and rank == 0))
    ok = open_peer_exchange(be)
    out_q.put((rank, ok, be.opened, be.closed))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("scenario", ["ok", "handle", "open"])
def test_peer_exchange_handshake_is_all_or_nothing(scenario):
    """Every rank's 64 handle bytes reach every rank in rank order; if ANY rank cannot export or map a mailbox, EVERY rank closes
    again and reports False, so that no rank waits in a mailbox the others never write (the exchange then stays on NCCL)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, scenario, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    if scenario == "ok":
        for rank, ok, opened, closed in got:
            assert ok and closed == 0 and opened == ([bytes([0]) * 64, bytes([1]) * 64], rank)
    else:
        assert all(not ok and closed == 1 for _, ok, _, closed in got), got
