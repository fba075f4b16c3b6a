"""Multi-GPU plumbing: pixel-row sharding + the one exchange step (an integer all-reduce of the
per-candidate result words) over torch.distributed (NCCL on GPUs, gloo in CPU tests).

One process per GPU.  Each rank uploads only its rows; every rank runs the identical host-side
annealing loop with the identical seed, so after the all-reduce all ranks take the same
accept/reject decisions — the trajectory does not depend on the number of GPUs because the
reduced quantities are exact integers.
"""
from __future__ import annotations

import numpy as np


def row_shard(height: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous block of rows [r0, r1) of rank `rank`: rows r*H/G .. (r+1)*H/G."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world size")
    return height * rank // world_size, height * (rank + 1) // world_size


def row_shard_with_halo(height: int, world_size: int, rank: int, halo: int) -> tuple[int, int, int, int]:
    """(r0, r1, halo_top, halo_bottom): the rank's own rows [r0, r1) and how many neighbour rows to upload with
    them for the S-CIELAB stage (halo = taps // 2; fewer at the global borders)."""
    r0, r1 = row_shard(height, world_size, rank)
    return r0, r1, min(halo, r0), min(halo, height - r1)


def allreduce_words(t):
    """In-place SUM all-reduce of an int64 tensor of result words (no-op when not distributed)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


class _DevWords:
    """Exposes a raw device pointer to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr: int, n_words: int):
        self.__cuda_array_interface__ = {"shape": (n_words,), "typestr": "<i8", "data": (ptr, False), "version": 2}


def install_nccl_allreduce(backend) -> None:
    """Installs the backend's all-reduce hook: sums the device result words over all ranks with
    torch.distributed (NCCL), ordered on the library's own stream."""
    import torch
    import torch.distributed as dist

    if not (dist.is_initialized() and dist.get_world_size() > 1):
        backend.setAllreduce(None)
        return

    def hook(d_ptr: int, n_words: int, stream: int) -> int:
        t = torch.as_tensor(_DevWords(d_ptr, n_words), device=torch.device("cuda", torch.cuda.current_device()))
        ext = torch.cuda.ExternalStream(stream) if stream else torch.cuda.current_stream()
        with torch.cuda.stream(ext):
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return 0

    backend.setAllreduce(hook)


def install_native_nccl(backend) -> dict:
    """The exchange step inside the library: rank 0 asks libhq_b200 for an NCCL unique id, torch.distributed only ships those
    128 bytes to the other ranks, and every rank's context joins the library's OWN communicator (hq_comm_init_rank).  From
    then on hq_eval_palettes* / the search all-reduce the integer result words with ncclAllReduce on the context's stream —
    no Python in the data path.  Returns hq_comm_info."""
    import torch.distributed as dist

    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return backend.commInfo()
    box = [backend.commUniqueId() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(box, src=0)
    backend.setAllreduce(None)
    backend.commInitRank(box[0], dist.get_world_size(), dist.get_rank())
    open_peer_exchange(backend)
    return backend.commInfo()


def open_peer_exchange(backend) -> bool:
    """Small exchanges over NVLink peer memory (hq_comm_open_peers): every rank's mailbox handle (64 bytes) goes to every rank
    through torch.distributed, each rank maps the others' mailboxes with CUDA IPC.  Collective; if ANY rank cannot (no IPC
    between the processes, HQ_PEER_EXCHANGE=0, more than 16 ranks) every rank closes again and the exchange stays on
    ncclAllReduce.  Returns whether the peer path is open."""
    import torch.distributed as dist

    from ._lib import HqError

    world, rank = dist.get_world_size(), dist.get_rank()
    ok, handle = True, b""
    try:
        handle = backend.commPeerHandle()
    except HqError:
        ok = False
    handles = [None] * world
    dist.all_gather_object(handles, handle if ok else None)
    if ok and all(h is not None for h in handles):
        try:
            backend.commOpenPeers(handles, rank)
        except HqError:
            ok = False
    else:
        ok = False
    oks = [None] * world
    dist.all_gather_object(oks, ok)
    if not all(oks):
        backend.commClosePeers()
        return False
    return True


def close_peer_exchange(backend) -> None:
    """Unmaps the other ranks' mailboxes before any rank frees its own (call on every rank before close())."""
    import torch.distributed as dist

    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
        backend.commClosePeers()
        dist.barrier()


def reduce_partials_numpy(parts: list[dict]) -> dict:
    """Host-side statement of what the all-reduce computes (used by the CPU gloo tests)."""
    out = {k: np.zeros_like(parts[0][k]) for k in ("err_fx", "counts", "sums_fx") if parts[0].get(k) is not None}
    for p in parts:
        for k in out:
            out[k] += p[k]
    return out
