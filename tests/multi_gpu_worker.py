"""torchrun worker for tests/test_gpu_multi.py: every rank holds a row shard, the integer result
words are all-reduced by the library's own NCCL communicator (hq_comm_init_rank; the torch hook of
hq_set_allreduce is exercised once beside it), and the totals plus a full sharded SWASA run must
equal the single-GPU result computed on rank 0 over the whole image."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from hybridquantization_b200 import COST_SCIELAB, EVAL_ALLREDUCE, EVAL_PRUNE, PRUNE_OFF, SPACE_SRGB, SWASA, ImageManipulation, synth  # noqa: E402
from hybridquantization_b200.dist import close_peer_exchange, install_native_nccl, install_nccl_allreduce, row_shard, row_shard_with_halo  # noqa: E402


def main():
    out_path = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w, h, K, B = 1031, 517, 64, 5
    img = synth.synth_image(w, h, 31337, smooth=True)
    pal = synth.synth_palettes(B, K)
    r0, r1 = row_shard(h, world, rank)
    be = ImageManipulation("CIE76", False, True, local)
    be.setImage(img[r0:r1])
    install_nccl_allreduce(be)                                # round-1 path: torch.distributed through the C ABI's hook
    got_hook = be.evalPalettes(pal, sums=True)
    info = install_native_nccl(be)                            # the library's own communicator from here on
    got = be.evalPalettes(pal, sums=True)                     # totals over all ranks
    got_pruned = be.evalPalettes(pal, sums=True, flags=EVAL_PRUNE)   # the exact pruned kernel on every shard, same all-reduce
    # a population whose palettes / results exceed the direct host I/O thresholds (64 KB / 32 KB): the DMA-copy path, same hook
    big_pal = synth.synth_palettes(24, 300, seed=9)
    got_big = be.evalPalettes(big_pal, sums=False)
    sw = SWASA(population=4, imax=60, seed=2024)
    best, err, tr, its = be.findBestQuantization(K, sw, n_total=w * h, trace=True)
    # the plugin's default palette size: ONE launch per rank and evaluation, the exchange inside its last CTA (peer memory)
    pal8 = synth.synth_palettes(4, 8, seed=5)
    got8 = be.evalPalettes(pal8, sums=True)
    best8, err8, tr8, its8 = be.findBestQuantization(8, SWASA(population=4, imax=80, seed=99), n_total=w * h, trace=True)
    # the device-pointer entry with HQ_EVAL_ALLREDUCE: K = 16 (the scoring kernel's own tail) and K = 64 (a one-CTA launch behind it)
    dev_ok = True
    for Kd in (16, 64):
        pd = synth.synth_palettes(3, Kd, seed=11)
        want_d = be.evalPalettes(pd, sums=True)
        d_pal = torch.from_numpy(pd).cuda()
        d_res = torch.zeros((3, be.resultWords(Kd, 1)), dtype=torch.int64, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(3):
            be.evalPalettesDevice(d_pal.data_ptr(), 3, Kd, d_res.data_ptr(), 0, 1 | EVAL_ALLREDUCE, st)
        torch.cuda.synchronize()
        hres = d_res.cpu().numpy()
        dev_ok = dev_ok and np.array_equal(hres[:, 0], want_d["err_fx"]) and np.array_equal(hres[:, 1:1 + Kd].astype(np.uint64), want_d["counts"]) \
            and np.array_equal(hres[:, 1 + Kd:].reshape(3, Kd, 3), want_d["sums_fx"])
    res = {"rank": rank, "ok": True, "comm": info, "device_allreduce_equals_host_call": bool(dev_ok)}
    res["hook_equals_native"] = all(np.array_equal(got[k], got_hook[k]) for k in ("err_fx", "counts", "sums_fx"))
    res["pruned_equals_exhaustive"] = all(np.array_equal(got[k], got_pruned[k]) for k in ("err_fx", "counts", "sums_fx"))
    # every rank must hold identical totals / trajectory
    blob = torch.from_numpy(np.concatenate([got["err_fx"], got["counts"].astype(np.int64).ravel(), got["sums_fx"].ravel(),
                                            tr.view(np.int64).ravel(), best.view(np.int32).astype(np.int64).ravel()])).cuda()
    ref = blob.clone()
    dist.broadcast(ref, 0)
    res["same_on_all_ranks"] = bool(torch.equal(blob, ref))
    if rank == 0:
        single = ImageManipulation("CIE76", False, True, local)
        single.setImage(img)
        single.setPruning(PRUNE_OFF)   # the single-GPU reference run scores exhaustively; the sharded run above prunes (AUTO)
        want = single.evalPalettes(pal, sums=True)
        want_big = single.evalPalettes(big_pal, sums=False)
        res["large_population_equal_single_gpu"] = all(np.array_equal(got_big[k], want_big[k]) for k in ("err_fx", "counts"))
        sbest, serr, str_, _ = single.findBestQuantization(K, SWASA(population=4, imax=60, seed=2024), trace=True)
        want8 = single.evalPalettes(pal8, sums=True)
        sbest8, serr8, str8, _ = single.findBestQuantization(8, SWASA(population=4, imax=80, seed=99), trace=True)
        res["small_k_equal_single_gpu"] = bool(all(np.array_equal(got8[k], want8[k]) for k in ("err_fx", "counts", "sums_fx")) and
                                               np.array_equal(tr8.view(np.uint64), str8.view(np.uint64)) and err8 == serr8 and
                                               np.array_equal(best8.view(np.uint32), sbest8.view(np.uint32)))
        single.close()
        res["totals_equal_single_gpu"] = all(np.array_equal(got[k], want[k]) for k in ("err_fx", "counts", "sums_fx"))
        res["trajectory_equal_single_gpu"] = bool(np.array_equal(tr.view(np.uint64), str_.view(np.uint64)) and err == serr and
                                                  np.array_equal(best.view(np.uint32), sbest.view(np.uint32)))
        res["iterations"] = its
    # the same totals with the peer mailboxes closed again: everything on ncclAllReduce
    close_peer_exchange(be)
    got_nccl = be.evalPalettes(pal, sums=True)
    got8_nccl = be.evalPalettes(pal8, sums=True)
    res["peer_equals_nccl"] = bool(all(np.array_equal(got[k], got_nccl[k]) for k in ("err_fx", "counts", "sums_fx")) and
                                   all(np.array_equal(got8[k], got8_nccl[k]) for k in ("err_fx", "counts", "sums_fx")))
    res["peers_closed"] = not be.commInfo()["peer_exchange"]
    be.close()
    # ---- the S-CIELAB stage on row shards with halo rows, all-reduced over NCCL
    r0, r1, top, bot = row_shard_with_halo(h, world, rank, 10)
    sc = ImageManipulation("CIE76", False, True, local)
    sc.setImageSharded(img[r0 - top:r1 + bot], top, bot, r0, h)
    sc.scielabConfigure(72, 45.0)
    install_native_nccl(sc)
    sc_tot = sc.evalPalettesScielab(pal[:2, :32])
    sw2 = SWASA(population=3, imax=25, seed=7, space=SPACE_SRGB, costModel=COST_SCIELAB)
    sbest2, serr2, str2, _ = sc.findBestQuantization(32, sw2, n_total=w * h, trace=True)
    close_peer_exchange(sc)
    sc.close()
    if rank == 0:
        one = ImageManipulation("CIE76", False, True, local)
        one.setImage(img)
        one.scielabConfigure(72, 45.0)
        ref = one.evalPalettesScielab(pal[:2, :32])
        obest2, oerr2, otr2, _ = one.findBestQuantization(32, SWASA(population=3, imax=25, seed=7, space=SPACE_SRGB, costModel=COST_SCIELAB), trace=True)
        one.close()
        res["scielab_totals_equal_single_gpu"] = bool(np.array_equal(sc_tot["err_fx"], ref["err_fx"]) and np.array_equal(sc_tot["counts"], ref["counts"]))
        res["scielab_trajectory_equal_single_gpu"] = bool(np.array_equal(str2.view(np.uint64), otr2.view(np.uint64)) and serr2 == oerr2 and
                                                          np.array_equal(sbest2.view(np.uint32), obest2.view(np.uint32)))
    gathered = [None] * world
    dist.all_gather_object(gathered, res)
    if rank == 0:
        json.dump(gathered, open(out_path, "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
