"""torchrun worker for tests/test_gpu_multi_soak.py: the exchange over NVLink peer memory under sustained load — long sharded
searches (one fused launch per rank and iteration at K = 8; scoring kernel + one-CTA exchange launch at K = 64) and thousands of
back-to-back device-pointer evaluations with HQ_EVAL_ALLREDUCE and NO host synchronisation between them (ranks drift apart by up to
one exchange: the two parity sets of the mailboxes are what keeps that safe).  Everything must equal the single-GPU result."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from hybridquantization_b200 import EVAL_ALLREDUCE, PRUNE_OFF, SWASA, ImageManipulation, synth  # noqa: E402
from hybridquantization_b200.dist import close_peer_exchange, install_native_nccl, row_shard  # noqa: E402


def main():
    out_path, iters, launches = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w, h = 512, 509
    img = synth.synth_image(w, h, 4242, smooth=True)
    r0, r1 = row_shard(h, world, rank)
    be = ImageManipulation("CIE76", False, True, local)
    be.setImage(img[r0:r1])
    be.setPruning(PRUNE_OFF)
    info = install_native_nccl(be)
    res = {"rank": rank, "peer_exchange": info["peer_exchange"]}
    runs = {}
    for K in (8, 64):
        runs[K] = be.findBestQuantization(K, SWASA(population=4, imax=iters, seed=1000 + K), n_total=w * h, trace=True)
    # back-to-back launches on one stream, no host synchronisation: alternating palette sizes, every result kept
    st = torch.cuda.Stream()
    torch.cuda.set_stream(st)
    pals = {K: torch.from_numpy(synth.synth_palettes(3, K, seed=K)).cuda() for K in (8, 16, 64)}
    outs = {K: torch.zeros((launches, 3, be.resultWords(K, 0)), dtype=torch.int64, device="cuda") for K in pals}
    torch.cuda.synchronize()
    for i in range(launches):
        for K in pals:
            be.evalPalettesDevice(pals[K].data_ptr(), 3, K, outs[K][i].data_ptr(), 0, EVAL_ALLREDUCE, st.cuda_stream)
    torch.cuda.synchronize()
    same_every_launch = all(bool((outs[K] == outs[K][0:1]).all().item()) for K in pals)
    if rank == 0:
        one = ImageManipulation("CIE76", False, True, local)
        one.setImage(img)
        one.setPruning(PRUNE_OFF)
        ok = True
        for K in (8, 64):
            best, err, tr, its = runs[K]
            sbest, serr, str_, sits = one.findBestQuantization(K, SWASA(population=4, imax=iters, seed=1000 + K), trace=True)
            ok = ok and its == sits == iters and err == serr and np.array_equal(tr.view(np.uint64), str_.view(np.uint64)) \
                and np.array_equal(best.view(np.uint32), sbest.view(np.uint32))
        res["searches_equal_single_gpu"] = bool(ok)
        dev_ok = True
        for K in pals:
            want = one.evalPalettes(pals[K].cpu().numpy())
            got = outs[K][0].cpu().numpy()
            dev_ok = dev_ok and np.array_equal(got[:, 0], want["err_fx"]) and np.array_equal(got[:, 1:1 + K].astype(np.uint64), want["counts"])
        res["device_launches_equal_single_gpu"] = bool(dev_ok)
        one.close()
    res["same_every_launch"] = same_every_launch
    close_peer_exchange(be)
    be.close()
    gathered = [None] * world
    dist.all_gather_object(gathered, res)
    if rank == 0:
        json.dump(gathered, open(out_path, "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
