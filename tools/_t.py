import sys, time, numpy as np
sys.path.insert(0, "/root/repo")
from hybridquantization_b200 import ImageManipulation, synth, EVAL_PRUNE, SWASA, PRUNE_AUTO, PRUNE_OFF
be = ImageManipulation("CIE76", False, True, 0)
for (w,h,smooth,P,imax) in [(1920,1080,True,4,1000),(3840,2160,False,64,100)]:
    img = synth.synth_image(w,h,synth.SEED_BASE+2,smooth)
    be.setImage(img)
    for rep in range(5):
        be.setPruning(PRUNE_AUTO)
        t0=time.perf_counter(); best, err, _, its = be.findBestQuantization(256, SWASA(population=P, imax=imax, seed=77760)); dt=time.perf_counter()-t0
        pal = synth.synth_palettes(P, 256)
        be.setProfiling(True)
        ms=[]
        for i in range(5):
            be.evalPalettes(pal, flags=EVAL_PRUNE); ms.append(be.lastAssignMs())
        be.setProfiling(False)
        print(w,h,smooth,"search %.3f s"%dt, "kernel ms (random palettes):", [round(m,3) for m in ms])
