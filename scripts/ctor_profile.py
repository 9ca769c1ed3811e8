import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from audio_suite_b200 import configs, engine, tables as T
dev = engine.CudaDevice(0)
ir = configs.synth_ir(5.0, 48000, 303)
params = [configs.c5_params(i, shared_ir=ir) for i in range(4096)]
host = torch.empty(2 * 4096 * 96000, dtype=torch.float32).pin_memory()
engine.render_batch(params, device=dev, host_out=host)
torch.cuda.synchronize()
for rep in range(2):
    k = 0
    keep = []
    for tb in T.plan_stream(params, 512):
        if k in (0, 3):
            pr = cProfile.Profile(); pr.enable()
            br = engine.BatchRenderer(device=dev, tables=tb)
            pr.disable()
            print("==== rep", rep, "chunk", k)
            pstats.Stats(pr).sort_stats('tottime').print_stats(12)
        else:
            br = engine.BatchRenderer(device=dev, tables=tb)
        br.run(); keep.append(br); k += 1
    torch.cuda.synchronize()
    for b in keep: b.close()
    del keep
