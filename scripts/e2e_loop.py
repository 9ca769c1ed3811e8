import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from audio_suite_b200 import configs, engine
R = 4096
dev = engine.CudaDevice(0)
ir = configs.synth_ir(5.0, 48000, 303)
params = [configs.c5_params(i, shared_ir=ir) for i in range(R)]
host = torch.empty(2 * R * 96000, dtype=torch.float32).pin_memory()
for kw in (dict(chunk=512), dict(chunk=512, ramp=False), dict(chunk=512, workers=12)):
    ts = []
    for rep in range(6):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        engine.render_batch(params, device=dev, host_out=host, **kw)
        torch.cuda.synchronize(); ts.append(1e3 * (time.perf_counter() - t0))
    print(kw, " ".join("%.1f" % t for t in ts))
