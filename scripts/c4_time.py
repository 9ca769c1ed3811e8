"""C4 (long-form render, BASELINE.json configs[3]) on one GPU: kernels-only and end-to-end times."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from audio_suite_b200 import configs, engine
dev = engine.CudaDevice(0)
p = configs.canonical("C4")
t0 = time.perf_counter(); br = engine.BatchRenderer([p], device=dev); torch.cuda.synchronize(); t1 = time.perf_counter()
print("plan+upload %.1f ms (plan %.1f)" % (1e3 * (t1 - t0), 1e3 * br.t_plan))
for rep in range(3):
    names, evs = [], [torch.cuda.Event(enable_timing=True)]
    evs[0].record()
    def mark(n):
        e = torch.cuda.Event(enable_timing=True); e.record(); evs.append(e); names.append(n)
    br.run(mark); torch.cuda.synchronize()
    print("run %.1f ms:" % evs[0].elapsed_time(evs[-1]), {n: round(a.elapsed_time(b), 2) for n, a, b in zip(names, evs[:-1], evs[1:])})
host = torch.empty(2 * 57_600_000, dtype=torch.float32).pin_memory()
t0 = time.perf_counter(); host.copy_(br.outputs_device()); torch.cuda.synchronize(); print("d2h %.1f ms" % (1e3 * (time.perf_counter() - t0)))
print("alg", br.tables.alg)
br.close()
for rep in range(2):
    t0 = time.perf_counter(); out, meta = engine.render(p, device=dev); print("engine.render e2e %.1f ms" % (1e3 * (time.perf_counter() - t0)), out.shape, out.dtype)
