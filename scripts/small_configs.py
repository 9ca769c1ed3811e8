"""Latency of single renders C1b/C1/C2/C3 through render(): GPU path vs the numpy oracle on the host."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from audio_suite_b200 import configs, engine
from oracle import microsound_np as O
dev = engine.CudaDevice(0)
for name in ("C1b", "C1", "C2", "C3"):
    p = configs.canonical(name)
    for _ in range(3):
        engine.render(p, device=dev)
    ts = []
    for _ in range(10):
        t = time.perf_counter(); out, _ = engine.render(p, device=dev); ts.append(time.perf_counter() - t)
    tc = []
    for _ in range(3):
        t = time.perf_counter(); ref, _ = O.render(p); tc.append(time.perf_counter() - t)
    br = engine.BatchRenderer([p], device=dev)
    br.run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        br.run()
    e1.record(); torch.cuda.synchronize()
    print("%s: render() %.2f ms (min %.2f) | kernels only %.3f ms | numpy oracle %.2f ms | max-abs %.2e" % (
        name, 1e3 * np.median(ts), 1e3 * min(ts), e0.elapsed_time(e1) / 20, 1e3 * min(tc), np.max(np.abs(out - ref))))
    br.close()
