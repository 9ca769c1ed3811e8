import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from audio_suite_b200 import configs, engine, tables as T
dev = engine.CudaDevice(0)
ir = configs.synth_ir(5.0, 48000, 303)
params = [configs.c5_params(i, shared_ir=ir) for i in range(4096)]
brs = [engine.BatchRenderer(device=dev, tables=tb) for tb in T.plan_stream(params, 512)]
torch.cuda.synchronize(); [b.close() for b in brs]; del brs
mode = sys.argv[1]
if len(sys.argv) > 2 and sys.argv[2] == "mallopt":
    import ctypes
    libc = ctypes.CDLL("libc.so.6")
    print("mallopt", libc.mallopt(-3, 1 << 30), libc.mallopt(-1, 1 << 30))     # M_MMAP_THRESHOLD, M_TRIM_THRESHOLD
for rep in range(2):
    t0 = time.perf_counter()
    gen = T.plan_stream(params, 512)
    tb = next(gen)
    t1 = time.perf_counter()
    if mode == "sleep":
        time.sleep(0.2)
    elif mode == "pyloop":
        for k in range(8):
            t = time.perf_counter(); x = 0
            for i in range(100000): x += i
            print("   pyloop 100k iters: %.1f ms at t=%.1f" % (1e3 * (time.perf_counter() - t), 1e3 * (time.perf_counter() - t0)))
    elif mode == "torchloop":
        for k in range(8):
            t = time.perf_counter()
            for i in range(20):
                h = torch.from_numpy(np.zeros(100000, np.uint8)).pin_memory().to(dev.dev, non_blocking=True)
            print("   20 uploads: %.1f ms at t=%.1f" % (1e3 * (time.perf_counter() - t), 1e3 * (time.perf_counter() - t0)))
    elif mode == "nopin":
        for k in range(8):
            t = time.perf_counter()
            for i in range(20):
                h = torch.from_numpy(np.zeros(100000, np.uint8)).to(dev.dev, non_blocking=True)
            print("   20 pageable uploads: %.1f ms at t=%.1f" % (1e3 * (time.perf_counter() - t), 1e3 * (time.perf_counter() - t0)))
    t2 = time.perf_counter()
    os.environ["MS_TRACE"] = "1"
    br = engine.BatchRenderer(device=dev, tables=tb)
    t3 = time.perf_counter()
    print("mode %s: first yield %.1f ms, ctor %.1f ms (started at %.1f)" % (mode, 1e3 * (t1 - t0), 1e3 * (t3 - t2), 1e3 * (t2 - t0)))
    rest = list(gen)
    print("   all planned at %.1f" % (1e3 * (time.perf_counter() - t0)))
    br.close(); torch.cuda.synchronize()
