"""Host timeline of the streamed end-to-end path (where does the main thread wait?)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from collections import deque
from audio_suite_b200 import configs, engine, tables as T
R = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = engine.CudaDevice(0)
ir = configs.synth_ir(5.0, 48000, 303)
params = [configs.c5_params(i, shared_ir=ir) for i in range(R)]
host = torch.empty(2 * R * 96000, dtype=torch.float32).pin_memory()
W = int(os.environ.get("W", "16")); SI = float(os.environ.get("SI", "0.005"))
sys.setswitchinterval(SI)
engine.render_batch(params, device=dev, host_out=host, chunk=chunk, workers=W)
torch.cuda.synchronize()
print("workers", W, "switchinterval", SI)
for rep in range(2):
    main = torch.cuda.current_stream(); copy = torch.cuda.Stream()
    t0 = time.perf_counter(); rows = []; live = deque(); at = 0
    tw = time.perf_counter()
    for tb in T.plan_stream(params, chunk, workers=W):
        t1 = time.perf_counter()
        br = engine.BatchRenderer(device=dev, tables=tb)
        t2 = time.perf_counter()
        br.run()
        t3 = time.perf_counter()
        ev = torch.cuda.Event(); ev.record(main)
        with torch.cuda.stream(copy):
            copy.wait_event(ev)
            host[at:at + 2 * tb.frames].copy_(br.out[:2 * tb.frames], non_blocking=True)
            dr = torch.cuda.Event(); dr.record(copy)
        at += 2 * tb.frames
        live.append((br, dr))
        t4 = time.perf_counter()
        while len(live) > 3:
            old, e = live.popleft(); e.synchronize(); old.close()
        t5 = time.perf_counter()
        rows.append((t1 - tw, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t5 - t0))
        tw = time.perf_counter()
    while live:
        old, e = live.popleft(); e.synchronize(); old.close()
    torch.cuda.synchronize()
    print("total %.1f ms" % (1e3 * (time.perf_counter() - t0)))
    for r in rows:
        print("  wait_plan %.1f ctor %.1f run %.1f d2h_enq %.1f retire %.1f | t=%.1f" % tuple(1e3 * x for x in r))
