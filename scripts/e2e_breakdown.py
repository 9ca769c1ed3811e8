"""Where the end-to-end time of the C5 sweep goes (host planning / table upload / kernels / D2H)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from audio_suite_b200 import configs, engine
R = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = engine.CudaDevice(0)
ir = configs.synth_ir(5.0, 48000, 303)
params = [configs.c5_params(i, shared_ir=ir) for i in range(R)]
host = torch.empty(2 * R * 96000, dtype=torch.float32).pin_memory()
print("cores", len(os.sched_getaffinity(0)))
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    br = engine.BatchRenderer(params, device=dev)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    br.run(); torch.cuda.synchronize(); t3 = time.perf_counter()
    host.copy_(br.outputs_device(), non_blocking=True); torch.cuda.synchronize(); t4 = time.perf_counter()
    print(f"rep{rep}: plan {br.t_plan*1e3:.1f} pack/upload/create {br.t_pack*1e3:.1f} (ctor {1e3*(t1-t0):.1f}, drain {1e3*(t2-t1):.1f}) run {1e3*(t3-t2):.1f} d2h {1e3*(t4-t3):.1f} ms  ({host.numel()*4/(t4-t3)/1e9:.1f} GB/s)")
    br.close(); del br
