import os, sys, time, subprocess
import numpy as np, torch
torch.cuda.init(); dev = torch.device("cuda", 0)
x = torch.zeros(1 << 20, device=dev); torch.cuda.synchronize()
pin = torch.zeros(1 << 20, dtype=torch.uint8).pin_memory()
dst = torch.zeros(1 << 20, dtype=torch.uint8, device=dev)
def T(name, fn, n=3):
    out = []
    for i in range(n):
        t = time.perf_counter(); fn(); out.append(1e3 * (time.perf_counter() - t))
    print("   %-28s %s" % (name, " ".join("%.2f" % v for v in out)), flush=True)
def prims(tag):
    print(tag, flush=True)
    T("launch (x.add_)", lambda: x.add_(1))
    T("sync", lambda: torch.cuda.synchronize())
    T("h2d from pinned async", lambda: dst.copy_(pin, non_blocking=True))
    T("sync", lambda: torch.cuda.synchronize())
    T("pin_memory(100KB)", lambda: torch.zeros(100000, dtype=torch.uint8).pin_memory())
    T("pin_memory(3MB)", lambda: torch.zeros(3000000, dtype=torch.uint8).pin_memory())
    T("empty cuda 1MB", lambda: torch.empty(1 << 20, device=dev))
    T("event record+query", lambda: (lambda e: (e.record(), e.query()))(torch.cuda.Event()))
    T("sync", lambda: torch.cuda.synchronize())
prims("idle")
prims("idle again")
t0 = time.perf_counter()
kind = sys.argv[1]
if kind == "spin":
    code = "import time\nt=time.time()\nwhile time.time()-t<0.4: pass"
elif kind == "sleep":
    code = "import time\ntime.sleep(0.4)"
elif kind == "numpy":
    code = "import time\nimport numpy as np\nt=time.time()\nwhile time.time()-t<0.4: np.random.default_rng(1).standard_normal(100000)"
ps = [subprocess.Popen([sys.executable, "-c", code]) for _ in range(int(sys.argv[2]))]
print("spawned in %.1f ms" % (1e3 * (time.perf_counter() - t0)))
time.sleep(0.2)
prims("children running (%s) t=%.0f" % (kind, 1e3 * (time.perf_counter() - t0)))
print("t=%.0f" % (1e3 * (time.perf_counter() - t0)))
[p.wait() for p in ps]
prims("children gone")
