import os, sys, time, subprocess, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
torch.cuda.init(); dev = torch.device("cuda", 0)
x = torch.zeros(10, device=dev); torch.cuda.synchronize()
mode = sys.argv[1]
def uploads(tag, t0):
    for k in range(6):
        t = time.perf_counter()
        for i in range(20):
            h = torch.from_numpy(np.zeros(100000, np.uint8)).pin_memory().to(dev, non_blocking=True)
        print("   %s 20 uploads: %.1f ms at t=%.1f" % (tag, 1e3 * (time.perf_counter() - t), 1e3 * (time.perf_counter() - t0)), flush=True)
uploads("warm", time.perf_counter())
torch.cuda.synchronize()
if mode == "busyprocs":
    t0 = time.perf_counter()
    NP = int(sys.argv[2]); NICE = int(sys.argv[3])
    ps = [subprocess.Popen([sys.executable, "-c", "import time,os\nos.nice(%d)\nt=time.time()\nwhile time.time()-t<0.3: pass" % NICE]) for _ in range(NP)]
    time.sleep(0.15)     # let them start
    uploads("busy", t0)
    [p.wait() for p in ps]
elif mode == "threads":
    # 16 threads that block on pipe reads from sleeping children
    t0 = time.perf_counter()
    ps = [subprocess.Popen([sys.executable, "-c", "import sys,time\nfor i in range(30):\n  sys.stdout.buffer.write(b'x'*300000); sys.stdout.buffer.flush(); time.sleep(0.005)"], stdout=subprocess.PIPE) for _ in range(16)]
    def serve(p):
        while True:
            b = p.stdout.read(300000)
            if not b: return
    th = [threading.Thread(target=serve, args=(p,)) for p in ps]
    [t.start() for t in th]
    time.sleep(0.1)
    uploads("pipes", t0)
    [t.join() for t in th]
elif mode == "newthreads":
    t0 = time.perf_counter()
    ev = threading.Event()
    th = [threading.Thread(target=ev.wait) for _ in range(16)]
    [t.start() for t in th]
    uploads("after 16 new idle threads", t0)
    ev.set(); [t.join() for t in th]
    uploads("after join", t0)
