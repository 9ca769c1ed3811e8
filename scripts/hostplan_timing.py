"""Host-side timing of the native planner on this box: threads x slice size, marshalling, and the set-up of one slice."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_suite_b200 import configs, hostplan
ps = [configs.c5_params(i) for i in range(1024)]
l = hostplan.lib()
print("cpus", len(os.sched_getaffinity(0)))
os.system("lscpu | egrep 'Model name|Thread|Core|Socket|MHz' | head -8")
hostplan.plan_chunk(ps[:64])
for n in (128, 512):
    t = time.perf_counter(); r = hostplan._marshal(ps[:n]); print("marshal", n, round((time.perf_counter() - t) * 1e3, 2), "ms")
    rows, lane_ptr, lane_t, lane_v, irs, ir_len, btab = r
    for thr in (1, 2, 4, 8):
        best = 1e9
        for k in range(5):
            t = time.perf_counter()
            h = l.ms_hp_plan(rows.ctypes.data, n, lane_ptr.ctypes.data, lane_t.ctypes.data, lane_v.ctypes.data, ir_len.ctypes.data, btab.ctypes.data, int(btab.shape[1]), thr)
            best = min(best, time.perf_counter() - t); l.ms_hp_free(h)
        print("native n=%d threads=%d: %.2f ms" % (n, thr, best * 1e3))
    for thr in (1, 4):
        best = 1e9
        for k in range(5):
            t = time.perf_counter(); hostplan.plan_chunk(ps[:n], thr); best = min(best, time.perf_counter() - t)
        print("plan_chunk n=%d threads=%d: %.2f ms" % (n, thr, best * 1e3))
