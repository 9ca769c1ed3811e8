import os, time, multiprocessing as mp
def burn(_):
    t=time.perf_counter(); x=0
    for i in range(3_000_000): x+=i*i
    return time.perf_counter()-t
if __name__=="__main__":
    for f in ("/sys/fs/cgroup/cpu.max","/sys/fs/cgroup/cpu.stat","/sys/fs/cgroup/cpu/cpu.cfs_quota_us","/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
        try: print(f, open(f).read().strip().replace("\n"," | "))
        except Exception as e: print(f, "n/a")
    print("affinity", len(os.sched_getaffinity(0)), "cpu_count", os.cpu_count())
    for n in (1,2,4,8,12,16,24,32):
        with mp.get_context("fork").Pool(n) as p:
            t=time.perf_counter(); r=p.map(burn, range(n)); w=time.perf_counter()-t
        print(n, "procs: wall %.3f mean each %.3f"%(w, sum(r)/n))
    print(open("/sys/fs/cgroup/cpu.stat").read().strip().replace("\n"," | ") if os.path.exists("/sys/fs/cgroup/cpu.stat") else "")
