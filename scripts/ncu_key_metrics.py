"""Key counters of every kernel in an `ncu --set full` capture, from its `--page raw --csv` export, as a markdown table.
usage: ncu_key_metrics.py raw.csv > table.md"""
import csv
import sys

COLS = [("ms", "gpu__time_duration.sum", 1),
        ("regs", "launch__registers_per_thread", 1),
        ("smem KB", "launch__shared_mem_per_block_dynamic", 1),
        ("occ %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1),
        ("issue %", "sm__inst_issued.avg.pct_of_peak_sustained_active", 1),
        ("LSU wavefronts %", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", 1),
        ("..shared %", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", 1),
        ("fp64 pipe %", "sm__pipe_fp64_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", 1),
        ("DRAM %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1),
        ("DRAM MB", None, 1),
        ("L2 hit %", "lts__t_sector_hit_rate.pct", 1),
        ("smem conflicts / wavefront", None, 1),
        ("local ld+st wavefronts %", None, 1)]
STALLS = "smsp__average_warps_issue_stalled_"
STALL_END = "_per_issue_active.ratio"


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    scale_of = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Kbyte/block": 1.0}
    ix = {}
    for i, h in enumerate(hdr):
        ix.setdefault(h, i)
        ix.setdefault(h.split(".", 2)[-1] if h.startswith(("SM_", "TPC.", "GPC.")) else h, i)

    def get(r, name, default=0.0):
        """value in ms (times), bytes (sizes) or as reported (everything else)"""
        for k in (name, "TPC.TriageCompute." + name, "SM_A.TriageCompute." + name):
            if k in ix:
                try:
                    return float(r[ix[k]].replace(",", "")) * scale_of.get(units[ix[k]], 1.0)
                except ValueError:
                    return default
        return default
    stall_cols = [(h[len(STALLS):-len(STALL_END)], i) for i, h in enumerate(hdr) if h.startswith(STALLS) and h.endswith(STALL_END)
                  and "selected" not in h]
    print("| kernel | grid | " + " | ".join(c[0] for c in COLS) + " | top stall reasons (warps stalled per issue cycle) |")
    print("|---|---|" + "---|" * (len(COLS) + 1))
    for r in data:
        name = r[ix["Kernel Name"]].replace("void ms_kernel<", "").split(", const")[0].replace("msd::", "").replace("msf::", "")
        if name.startswith("void "):
            name = name[5:].split("(")[0]
        vals = []
        for label, key, scale in COLS:
            if label == "DRAM MB":
                v = (get(r, "dram__bytes_read.sum") + get(r, "dram__bytes_write.sum")) / 1e6
            elif label == "smem conflicts / wavefront":
                v = get(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum") / max(1.0, get(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"))
            elif label == "local ld+st wavefronts %":
                v = get(r, "l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_ld.sum.pct_of_peak_sustained_elapsed") + get(r, "l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_st.sum.pct_of_peak_sustained_elapsed")
            else:
                v = get(r, key) * scale
            vals.append("%.3f" % v if abs(v) < 10 else "%.1f" % v)
        st = sorted(((float(r[i].replace(",", "") or 0), n) for n, i in stall_cols), reverse=True)[:3]
        print("| %s | %s | %s | %s |" % (name, r[ix["Grid Size"]], " | ".join(vals), ", ".join("%s %.2f" % (n, v) for v, n in st)))


if __name__ == "__main__":
    main()
