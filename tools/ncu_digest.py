"""Digest an .ncu-rep: headline metrics + per-region stall samples.  usage: ncu_digest.py file.ncu-rep [kernel-regex]"""
import csv, io, subprocess, sys

def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h = rows[0]
    return [dict(zip(h, r)) for r in rows[2:]]

KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_pipe_tmem.sum",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum"]

def main():
    path = sys.argv[1]
    for d in raw(path):
        for k in KEYS:
            if k in d:
                print(f"  {k:75s} {d[k]}")
        st = {k: float(v) for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio")}
        tot = sum(st.values())
        print("  stalls per issue:", ", ".join(f"{k.split('stalled_')[1].split('_per')[0]} {v:.2f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:9]), f"(sum {tot:.2f})")
        print()

if __name__ == "__main__":
    main()
