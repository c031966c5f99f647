"""Summarise an .ncu-rep (read here, without a GPU): duration, DRAM bytes, throughput %, pipe utilisation, stalls.

    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep > profiles/rNN_x.txt
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum"]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[head.index("Kernel Name")]
        print(f"kernel: {name[:110]}")
        for i, k in enumerate(head):
            if k in KEYS:
                print(f"  {k:<72} {r[i]:>16} {units[i]}")
        pipes = [(float(r[i]), k) for i, k in enumerate(head)
                 if "pipe_" in k and k.endswith("pct_of_peak_sustained_active") and k not in KEYS and r[i]]
        for v, k in sorted(pipes, reverse=True)[:8]:
            if v >= 1.0:
                print(f"  {k:<72} {v:>16.2f} %")
        stalls = [(float(r[i]), k) for i, k in enumerate(head)
                  if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and r[i]]
        print("  top stall reasons (warps per issue-active cycle):")
        for v, k in sorted(stalls, reverse=True)[:5]:
            print(f"    {k.split('issue_stalled_')[1].replace('_per_issue_active.ratio', ''):<28} {v:.2f}")


if __name__ == "__main__":
    main(sys.argv[1])
