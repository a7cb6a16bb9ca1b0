"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/ (tracked).

  python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r01_launches.txt
  python tools/ncu_summary.py full gpurun_out/prof.ncu-rep profiles/r01_step_kernel_full.txt
"""
import collections
import csv
import subprocess
import sys

FULL_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "lts__t_sector_hit_rate.pct", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_fma.sum",
    "smsp__inst_executed_pipe_lsu.sum", "smsp__inst_executed_pipe_xu.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def launches(src, dst):
    rows = list(csv.reader(open(src, errors="replace")))
    hdr, agg, order = None, collections.defaultdict(lambda: [0, 0.0]), []
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get("Metric Name") == "gpu__time_duration.sum":
                v = float(d["Metric Value"].replace(",", ""))
                v = v / 1000 if d["Metric Unit"] == "ns" else v
                agg[d["Kernel Name"]][0] += 1
                agg[d["Kernel Name"]][1] += v
    tot = sum(t for _, t in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none  ({src}); cold-cache, serialised launches\n")
        f.write(f"# total {tot:.1f} us over {sum(c for c, _ in agg.values())} launches\n")
        f.write(f"{'share':>7} {'total_us':>11} {'launches':>8} {'us/launch':>10}  kernel\n")
        for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{100 * t / tot:6.1f}% {t:11.1f} {c:8d} {t / c:10.2f}  {k[:110]}\n")


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on  ({src}); one column per captured launch\n")
        ki = hdr.index("Kernel Name")
        f.write("kernel: " + " | ".join(r[ki][:70] for r in rows[2:]) + "\n")
        for m in FULL_METRICS:
            if m in hdr:
                i = hdr.index(m)
                f.write(f"{m:92s} [{rows[1][i]}] " + "  ".join(r[i] for r in rows[2:]) + "\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
