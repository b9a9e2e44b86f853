#!/usr/bin/env python
"""Summarise an ncu report (read here, no GPU): headline raw metrics + instructions / stall samples
per source line.   python profiles/summarise.py gpurun_out/x.ncu-rep [launch-id] > profiles/x.txt"""
import csv
import io
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
       "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
       "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "sm__throughput.avg.pct_of_peak_sustained_elapsed",
       "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
       "lts__throughput.avg.pct_of_peak_sustained_elapsed",
       "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
       "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
       "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum",
       "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
       "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]


def ncu(*a):
    return subprocess.run(["ncu", "-i", *a], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# {rep}: {len(data)} launch(es)")
    for r in data:
        print("##", r[col["Kernel Name"]][:110], "grid", r[col["Grid Size"]], "block", r[col["Block Size"]])
        for m in RAW:
            if m in col:
                print(f"{m:95s} {r[col[m]]:>16s} {units[col[m]]}")
    src = ncu(rep, "--page", "source", "--csv", "--print-source", "sass,cuda")
    lines = src.splitlines()
    # the source page is one table per launch; take the first
    start = next(i for i, l in enumerate(lines) if l.startswith('"Line No"'))
    end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"File Path"')), len(lines))
    tab = list(csv.reader(io.StringIO("\n".join(lines[start:end]))))
    h = tab[0]
    c_line, c_src = h.index("Line No"), h.index("Source")
    c_inst, c_smp = h.index("Instructions Executed"), h.index("# Samples")
    c_wf = h.index("L1 Tag Requests Global") if "L1 Tag Requests Global" in h else None
    def num(x):
        try:
            return int(x)
        except ValueError:
            return 0
    body = [r for r in tab[1:] if len(r) > c_inst and r[c_line] != "" and num(r[c_inst]) > 0]
    tot_i = sum(num(r[c_inst]) for r in body) or 1
    tot_s = sum(num(r[c_smp]) for r in body) or 1
    print(f"\n## per source line (first launch): instructions executed {tot_i}, stall samples {tot_s}")
    print(f"{'line':>5s} {'inst%':>6s} {'smp%':>6s} {'L1 tag req':>11s}  source")
    body.sort(key=lambda r: -num(r[c_inst]))
    for r in body[:45]:
        print(f"{r[c_line]:>5s} {100 * num(r[c_inst]) / tot_i:6.2f} {100 * num(r[c_smp]) / tot_s:6.2f} "
              f"{(r[c_wf] if c_wf is not None else ''):>11s}  {r[c_src].strip()[:100]}")


if __name__ == "__main__":
    main()
