#!/usr/bin/env python3
"""Summarise an ncu report (CPU side, no GPU needed) into profiles/:
   python tools/summarize_ncu.py gpurun_out/X.ncu-rep profiles/rNN_name
writes <out>_metrics.txt (per-kernel key metrics), <out>_sass_hist.txt (execution-weighted opcode mix)
and updates profiles/ncu_traffic.json (DRAM bytes per launch, keyed by bench.py's kernel names)."""
import csv
import io
import json
import os
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "sm__cycles_elapsed.max",
    "lts__t_sector_hit_rate.pct", "smsp__sass_inst_executed_op_local_ld.sum",
]
NAMES = {"(int)0": "single", "(int)1": "pass1", "(int)2": "pass2"}


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    lines, traffic = [], {}
    for r in data:
        lines.append(f"== {r[ki]}")
        rd = wr = None
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append(f"   {k:85s} {r[i]:>18s} {units[i]}")
                if k == "dram__bytes_read.sum":
                    rd = float(r[i]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[units[i]]
                if k == "dram__bytes_write.sum":
                    wr = float(r[i]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[units[i]]
        stalls = []
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h:
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v > 0.05:
                    stalls.append((v, h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
        if stalls:
            lines.append("   stall cycles per issued instruction: " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)))
        import re
        m = re.search(r"ntt_tile_kernel<[^,]+, *(?:\(int\))?(\d), *(?:\(bool\))?(\d)", r[ki])
        m3 = re.search(r"ntt3_kernel<[^,]+, *(?:\(int\))?(\d), *(?:\(bool\))?(\d)", r[ki])
        if m3 and rd is not None:
            traffic[("intt3_pass" if m3.group(2) == "1" else "ntt3_pass") + m3.group(1)] = rd + wr
        if m and rd is not None:
            nm = {"0": "single", "1": "pass1", "2": "pass2"}[m.group(1)]
            traffic[("intt_" if m.group(2) == "1" else "ntt_") + nm] = rd + wr
    with open(out + "_metrics.txt", "w") as f:
        f.write(f"# ncu --set full --clock-control none, report {os.path.basename(rep)} (cold-cache, serialised replays)\n")
        f.write("\n".join(lines) + "\n")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    tmp = out + "_sass.csv.tmp"
    open(tmp, "w").write(src)
    hist = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "sass_hist.py"), tmp, "24"], capture_output=True, text=True).stdout
    os.remove(tmp)
    with open(out + "_sass_hist.txt", "w") as f:
        f.write("# execution-weighted SASS opcode mix (ncu source page)\n" + hist)
    tpath = os.path.join(os.path.dirname(out), "ncu_traffic.json")
    cur = json.load(open(tpath)) if os.path.exists(tpath) else {}
    cur.update(traffic)
    cur["_source"] = os.path.basename(rep) + " (dram__bytes_read.sum + dram__bytes_write.sum per launch)"
    json.dump(cur, open(tpath, "w"), indent=1)
    print("wrote", out + "_metrics.txt", out + "_sass_hist.txt", tpath, traffic)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
