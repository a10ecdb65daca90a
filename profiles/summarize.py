#!/usr/bin/env python
"""Turns an .ncu-rep (ncu --set full) into a small markdown summary that can be committed.
usage: python profiles/summarize.py gpurun_out/prof.ncu-rep "title" > profiles/<name>.md"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_xu.sum", "XU (MUFU) instructions"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe (instructions) %"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1/shared data-pipe wavefronts % of peak"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / scheduler / cycle"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_static", "static smem / block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
]


def main():
    rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# {title}\n\nsource: `{rep}` (ncu --set full --clock-control none, one B200); values per launch.\n")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0]
        print(f"## {name}\n\n| metric | value |\n|---|---|")
        for key, label in WANT:
            if key in hdr:
                i = hdr.index(key)
                print(f"| {label} (`{key}`) | {r[i]} {units[i]} |")
        stalls = []
        for i, k in enumerate(hdr):
            if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and "not_issued" not in k:
                try:
                    stalls.append((float(r[i].replace(",", "")), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        if stalls:
            top = ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)[:6])
            print(f"| warps stalled per issue, top reasons | {top} |")
        i0, i1 = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        try:
            tot = float(r[i0]) + float(r[i1])
            print(f"| DRAM traffic read+write | {tot:.1f} {units[i0]} |")
        except ValueError:
            pass
        print()


if __name__ == "__main__":
    main()
