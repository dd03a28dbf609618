"""Compact per-kernel summary of an ncu report: python scripts/ncu_summary.py rep.ncu-rep > out.md
(needs only `ncu -i`, no GPU)."""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_thr%"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_pipe%"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
    ("sm__cycles_elapsed.max", "cycles"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    groups = OrderedDict()
    for r in data:
        name = r[idx["Kernel Name"]].split("(")[0]
        groups.setdefault(name, []).append(r)
    print(f"# ncu summary of {rep} (per kernel: mean over captured launches)\n")
    for name, rs in groups.items():
        print(f"## {name}  ({len(rs)} launches)")
        for m, short in METRICS:
            if m not in idx:
                continue
            vals = []
            for r in rs:
                try:
                    vals.append(float(r[idx[m]].replace(",", "")))
                except ValueError:
                    pass
            if vals:
                print(f"- {short}: {sum(vals) / len(vals):.6g} {units[idx[m]]}")
        print()


if __name__ == "__main__":
    main()
