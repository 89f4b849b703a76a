"""Summarise an .ncu-rep (ncu --set full) into the markdown table kept under profiles/.

    python tools/ncu_summary.py report.ncu-rep "title / what changed" >> profiles/rNN_xxx.md
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
]


def main():
    rep, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_col = hdr.index("Kernel Name")
    names = []
    for r in data:
        n = r[name_col].split("(")[0].replace("void ", "").replace("mdc::", "")
        names.append(n + (f" #{names.count(n) + 1}" if n in names else ""))
    print(f"\n## {rep.split('/')[-1]} - {title}\n")
    print("| metric | unit | " + " | ".join(names) + " |")
    print("|---|---|" + "---|" * len(names))
    for m in METRICS:
        if m not in hdr:
            continue
        i = hdr.index(m)
        print(f"| {m} | {units[i]} | " + " | ".join(r[i] for r in data) + " |")


if __name__ == "__main__":
    main()
