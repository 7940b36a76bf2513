"""Summarise an .ncu-rep (ncu --set full) into markdown: per captured launch the numbers DESIGN.md / bench.py cite.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_kernel.md"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (elapsed)"),
    ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "HMMA (mma.sync) pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__occupancy_limit_registers", "CTA/SM limit (registers)"),
    ("launch__occupancy_limit_shared_mem", "CTA/SM limit (smem)"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu summary of `{path}`\n")
    print("Captured with `ncu --set full --clock-control none --import-source on` (per-launch values are cold-cache and "
          "serialised: compare shares, not absolutes).\n")
    for n, row in enumerate(rows[2:]):
        print(f"## launch {n}: `{row[ix['Kernel Name']]}`\n")
        print("| metric | value |")
        print("|---|---|")
        for k, label in KEYS:
            if k in ix:
                print(f"| {label} | {row[ix[k]]} {units[ix[k]]} |")
        stalls = []
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    v = float(row[ix[h]])
                except ValueError:
                    continue
                if v >= 0.2:
                    stalls.append((v, h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
        stalls.sort(reverse=True)
        print("| warp stall reasons (per issue) | " + ", ".join(f"{nme} {v:.2f}" for v, nme in stalls) + " |")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
