#!/usr/bin/env python
"""Print the judged subset of an `ncu --set full` report (.ncu-rep) as text: one block per captured kernel launch.

Usage: python tools/ncu_summary.py gpurun_out/prof_gemm_r1.ncu-rep > profiles/r1_ncu_gemm_summary.txt
"""
import _path  # noqa: F401  (repository root on sys.path)
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    name_i = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("----")
        print(f"{'Kernel Name':100s}{r[name_i][:120]}")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"{m:83s}{units[i]:17s}{r[i]}")
    print("----")


if __name__ == "__main__":
    main(sys.argv[1])
