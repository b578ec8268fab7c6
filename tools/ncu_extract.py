#!/usr/bin/env python
"""Extracts the handful of metrics DESIGN.md / profiles/ quote from an .ncu-rep (ncu --set full).
Usage: python tools/ncu_extract.py gpurun_out/x.ncu-rep profiles/out.csv"""
import csv
import io
import re
import subprocess
import sys

KEEP = re.compile(
    r"^(Kernel Name|Grid Size|Block Size|gpu__time_duration.sum|dram__bytes_(read|write).sum|"
    r"gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|l1tex__m_xbar2l1tex_read_bytes.sum|"
    r"lts__t_sector_hit_rate.pct|lts__throughput.avg.pct_of_peak_sustained_elapsed|"
    r"l1tex__throughput.avg.pct_of_peak_sustained_elapsed|"
    r"sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_(active|elapsed)|"
    r"sm__warps_active.avg.per_cycle_active|launch__registers_per_thread|smsp__inst_executed.sum|"
    r"smsp__issue_active.avg.pct_of_peak_sustained_active|smsp__average_warp_latency_per_inst_issued.ratio|"
    r"smsp__average_warps_issue_stalled_(long_scoreboard|short_scoreboard|wait|branch_resolving|barrier)"
    r"_per_issue_active.ratio|sm__cycles_active.avg|"
    r"l1tex__data_pipe_lsu_wavefronts_mem_shared.sum|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum)$")


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = [i for i, h in enumerate(hdr) if KEEP.match(h)]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx])
    for r in rows[2:]:
        print({hdr[i]: r[i] for i in idx if hdr[i] in ("gpu__time_duration.sum", "Grid Size", "dram__bytes_read.sum",
              "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
              "smsp__average_warp_latency_per_inst_issued.ratio", "l1tex__m_xbar2l1tex_read_bytes.sum")})


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
