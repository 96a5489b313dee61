#!/usr/bin/env python
"""ncu_summary.py -- the metrics DESIGN.md and bench.py quote, out of an `ncu --set full` report:

    python tools/ncu_summary.py gpurun_out/<name>.ncu-rep profiles/<name>_summary.json [kernel-name-substring]

Runs `ncu -i <rep> --page raw --csv` (no GPU needed) and keeps one column of values per captured launch."""
import csv
import io
import json
import subprocess
import sys

KEEP = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
        "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg",
        "smsp__inst_executed.sum", "sm__cycles_active.avg"]

rep, out = sys.argv[1], sys.argv[2]
needle = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
header, units, data = rows[0], rows[1], [r for r in rows[2:] if needle in r[rows[0].index("Kernel Name")]]
summary = {}
for name in KEEP:
    if name in header:
        i = header.index(name)
        summary[name] = {"unit": units[i], "values": [r[i] for r in data]}
# any tensor-pipe metric the report has, whatever this ncu version calls it
for i, name in enumerate(header):
    if ("pipe_tc" in name or "tmem" in name) and name not in summary and len(summary) < 60:
        summary[name] = {"unit": units[i], "values": [r[i] for r in data]}
json.dump(summary, open(out, "w"), indent=1)
print("%d launches, %d metrics -> %s" % (len(data), len(summary), out))
