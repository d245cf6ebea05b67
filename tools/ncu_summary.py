#!/usr/bin/env python
"""ncu -i X.ncu-rep --page raw --csv | python tools/ncu_summary.py label1 label2 ...  -> JSON list of the key metrics per launch.

Labels are attached to the captured launches in order; the output is what profiles/r1_ncu_full_summary.json stores
(bench.py reads the DRAM bytes of the first entry as `roofline.traffic`)."""
import csv, json, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "lts__t_sector_hit_rate.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__shared_mem_per_block_dynamic"]
rows = list(csv.reader(sys.stdin))
hdr, units, data = rows[0], rows[1], rows[2:]
labels = sys.argv[1:]
out = []
for n, r in enumerate(data):
    e = {"launch": labels[n] if n < len(labels) else r[hdr.index("Kernel Name")][:80]}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            e[f"{k} [{units[i]}]"] = r[i]
    out.append(e)
print(json.dumps(out, indent=1))
