"""Summarise `ncu -i X.ncu-rep --page raw --csv` (stdin or file): one block per kernel launch with the metrics DESIGN.md cites."""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__maximum_warps_per_active_cycle_pct"]
rows = list(csv.reader(open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
names, units = rows[hdr], rows[hdr + 1]
col = {n: i for i, n in enumerate(names)}
seen = {}
for r in rows[hdr + 2:]:
    if len(r) < len(names):
        continue
    k = r[col["Kernel Name"]]
    seen[k] = seen.get(k, 0) + 1
    if seen[k] > (int(sys.argv[2]) if len(sys.argv) > 2 else 1):
        continue
    print(f"kernel: {k}  (launch #{seen[k]})")
    rd = wr = None
    for key in KEYS:
        if key in col:
            print(f"  {key:72s} {r[col[key]]} {units[col[key]]}")
            if key == "dram__bytes_read.sum":
                rd = (float(r[col[key]].replace(",", "")), units[col[key]])
            if key == "dram__bytes_write.sum":
                wr = (float(r[col[key]].replace(",", "")), units[col[key]])
    scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
    if rd and wr:
        print(f"  => DRAM traffic per launch: {rd[0] * scale.get(rd[1], 1) + wr[0] * scale.get(wr[1], 1):.1f} MB")
    print()
