"""profiles/ncu_traffic.json from an `ncu --set full` capture: DRAM bytes (read + write) per launch of the render kernels, keyed by the
stage names bench.py reports (its `roofline.traffic` is looked up here).  usage: ncu -i X.ncu-rep --page raw --csv > raw.csv;
python scripts/ncu_traffic.py raw.csv <label of the capture>"""
import csv
import json
import os
import sys

STAGE = {"blend_backward_ring_kernel": "blend_backward_kernel", "blend_backward_kernel": "blend_backward_kernel",
         "blend_forward_ring_kernel": "blend_forward_kernel", "blend_forward_kernel": "blend_forward_kernel",
         "preprocess_raw_kernel": "preprocess_kernel", "preprocess_kernel": "preprocess_kernel",
         "preprocess_backward_raw_kernel": "preprocess_backward_kernel", "preprocess_backward_compact_kernel": "preprocess_backward_kernel(compact)",
         "kback_zero_flag_kernel": "preprocess_backward_kernel(flag)", "tile_scatter_kernel": "tile_scatter_kernel",
         "tile_count_kernel": "tile_count_kernel"}
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
names, units = rows[hdr], rows[hdr + 1]
col = {n: i for i, n in enumerate(names)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
acc = {}
for r in rows[hdr + 2:]:
    if len(r) < len(names):
        continue
    import re
    k = re.sub(r"^void\s+", "", r[col["Kernel Name"]]).replace("<unnamed>::", "")
    k = re.split(r"[<(]", k)[0].strip()
    st = STAGE.get(k)
    if st is None:
        continue
    b = sum(float(r[col[m]].replace(",", "")) * scale.get(units[col[m]], 1.0) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    acc.setdefault(st, []).append(b)
out = {"source": sys.argv[2] if len(sys.argv) > 2 else sys.argv[1],
       "dram_bytes_per_launch": {k: sum(v) / len(v) for k, v in sorted(acc.items())}}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
