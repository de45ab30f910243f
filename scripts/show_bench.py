"""Pretty-print the last JSON line of a bench log: headline + per-stage table."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", round(d["value"], 1), d["unit"], "| ms/step", round(d["ms_per_step"], 3), "| e2e", d.get("e2e") and round(d["e2e"]["value"], 1),
      "| n_gpus", d["n_gpus"], "| launches", d.get("gpu_launches"), "| clocks", d.get("clocks"))
if "stages" in d:
    tot = 0
    for k, v in d["stages"].items():
        tot += v["ms_per_launch"]
        print(f"  {k:32s} {v['ms_per_launch']:.4f} ms  {v['gbs'] and round(v['gbs'])} GB/s")
    print("  native sum", round(tot, 3), "ms; roofline:", {k: d["roofline"][k] for k in ("kernel", "achieved", "frac", "share_of_native_time")})
    print("  whole view:", d["roofline"]["whole_view"])
if "cpu_baseline" in d:
    print("  cpu_baseline:", d["cpu_baseline"])
for k in ("significance_pass", "loss_pass", "iteration_pass", "vq_pass"):
    if k in d:
        print(f"  {k}:", {a: (round(b, 3) if isinstance(b, float) else b) for a, b in d[k].items() if a != "what"})
