"""Prints the interesting parts of a bench.py JSON line.   python scripts/show_bench.py gpurun_out/x.json"""
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
for k in ("value", "ms_per_step", "n_gpus", "per_rank_ms_per_step", "comm", "e2e", "c3_candidates", "c1_single_video", "c4_long_video", "train_step"):
    print(k, json.dumps(d.get(k))[:1200])
if "full_inference" in d:
    print("full", json.dumps({k: v for k, v in d["full_inference"].items() if k.endswith("ms") or k == "ms_per_step" or k == "error"}), d["full_inference"].get("roofline", {}).get("frac"))
print("roofline", d["roofline"]["frac"], d["roofline"]["ms_per_launch"])
if "cpu_baseline" in d:
    print("cpu", json.dumps(d["cpu_baseline"])[:1800])
if "masks" in d:
    print("masks roof", d["masks"].get("roofline"), d["masks"].get("flint_fused", {}).get("roofline"))
