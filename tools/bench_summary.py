"""One-screen summary of a bench.py JSON line: python tools/bench_summary.py gpurun_out/bench_final.log"""
import json, sys
d = json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "steps", "warmup", "gpu_launches", "n_gpus")})
print("e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "achieved TF", d["roofline"]["achieved"], "traffic", d["roofline"]["traffic"])
print("latency", d.get("latency"))
print("cpu", d["cpu_baseline"] and (d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["cpu_baseline"]["kind"]))
for k in ("full_step", "mixed_gait", "horizon30"):
    if k in d: print(k, d[k]["value"], d[k].get("ms_per_step"), d[k].get("rounds_mean"))
if "wbc" in d: print("wbc", d["wbc"]["batch_1024"]["value"], d["wbc"]["batch_65536"]["value"])
print("clocks", d["clocks"], "rounds", d["config"]["polish_rounds_mean"], "not_converged", d["config"]["not_converged"])
