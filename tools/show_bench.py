"""Print the parts of a bench.py JSON line a human wants to see.  usage: python tools/show_bench.py line.json [key ...]"""
import json
import sys


def short(v, depth=0):
    if isinstance(v, float):
        return round(v, 4)
    if isinstance(v, dict):
        return {k: short(x, depth + 1) for k, x in v.items() if k not in ("how", "note", "calls", "what", "sample", "peak_source", "profile")}
    if isinstance(v, list):
        return [short(x, depth + 1) for x in v[:12]]
    return v


d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
keys = sys.argv[2:] or ["value", "ms_per_step", "n_gpus", "gpu_launches", "clocks", "e2e", "e2e_u8", "parity_check", "phases_ms", "other_gather", "roofline", "cpu_baseline", "extras"]
for k in keys:
    if k in d:
        print(k, "=", json.dumps(short(d[k]))[:1800])
