"""Prints a short summary of one or more bench.py JSON lines."""
import json, sys
for path in sys.argv[1:]:
    try:
        d = json.load(open(path))
    except Exception as e:  # noqa
        print(path, "ERR", e)
        continue
    print(f"== {path}: {d['config']['workload'][:60]}")
    print("  value", round(d["value"]), "audio-s/s  ms/step", round(d["ms_per_step"], 3), " eager", round(d.get("eager_ms_per_step") or 0, 3),
          " e2e", round(d["e2e"]["value"]), " gemm frac", d["roofline"]["frac"] if d.get("roofline") else None,
          " launches/step", d["gpu_launches"] // d["steps"], " n_gpus", d["n_gpus"], " clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    for k, v in (d.get("kernels") or {}).items():
        print(f"    {k:24s} {v['launches_per_step']:4d} {v['ms_per_step']:8.3f} ms  {v.get('achieved', '')} {v.get('unit', '')} {v.get('frac', '')}  (raw brackets {v.get('ms_per_step_raw_brackets', '')} ms, {v.get('frac_raw_brackets', '')})")
    if d.get("cpu_baseline"):
        print("  cpu", round(d["cpu_baseline"]["value"], 1), d["cpu_baseline"]["cores"], "cores")
