"""Prints the headline fields of a bench.py JSON line (gpurun_out/*.json)."""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.load(open(f))
    except Exception as e:  # noqa: BLE001
        print(f, "ERR", e)
        continue
    print("==", f)
    for k in ("value", "ms_per_step", "scaling", "n_gpus", "gpu_launches", "placement"):
        print(" ", k, d.get(k))
    strip = lambda x: {k: v for k, v in x.items() if k != "note"} if isinstance(x, dict) else x
    if d.get("e2e"):
        print("  e2e", strip(d["e2e"]))
    r = d.get("roofline", {})
    print("  roofline frac", r.get("frac"), "kernel_ms", r.get("kernel_ms"), "traffic", r.get("traffic"), "|", r.get("traffic_source"))
    for k in ("phases", "cpu_baseline", "parity", "vbhem_iter", "weak", "single_process", "grid_em"):
        if d.get(k):
            v = strip(d[k])
            if k == "cpu_baseline":
                v = {kk: vv for kk, vv in v.items() if kk != "sample"}
            print(" ", k, v)
