import json,sys
for f in sys.argv[1:]:
    try:
        d=json.load(open(f))
        print(f, "ms/step %.4f"%d["ms_per_step"], {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["phases"].items() if k!="note"}, "frac %.4f"%d["roofline"]["frac"], "launches", d["gpu_launches"], "iter_ms", d.get("vbhem_iter",{}).get("ms"))
    except Exception as e:
        print(f, "ERR", e)
