# Round 2, final state (profiles/r02d_*): GPU tests, the default bench line, the launch list and ONE --set full capture of
# the first and second pass.  Run under gpurun from the repo root.
set -x
P=${PFX:-r02d}
python -m pytest tests -m gpu -x -q > gpurun_out/${P}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${P}_tests.log
tail -3 gpurun_out/${P}_tests.log
python bench.py > gpurun_out/${P}_bench_default.json 2> gpurun_out/${P}_bench_default.err || exit 1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extras"
$CMD > gpurun_out/${P}_plain.json 2> gpurun_out/${P}_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${P}_launches.csv $CMD > gpurun_out/${P}_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lattice_bwd -s 3 -c 1 -o gpurun_out/${P}_bwd -f $CMD > gpurun_out/${P}_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lattice_full -s 3 -c 1 -o gpurun_out/${P}_full -f $CMD > gpurun_out/${P}_ncu_f.log 2>&1
SRC="vbhem_lattice.cu vbhem_math.cuh vbhem_internal.cuh"
python scripts/ncu_summary.py gpurun_out/${P}_bwd.ncu-rep $SRC > gpurun_out/${P}_k_lattice_bwd_ncu.json
python scripts/ncu_summary.py gpurun_out/${P}_full.ncu-rep $SRC > gpurun_out/${P}_k_lattice_full_ncu.json
python - <<PY
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/${P}_launches.csv") if l.startswith('"')))
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    try:
        tot[r[ki]] += float(r[vi].replace(",", "")); cnt[r[ki]] += 1
    except ValueError:
        pass
ours = {k: v for k, v in tot.items() if k.startswith(("k_", "void vbhem", "vbhem")) and "k_peak" not in k}
s = sum(ours.values())
with open("gpurun_out/${P}_launches_summary.csv", "w") as f:
    f.write("# ncu launch list of: python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extras (N=100k, K=64); k_peak_* (fp64 peak probes) and torch fill kernels excluded from the shares\n")
    f.write("# cold-cache, serialised launches: compare SHARES, not absolutes\nkernel,launches,total_ms,share_pct\n")
    for k, v in sorted(ours.items(), key=lambda kv: -kv[1]):
        f.write(f"{k.split('(')[0]},{cnt[k]},{v / 1e6:.3f},{100 * v / s:.2f}\n")
print(open("gpurun_out/${P}_launches_summary.csv").read())
PY
rm -f gpurun_out/${P}_full.ncu-rep
cat gpurun_out/${P}_bench_default.json
