# ncu captures of round 2 (profiles/r02_*): run under gpurun from the repo root.
# Per B200_PROFILING.md: first the plain command (must exit 0), then the launch list, then ONE --set full capture per kernel.
set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extras"
$CMD > gpurun_out/r02_plain.json 2> gpurun_out/r02_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lattice_bwd -s 3 -c 1 -o gpurun_out/r02_bwd -f $CMD > gpurun_out/r02_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lattice_full -s 3 -c 1 -o gpurun_out/r02_full -f $CMD > gpurun_out/r02_ncu_f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_zhat -s 3 -c 1 -o gpurun_out/r02_zhat -f $CMD > gpurun_out/r02_ncu_z.log 2>&1
for r in r02_bwd r02_full r02_zhat; do python scripts/ncu_summary.py gpurun_out/$r.ncu-rep > gpurun_out/$r.summary.json; done
python - <<'PY'
import csv, collections, json
rows = list(csv.reader(l for l in open("gpurun_out/r02_launches.csv") if l.startswith('"')))
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
with open("gpurun_out/r02_launches_summary.csv", "w") as f:
    f.write("# ncu launch list of: python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extras (N=100k, K=64); k_peak_* (fp64 peak probes) and torch fill kernels excluded from the shares\n")
    f.write("# cold-cache, serialised launches: compare SHARES, not absolutes\nkernel,launches,total_ms,share_pct\n")
    for k, v in sorted(ours.items(), key=lambda kv: -kv[1]):
        f.write(f"{k.split('(')[0]},{cnt[k]},{v / 1e6:.3f},{100 * v / s:.2f}\n")
print(open("gpurun_out/r02_launches_summary.csv").read())
PY
rm -f gpurun_out/r02_full.ncu-rep gpurun_out/r02_zhat.ncu-rep
ls -la gpurun_out/ | tail -20
# BASELINE configs[4]: the forward-backward kernel
python bench.py --workload fb --steps 1 --no-cpu --no-e2e > gpurun_out/r02_plain_fb.json 2> gpurun_out/r02_plain_fb.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:"k_fb<" -s 3 -c 1 -o gpurun_out/r02_fb -f python bench.py --workload fb --steps 1 --no-cpu --no-e2e > gpurun_out/r02_ncu_fb.log 2>&1
python scripts/ncu_summary.py gpurun_out/r02_fb.ncu-rep > gpurun_out/r02_fb.summary.json
rm -f gpurun_out/r02_fb.ncu-rep
