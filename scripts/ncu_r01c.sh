# ncu captures of the third session (profiles/r01c_*): run under gpurun, summaries via scripts/ncu_summary.py
set -x
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r03_plain_default.json 2> gpurun_out/r03_plain_default.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r03_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r03_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lattice_bwd -s 3 -c 1 -o gpurun_out/r03_bwd -f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r03_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lattice_full -s 3 -c 1 -o gpurun_out/r03_full -f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r03_ncu_f2.log 2>&1
# (K,S) grid: the five first-pass launches of one E-step (tile classes)
python bench.py --workload grid --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/r03_plain_grid.json 2> gpurun_out/r03_plain_grid.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_lattice_bwd -s 15 -c 5 -o gpurun_out/r03_bwd_grid -f python bench.py --workload grid --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/r03_ncu_g.log 2>&1
# batched VB-HMM learning: forward-backward, sums, update, GMM fit
python bench.py --workload fb --steps 1 --no-cpu --no-e2e > gpurun_out/r03_plain_fb.json 2> gpurun_out/r03_plain_fb.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:"k_fb_batch|k_batch_sums|k_batch_update|k_gmm_fit" -s 8 -c 4 -o gpurun_out/r03_batch -f python bench.py --workload fb --steps 1 --no-cpu --no-e2e > gpurun_out/r03_ncu_f.log 2>&1
# summaries on the box; only the first-pass report travels back (gpurun_out/ is capped at 64 MiB)
for r in r03_bwd r03_full r03_bwd_grid r03_batch; do python scripts/ncu_summary.py gpurun_out/$r.ncu-rep > gpurun_out/$r.summary.json; done
rm -f gpurun_out/r03_full.ncu-rep gpurun_out/r03_bwd_grid.ncu-rep gpurun_out/r03_batch.ncu-rep
ls -la gpurun_out/
