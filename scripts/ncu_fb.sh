# ncu --set full of the forward-backward kernel on the configs[4] shape (scripts/fb_probe.py: 1 M sequences x T = 50, S = 5, d = 2)
P=${PFX:-r02d}
python scripts/fb_probe.py > gpurun_out/${P}_fb_probe.json 2> gpurun_out/${P}_fb_probe.err || exit 1
ncu --set full --clock-control none --import-source on -k k_fb -s 2 -c 1 -o gpurun_out/${P}_fb -f python scripts/fb_probe.py > gpurun_out/${P}_ncu_fb.log 2>&1
python scripts/ncu_summary.py gpurun_out/${P}_fb.ncu-rep vbhmm_fb.cu vbhem_math.cuh > gpurun_out/${P}_k_fb_ncu.json
cat gpurun_out/${P}_fb_probe.json
