# Accuracy / speed of the first pass for the Taylor degrees of the kernel-side exp / log (csrc/vbhem_math.cuh):
# every <pkg>/_build/lib_exp*log*.so (built with -DVB_EXP_DEG=.. -DVB_LOG_DEG=..) against the default library.
P=clustering-hidden-markov-models-with-variational-bayesian-hierarchical-em_b200
for f in $P/libvbhem_b200.so $P/_build/lib_exp*log*.so; do
  echo "== $f"
  VBHEM_LIB=$PWD/$f python -m pytest tests/test_gpu_parity.py -q -s -k "full_size or golden or fused_estep_matches" 2>&1 | grep -E "full-size|passed|failed"
  VBHEM_LIB=$PWD/$f python bench.py --steps 20 --no-cpu --no-e2e --no-extras 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('step %.3f ms first %.3f ms' % (d['ms_per_step'], d['phases']['first_pass_ms']))"
done
