# First-pass time of the default library and of every experimental build <pkg>/_build/lib_*.so (N = 100 000 and a
# 12 500-base shard), plus the full-shape parity line of each.
P=clustering-hidden-markov-models-with-variational-bayesian-hierarchical-em_b200
for f in $P/libvbhem_b200.so $P/_build/lib_*.so; do
  echo "== $f"
  VBHEM_LIB=$PWD/$f python -m pytest tests/test_gpu_parity.py -q -s -k "full_size" 2>&1 | grep -E "full-size|passed|failed" | cut -c1-200
  for nb in 100000 12500; do
  VBHEM_LIB=$PWD/$f python bench.py --bases $nb --steps 20 --no-cpu --no-e2e --no-extras 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bases $nb step %.4f ms first %.4f ms' % (d['ms_per_step'], d['phases']['first_pass_ms']))"
  done
done
