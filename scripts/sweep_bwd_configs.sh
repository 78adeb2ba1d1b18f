# Times bench.py's first pass for every library variant in <pkg>/_build/lib_*.so (build experiments).
P=clustering-hidden-markov-models-with-variational-bayesian-hierarchical-em_b200
cp $P/libvbhem_b200.so /tmp/orig.so
for f in $P/_build/lib_*.so; do
  cp $f $P/libvbhem_b200.so
  for tau in ${TAUS:-10}; do
  python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e --tau $tau 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$f'.split('/')[-1], 'tau $tau', 'step %.2f ms first %.2f ms fallback %d' % (d['ms_per_step'], d['phases']['first_pass_ms'], d['phases']['fallback_pairs']))"
  done
done
cp /tmp/orig.so $P/libvbhem_b200.so
