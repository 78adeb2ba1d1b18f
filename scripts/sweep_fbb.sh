# Times the batched VB-HMM learning run (bench.py --workload fb: learn_batch.device_ms) for every library variant in
# <pkg>/_build/libfbb_*.so (build experiments of k_fb_batch).
P=clustering-hidden-markov-models-with-variational-bayesian-hierarchical-em_b200
cp $P/libvbhem_b200.so /tmp/orig.so
for f in $P/_build/libfbb_*.so; do
  cp $f $P/libvbhem_b200.so
  python bench.py --workload fb --steps 1 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$f'.split('/')[-1], 'learn_batch %.2f ms' % d['learn_batch']['device_ms'])"
done
cp /tmp/orig.so $P/libvbhem_b200.so
