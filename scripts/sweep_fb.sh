# Times scripts/fb_probe.py for every library variant in <pkg>/_build/libfb_*.so (build experiments).
P=clustering-hidden-markov-models-with-variational-bayesian-hierarchical-em_b200
cp $P/libvbhem_b200.so /tmp/orig.so
for f in $P/_build/libfb_*.so; do
  cp $f $P/libvbhem_b200.so
  python scripts/fb_probe.py ${FBN:-1000000} 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$f'.split('/')[-1], 'best %.3f ms  %.0f GB/s algorithmic' % (d['best_ms'], d['achieved_GBs']))"
done
cp /tmp/orig.so $P/libvbhem_b200.so
