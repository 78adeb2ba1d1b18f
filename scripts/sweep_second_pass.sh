# Times the second pass (north-star shape and the (K,S) grid) for every library variant in <pkg>/_build/lib_*.so.
P=clustering-hidden-markov-models-with-variational-bayesian-hierarchical-em_b200
cp $P/libvbhem_b200.so /tmp/orig.so
for f in $P/_build/lib_*.so; do
  cp $f $P/libvbhem_b200.so
  a=$(python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.4f' % d['phases']['second_pass_ms'])")
  b=$(python bench.py --workload grid --steps 3 --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.3f' % d['phases']['second_pass_ms'])")
  echo "$(basename $f) second pass: north-star $a ms, grid $b ms"
done
cp /tmp/orig.so $P/libvbhem_b200.so
