# e2e (host buffers in and out, streamed in slices) of the default workload for several slice counts
for s in ${SLICES:-3 4 6 8 10 12 16}; do
  python bench.py --no-cpu --no-extras --steps 5 --slices $s 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('slices', d['e2e']['slices'], 'e2e %.3f ms' % d['e2e']['ms_per_step'], 'device %.3f ms' % d['ms_per_step'])"
done
