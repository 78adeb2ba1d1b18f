# Builds <pkg>/_build/libfb_<tag>.so for a list of "tag:flags" variants of vbhmm_fb.cu (the other objects are compiled once).
# Usage: scripts/build_fb_variants.sh "m2:-DVB_FB_MINB=2" "m3:-DVB_FB_MINB=3" ...   then scripts/sweep_fb.sh under gpurun
set -e
P=$(cd "$(dirname "$0")/.." && pwd)/clustering-hidden-markov-models-with-variational-bayesian-hierarchical-em_b200
O=$P/_build/obj; mkdir -p $O
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC"
cd $P/csrc
for f in vbhem_capi vbhem_general vbhem_lattice vbhem_aux vbhem_em; do
  if [ ! -f $O/$f.o ] || [ $f.cu -nt $O/$f.o ]; then $NV -c $f.cu -o $O/$f.o & fi
done
wait
rm -f $P/_build/libfb_*.so
for v in "$@"; do
  tag=${v%%:*}; flags=${v#*:}
  ( $NV $flags -c vbhmm_fb.cu -o $O/fb_$tag.o && $NV -shared -o $P/_build/libfb_$tag.so $O/vbhem_capi.o $O/vbhem_general.o $O/vbhem_lattice.o $O/vbhem_aux.o $O/vbhem_em.o $O/fb_$tag.o ) &
done
wait
ls -la $P/_build/libfb_*.so
