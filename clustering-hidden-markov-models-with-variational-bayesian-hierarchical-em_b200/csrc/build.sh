#!/bin/sh
# Builds libvbhem_b200.so (sm_100a) next to the package.  Usage: [VBHEM_OUT=path.so] csrc/build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
nvcc --threads 0 -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -shared \
     "$@" -o "${VBHEM_OUT:-../libvbhem_b200.so}" vbhem_capi.cu vbhem_general.cu vbhem_lattice.cu vbhem_aux.cu vbhem_em.cu vbhmm_fb.cu
