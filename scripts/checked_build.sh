#!/bin/sh
# The memcheck stand-in (compute-sanitizer is closed on this GPU pool, see profiles/README.md): build the library with
# device-side bounds assertions (-DVBHEM_CHECKED: every list position, partial slot and pair index computed from
# device-resident counts) and run the GPU parity tests against it.  A failed assertion aborts the kernel with
# "device-side assert triggered" and the test fails.
#   here (no GPU):   sh scripts/checked_build.sh build
#   on the GPU box:  sh scripts/checked_build.sh run      (under gpurun; log in gpurun_out/checked_build.log)
set -e
P=clustering-hidden-markov-models-with-variational-bayesian-hierarchical-em_b200
case "$1" in
  build)
    mkdir -p $P/_build
    VBHEM_OUT=../_build/libvbhem_b200_checked.so sh $P/csrc/build.sh -DVBHEM_CHECKED
    ;;
  run)
    VBHEM_LIB=$PWD/$P/_build/libvbhem_b200_checked.so python -m pytest tests/test_gpu_parity.py tests/test_em_loop.py \
      tests/test_multi_gpu.py -q -m gpu > gpurun_out/checked_build.log 2>&1 || true
    tail -5 gpurun_out/checked_build.log
    ;;
  *) echo "usage: $0 build|run"; exit 2 ;;
esac
