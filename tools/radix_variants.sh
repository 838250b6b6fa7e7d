#!/bin/bash
# builds onesweep variants into variants (run on the GPU box or here) and times each
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
build() { # name flags...
  name=$1; shift
  touch mauvealigner_b200/csrc/kernels_radix.cu
  make -s -C mauvealigner_b200/csrc EXTRA="$*" > /dev/null
  cp mauvealigner_b200/libmauve_b200.so variants/lib_$name.so
}
if [ "$1" == "build" ]; then
  build cur
  build nolook -DRS_NOLOOK
else
  for f in variants/lib_*.so; do echo -n "$(basename $f): "; MAUVE_B200_LIB=$f python tools/bench_radix.py 40000000 27 30 5; done
fi
