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
  build early_ballot -DRS_EARLY -DRS_BALLOT
  build eb_minb3 -DRS_EARLY -DRS_BALLOT -DRS_MINB=3
  build eb_512x8 -DRS_EARLY -DRS_BALLOT -DRS_IPT=8
  build eb_512x8_minb3 -DRS_EARLY -DRS_BALLOT -DRS_IPT=8 -DRS_MINB=3
  build eb_512x10 -DRS_EARLY -DRS_BALLOT -DRS_IPT=10
  build eb_512x14 -DRS_EARLY -DRS_BALLOT -DRS_IPT=14
  build eb_384x12 -DRS_EARLY -DRS_BALLOT -DRS_NT=384 -DRS_IPT=12
  build eb_384x12_minb4 -DRS_EARLY -DRS_BALLOT -DRS_NT=384 -DRS_IPT=12 -DRS_MINB=4
  build eb_256x16_minb6 -DRS_EARLY -DRS_BALLOT -DRS_NT=256 -DRS_IPT=16 -DRS_MINB=6
else
  for f in variants/lib_*.so; do echo -n "$(basename $f): "; MAUVE_B200_LIB=$f python tools/bench_radix.py 40000000 27 30 5; done
fi
