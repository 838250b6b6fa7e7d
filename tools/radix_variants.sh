#!/bin/bash
# Times one radix pass (tools/bench_radix.py) with the product library and with every tuning build under build/variants/.
out=${1:-gpurun_out/radix_variants.txt}
: > $out
echo "default: $(python tools/bench_radix.py 40000000 27 30 5 2>&1 | tail -1)" >> $out
echo "default C5 size: $(python tools/bench_radix.py 320000000 30 30 3 2>&1 | tail -1)" >> $out
for so in build/variants/*.so; do
  echo "$(basename $so): $(MAUVE_B200_LIB=$so timeout 120 python tools/bench_radix.py 40000000 27 30 5 2>&1 | tail -1)" >> $out
done
cat $out
