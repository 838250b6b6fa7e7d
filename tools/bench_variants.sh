#!/bin/bash
# bench.py (C5 only, no CPU leg) with the product library and with every tuning build under build/variants/.
out=${1:-gpurun_out/bench_variants.txt}
: > $out
python bench.py --no-cpu-baseline --no-extra --steps 5 > gpurun_out/bv_default.json 2> gpurun_out/bv_default.err && echo "default: $(python tools/show_bench.py gpurun_out/bv_default.json | head -1)" >> $out
for so in build/variants/*.so; do
  b=$(basename $so .so)
  MAUVE_B200_LIB=$so python bench.py --no-cpu-baseline --no-extra --steps 5 > gpurun_out/bv_$b.json 2> gpurun_out/bv_$b.err && echo "$b: $(python tools/show_bench.py gpurun_out/bv_$b.json | head -1)" >> $out
done
cat $out
