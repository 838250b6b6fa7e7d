"""Radix-pass micro-benchmark through the C ABI (mb_debug_radix): python tools/bench_radix.py [N] [SHIFT] [KBITS] [REPS]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mauvealigner_b200 as mb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40_000_000
shift = int(sys.argv[2]) if len(sys.argv) > 2 else 27
kbits = int(sys.argv[3]) if len(sys.argv) > 3 else 30
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
ctx = mb.Context(0)
out = (C.c_float * 2)()
rc = mb.lib().mb_debug_radix(ctx._h, n, shift, kbits, reps, out)
if rc != 0:
    print(f'WARNING rc={rc} (order check failed)', end=' ')
gbs = 16.0 * n / (out[0] * 1e-3) / 1e9
print(f"n={n} shift={shift} kbits={kbits}: {out[0]:.4f} ms/pass ({gbs:.0f} GB/s algorithmic, {gbs / 6553.3:.3f} of 6553 GB/s), sort {out[1]:.4f} ms")
