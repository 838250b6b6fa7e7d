"""Timing of inputs with very long window-consistent runs (DESIGN.md §4 'known cliff'): identical genomes, and two
identical genomes among diverged ones.  python tools/stress_cliff.py [LEN [CASE-SUBSTRING]]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mauvealigner_b200 as mb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
rng = np.random.default_rng(7)
alpha = np.frombuffer(b"ACGT", dtype=np.uint8)
a = alpha[rng.integers(0, 4, size=n)]


def mutate(x, rate):
    y = x.copy()
    idx = np.flatnonzero(rng.random(x.size) < rate)
    y[idx] = alpha[(np.searchsorted(alpha, y[idx]) + rng.integers(1, 4, size=idx.size)) % 4]
    return y


cases = {"2 identical": [a, a.copy()], "3 identical": [a, a.copy(), a.copy()],
         "2 identical + 3 at 3%": [a, a.copy(), mutate(a, 0.03), mutate(a, 0.03), mutate(a, 0.03)],
         "1 SNP per 10 kb": [a, mutate(a, 1e-4)]}
ctx = mb.Context(0)
ctx.set_seed(mb.get_seed(15, 0))
for name, seqs in cases.items():
    if len(sys.argv) > 2 and sys.argv[2] not in name:
        continue
    ctx.clear_sequences()
    for s in seqs:
        ctx.add_sequence(s)
    ctx.find_device(mb.MODE_UNIQUE)  # warm-up: the first search of a size allocates the workspaces
    ctx.fetch()
    t0 = time.perf_counter()
    ctx.find_device(mb.MODE_UNIQUE)
    r = ctx.fetch()
    dt = time.perf_counter() - t0
    st = ctx.stats()
    print(f"{name:24s} len {n}: {dt * 1e3:9.1f} ms wall, device {st['ms_total_device']:9.2f} ms (dedup {st['ms_dedup']:.2f}), "
          f"{r['n_matches']} matches, reps {st['n_extended']}, rounds {st['dedup_iters']}", flush=True)
