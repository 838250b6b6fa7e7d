"""Many small problems: searches per second for problems of K sequences x LEN bases, one
context one at a time against a pool of contexts on concurrent streams.
usage: python tools/bench_small.py [LEN ...]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mauvealigner_b200 as mb  # noqa: E402
from toygen import family  # noqa: E402


def main():
    lens = [int(x) for x in sys.argv[1:]] or [500, 5000, 50000]
    rng = np.random.default_rng(1)
    pattern = mb.get_seed(11, 0)
    for n in lens:
        problems = [family(rng, n, 2, sub=0.05, indel=0.005, inv=0) for _ in range(64)]
        problems = [[np.frombuffer(s.encode(), dtype=np.uint8) for s in p] for p in problems] * 8  # 512 searches
        line = f"2 x {n} bp:"
        for streams in (1, 4, 16):
            pool = mb.ContextPool(streams)
            pool.find_many(problems[:2 * streams], pattern)  # warm-up: workspaces allocated
            t0 = time.perf_counter()
            res = pool.find_many(problems, pattern)
            dt = time.perf_counter() - t0
            pool.close()
            line += f"  {streams:2d} context(s) {len(problems) / dt:8.0f} searches/s ({1e3 * dt / len(problems):.3f} ms each)"
        print(line + f"  [{sum(r['n_matches'] for r in res)} matches]", flush=True)


if __name__ == "__main__":
    main()
