"""Small end-to-end runs of every mode and of the emulated multi-rank path, for compute-sanitizer
(memcheck / initcheck / racecheck / synccheck):  compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mauvealigner_b200 as mb  # noqa: E402
from mauvealigner_b200 import dist  # noqa: E402
from toygen import family, revcomp  # noqa: E402


def main():
    rng = np.random.default_rng(5)
    seqs = family(rng, 6000, 4, sub=0.03, indel=0.003, inv=1)
    seqs[1] = revcomp(seqs[1])
    ctx = mb.Context(0)
    for s in seqs:
        ctx.add_sequence(s)
    for pattern in (0b110111011, mb.get_seed(15, 0), (1 << 40) - 1):
        ctx.set_seed(pattern)
        for mode, kw in ((mb.MODE_UNIQUE, {}), (mb.MODE_PAIRWISE, {}), (mb.MODE_UNIQUE_COUNT, {}),
                         (mb.MODE_SEED_ENUM, dict(min_multi=2, max_multi=50))):
            r = ctx.find(mode, **kw)
            print("mode", mode, "pattern", bin(pattern)[:14], "matches", r["n_matches"], flush=True)
    ctx.close()
    for p2p in (0, 1, 2):
        r = dist.find_unique_emulated(seqs, 0b1101110111110111011, 3, p2p=p2p)
        print("emulated world 3 p2p", p2p, "matches", r["n_matches"], flush=True)


if __name__ == "__main__":
    main()
