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
    wide = sum(1 << (40 - j) for j in (0, 5, 13, 19, 20, 21, 27, 35, 40))  # L = 41 (> 32: the window-by-window extension), weight 9
    for pattern in (0b110111011, mb.get_seed(15, 0), wide):
        ctx.set_seed(pattern)
        for mode, kw in ((mb.MODE_UNIQUE, {}), (mb.MODE_PAIRWISE, {}), (mb.MODE_UNIQUE_COUNT, {})):
            r = ctx.find(mode, **kw)
            print("mode", mode, "pattern", bin(pattern)[:14], "matches", r["n_matches"], flush=True)
    # the table persisting across searches (seed family), compact fetch
    ctx.accumulate(True)
    for pattern in (mb.get_seed(11, 2), mb.get_seed(11, 1), mb.get_seed(9, 0)):
        ctx.set_seed(pattern)
        r = ctx.find(mb.MODE_UNIQUE, compact=True)
        print("family pass", bin(pattern)[:14], "matches", r["n_matches"], flush=True)
    ctx.accumulate(False)
    # one sequence: seed enumeration, RepeatHash, the position lookup table
    ctx.clear_sequences()
    ctx.add_sequence(seqs[0] + revcomp(seqs[0][1000:1400]) + seqs[0][2000:2300] + "AC" * 40)
    ctx.set_seed(0b110111011)
    for mode, kw in ((mb.MODE_SEED_ENUM, dict(min_multi=2, max_multi=50)), (mb.MODE_REPEAT, dict(max_multi=255))):
        r = ctx.find(mode, **kw)
        mo, co = ctx.position_table()
        print("mode", mode, "matches", r["n_matches"], "table entries", int((mo != 0xFFFFFFFF).sum()), flush=True)
    # many small problems in one pass
    probs = [family(rng, int(rng.integers(0, 400)), 3, sub=0.04, indel=0.004, inv=0) for _ in range(40)]
    probs[3][2] = ""
    ctx.set_seed(0b1011101)
    res = ctx.find_batch(probs, mb.MODE_UNIQUE)
    print("batch of", len(probs), "matches", sum(r["n_matches"] for r in res), flush=True)
    ctx.close()
    for p2p in (0, 1, 2):
        r = dist.find_unique_emulated(seqs, 0b1101110111110111011, 3, p2p=p2p)
        print("emulated world 3 p2p", p2p, "matches", r["n_matches"], flush=True)


if __name__ == "__main__":
    main()
