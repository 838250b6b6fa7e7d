"""Randomised parity sweep of the CUDA path against the CPU oracle (test infrastructure): random genome counts, lengths
(around the tile sizes of the kernels: 2048, 4096, 7168 and multiples), seed patterns, policies and relations between the
genomes.  usage: python tools/fuzz_parity.py [CASES] [SEED]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mauvealigner_b200 as mb  # noqa: E402
import oracle_lib as O  # noqa: E402
from toygen import family, mutate, rand_seq, revcomp  # noqa: E402


def random_pattern(rng):
    while True:
        L = int(rng.integers(3, 40))
        half = [int(rng.random() < 0.7) for _ in range(L // 2)]
        mid = [1] if L % 2 else []
        bits = half + mid + half[::-1]
        bits[0] = bits[-1] = 1
        p = int("".join(map(str, bits)), 2)
        w = bin(p).count("1")
        if w % 2 == 1 and 3 <= w <= 31 and mb.seed_valid(p):
            return p


def same(a, b):
    return a["n_matches"] == b["n_matches"] and all(np.array_equal(np.asarray(a[k], dtype=np.int64), np.asarray(b[k], dtype=np.int64))
                                                    for k in ("length", "comp_off", "comp_seq", "comp_start"))


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    ctx = mb.Context(0)
    sizes = [0, 1, 5, 40, 300, 2047, 2048, 2049, 4095, 4097, 7167, 7168, 7169, 14336, 14337, 21000, 30000]
    bad = 0
    for it in range(cases):
        k = int(rng.integers(1, 7))
        n = int(sizes[int(rng.integers(0, len(sizes)))]) + int(rng.integers(0, 3))
        kind = int(rng.integers(0, 4))
        if kind == 0 or n < 60:
            seqs = [rand_seq(rng, max(0, n + int(rng.integers(-3, 4)))) for _ in range(k)]
        elif kind == 1:
            seqs = family(rng, n, k, sub=float(rng.choice([0.0, 0.01, 0.05])), indel=float(rng.choice([0.0, 0.004])), inv=int(rng.integers(0, 2)))
        elif kind == 2:
            u = rand_seq(rng, int(rng.integers(5, 60)))
            seqs = [(u * (n // len(u) + 1))[:n] if g % 2 == 0 else mutate(rng, (u * (n // len(u) + 1))[:n], sub=0.02, indel=0.0) for g in range(k)]
        else:
            base = rand_seq(rng, n)
            seqs = [base if g % 3 == 0 else (revcomp(base) if g % 3 == 1 else mutate(rng, base, sub=0.03, indel=0.002)) for g in range(k)]
        pattern = random_pattern(rng)
        modes = [(mb.MODE_UNIQUE, {}), (mb.MODE_UNIQUE_COUNT, {})]
        if k <= 4:
            modes.append((mb.MODE_PAIRWISE, {}))
        if k == 1:
            modes += [(mb.MODE_SEED_ENUM, dict(min_multi=2, max_multi=int(rng.choice([3, 50, 1000])), direct_only=bool(rng.integers(0, 2)))),
                      (mb.MODE_REPEAT, dict(min_multi=2, max_multi=int(rng.choice([3, 50, 255]))))]
        if k >= 3 and rng.random() < 0.3:
            modes.append((mb.MODE_UNIQUE, dict(nway_mask=int(rng.integers(1, 1 << k)))))
        ctx.clear_sequences()
        for s in seqs:
            ctx.add_sequence(s)
        ctx.set_seed(pattern)
        for mode, kw in modes:
            got = ctx.find(mode, **kw)
            want = O.find(seqs, pattern, mode, **kw)
            ok = same(got, want) and (mode != mb.MODE_UNIQUE_COUNT or int(got["unique_mers"]) == int(want["unique_mers"]))
            if not ok:
                bad += 1
                print(f"MISMATCH case {it}: k={k} n={n} kind={kind} pattern={bin(pattern)} mode={mode} kw={kw} got {got['n_matches']} want {want['n_matches']}", flush=True)
    ctx.close()
    print(f"FUZZ {cases} cases, {bad} mismatches", flush=True)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
