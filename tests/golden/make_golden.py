"""Generates tests/golden/*.json — small committed input/output vectors of the seed-match path.

PARITY UNPINNED: the reference tree (/root/reference) holds no golden vectors, fixtures or tests for this path and
libMems (where the arithmetic lives) is not available, so these vectors pin OUR semantics (SURVEY.md Appendix A
D1-D18): they are produced by the CPU oracle (oracle/oracle.cpp) and, before being written, cross-checked against
the independent naive restatement in tests/brute.py (no sort, no hash table, straight from the ASCII).  They keep
the oracle, the brute-force checker and the CUDA path from drifting together unnoticed.

    python tests/golden/make_golden.py        (re-run only when a D-rule is changed on purpose)"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import brute  # noqa: E402
import oracle_lib as O  # noqa: E402
from toygen import family, mutate, rand_seq, revcomp  # noqa: E402


def case(name, seqs, pattern, mode, **kw):
    res = O.find(seqs, pattern, mode, **kw)
    got = O.matches_as_list(res)
    if sum(len(s) for s in seqs) <= 4000:  # independent naive restatement (quadratic): small cases only
        want = brute.find(seqs, pattern, mode, **kw)
        assert got == want["matches"], name
        assert int(res["unique_mers"]) == want["unique_mers"], name
    return dict(name=name, seqs=seqs, pattern=pattern, mode=mode, params=kw, n_matches=int(res["n_matches"]),
                length=[int(x) for x in res["length"]], comp_off=[int(x) for x in res["comp_off"]],
                comp_seq=[int(x) for x in res["comp_seq"]], comp_start=[int(x) for x in res["comp_start"]],
                unique_mers=int(res["unique_mers"]), unique_mers_per_seq=[int(x) for x in res["unique_mers_per_seq"]])


def main():
    rng = np.random.default_rng(20261018)
    cases = []
    a = rand_seq(rng, 600)
    cases.append(case("two_genomes_subst", [a, mutate(rng, a, sub=0.03, indel=0.0)], 0b1101110111110111011, O.MODE_UNIQUE))
    cases.append(case("three_genomes_indel_inversion", family(rng, 900, 3, sub=0.02, indel=0.004, inv=1), 0b110111011, O.MODE_UNIQUE))
    f = family(rng, 700, 4, sub=0.03, indel=0.003, inv=0)
    f[2] = revcomp(f[2])
    cases.append(case("four_genomes_one_reverse", f, 0b110110110111011011011, O.MODE_UNIQUE))
    cases.append(case("nway_mask", f, 0b1011101, O.MODE_UNIQUE, nway_mask=0b0101))
    cases.append(case("pairwise_four_genomes", f, 0b110111011, O.MODE_PAIRWISE))
    cases.append(case("identical_pair_solid", [a, a], 0b1111111, O.MODE_UNIQUE))
    cases.append(case("short_and_empty", ["ACG", "ACGTTGCAGT", "", "TTGCAGTACG"], 0b11111, O.MODE_UNIQUE))
    unit = rand_seq(rng, 120)
    rep = rand_seq(rng, 700) + unit + rand_seq(rng, 200) + mutate(rng, unit, sub=0.04, indel=0) + revcomp(unit) + "AT" * 60 + unit
    cases.append(case("repeats_enum", [rep], 0b110111011, O.MODE_SEED_ENUM, min_multi=2, max_multi=500))
    cases.append(case("repeats_enum_direct_only", [rep], 0b11111, O.MODE_SEED_ENUM, min_multi=2, max_multi=6, direct_only=True))
    cases.append(case("unique_count", [rep, a], 0b1011101, O.MODE_UNIQUE_COUNT))
    mers = dict(name="mers_and_sml", seq=rep[:400], pattern=0b1101110111110111011)
    mers["mers"] = [int(x) for x in O.mers(mers["seq"], mers["pattern"])]
    mers["sml"] = [int(x) for x in O.sml(mers["seq"], mers["pattern"])]
    with open(os.path.join(HERE, "seed_match_golden.json"), "w") as fh:
        json.dump(dict(cases=cases, mers=mers), fh, separators=(",", ":"))
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
