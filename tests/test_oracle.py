"""CPU tests pinning oracle/oracle.cpp against (i) hand-derivable toy cases, (ii) the naive
restatement in tests/brute.py, (iii) invariants (SURVEY.md §8c "oracle self-validation").
The reference holds no golden vectors for this path (parity unpinned)."""
import numpy as np
import pytest

import brute
import oracle_lib as O
from toygen import family, mutate, rand_seq, revcomp

PATTERNS = [0b111, 0b11111, 0b1011101, 0b110111011, 0b11011011111011011, 0b1101110111110111011]


def as_brute_list(res):
    return [(ln, comps) for ln, comps in O.matches_as_list(res)]


def test_seed_validity():
    L = O.lib()
    assert L.orc_seed_valid(0b111) == 1
    assert L.orc_seed_valid(0b1111) == 0  # even weight
    assert L.orc_seed_valid(0b1101) == 0  # not palindromic
    assert L.orc_seed_valid(0b10101) == 1
    assert L.orc_seed_length(0b1011101) == 7 and L.orc_seed_weight(0b1011101) == 5


def test_pack_layout():
    w = O.pack("ACGT" * 9)  # 36 bases -> 2 words
    assert w.size == 2
    assert int(w[0]) == int("00011011" * 8, 2)
    assert int(w[1]) == int("00011011", 2) << 56


def test_mers_known_answer():
    # solid weight-3 seed: window ACG -> fwd 0b000110=6, rc CGT=0b011011=27 -> key 6, strand 0
    m = O.mers("ACGT", 0b111)
    assert m.size == 2
    assert int(m[0]) == (6 << (64 - 6)) | 0
    # window CGT: fwd 27, rc ACG 6 -> key 6, strand 1
    assert int(m[1]) == (6 << (64 - 6)) | 1


@pytest.mark.parametrize("pattern", PATTERNS)
def test_mers_vs_brute(pattern):
    rng = np.random.default_rng(pattern)
    s = rand_seq(rng, 300)
    m = O.mers(s, pattern)
    w = bin(pattern).count("1")
    ref = brute.all_mers([s], pattern)[0]
    assert m.size == len(ref)
    for p, (key, strand) in enumerate(ref):
        assert int(m[p]) == (key << (64 - 2 * w)) | strand


def test_mers_strand_symmetry():
    rng = np.random.default_rng(5)
    s = rand_seq(rng, 500)
    for pattern in PATTERNS:
        Ls = pattern.bit_length()
        a, b = O.mers(s, pattern), O.mers(revcomp(s), pattern)
        # window p of s is window len-L-p of revcomp(s), on the other strand
        assert np.array_equal(a >> 1, (b >> 1)[::-1])
        assert np.array_equal(a & 1, 1 - (b & 1)[::-1])


def test_sml_sorted():
    rng = np.random.default_rng(7)
    s = rand_seq(rng, 2000, alphabet=2)
    for pattern in PATTERNS[:4]:
        m = O.mers(s, pattern)
        pos = O.sml(s, pattern)
        keys = (m >> 1)[pos]
        assert np.all(keys[:-1] <= keys[1:])
        same = keys[:-1] == keys[1:]
        assert np.all(pos[:-1][same] < pos[1:][same])
        assert sorted(pos.tolist()) == list(range(m.size))


def test_planted_match_solid():
    # two genomes sharing exactly one 40-mer, everything else unrelated (alphabet split)
    core = "ACGTTGCATGCAAGCTTAGCCGATATCGGCTAAGGCTTACG"[:40]
    g0 = "A" * 30 + "C" + core + "G" + "A" * 25  # flanks differ at both ends -> maximal
    g1 = "T" * 11 + "G" + core + "T" + "T" * 40
    r = O.find([g0, g1], 0b1111111, O.MODE_UNIQUE)
    ms = O.matches_as_list(r)
    want = (40, [(0, 32), (1, 13)])
    assert want in ms
    # and on the reverse strand of genome 1
    r = O.find([g0, revcomp(g1)], 0b1111111, O.MODE_UNIQUE)
    ms = O.matches_as_list(r)
    n1 = len(g1)
    left = n1 - (13 + 40 - 1) + 1
    assert (40, [(0, 32), (1, -left)]) in ms


def test_identical_genomes_single_match():
    rng = np.random.default_rng(11)
    s = rand_seq(rng, 3000)
    for pattern in (0b1111111, 0b1101110111110111011):
        r = O.find([s, s], pattern, O.MODE_UNIQUE)
        ms = O.matches_as_list(r)
        # all unique seeds lie on the main diagonal and the first one extends over everything
        assert (len(s), [(0, 1), (1, 1)]) in ms
        assert r["n_matches"] == 1
        assert r["n_contained"] == r["n_candidates"] - 1


@pytest.mark.parametrize("seed", range(12))
def test_unique_vs_brute(seed):
    rng = np.random.default_rng(100 + seed)
    pattern = PATTERNS[seed % len(PATTERNS)]
    k = int(rng.integers(2, 5))
    n = int(rng.integers(150, 500))
    seqs = family(rng, n, k, sub=0.04, indel=0.01, inv=seed % 3)
    if seed % 4 == 0:
        seqs[1] = revcomp(seqs[1])
    got = as_brute_list(O.find(seqs, pattern, O.MODE_UNIQUE))
    want = brute.find(seqs, pattern, brute.MODE_UNIQUE)["matches"]
    assert got == want


@pytest.mark.parametrize("seed", range(6))
def test_long_consistent_runs_vs_brute(seed):
    """Long window-consistent runs holding many candidates of ONE group (identical genomes, two identical genomes among
    diverged ones, sparse SNPs, a reverse-complemented copy): the structure behind the shared long walks of the CUDA path
    (kernels_dedup.cu, k_extend_long_classes) — the checker itself is pinned against the brute force here."""
    rng = np.random.default_rng(900 + seed)
    pattern = PATTERNS[seed % len(PATTERNS)]
    a = rand_seq(rng, int(rng.integers(500, 900)))
    cases = [[a, a], [a, a, mutate(rng, a, sub=0.03, indel=0.0), mutate(rng, a, sub=0.03, indel=0.0)], [a, mutate(rng, a, sub=0.004, indel=0.0)],
             [a, revcomp(a)], [a, mutate(rng, a, sub=0.004, indel=0.0), revcomp(mutate(rng, a, sub=0.004, indel=0.0))]]
    for seqs in cases:
        got = as_brute_list(O.find(seqs, pattern, O.MODE_UNIQUE))
        assert got == brute.find(seqs, pattern, brute.MODE_UNIQUE)["matches"]


@pytest.mark.parametrize("seed", range(6))
def test_unique_low_complexity_vs_brute(seed):
    # 2-letter alphabet + short seeds: many non-unique buckets, ties, overlapping diagonals
    rng = np.random.default_rng(200 + seed)
    pattern = PATTERNS[seed % 3]
    a = rand_seq(rng, 120, alphabet=2 + seed % 2)
    seqs = [a, mutate(rng, a, sub=0.05, indel=0.02), mutate(rng, a, sub=0.1, indel=0.0, inv=1)]
    r = O.find(seqs, pattern, O.MODE_UNIQUE)
    b = brute.find(seqs, pattern, brute.MODE_UNIQUE)
    assert as_brute_list(r) == b["matches"]
    assert r["unique_mers"] == b["unique_mers"]
    assert r["unique_mers_per_seq"].tolist() == b["unique_per_seq"]


@pytest.mark.parametrize("seed", range(4))
def test_pairwise_and_nway_vs_brute(seed):
    rng = np.random.default_rng(300 + seed)
    pattern = PATTERNS[1 + seed % 4]
    seqs = family(rng, 300, 4, sub=0.03, indel=0.005, inv=1)
    got = as_brute_list(O.find(seqs, pattern, O.MODE_PAIRWISE))
    want = brute.find(seqs, pattern, brute.MODE_PAIRWISE)["matches"]
    assert got == want
    got = as_brute_list(O.find(seqs, pattern, O.MODE_UNIQUE, nway_mask=0b1111))
    want = brute.find(seqs, pattern, brute.MODE_UNIQUE, nway_mask=0b1111)["matches"]
    assert got == want
    for ln, comps in got:
        assert [g for g, _ in comps] == [0, 1, 2, 3]


@pytest.mark.parametrize("seed", range(8))
def test_seed_enum_vs_brute(seed):
    rng = np.random.default_rng(400 + seed)
    pattern = PATTERNS[seed % 4]
    unit = rand_seq(rng, 40)
    s = rand_seq(rng, 100) + unit + rand_seq(rng, 50) + mutate(rng, unit, sub=0.05, indel=0) + revcomp(unit) + rand_seq(rng, 60) + "AT" * 20
    kw = dict(min_multi=2 + seed % 2, max_multi=[1000, 5, 3][seed % 3], direct_only=bool(seed & 1))
    got = as_brute_list(O.find([s], pattern, O.MODE_SEED_ENUM, **kw))
    want = brute.find([s], pattern, brute.MODE_SEED_ENUM, **kw)["matches"]
    assert got == want
    assert len(got) > 0


@pytest.mark.parametrize("seed", range(8))
def test_repeat_hash_vs_brute(seed):
    """RepeatHash (mauveAligner --repeats, src/mauveAligner.cpp:480-487): one sequence, every occurrence a component,
    extension + containment de-dup; direct, diverged and inverted copies, a tandem array and a low-complexity tail"""
    rng = np.random.default_rng(900 + seed)
    pattern = PATTERNS[seed % 4]
    unit = rand_seq(rng, 60)
    s = (rand_seq(rng, 80) + unit + rand_seq(rng, 40) + mutate(rng, unit, sub=0.04, indel=0) + rand_seq(rng, 30) + revcomp(unit) +
         rand_seq(rng, 50) + unit[:35] + rand_seq(rng, 20) + (rand_seq(rng, 7) * 6) + "AT" * 15 + rand_seq(rng, 30))
    kw = dict(min_multi=2 + seed % 2, max_multi=[1000, 6, 3][seed % 3])
    got = as_brute_list(O.find([s], pattern, O.MODE_REPEAT, **kw))
    want = brute.find([s], pattern, brute.MODE_REPEAT, **kw)["matches"]
    assert got == want
    assert len(got) > 0
    assert brute.seed_len(pattern) < 7 or any(ln > brute.seed_len(pattern) for ln, _ in got)  # something was extended
    assert O.find([s, s], pattern, O.MODE_REPEAT)["n_matches"] == 0  # one sequence only


def test_seed_enum_needs_one_sequence():
    # SeedMatchEnumerator::CreateMatches is a no-op unless seq_count == 1 (SeedMatchEnumerator.h:59-65)
    r = O.find(["ACGTACGTACGTAAC", "ACGTACGTACGTAAC"], 0b111, O.MODE_SEED_ENUM)
    assert r["n_matches"] == 0


def test_unique_single_genome_is_empty():
    rng = np.random.default_rng(3)
    s = rand_seq(rng, 500)
    r = O.find([s + s], 0b11111, O.MODE_UNIQUE)
    assert r["n_matches"] == 0  # UniqueMatchFinder.cpp:57 needs >= 2 distinct genomes


def test_short_and_empty_sequences():
    r = O.find(["ACG", "ACGTTGCA"], 0b11111, O.MODE_UNIQUE)
    assert r["n_matches"] == 0 and r["n_seeds"] == 4
    r = O.find(["", "ACGTTGCAGG"], 0b111, O.MODE_UNIQUE_COUNT)
    assert r["unique_mers_per_seq"][0] == 0


def test_invariants_on_output():
    rng = np.random.default_rng(77)
    pattern = 0b1101110111110111011
    Ls = pattern.bit_length()
    seqs = family(rng, 4000, 3, sub=0.02, indel=0.002, inv=2)
    r = O.find(seqs, pattern, O.MODE_UNIQUE)
    ms = O.matches_as_list(r)
    assert len(ms) > 10
    mers = [O.mers(s, pattern) for s in seqs]
    for ln, comps in ms:
        assert ln >= Ls and len(comps) >= 2 and comps[0][1] > 0
        # both terminal windows satisfy the window predicate (they were visited by the extension)
        for left in (True, False):
            ref = None
            for g, st in comps:
                rev = st < 0
                a = abs(st)
                pos = a if (left != rev) else a + ln - Ls
                m = int(mers[g][pos - 1])
                v = (m >> 1, (m & 1) ^ int(rev))
                ref = ref or v
                assert v == ref
    # canonical order (D18)
    keys = [([abs(dict(c).get(g, 0)) for g in range(3)], [dict(c).get(g, 0) < 0 for g in range(3)], ln) for ln, c in ms]
    assert keys == sorted(keys)


def test_genome_permutation_equivariance_two_genomes():
    rng = np.random.default_rng(5150)
    pattern = 0b110111011
    a, b = family(rng, 1500, 2, sub=0.03, indel=0.003, inv=0)
    r1 = O.matches_as_list(O.find([a, b], pattern, O.MODE_UNIQUE))
    r2 = O.matches_as_list(O.find([b, a], pattern, O.MODE_UNIQUE))
    # swapping genomes swaps columns; the first present genome is always the positive one
    s1 = sorted((ln, tuple(sorted((g, s) for g, s in c))) for ln, c in r1)
    def swap(c):
        d = sorted((1 - g, s) for g, s in c)
        return tuple((g, -s) for g, s in d) if d[0][1] < 0 else tuple(d)

    s2 = sorted((ln, swap(c)) for ln, c in r2)
    assert s1 == s2


def _normalised(ln, comps):
    """a match as a set element: components by genome, the first one positive (orientation is relative to it)"""
    d = sorted(comps)
    if d[0][1] < 0:
        d = [(g, -s) for g, s in d]
    return (ln, tuple(d))


@pytest.mark.parametrize("seed", range(3))
def test_reverse_complement_equivariance(seed):
    """Replacing one genome by its reverse complement maps the multi-MUM set onto itself: that genome's components
    change strand and their left ends mirror (n - start - length + 2); everything else is unchanged.  Holds for the first
    genome too (the whole match is then read on the other strand), because canonical seeds, the ascending-seed
    de-dup order and the four-phase extension are all strand-symmetric."""
    rng = np.random.default_rng(4000 + seed)
    for pattern in (0b110111011, 0b1101110111110111011):
        seqs = family(rng, 2500, 3, sub=0.03, indel=0.003, inv=1)
        base = sorted(_normalised(ln, c) for ln, c in O.matches_as_list(O.find(seqs, pattern, O.MODE_UNIQUE)))
        assert len(base) > 10
        for k in range(3):
            s2 = list(seqs)
            s2[k] = revcomp(seqs[k])
            n = len(seqs[k])
            back = []
            for ln, comps in O.matches_as_list(O.find(s2, pattern, O.MODE_UNIQUE)):
                back.append(_normalised(ln, [(g, (-1 if s > 0 else 1) * (n - abs(s) - ln + 2)) if g == k else (g, s) for g, s in comps]))
            assert sorted(back) == base, (bin(pattern), k)


def test_genome_permutation_equivariance_three_genomes():
    """Relabelling the genomes permutes the columns of every match and nothing else (the seed order, hence the
    de-dup order, does not depend on the labels)."""
    import itertools
    rng = np.random.default_rng(4100)
    pattern = 0b1101110111110111011
    seqs = family(rng, 2000, 3, sub=0.03, indel=0.003, inv=1)
    base = sorted(_normalised(ln, c) for ln, c in O.matches_as_list(O.find(seqs, pattern, O.MODE_UNIQUE)))
    for perm in itertools.permutations(range(3)):
        got = O.matches_as_list(O.find([seqs[p] for p in perm], pattern, O.MODE_UNIQUE))  # new genome i = old genome perm[i]
        assert sorted(_normalised(ln, [(perm[g], s) for g, s in comps]) for ln, comps in got) == base, perm


def test_property_helpers_on_scaled_baseline_configs():
    """tests/properties.py (the checks the CUDA path must pass at full size, tests/test_gpu_zz_properties.py) on the
    oracle, over the synthetic generator's C1 and C2 at reduced length."""
    import mauvealigner_b200 as mb
    import properties as P

    def rc(a):
        lut = np.arange(256, dtype=np.uint8)
        for x, y in zip(b"ACGT", b"TGCA"):
            lut[x] = y
        return np.ascontiguousarray(lut[np.asarray(a, dtype=np.uint8)][::-1])

    seqs = mb.synth_genomes(1, 50)
    pat = mb.get_seed(15, 0)
    assert P.check_reverse_complement_equivariance(lambda s: O.find(s, pat, O.MODE_UNIQUE), seqs, rc) > 1000
    seqs = mb.synth_genomes(2, 100)
    pat = mb.get_seed(15, mb.CODING_SEED)
    find = lambda s: O.find(s, pat, O.MODE_UNIQUE)  # noqa: E731
    assert P.check_reverse_complement_equivariance(find, seqs, rc, genomes=(0, 5)) > 5000
    assert P.check_permutation_equivariance(find, seqs, (3, 1, 7, 0, 2, 6, 5, 4)) > 5000


@pytest.mark.parametrize("seed", range(6))
def test_seed_family_vs_brute(seed):
    """Seed-family search (src/progressiveMauve.cpp:503-548): one persistent de-dup table across
    the patterns, longest seed first.  Oracle against the naive restatement; the GPU path does not have this mode yet."""
    rng = np.random.default_rng(7000 + seed)
    seqs = family(rng, 900, 3, sub=0.04, indel=0.004, inv=1 if seed % 2 else 0)
    if seed == 3:
        seqs[2] = revcomp(seqs[2])
    fams = [[0b1101110111110111011, 0b110111011, 0b11111], [0b110010110011010011, 0b1011101, 0b111]]
    # palindromic patterns of odd weight only
    pats = [p for p in fams[seed % 2] if O.lib().orc_seed_valid(p)]
    assert len(pats) >= 2
    pats.sort(key=lambda p: -p.bit_length())
    got = as_brute_list(O.find_family(seqs, pats))
    want = brute.find_family(seqs, pats)["matches"]
    assert got == want
    assert len(got) > 5
    # one pattern alone = the plain search; the family finds at least the matches of its first pattern
    single = as_brute_list(O.find(seqs, pats[0], O.MODE_UNIQUE))
    assert as_brute_list(O.find_family(seqs, pats[:1])) == single
    assert set((ln, tuple(c)) for ln, c in single) <= set((ln, tuple(c)) for ln, c in got)


def test_seed_family_later_patterns_are_contained():
    """two identical genomes: the first (longest) pattern's single match contains every candidate of the later ones"""
    rng = np.random.default_rng(11)
    s = rand_seq(rng, 600)
    pats = [0b1101110111110111011, 0b110111011, 0b111]
    r = O.find_family([s, s], pats)
    assert as_brute_list(r) == [(600, [(0, 1), (1, 1)])]
    assert r["n_contained"] > 1000
