"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.
Bit-exact: integer / index work only (no floating point on this path, no tolerance)."""
import numpy as np
import pytest

import oracle_lib as O
from toygen import family, mutate, rand_seq, revcomp

pytestmark = pytest.mark.gpu

PATTERNS = [0b111, 0b11111, 0b1011101, 0b110111011, 0b11011011111011011, 0b1101110111110111011,
            0b110110110111011011011, 0b1011101110111110111011101]


def wide_pattern():
    # L = 41, weight 9, palindromic
    bits = [0] * 41
    for j in (0, 5, 13, 19, 20, 21, 27, 35, 40):
        bits[j] = 1
    p = 0
    for b in bits:
        p = (p << 1) | b
    return p


@pytest.fixture(scope="module")
def ctx():
    import mauvealigner_b200 as mb
    c = mb.Context(0)
    yield c
    c.close()


def run(ctx, seqs, pattern, mode, **kw):
    ctx.clear_sequences()
    for s in seqs:
        ctx.add_sequence(s)
    ctx.set_seed(pattern)
    return ctx.find(mode, **kw)


def assert_same(got, want, what=""):
    assert got["n_matches"] == want["n_matches"], (what, got["n_matches"], want["n_matches"])
    assert got["n_comps"] == want["n_comps"], what
    for k in ("length", "comp_off", "comp_seq", "comp_start"):
        assert np.array_equal(np.asarray(got[k], dtype=np.int64), np.asarray(want[k], dtype=np.int64)), (what, k)


@pytest.mark.parametrize("pattern", PATTERNS + [wide_pattern()])
def test_mers_and_sml_record_by_record(ctx, pattern):
    rng = np.random.default_rng(pattern & 0xFFFF)
    seqs = [rand_seq(rng, 9000), rand_seq(rng, 40, alphabet=2), rand_seq(rng, 4097 + pattern.bit_length())]
    ctx.clear_sequences()
    for s in seqs:
        ctx.add_sequence(s)
    ctx.set_seed(pattern)
    for i, s in enumerate(seqs):
        assert np.array_equal(ctx.mers(i, len(s)), O.mers(s, pattern)), i
    ctx.find(2)
    for i, s in enumerate(seqs):
        assert np.array_equal(ctx.sml(i, len(s)), O.sml(s, pattern)), i


def test_non_acgt_and_lowercase(ctx):
    s = "ACGTNNNNacgtRYKMacgtACGTTTGACCA" * 10
    ctx.clear_sequences()
    ctx.add_sequence(s)
    ctx.set_seed(0b11111)
    assert np.array_equal(ctx.mers(0, len(s)), O.mers(s, 0b11111))


def test_unique_count(ctx):
    rng = np.random.default_rng(1)
    seqs = [rand_seq(rng, 5000, alphabet=3), rand_seq(rng, 3000), ""]
    for pattern in PATTERNS[:5]:
        got = run(ctx, seqs, pattern, 2)
        want = O.find(seqs, pattern, O.MODE_UNIQUE_COUNT)
        assert got["unique_mers"] == want["unique_mers"]
        assert got["unique_mers_per_seq"].tolist() == want["unique_mers_per_seq"].tolist()


@pytest.mark.parametrize("seed", range(16))
def test_unique_vs_oracle_small(ctx, seed):
    rng = np.random.default_rng(1000 + seed)
    pattern = PATTERNS[seed % len(PATTERNS)]
    k = int(rng.integers(2, 7))
    n = int(rng.integers(300, 3000))
    seqs = family(rng, n, k, sub=0.03, indel=0.005, inv=seed % 3)
    if seed % 4 == 0:
        seqs[1] = revcomp(seqs[1])
    got = run(ctx, seqs, pattern, 0)
    want = O.find(seqs, pattern, O.MODE_UNIQUE)
    assert_same(got, want, f"seed {seed}")
    assert got["unique_mers"] == want["unique_mers"]


@pytest.mark.parametrize("seed", range(6))
def test_unique_low_complexity(ctx, seed):
    rng = np.random.default_rng(2000 + seed)
    pattern = PATTERNS[seed % 4]
    a = rand_seq(rng, 2000, alphabet=2 + seed % 2)
    seqs = [a, mutate(rng, a, sub=0.05, indel=0.02), mutate(rng, a, sub=0.1, indel=0.0, inv=1)]
    assert_same(run(ctx, seqs, pattern, 0), O.find(seqs, pattern, O.MODE_UNIQUE))


def test_identical_and_near_identical_genomes(ctx):
    rng = np.random.default_rng(42)
    s = rand_seq(rng, 200000)
    for pattern in (0b111111111111111, 0b1101110111110111011):
        got = run(ctx, [s, s], pattern, 0)
        assert_same(got, O.find([s, s], pattern, O.MODE_UNIQUE))
        assert got["n_matches"] == 1 and got["length"][0] == len(s)
        t = s[:100000] + "A" + s[100001:] if s[100000] != "A" else s[:100000] + "C" + s[100001:]
        assert_same(run(ctx, [s, t, revcomp(s)], pattern, 0), O.find([s, t, revcomp(s)], pattern, O.MODE_UNIQUE))


def test_substitutions_only_long_diagonal(ctx):
    rng = np.random.default_rng(43)
    s = rand_seq(rng, 300000)
    t = mutate(rng, s, sub=0.02, indel=0.0)
    for pattern in (0b111111111, 0b110110110111011011011):
        assert_same(run(ctx, [s, t], pattern, 0), O.find([s, t], pattern, O.MODE_UNIQUE))


@pytest.mark.parametrize("which", range(4))
def test_wide_records_all_modes(ctx, which):
    """Seeds of weight >= 29 do not fit key + genome + position in 64 bits: 16-byte records (key word + value word)."""
    import mauvealigner_b200 as mb
    pattern = [(1 << 31) - 1, (1 << 29) - 1, mb.get_seed(31, 0), mb.get_seed(29, mb.CODING_SEED)][which]
    rng = np.random.default_rng(pattern & 0xFFFF)
    seqs = family(rng, 30000, 3, sub=0.01, indel=0.001, inv=1)
    seqs[1] = revcomp(seqs[1])
    assert_same(run(ctx, seqs, pattern, 0), O.find(seqs, pattern, O.MODE_UNIQUE), "unique")
    got = run(ctx, seqs, pattern, 2)
    want = O.find(seqs, pattern, O.MODE_UNIQUE_COUNT)
    assert got["unique_mers"] == want["unique_mers"] and got["unique_mers_per_seq"].tolist() == want["unique_mers_per_seq"].tolist()
    unit = rand_seq(rng, 400)
    s = rand_seq(rng, 5000) + unit + rand_seq(rng, 777) + unit + revcomp(unit) + rand_seq(rng, 100)
    assert_same(run(ctx, [s], pattern, 1, min_multi=2, max_multi=100), O.find([s], pattern, O.MODE_SEED_ENUM, min_multi=2, max_multi=100), "enum")
    for i, q in enumerate([s]):
        assert np.array_equal(ctx.sml(i, len(q)), O.sml(q, pattern))


@pytest.mark.parametrize("k", [2, 3, 4, 6])
def test_pairwise_vs_oracle(ctx, k):
    """PairwiseMatchFinder policy: several candidates of one bucket share their first component."""
    rng = np.random.default_rng(600 + k)
    seqs = family(rng, 6000, k, sub=0.02, indel=0.003, inv=1)
    if k > 2:
        seqs[2] = revcomp(seqs[2])
    for pattern in (0b110111011, 0b1101110111110111011):
        got = run(ctx, seqs, pattern, 3)
        want = O.find(seqs, pattern, O.MODE_PAIRWISE)
        assert_same(got, want, f"pairwise k={k}")
        assert got["n_matches"] > 0 and set(np.diff(got["comp_off"]).tolist()) == {2}


def test_nway_mask(ctx):
    rng = np.random.default_rng(44)
    seqs = family(rng, 5000, 4, sub=0.02, indel=0.002, inv=1)
    pattern = 0b110111011
    for mask in (0b1111, 0b0101, 0b0011):
        assert_same(run(ctx, seqs, pattern, 0, nway_mask=mask), O.find(seqs, pattern, O.MODE_UNIQUE, nway_mask=mask), f"mask {mask}")


def test_many_genomes(ctx):
    rng = np.random.default_rng(45)
    seqs = family(rng, 3000, 40, sub=0.01, indel=0.001, inv=0)
    pattern = 0b1101110111110111011
    assert_same(run(ctx, seqs, pattern, 0), O.find(seqs, pattern, O.MODE_UNIQUE))


@pytest.mark.parametrize("seed", range(8))
def test_seed_enum_vs_oracle(ctx, seed):
    rng = np.random.default_rng(3000 + seed)
    pattern = PATTERNS[seed % 6]
    unit = rand_seq(rng, 300)
    s = rand_seq(rng, 3000) + unit + rand_seq(rng, 500) + mutate(rng, unit, sub=0.05, indel=0) + revcomp(unit) + rand_seq(rng, 600) + "AT" * 300 + unit
    kw = dict(min_multi=2 + seed % 2, max_multi=[1000, 5, 3, 500][seed % 4], direct_only=bool(seed & 1))
    got = run(ctx, [s], pattern, 1, **kw)
    want = O.find([s], pattern, O.MODE_SEED_ENUM, **kw)
    assert_same(got, want, f"seed {seed}")
    if pattern.bit_length() > 3:
        assert got["n_matches"] > 0


@pytest.mark.parametrize("seed", range(6))
def test_repeat_hash_vs_oracle(ctx, seed):
    """MB_MODE_REPEAT (RepeatHash, src/mauveAligner.cpp:480-487) against the oracle: repeat families with diverged and
    inverted copies, tandem arrays, low-complexity tracts; multiplicity windows"""
    import mauvealigner_b200 as mb
    rng = np.random.default_rng(7100 + seed)
    units = [rand_seq(rng, int(rng.integers(60, 400))) for _ in range(6)]
    parts = []
    for _ in range(60):
        parts.append(rand_seq(rng, int(rng.integers(50, 600))))
        u = units[int(rng.integers(0, len(units)))]
        c = mutate(rng, u, sub=0.03 * (seed % 3), indel=0.002 * (seed % 2))
        parts.append(revcomp(c) if rng.random() < 0.3 else c)
    parts.append(rand_seq(rng, 9) * 12 + "ACACACACACACACACACACACACACACAC" + "A" * 40)
    s = "".join(parts)
    pattern = [mb.get_seed(9, 0), mb.get_seed(11, 1), mb.get_seed(11, 2), mb.get_seed(13, 0), mb.get_seed(9, 2), 0b1011101][seed]
    kw = dict(min_multi=2 + seed % 2, max_multi=[255, 1000, 4, 17, 255, 40][seed])
    got = run(ctx, [s], pattern, mb.MODE_REPEAT, **kw)
    want = O.find([s], pattern, O.MODE_REPEAT, **kw)
    assert_same(got, want, f"repeat {seed}")
    assert want["n_matches"] > 0 or seed == 5
    # RepeatHash through the Python mirror
    ml = mb.MatchList()
    ml.seq_table = [s]
    ml.seed_pattern = pattern
    rh = mb.RepeatHash()
    assert rh.FindMatches(ml, **kw)
    assert [(m.Length(), [m.Start(i) for i in range(m.SeqCount())]) for m in ml] == [(ln, [st for _, st in comps]) for ln, comps in O.matches_as_list(want)]


def test_position_lookup_table(ctx):
    """mb_position_table against repeatoire's own three steps (src/repeatoire.cpp:1920-1966) applied to the oracle's list:
    records sorted by LeftEnd(0), components sorted by left end, table[left end] = (record, component)"""
    import mauvealigner_b200 as mb
    rng = np.random.default_rng(99)
    u = rand_seq(rng, 300)
    s = rand_seq(rng, 2000) + u + rand_seq(rng, 500) + revcomp(u) + rand_seq(rng, 800) + mutate(rng, u, sub=0.02, indel=0) + "ACG" * 30
    for mode, kw in ((mb.MODE_SEED_ENUM, dict(min_multi=2, max_multi=500)), (mb.MODE_REPEAT, dict(max_multi=255))):
        got = run(ctx, [s], mb.get_seed(9, 0), mode, **kw)
        want = O.find([s], mb.get_seed(9, 0), mode, **kw)
        assert_same(got, want)
        recs = sorted(((abs(comps[0][1]), j) for j, (_, comps) in enumerate(O.matches_as_list(want))))  # seed_sort_list
        mplt = sorted((abs(st), order, k) for order, (_, j) in enumerate(recs) for k, (_, st) in enumerate(O.matches_as_list(want)[j][1]))
        exp_m = np.full(len(s) + 1, 0xFFFFFFFF, dtype=np.uint32)
        exp_c = np.full(len(s) + 1, 0xFFFFFFFF, dtype=np.uint32)
        for left, order, k in mplt:
            exp_m[left], exp_c[left] = order, k
        mo, co = ctx.position_table()
        assert np.array_equal(mo, exp_m) and np.array_equal(co, exp_c), (mode, np.flatnonzero(mo != exp_m)[:5], np.flatnonzero(co != exp_c)[:5])
        assert (mo != 0xFFFFFFFF).sum() == want["n_comps"] or mode == mb.MODE_REPEAT


def test_compact_result_form(ctx):
    """mb_fetch_result_compact / mb_find_compact: the layout the device keeps (u8 sequence, i32 start) carries the same
    values as the wide arrays of mb_result"""
    import mauvealigner_b200 as mb
    rng = np.random.default_rng(55)
    seqs = family(rng, 20000, 5, sub=0.02, indel=0.003, inv=1)
    seqs[2] = revcomp(seqs[2])
    wide = run(ctx, seqs, mb.get_seed(11, 0), mb.MODE_UNIQUE)
    compact = ctx.fetch(compact=True)
    assert compact["comp_seq"].dtype == np.uint8 and compact["comp_start"].dtype == np.int32
    assert_same(compact, wide)
    again = ctx.find(mb.MODE_UNIQUE, compact=True)
    assert_same(again, O.find(seqs, mb.get_seed(11, 0), O.MODE_UNIQUE))
    assert (np.asarray(again["comp_start"]) < 0).any()


def test_edge_cases(ctx):
    import mauvealigner_b200 as mb
    # sequences shorter than the seed, empty sequences
    got = run(ctx, ["ACG", "ACGTTGCA", ""], 0b11111, 0)
    assert got["n_matches"] == 0
    # SEED_ENUM needs exactly one sequence
    with pytest.raises(mb.MauveError):
        run(ctx, ["ACGTACGTACGTAAC", "ACGTACGTACGTAAC"], 0b111, 1)
    # invalid seeds are rejected
    with pytest.raises(mb.MauveError):
        ctx.set_seed(0b1111)
    with pytest.raises(mb.MauveError):
        ctx.set_seed(0b1101)
    # a single genome can never produce a UNIQUE match
    rng = np.random.default_rng(3)
    s = rand_seq(rng, 5000)
    assert run(ctx, [s + s], 0b11111, 0)["n_matches"] == 0


@pytest.mark.parametrize("config,scale", [(1, 50), (2, 50), (3, 200), (4, 400), (5, 100)])
def test_baseline_configs_scaled(ctx, config, scale):
    import mauvealigner_b200 as mb
    seqs = mb.synth_genomes(config, scale)
    if config in (1,):
        pattern, mode, kw = mb.get_seed(15, 0), 0, {}
    elif config in (2, 5):
        pattern, mode, kw = mb.get_seed(15, mb.CODING_SEED), 0, {}
    elif config == 3:
        pattern, mode, kw = mb.get_seed(19, 0), 2, {}
    else:
        pattern, mode, kw = mb.get_seed(15, 0), 1, dict(min_multi=2, max_multi=500)
    got = run(ctx, seqs, pattern, mode, **kw)
    want = O.find(seqs, pattern, mode, **kw)
    assert_same(got, want, f"C{config}")
    assert got["unique_mers"] == want["unique_mers"]
    if config == 3:
        assert got["unique_mers_per_seq"].tolist() == want["unique_mers_per_seq"].tolist()
        # weight-19 seeds on one genome: the UNIQUE match set is empty by construction (UniqueMatchFinder.cpp:57)
        assert run(ctx, seqs, pattern, 0)["n_matches"] == 0


def test_cpp_host_api_matches_oracle(tmp_path):
    """The reference-shaped C++ classes (include/mems_compat) drive the same C ABI: compile the little driver,
    run UniqueMatchFinder / SeedMatchEnumerator / UniqueMerCount, compare the printed lists with the oracle."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "compat_driver"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(root, "include"),
                           os.path.join(root, "tests", "cpp", "compat_driver.cpp"), "-o", str(exe),
                           "-L", os.path.join(root, "mauvealigner_b200"), "-lmauve_b200",
                           "-Wl,-rpath," + os.path.join(root, "mauvealigner_b200")])
    import mauvealigner_b200 as mb
    rng = np.random.default_rng(77)
    seqs = family(rng, 6000, 3, sub=0.02, indel=0.002, inv=1)
    files = []
    for i, s in enumerate(seqs):
        f = tmp_path / f"s{i}.txt"
        f.write_text(s + "\n")
        files.append(str(f))

    def parse(out):
        lines = out.strip().split("\n")
        k = [i for i, l in enumerate(lines) if l.startswith("MatchCount")][0]
        assert int(lines[k].split()[1]) == len(lines) - k - 1
        return [[int(x) for x in l.split("\t")] for l in lines[k + 1:]]

    out = subprocess.check_output([str(exe), "umf", "11", "0"] + files, text=True)
    want = O.find(seqs, mb.get_seed(11, 0), O.MODE_UNIQUE)
    rows = []
    for ln, comps in O.matches_as_list(want):
        d = dict(comps)
        rows.append([ln] + [d.get(g, 0) for g in range(3)])
    assert parse(out) == rows and len(rows) > 10
    out = subprocess.check_output([str(exe), "sme", "9", "0", files[0]], text=True)
    want = O.find(seqs[:1], mb.get_seed(9, 0), O.MODE_SEED_ENUM, min_multi=2, max_multi=500)
    assert parse(out) == [[ln] + [s for _, s in comps] for ln, comps in O.matches_as_list(want)]
    out = subprocess.check_output([str(exe), "repeats", "9", "0", files[0]], text=True)
    want = O.find(seqs[:1], mb.get_seed(9, 0), O.MODE_REPEAT, min_multi=2, max_multi=255)
    assert parse(out) == [[ln] + [s for _, s in comps] for ln, comps in O.matches_as_list(want)]
    out = subprocess.check_output([str(exe), "count", "9", "0", files[0]], text=True)
    assert int(out.split()[-1]) == O.find(seqs[:1], mb.get_seed(9, 0), O.MODE_UNIQUE_COUNT)["unique_mers"]
    # the same UniqueMatchFinder over several contexts (MAUVE_B200_DEVICES: here three ranks on device 0) = mb_find_multi
    multi = subprocess.check_output([str(exe), "umf", "11", "0"] + files, text=True, env=dict(os.environ, MAUVE_B200_DEVICES="0,0,0"))
    assert parse(multi) == rows
    # PairwiseMatchFinder
    out = subprocess.check_output([str(exe), "pmf", "11", "0"] + files, text=True)
    want = O.find(seqs, mb.get_seed(11, 0), O.MODE_PAIRWISE)
    rows = []
    for ln, comps in O.matches_as_list(want):
        d = dict(comps)
        rows.append([ln] + [d.get(g, 0) for g in range(3)])
    assert parse(out) == rows and len(rows) > 10
    # .sslist files: LoadSMLs creates them (second call loads), DNAFileSML::LoadFile + UniqueMerCount reads them back
    res = subprocess.run([str(exe), "sml", "11", "0"] + files, text=True, capture_output=True, check=True)
    assert res.stderr.count("Creating sorted mer list") == 3
    uniq = O.find(seqs[:1], mb.get_seed(11, 0), O.MODE_UNIQUE_COUNT)["unique_mers"]
    assert int(res.stdout.split()[-1]) == uniq
    out = subprocess.check_output([str(exe), "smlcount", files[0] + ".sslist"], text=True).split()
    assert int(out[0]) == uniq
    assert [int(x) for x in out[1:]] == O.sml(seqs[0], mb.get_seed(11, 0)).tolist()
    # the seed-family search, call for call as src/progressiveMauve.cpp:503-548 (one finder, three patterns, longest first)
    out = subprocess.check_output([str(exe), "family", "11", "0"] + files, text=True)
    pats = [mb.get_seed(11, r) for _, r in sorted(((mb.seed_length(mb.get_seed(11, r)), r) for r in (0, 1, 2)), reverse=True)]
    want = O.find_family(seqs, pats)
    rows = []
    for ln, comps in O.matches_as_list(want):
        d = dict(comps)
        rows.append([ln] + [d.get(g, 0) for g in range(3)])
    assert parse(out) == rows and len(rows) > 10
    # many small problems in one pass through the C++ mirror (MemHash::FindMatchesBatch): pieces of 400 bases as "gaps"
    out = subprocess.check_output([str(exe), "batch", "7", "0"] + files, text=True).strip().split("\n")
    k, gap = 0, 0
    while k < len(out):
        tag, gi, n = out[k].split("\t")
        assert tag == "Gap" and int(gi) == gap
        rows_g = [[int(x) for x in l.split("\t")] for l in out[k + 1:k + 1 + int(n)]]
        pieces = [s[gap * 400:(gap + 1) * 400] for s in seqs]
        want = O.find(pieces, mb.get_seed(7, 0), O.MODE_UNIQUE)
        exp = []
        for ln, comps in O.matches_as_list(want):
            d = dict(comps)
            exp.append([ln] + [d.get(g, 0) for g in range(3)])
        assert rows_g == exp, gap
        k += 1 + int(n)
        gap += 1
    assert gap == max((len(s) + 399) // 400 for s in seqs)
    # WriteList -> ReadList -> WriteList round trip
    first = subprocess.check_output([str(exe), "umf", "11", "0"] + files, text=True)
    lst = tmp_path / "matches.mums"
    lst.write_text(first)
    assert subprocess.check_output([str(exe), "readlist", str(lst)], text=True) == first


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_seed_family_vs_oracle(seed):
    """One finder, several seed patterns, the table persisting across the calls (src/progressiveMauve.cpp:503-548):
    the device filter + the host union against orc_find_family, for families of equal and of different seed lengths."""
    import mauvealigner_b200 as mb
    rng = np.random.default_rng(4200 + seed)
    k = 2 + seed
    seqs = family(rng, 8000, k, sub=0.03, indel=0.004, inv=seed % 2)
    if seed == 3:
        seqs[1] = seqs[0]  # identical genomes: long matches of the first pattern contain almost every later candidate
    families = [[mb.get_seed(11, 2), mb.get_seed(11, 1), mb.get_seed(11, 0)],
                [mb.get_seed(13, 0), mb.get_seed(9, 2), mb.get_seed(11, 1)],   # not ordered by length
                [mb.get_seed(9, 0), mb.get_seed(13, 1)],
                [mb.get_seed(9, 0), mb.get_seed(9, 1), mb.get_seed(9, 2), mb.get_seed(15, 0), 0b11111]][seed]
    want = O.find_family(seqs, families)
    ml = mb.MatchList()
    ml.seq_table = list(seqs)
    umf = mb.UniqueMatchFinder()
    for p in families:
        ml.seed_pattern = p
        umf.FindMatches(ml)
        umf.ClearSequences()
    out = mb.MatchList()
    umf.GetMatchList(out)
    got = [(m.Length(), [(g, m.Start(g)) for g in range(k) if m.Start(g) != 0]) for m in out]
    assert got == [(ln, list(comps)) for ln, comps in O.matches_as_list(want)]
    assert len(got) > 5
    if k >= 3:  # MaskedMemHash over a family: only matches present in exactly the masked genomes, table shared across the patterns
        mask = 0b101
        mmh = mb.MaskedMemHash()
        mmh.SetMask(mask)
        for p in families:
            ml.seed_pattern = p
            mmh.FindMatches(ml)
            mmh.ClearSequences()
        out = mb.MatchList()
        mmh.GetMatchList(out)
        want_m = O.find_family(seqs, families, nway_mask=mask)
        assert [(m.Length(), [(g, m.Start(g)) for g in range(k) if m.Start(g) != 0]) for m in out] == [(ln, list(c)) for ln, c in O.matches_as_list(want_m)]
    # Clear() forgets the table: the next search is a plain single-pattern one again
    umf.Clear()
    ml.seed_pattern = families[-1]
    umf.FindMatches(ml)
    single = O.find(seqs, families[-1], O.MODE_UNIQUE)
    assert [(m.Length(), [(g, m.Start(g)) for g in range(k) if m.Start(g) != 0]) for m in ml] == [(ln, list(c)) for ln, c in O.matches_as_list(single)]


def test_find_batch_many_small_problems_one_pass():
    """mb_find_batch (recursive anchoring: thousands of tiny searches, src/mauveAligner.cpp:94,698): all problems in ONE
    pass of the pipeline; every problem's result equals its own search (oracle), whatever its neighbours are"""
    import mauvealigner_b200 as mb
    rng = np.random.default_rng(314)
    ctx = mb.Context(0)
    for mode, k, pattern in ((mb.MODE_UNIQUE, 3, 0b1011101), (mb.MODE_UNIQUE, 2, mb.get_seed(9, 0)), (mb.MODE_PAIRWISE, 3, 0b110111011),
                             (mb.MODE_UNIQUE, 5, 0b1101011)):
        problems = []
        for i in range(150):
            n = int(rng.integers(0, 700)) if i % 17 else int(rng.integers(0, 12))  # some shorter than the seed, some empty
            seqs = family(rng, n, k, sub=0.04, indel=0.006, inv=i % 2) if n > 40 else [rand_seq(rng, n) for _ in range(k)]
            if i % 11 == 3:
                seqs[k - 1] = ""  # a problem without its last sequence
            if i % 13 == 5:
                seqs[0] = revcomp(seqs[0])
            if i % 29 == 7 and i > 0:
                seqs = list(problems[i - 1])  # the same problem twice, side by side: seeds must not meet across problems
            problems.append(seqs)
        ctx.set_seed(pattern)
        got = ctx.find_batch(problems, mode)
        assert len(got) == len(problems)
        total = 0
        for i, seqs in enumerate(problems):
            want = O.find(seqs, pattern, mode)
            assert_same(got[i], want, f"batch problem {i} mode {mode}")
            total += want["n_matches"]
        assert total > 200
    # one sequence per problem: repeats (MB_MODE_REPEAT) and seed enumeration
    singles = [[rand_seq(rng, 100) + u + rand_seq(rng, 30) + u + revcomp(u)] for u in (rand_seq(rng, 50) for _ in range(40))]
    ctx.set_seed(0b1011101)
    for mode, kw in ((mb.MODE_REPEAT, dict(max_multi=255)), (mb.MODE_SEED_ENUM, dict(max_multi=500))):
        got = ctx.find_batch(singles, mode, **kw)
        for i, seqs in enumerate(singles):
            assert_same(got[i], O.find(seqs, 0b1011101, mode, **kw), f"batch single {i} mode {mode}")
    ctx.close()


def test_long_consistent_runs_share_their_walks(ctx):
    """The 'cliff' of DESIGN.md §4: thousands of reps of ONE group inside one very long window-consistent run (identical
    genomes; two identical genomes among diverged ones; an exact duplication inside otherwise diverged genomes).  Reps
    whose seeds lie a multiple of the seed length apart share one walk (k_extend_long_classes): bit-exact against the
    oracle, and in bounded time."""
    import mauvealigner_b200 as mb
    rng = np.random.default_rng(2024)
    a = rand_seq(rng, 150000)
    dup = rand_seq(rng, 40000)
    cases = {"2 identical": [a, a],
             "2 identical + 2 at 3 %": [a, a, mutate(rng, a, sub=0.03, indel=0.0), mutate(rng, a, sub=0.03, indel=0.0)],
             "reverse-complemented copy": [a, revcomp(a)],
             "1 SNP per 4 kb": [a, mutate(rng, a, sub=0.00025, indel=0.0)],  # many runs: several walks per class, extents that differ by residue
             "1 SNP per 4 kb, three-way": [a, mutate(rng, a, sub=0.00025, indel=0.0), mutate(rng, a, sub=0.00025, indel=0.0)],
             "exact duplication in diverged genomes": [rand_seq(rng, 30000) + dup + rand_seq(rng, 20000), rand_seq(rng, 10000) + dup + rand_seq(rng, 45000),
                                                       rand_seq(rng, 50000) + mutate(rng, dup, sub=0.002, indel=0.0) + rand_seq(rng, 5000)]}
    import os
    for pattern in (mb.get_seed(15, 0), mb.get_seed(11, 1)):
        for name, seqs in cases.items():
            want = O.find(seqs, pattern, O.MODE_UNIQUE)
            for knob in (None, "1"):  # the default threshold (lists of >= 256 long reps share walks), then every long rep
                if knob:
                    os.environ["MB_LONG_CLASSES_MIN"] = knob
                try:
                    got = run(ctx, seqs, pattern, mb.MODE_UNIQUE)
                finally:
                    os.environ.pop("MB_LONG_CLASSES_MIN", None)
                st = ctx.stats()
                assert_same(got, want, name)
                assert st["ms_total_device"] < 60.0, (name, st["ms_total_device"], st["n_extended"])
    # the same with the table persisting across two patterns: the second search's long runs hold reps dropped before the de-dup
    seqs = cases["2 identical + 2 at 3 %"]
    fam = [mb.get_seed(15, 0), mb.get_seed(11, 1)]
    want = O.find_family(seqs, fam)
    ml = mb.MatchList()
    ml.seq_table = list(seqs)
    umf = mb.UniqueMatchFinder()
    for p in fam:
        ml.seed_pattern = p
        umf.FindMatches(ml)
        umf.ClearSequences()
    out = mb.MatchList()
    umf.GetMatchList(out)
    assert [(m.Length(), [(g, m.Start(g)) for g in range(4) if m.Start(g) != 0]) for m in out] == [(ln, list(c)) for ln, c in O.matches_as_list(want)]


def test_randomised_parity_sweep():
    """tools/fuzz_parity.py: random genome counts, lengths around the kernels' tile sizes, random valid seed patterns, all
    policies — 150 cases here (4000 were run for profiles/r02_fuzz_parity.txt)"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_parity.py"), "150", "3"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "0 mismatches" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
    # and with every unfinished extension sent through the shared long walks (normally only lists of 256 reps and more)
    env = dict(os.environ, MB_LONG_CLASSES_MIN="1")
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_parity.py"), "100", "11"], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0 and "0 mismatches" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_context_pool_many_small_problems():
    """f4 (many small problems, one search per inter-anchor gap): a pool of contexts on concurrent streams gives the
    results of the one-at-a-time searches, in order, for problems of mixed sizes (including empty ones)."""
    import mauvealigner_b200 as mb
    rng = np.random.default_rng(2024)
    problems = []
    for i in range(40):
        n = int(rng.integers(0, 3000))
        k = int(rng.integers(2, 5))
        seqs = family(rng, n, k, sub=0.03, indel=0.004, inv=0) if n > 50 else [rand_seq(rng, n) for _ in range(k)]
        if i % 5 == 0:
            seqs[-1] = revcomp(seqs[-1])
        problems.append(seqs)
    pattern = 0b110111011
    pool = mb.ContextPool(6)
    try:
        got = pool.find_many(problems, pattern, mb.MODE_UNIQUE)
        again = pool.find_many(problems[::-1], pattern, mb.MODE_UNIQUE)[::-1]
    finally:
        pool.close()
    for seqs, g, g2 in zip(problems, got, again):
        want = O.find(seqs, pattern, O.MODE_UNIQUE)
        assert_same(g, want, "pool")
        assert_same(g2, want, "pool, reversed order")
