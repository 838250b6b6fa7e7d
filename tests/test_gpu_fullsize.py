"""Full-size parity properties (BASELINE.json sizes, where the CPU oracle would take minutes): size-independent checks
of the CUDA result — canonical order, well-formed rows, window consistency and maximality of sampled matches straight
from the ASCII genomes, run-to-run identity, and identity of the single-GPU and the (emulated) multi-rank paths."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

COMP = np.zeros(256, dtype=np.uint8)
for a, b in zip(b"ACGT", b"TGCA"):
    COMP[a] = b


def care_offsets(pattern):
    L = pattern.bit_length()
    return L, np.array([j for j in range(L) if (pattern >> (L - 1 - j)) & 1])


def oriented_window(seqs, lens, g, start, length, L, o):
    """L bases of component (g, signed 1-based start) at match offset o, in match orientation; None if outside."""
    left = abs(int(start)) - 1
    if start > 0:
        a = left + o
        if a < 0 or a + L > lens[g]:
            return None
        return seqs[g][a:a + L]
    a = left + (length - L - o)
    if a < 0 or a + L > lens[g]:
        return None
    return COMP[seqs[g][a:a + L]][::-1]


def window_ok(seqs, lens, comps, length, L, care, o):
    """True / False = all components agree / disagree on the cared columns of the window at offset o; None = the window
    leaves some sequence."""
    ws = [oriented_window(seqs, lens, g, st, length, L, o) for g, st in comps]
    if any(w is None for w in ws):
        return None
    return all(np.array_equal(ws[0][care], w[care]) for w in ws[1:])


def check_result(seqs, pattern, res, sample=3000, seed=1):
    L, care = care_offsets(pattern)
    lens = [len(s) for s in seqs]
    off = res["comp_off"].astype(np.int64)
    n = res["n_matches"]
    assert n > 0 and off[0] == 0 and off[-1] == res["n_comps"]
    mult = np.diff(off)
    assert mult.min() >= 2
    first = off[:-1]
    f = res["comp_seq"][first].astype(np.int64)
    st = res["comp_start"][first]
    assert (st > 0).all()  # the first component is the forward one (SetDirection)
    # canonical order (D18): larger first-genome index first, then ascending start of the first component
    key = (-f) * (1 << 40) + st
    assert (np.diff(key) >= 0).all()
    assert (res["length"] >= L).all()
    # components of a match name strictly increasing sequences
    inner = np.ones(res["n_comps"], dtype=bool)
    inner[first] = False
    assert (np.diff(res["comp_seq"].astype(np.int64))[inner[1:]] > 0).all()
    rng = np.random.default_rng(seed)
    for i in rng.choice(n, size=min(sample, n), replace=False):
        a, b = off[i], off[i + 1]
        comps = list(zip(res["comp_seq"][a:b].tolist(), res["comp_start"][a:b].tolist()))
        length = int(res["length"][i])
        for g, s in comps:
            assert 1 <= abs(s) and abs(s) - 1 + length <= lens[g]
        # the first and the last window of the match are consistent, the windows one step outside are not (or leave
        # a sequence): the extent is bounded by a failing window or a sequence end on both sides
        assert window_ok(seqs, lens, comps, length, L, care, 0) is True
        assert window_ok(seqs, lens, comps, length, L, care, length - L) is True
        assert window_ok(seqs, lens, comps, length, L, care, -1) in (False, None)
        assert window_ok(seqs, lens, comps, length, L, care, length - L + 1) in (False, None)


def same(a, b):
    return a["n_matches"] == b["n_matches"] and all(np.array_equal(a[k], b[k]) for k in ("length", "comp_off", "comp_seq", "comp_start"))


@pytest.mark.parametrize("config", [1, 2, 5])
def test_full_size_properties(config):
    import mauvealigner_b200 as mb
    from mauvealigner_b200 import dist
    seqs = mb.synth_genomes(config, 1)
    pattern = mb.get_seed(15, 0) if config == 1 else mb.get_seed(15, mb.CODING_SEED)
    ctx = mb.Context(0)
    try:
        for s in seqs:
            ctx.add_sequence(s)
        ctx.set_seed(pattern)
        r1 = ctx.find(mb.MODE_UNIQUE)
        r2 = ctx.find(mb.MODE_UNIQUE)
    finally:
        ctx.close()
    assert same(r1, r2)
    check_result(seqs, pattern, r1)
    # the multi-rank path (two ranks emulated on this GPU) returns the same CSR
    r3 = dist.find_unique_emulated(seqs, pattern, 2)
    assert same(r1, r3)


def test_full_size_unique_count_and_enum_scaled():
    """C3 / C4 at a quarter of their size: counts are consistent between modes and runs."""
    import mauvealigner_b200 as mb
    ctx = mb.Context(0)
    try:
        seqs = mb.synth_genomes(3, 4)
        ctx.add_sequence(seqs[0])
        ctx.set_seed(mb.get_seed(19, 0))
        r = ctx.find(mb.MODE_UNIQUE_COUNT)
        n_seeds = len(seqs[0]) - mb.seed_length(mb.get_seed(19, 0)) + 1
        assert 0 < r["unique_mers"] <= n_seeds and int(r["unique_mers_per_seq"][0]) == r["unique_mers"]
        # every position is in exactly one run: the sorted mer list is a permutation
        sml = ctx.sml(0, len(seqs[0]))
        assert sml.size == n_seeds and np.array_equal(np.sort(sml), np.arange(n_seeds, dtype=sml.dtype))
        ctx.clear_sequences()
        seqs = mb.synth_genomes(4, 4)
        ctx.add_sequence(seqs[0])
        ctx.set_seed(mb.get_seed(15, 0))
        e = ctx.find(mb.MODE_SEED_ENUM, min_multi=2, max_multi=500)
        mult = np.diff(e["comp_off"].astype(np.int64))
        assert e["n_matches"] > 0 and mult.min() >= 2 and mult.max() <= 500
        first = e["comp_off"][:-1].astype(np.int64)
        assert (np.diff(e["comp_start"][first]) > 0).all()  # canonical order: by first position, unique per bucket
        assert (e["comp_start"][first] > 0).all()
    finally:
        ctx.close()
