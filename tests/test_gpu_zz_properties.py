"""Full-size parity through size-independent properties (run last): the CUDA path at BASELINE sizes must map the
multi-MUM set onto itself under a reverse complement of one genome and under a relabelling of the genomes — exactly
what the oracle does on small inputs (tests/test_oracle.py), where the two are compared bit for bit."""
import numpy as np
import pytest

import properties as P

pytestmark = pytest.mark.gpu


def _revcomp_u8(a):
    lut = np.arange(256, dtype=np.uint8)
    for x, y in zip(b"ACGTacgt", b"TGCAtgca"):
        lut[x] = y
    return np.ascontiguousarray(lut[np.asarray(a, dtype=np.uint8)][::-1])


def _finder(ctx, pattern, mode):
    def find(seqs):
        ctx.clear_sequences()
        for s in seqs:
            ctx.add_sequence(s)
        ctx.set_seed(pattern)
        return ctx.find(mode)
    return find


def test_full_size_c1_reverse_complement_and_permutation():
    import mauvealigner_b200 as mb
    seqs = mb.synth_genomes(1, 1)  # C1: 2 x 5 Mbp
    ctx = mb.Context(0)
    try:
        find = _finder(ctx, mb.get_seed(15, 0), mb.MODE_UNIQUE)
        n = P.check_reverse_complement_equivariance(find, seqs, _revcomp_u8)
        assert n > 100000
        P.check_permutation_equivariance(find, seqs, (1, 0))
    finally:
        ctx.close()


def test_quarter_size_c2_reverse_complement():
    import mauvealigner_b200 as mb
    seqs = mb.synth_genomes(2, 4)  # C2 at a quarter: 8 x 1.25 Mbp
    ctx = mb.Context(0)
    try:
        find = _finder(ctx, mb.get_seed(15, mb.CODING_SEED), mb.MODE_UNIQUE)
        n = P.check_reverse_complement_equivariance(find, seqs, _revcomp_u8, genomes=(0, 5))
        assert n > 100000
        P.check_permutation_equivariance(find, seqs, (3, 1, 7, 0, 2, 6, 5, 4))
    finally:
        ctx.close()
