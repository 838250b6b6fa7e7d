"""Size-independent properties of the multi-MUM set, written over a `find(seqs) -> result dict` callable so that the
same check runs on the CPU oracle (small inputs, tests/test_oracle.py) and on the CUDA path at the BASELINE sizes."""
import numpy as np


def match_set(res, relabel=None, mirror=None):
    """The result as a sorted list of (length, ((genome, signed start), ...)) with the first component positive.
    relabel: genome -> genome map applied to the components; mirror = (genome k, its length n): components of genome k
    are mapped back from reverse-complement coordinates (other strand, left end n - start - length + 2)."""
    off = np.asarray(res["comp_off"], dtype=np.int64)
    seq = np.asarray(res["comp_seq"], dtype=np.int64).copy()
    start = np.asarray(res["comp_start"], dtype=np.int64).copy()
    length = np.asarray(res["length"], dtype=np.int64)
    n = int(res["n_matches"])
    per_comp_len = np.repeat(length, np.diff(off))
    if mirror is not None:
        k, nk = mirror
        sel = seq == k
        start[sel] = -np.sign(start[sel]) * (nk - np.abs(start[sel]) - per_comp_len[sel] + 2)
    if relabel is not None:
        seq = np.asarray(relabel, dtype=np.int64)[seq]
    out = []
    for i in range(n):
        a, b = int(off[i]), int(off[i + 1])
        comps = sorted(zip(seq[a:b].tolist(), start[a:b].tolist()))
        if comps[0][1] < 0:
            comps = [(g, -s) for g, s in comps]
        out.append((int(length[i]), tuple(comps)))
    out.sort()
    return out


def check_reverse_complement_equivariance(find, seqs, revcomp, genomes=None):
    """Replacing genome k by its reverse complement maps the match set onto itself (k's components change strand and
    mirror their left ends).  Returns the number of matches."""
    base = match_set(find(seqs))
    for k in (range(len(seqs)) if genomes is None else genomes):
        s2 = list(seqs)
        s2[k] = revcomp(seqs[k])
        got = match_set(find(s2), mirror=(k, len(seqs[k])))
        assert len(got) == len(base), (k, len(got), len(base))
        assert got == base, k
    return len(base)


def check_permutation_equivariance(find, seqs, perm):
    """new genome i = old genome perm[i]: the columns of every match are permuted and nothing else changes."""
    base = match_set(find(seqs))
    got = match_set(find([seqs[p] for p in perm]), relabel=list(perm))
    assert got == base, perm
    return len(base)
