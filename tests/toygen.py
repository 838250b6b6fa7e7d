"""Small deterministic toy-genome helpers for the CPU tests (numpy only)."""
import numpy as np

_ALPHA = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = {ord("A"): ord("T"), ord("C"): ord("G"), ord("G"): ord("C"), ord("T"): ord("A")}


def rand_seq(rng, n, alphabet=4):
    return _ALPHA[rng.integers(0, alphabet, size=n)].tobytes().decode()


def revcomp(s):
    return s[::-1].translate(str.maketrans("ACGT", "TGCA"))


def mutate(rng, s, sub=0.02, indel=0.002, inv=0, maxindel=5):
    a = list(s)
    out = []
    i = 0
    while i < len(a):
        r = rng.random()
        if r < sub:
            out.append("ACGT"[(("ACGT".index(a[i])) + int(rng.integers(1, 4))) % 4])
            i += 1
        elif r < sub + indel:
            k = int(rng.integers(1, maxindel + 1))
            if rng.random() < 0.5:
                i += k
            else:
                out.extend(rand_seq(rng, k))
        else:
            out.append(a[i])
            i += 1
    t = "".join(out)
    for _ in range(inv):
        if len(t) < 20:
            break
        x = int(rng.integers(0, len(t) - 10))
        y = int(rng.integers(x + 5, min(len(t), x + max(6, len(t) // 3)) + 1))
        t = t[:x] + revcomp(t[x:y]) + t[y:]
    return t


def family(rng, n, k, **kw):
    anc = rand_seq(rng, n)
    return [anc] + [mutate(rng, anc, **kw) for _ in range(k - 1)]
