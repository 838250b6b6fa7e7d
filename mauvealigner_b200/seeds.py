"""Seed pattern selection — Python mirror of include/mauve_b200/seed_masks.h (libMems SeedMasks.h stand-in:
getSeed / getSeedLength / getDefaultSeedWeight, SOLID_SEED, CODING_SEED; call sites
/root/reference/src/mauveAligner.cpp:266-279,465 and src/progressiveMauve.cpp:215-224,446-451).
The table is this project's own (libMems is not in the reference tree); every API takes the raw pattern."""
import math

SOLID_SEED = 2 ** 31 - 1
CODING_SEED = 3


def _gap(rank, d):
    if rank == 0:
        return d % 4 == 3
    if rank == 1:
        return d % 5 in (2, 4)
    if rank == 2:
        return d % 7 in (2, 4, 5)
    if rank == CODING_SEED:
        return d % 3 == 2
    return False


def get_seed(weight, rank=0):
    if weight < 3:
        return 0
    weight = min(weight, 31)
    if weight % 2 == 0:
        weight -= 1
    if rank not in (0, 1, 2, CODING_SEED):
        rank = SOLID_SEED
    half, placed, d, gaps = (weight - 1) // 2, 0, 0, []
    while placed < half:
        d += 1
        g = _gap(rank, d)
        gaps.append(g)
        placed += 0 if g else 1
        if len(gaps) >= 31:
            return 0
    bits = [0 if g else 1 for g in reversed(gaps)] + [1] + [0 if g else 1 for g in gaps]
    p = 0
    for b in bits:
        p = (p << 1) | b
    return p


def seed_length(pattern):
    return pattern.bit_length()


def seed_weight(pattern):
    return bin(pattern).count("1")


def seed_valid(pattern):
    L, w = seed_length(pattern), seed_weight(pattern)
    if L == 0 or w % 2 == 0 or w < 3 or w > 31:
        return False
    return all(((pattern >> j) & 1) == ((pattern >> (L - 1 - j)) & 1) for j in range(L))


def default_seed_weight(avg_len):
    w = int(math.log2(max(avg_len, 1)) / 1.5)
    if w % 2 == 0:
        w += 1
    return max(5, min(31, w))


def pattern_text(pattern):
    return bin(pattern)[2:]
