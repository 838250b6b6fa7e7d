/*
 * seed_masks.h — spaced-seed pattern selection for the seed-match path.
 *
 * Replaces libMems' SeedMasks.h (getSeed / getSeedLength / getDefaultSeedWeight,
 * SOLID_SEED, CODING_SEED) as called from
 *   /root/reference/src/mauveAligner.cpp:266-279,465
 *   /root/reference/src/progressiveMauve.cpp:215-224,446-451,511-517
 *   /root/reference/src/repeatoire.cpp:1841-1852
 *
 * NON-AUTHORITATIVE TABLE.  libMems is not vendored in the reference tree, so
 * its pattern table cannot be read.  The
 * patterns below are this project's own; every API of the path takes the raw
 * 64-bit pattern, so a maintainer with libMems at hand can pass its patterns
 * unchanged as long as they are palindromic with odd weight.
 *
 * Pattern convention (matches getPatternText, progressiveMauve.cpp:197-213):
 * printed MSB first from the highest set bit; bit (L-1-j) governs window
 * offset j; 1 = care.
 */
#ifndef MAUVE_B200_SEED_MASKS_H
#define MAUVE_B200_SEED_MASKS_H
#include <stdint.h>
#include <limits.h>

#define MB_SOLID_SEED INT_MAX /* libMems SOLID_SEED */
#define MB_CODING_SEED 3      /* libMems CODING_SEED */
#define MB_MIN_SEED_WEIGHT 3
#define MB_MAX_SEED_WEIGHT 31

static inline int mb_seed_length(uint64_t pattern) {
    int L = 0;
    while (pattern) { ++L; pattern >>= 1; }
    return L;
}
static inline int mb_seed_weight(uint64_t pattern) {
    int w = 0;
    while (pattern) { w += (int)(pattern & 1); pattern >>= 1; }
    return w;
}
/* valid = palindromic, odd weight in [3,31], which forces odd L with a cared
 * centre base, so a masked window can never equal its own reverse complement. */
static inline int mb_seed_valid(uint64_t pattern) {
    int L = mb_seed_length(pattern), w = mb_seed_weight(pattern);
    if (L == 0 || (w & 1) == 0 || w < MB_MIN_SEED_WEIGHT || w > MB_MAX_SEED_WEIGHT) return 0;
    for (int j = 0; j < L; ++j)
        if (((pattern >> j) & 1) != ((pattern >> (L - 1 - j)) & 1)) return 0;
    return 1;
}

/* is distance d (>=1) from the centre a don't-care column for this rank? */
static inline int mb__seed_gap(int rank, int d) {
    switch (rank) {
    case 0: return d % 4 == 3;                       /* 11 0 111 0 111 ...      */
    case 1: return d % 5 == 2 || d % 5 == 4;          /* 1 0 1 0 11 0 1 0 ...    */
    case 2: return d % 7 == 2 || d % 7 == 4 || d % 7 == 5;
    case MB_CODING_SEED: return d % 3 == 2;           /* every third column free */
    default: return 0;                                /* solid                   */
    }
}

/* getSeed(weight, rank): even weights are lowered by one (the DNA seeds must
 * have odd weight).  Returns 0 for an unusable request. */
static inline uint64_t mb_get_seed(int weight, int rank) {
    if (weight < MB_MIN_SEED_WEIGHT) return 0;
    if (weight > MB_MAX_SEED_WEIGHT) weight = MB_MAX_SEED_WEIGHT;
    if ((weight & 1) == 0) weight -= 1;
    if (rank != 0 && rank != 1 && rank != 2 && rank != MB_CODING_SEED) rank = MB_SOLID_SEED;
    int half = (weight - 1) / 2, placed = 0, d = 0;
    int gaps[64];
    int ncol = 0;
    while (placed < half) {
        ++d;
        int gap = mb__seed_gap(rank, d);
        gaps[ncol++] = gap;
        if (!gap) ++placed;
        if (ncol >= 31) return 0; /* would exceed 64 columns */
    }
    /* columns: mirror(gaps) centre gaps */
    uint64_t p = 0;
    for (int i = ncol - 1; i >= 0; --i) p = (p << 1) | (uint64_t)(gaps[i] ? 0 : 1);
    p = (p << 1) | 1u;
    for (int i = 0; i < ncol; ++i) p = (p << 1) | (uint64_t)(gaps[i] ? 0 : 1);
    return p;
}

/* getDefaultSeedWeight(average sequence length): about log2(len)/1.5, made odd
 * (usage text mauveAligner.cpp:878). 5 Mbp -> 15. */
static inline int mb_default_seed_weight(uint64_t avg_len) {
    int lg = 0;
    double x = (double)(avg_len ? avg_len : 1), l2 = 0.0;
    while (x >= 2.0) { x *= 0.5; l2 += 1.0; ++lg; }
    /* fractional part of log2 by repeated squaring */
    double frac = 0.0, bit = 0.5;
    for (int i = 0; i < 20; ++i) { x *= x; if (x >= 2.0) { x *= 0.5; frac += bit; } bit *= 0.5; }
    int w = (int)((l2 + frac) / 1.5);
    if ((w & 1) == 0) w += 1;
    if (w < 5) w = 5;
    if (w > MB_MAX_SEED_WEIGHT) w = MB_MAX_SEED_WEIGHT;
    return w;
}

/* getPatternText equivalent: writes L chars + NUL, returns L. buf >= 65 bytes. */
static inline int mb_seed_pattern_text(uint64_t pattern, char* buf) {
    int L = mb_seed_length(pattern);
    for (int j = 0; j < L; ++j) buf[j] = ((pattern >> (L - 1 - j)) & 1) ? '1' : '0';
    buf[L] = 0;
    return L;
}
#endif
