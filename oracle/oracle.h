/*
 * oracle/oracle.h — CPU ORACLE for the Mauve seed-match anchoring path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under mauvealigner_b200/ or include/ may
 * include, link or call this.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, as the checker or
 * the timed CPU baseline, never as the shipped path.
 *
 * PARITY UNPINNED: the reference's arithmetic for this path lives in libMems
 * (pkg-config libMems-1.6, /root/reference/configure.ac:46), which is neither
 * vendored in /root/reference nor installed here, and the reference tree holds
 * no golden vectors, known-answer tests or fixtures.  This oracle restates the
 * in-tree policy code line by line (src/UniqueMatchFinder.cpp:36-60,
 * src/SeedMatchEnumerator.h:19-141, src/uniqueMerCount.cpp:39) and defines the
 * libMems behaviour by SURVEY.md Appendix A decisions D1-D18.
 */
#ifndef MAUVE_ORACLE_H
#define MAUVE_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    ORC_MODE_UNIQUE = 0,       /* UniqueMatchFinder / MemHash: unique filter + extend + dedup */
    ORC_MODE_SEED_ENUM = 1,    /* SeedMatchEnumerator: one un-extended match per bucket      */
    ORC_MODE_UNIQUE_COUNT = 2, /* SortedMerList::UniqueMerCount only                         */
    ORC_MODE_PAIRWISE = 3,     /* PairwiseMatchFinder: unique filter, HashMatch every pair   */
    ORC_MODE_REPEAT = 4        /* RepeatHash: one sequence, every occurrence a component, extend + dedup */
};

typedef struct orc_result {
    uint64_t n_matches;
    uint64_t n_comps;
    uint32_t* length;      /* [n_matches]            */
    uint64_t* comp_off;    /* [n_matches + 1]        */
    uint32_t* comp_seq;    /* [n_comps] genome index */
    int64_t* comp_start;   /* [n_comps] signed 1-based left end, <0 = reverse strand */
    uint64_t unique_mers;            /* distinct canonical seeds over all sequences */
    uint64_t* unique_mers_per_seq;   /* [nseq] SortedMerList::UniqueMerCount per SML */
    uint32_t nseq;
    /* statistics */
    uint64_t n_seeds;       /* total seed positions                      */
    uint64_t n_buckets;     /* equal-seed buckets of size >= 2           */
    uint64_t n_candidates;  /* buckets handed to HashMatch               */
    uint64_t n_contained;   /* candidates dropped by the containment test */
    /* wall-clock seconds per stage (single thread) */
    double t_mers, t_sort, t_match, t_total;
} orc_result;

/* Whole path: ASCII genomes -> canonical match set.  Returns 0 or a negative error. */
int orc_find(uint32_t nseq, const uint8_t* const* ascii, const uint64_t* lens,
             uint64_t pattern, int mode,
             uint64_t min_multi, uint64_t max_multi, int direct_only, uint64_t nway_mask,
             orc_result** out);
/* Seed-family search: MODE_UNIQUE once per pattern, in the given order, with ONE persistent MemHash table
 * (src/progressiveMauve.cpp:503-548).  The statistics fields other than n_seeds / n_buckets / n_candidates /
 * n_contained are not filled. */
int orc_find_family(uint32_t nseq, const uint8_t* const* ascii, const uint64_t* lens,
                    const uint64_t* patterns, uint32_t npat, uint64_t nway_mask, orc_result** out);
void orc_result_free(orc_result* r);

/* Per-position canonical seed mers of one genome: out[p] = key<<(64-2w) | strand,
 * p in [0, len-L].  out must hold len-L+1 entries.  Returns the count or <0. */
int64_t orc_mers(const uint8_t* ascii, uint64_t len, uint64_t pattern, uint64_t* out);

/* D2 packing of one genome into 64-bit words, first base in the top bits.
 * out must hold (len+31)/32 words. */
void orc_pack(const uint8_t* ascii, uint64_t len, uint64_t* out);

/* Positions sorted by (key, position): the sorted mer list of one genome (a4). */
int64_t orc_sml(const uint8_t* ascii, uint64_t len, uint64_t pattern, uint32_t* out_pos);

int orc_seed_length(uint64_t pattern);
int orc_seed_weight(uint64_t pattern);
int orc_seed_valid(uint64_t pattern); /* 1 iff palindromic, odd weight <= 31 */

#ifdef __cplusplus
}
#endif
#endif
