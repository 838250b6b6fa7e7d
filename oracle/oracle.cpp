/*
 * oracle/oracle.cpp — CPU ORACLE for the Mauve seed-match anchoring path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  PARITY UNPINNED: libMems, which
 * holds the reference's arithmetic for this path, is not in /root/reference
 * and not installed; the reference tree has no golden vectors.  This file
 * follows the in-tree policy code line by line and SURVEY.md Appendix A
 * (D1-D18, A.2 pseudo-code) for the libMems part.
 *
 * Structure mirrors the reference's single-threaded streaming design, NOT the
 * GPU design: per-genome sorted mer lists (a4) -> N-way merge in ascending
 * seed order (a6) -> per-bucket policy (a7/a8) -> HashMatch / containment
 * lookup / ExtendMatch (a9-a11) -> match list (a12).
 */
#include "oracle.h"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <map>
#include <queue>
#include <unordered_map>
#include <vector>

typedef unsigned __int128 u128;

namespace {

/* ---------------------------------------------------------------- seeds (D3) */
struct Seed {
    uint64_t pattern;
    int L, w;
    u128 mask2;    /* care mask expanded to 2 bits per base, 2L bits wide */
    u128 full2;    /* 2L ones */
    int nrun;
    int run_shift[64], run_bits[64]; /* care runs, most significant first */
};

static int seed_length(uint64_t p) { int L = 0; while (p) { ++L; p >>= 1; } return L; }
static int seed_weight(uint64_t p) { return __builtin_popcountll(p); }
static int seed_valid(uint64_t p) {
    int L = seed_length(p), w = seed_weight(p);
    if (L == 0 || !(w & 1) || w < 3 || w > 31) return 0;
    for (int j = 0; j < L; ++j)
        if (((p >> j) & 1) != ((p >> (L - 1 - j)) & 1)) return 0;
    return 1;
}
static void seed_init(Seed& s, uint64_t pattern) {
    s.pattern = pattern; s.L = seed_length(pattern); s.w = seed_weight(pattern);
    s.mask2 = 0;
    for (int b = 0; b < s.L; ++b)
        if ((pattern >> b) & 1) s.mask2 |= ((u128)3) << (2 * b);
    s.full2 = (s.L == 64) ? ~(u128)0 : ((((u128)1) << (2 * s.L)) - 1);
    s.nrun = 0;
    int b = s.L - 1;
    while (b >= 0) {
        if (!((pattern >> b) & 1)) { --b; continue; }
        int hi = b;
        while (b >= 0 && ((pattern >> b) & 1)) --b;
        int lo = b + 1; /* run covers pattern bits [lo, hi] */
        s.run_shift[s.nrun] = 2 * lo;
        s.run_bits[s.nrun] = 2 * (hi - lo + 1);
        ++s.nrun;
    }
}
/* D4: concatenate the cared bases, first cared base most significant. */
static inline uint64_t seed_gather(const Seed& s, u128 win) {
    uint64_t k = 0;
    for (int r = 0; r < s.nrun; ++r) {
        uint64_t part = (uint64_t)(win >> s.run_shift[r]) & ((s.run_bits[r] == 64) ? ~0ull : ((1ull << s.run_bits[r]) - 1));
        k = (k << s.run_bits[r]) | part;
    }
    return k;
}

/* D1 */
static inline unsigned base_code(uint8_t c) {
    switch (c) {
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 0;
    }
}

/* D4 + SortedMerList::FillDnaSeedSML: mer[p] = key << (64-2w) | strand */
static int64_t compute_mers(const uint8_t* seq, uint64_t len, const Seed& s, uint64_t* out) {
    if (len < (uint64_t)s.L) return 0;
    u128 F = 0, R = 0;
    const int topshift = 2 * (s.L - 1);
    const int keyshift = 64 - 2 * s.w;
    uint64_t n = len - s.L + 1;
    for (uint64_t i = 0; i < len; ++i) {
        unsigned b = base_code(seq[i]);
        F = ((F << 2) | b) & s.full2;
        R = (R >> 2) | (((u128)(3 - b)) << topshift);
        if (i + 1 >= (uint64_t)s.L) {
            u128 fm = F & s.mask2, rm = R & s.mask2;
            unsigned strand = rm < fm;
            uint64_t key = seed_gather(s, strand ? rm : fm);
            out[i + 1 - s.L] = (key << keyshift) | strand;
        }
    }
    return (int64_t)n;
}

/* a4: positions sorted by (key, position).  LSD radix on the key bits; stable,
 * so equal keys keep ascending position. */
static void sort_sml(const uint64_t* mers, uint64_t n, const Seed& s, std::vector<uint32_t>& pos) {
    pos.resize(n);
    for (uint64_t i = 0; i < n; ++i) pos[i] = (uint32_t)i;
    if (n < 2) return;
    std::vector<uint32_t> tmp(n);
    const int keyshift = 64 - 2 * s.w;
    const int RB = 11;
    for (int bit = 0; bit < 2 * s.w; bit += RB) {
        int nb = std::min(RB, 2 * s.w - bit);
        uint32_t m = (1u << nb) - 1;
        std::vector<uint64_t> cnt((size_t)1 << nb, 0);
        for (uint64_t i = 0; i < n; ++i) ++cnt[(mers[pos[i]] >> (keyshift + bit)) & m];
        uint64_t acc = 0;
        for (auto& c : cnt) { uint64_t t = c; c = acc; acc += t; }
        for (uint64_t i = 0; i < n; ++i) tmp[cnt[(mers[pos[i]] >> (keyshift + bit)) & m]++] = pos[i];
        pos.swap(tmp);
    }
}

/* ------------------------------------------------------------ match records */
struct MatchRec {
    uint32_t length;
    std::vector<int64_t> start; /* MODE_UNIQUE/PAIRWISE: N entries, 0 = absent; SEED_ENUM: m entries */
};

struct idmer { uint32_t position; uint64_t mer; uint32_t id; };

struct Ctx {
    uint32_t nseq;
    std::vector<uint64_t> lens;
    Seed seed;
    std::vector<std::vector<uint64_t>> mers;   /* per genome, GetSeedMer(pos) */
    std::vector<std::vector<uint32_t>> sml;    /* per genome sorted positions */
    int mode;
    uint64_t min_multi, max_multi, nway_mask;
    int direct_only;
    std::vector<MatchRec> out;
    uint64_t n_buckets = 0, n_candidates = 0, n_contained = 0;

    /* MemHash table: group (genome set, strand vector, diagonal) -> accepted
     * extents on that diagonal, as an antichain map first-genome start -> end
     * (exclusive).  Matches swallowed by a later, larger match stay in `out`
     * but leave the lookup map: anything they contain the larger one contains. */
    struct Group { std::vector<int64_t> key; std::map<int64_t, int64_t> ext; };
    std::unordered_map<uint64_t, std::vector<Group>> table;
};

static uint64_t hash_vec(const std::vector<int64_t>& v) {
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (int64_t x : v) {
        h ^= (uint64_t)x + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
        h *= 0xFF51AFD7ED558CCDull; h ^= h >> 33;
    }
    return h;
}

/* genome behind column `col` of a match: the column itself, except for RepeatHash matches (ORC_MODE_REPEAT), whose
 * columns are the occurrences inside the one sequence (as for SeedMatchEnumerator, GetSar(i) = SML 0) */
static inline uint32_t gm(const Ctx& c, uint32_t col) { return c.mode == ORC_MODE_REPEAT ? 0u : col; }

/* D16 group key: invariant of a match under extension. */
static void group_key(const Ctx& c, const std::vector<int64_t>& st, uint32_t length, std::vector<int64_t>& key) {
    (void)c;
    key.clear();
    int first = -1;
    for (uint32_t g = 0; g < st.size(); ++g) {
        if (st[g] == 0) continue;
        if (first < 0) first = (int)g;
        key.push_back((int64_t)g);
        if (st[g] > 0) { key.push_back(0); key.push_back(st[g] - st[first]); }
        else { key.push_back(1); key.push_back(st[first] + (-st[g]) + (int64_t)length - 1); }
    }
}

/* a11, SURVEY A.2 ExtendMatch: four phases with Invert() between them. */
static void extend_match(const Ctx& c, std::vector<int64_t>& st, int64_t& length) {
    const int L = c.seed.L;
    const uint64_t mer_mask = ~1ull; /* key bits (left aligned) without the strand bit */
    std::vector<uint32_t> used;
    for (uint32_t g = 0; g < st.size(); ++g) if (st[g] != 0) used.push_back(g);
    int64_t jump = L;
    for (int dir = 0; dir < 4; ++dir) {
        int64_t maxlen = (dir < 2) ? INT64_MAX : (int64_t)L;
        for (uint32_t g : used) {
            int64_t room = st[g] < 0 ? (int64_t)c.lens[gm(c, g)] - length + st[g] + 1 : st[g] - 1;
            if (room < maxlen) maxlen = room;
        }
        while (maxlen - jump >= 0) {
            length += jump; maxlen -= jump;
            for (uint32_t g : used) if (st[g] > 0) st[g] -= jump;
            bool ok = true;
            uint64_t ref_mer = 0; bool ref_par = false;
            for (size_t i = 0; i < used.size(); ++i) {
                uint32_t g = used[i];
                int64_t pos = st[g] > 0 ? st[g] : -st[g] + length - L;
                uint64_t m = c.mers[gm(c, g)][(size_t)(pos - 1)];
                bool par = st[g] < 0 ? (m & 1) : !(m & 1);
                m &= mer_mask;
                if (i == 0) { ref_mer = m; ref_par = par; }
                else if (m != ref_mer || par != ref_par) { ok = false; break; }
            }
            if (!ok) {
                length -= jump;
                for (uint32_t g : used) if (st[g] > 0) st[g] += jump;
                break;
            }
        }
        for (uint32_t g : used) st[g] = -st[g]; /* Invert */
        if (dir >= 1) jump = 1;
    }
}

/* a10 AddHashEntry with the order-free containment predicate of D16. */
static void add_hash_entry(Ctx& c, std::vector<int64_t>& st) {
    const int L = c.seed.L;
    ++c.n_candidates;
    std::vector<int64_t> key;
    group_key(c, st, (uint32_t)L, key);
    uint64_t h = hash_vec(key);
    auto& bucket = c.table[h];
    Ctx::Group* grp = nullptr;
    for (auto& g : bucket) if (g.key == key) { grp = &g; break; }
    int first = 0; while (st[first] == 0) ++first;
    int64_t x = st[first];
    if (grp) {
        auto it = grp->ext.upper_bound(x);
        if (it != grp->ext.begin()) {
            --it;
            if (it->first <= x && x + L <= it->second) { ++c.n_contained; return; }
        }
    } else {
        bucket.push_back(Ctx::Group());
        grp = &bucket.back();
        grp->key = key;
    }
    int64_t length = L;
    extend_match(c, st, length);
    int64_t s = st[first], e = s + length;
    /* keep the lookup map an antichain: drop extents the new one swallows */
    auto it = grp->ext.lower_bound(s);
    while (it != grp->ext.end() && it->second <= e) it = grp->ext.erase(it);
    grp->ext[s] = e;
    MatchRec m; m.length = (uint32_t)length; m.start = st;
    c.out.push_back(std::move(m));
}

/* a9 MemHash::HashMatch + SetDirection (same body as SeedMatchEnumerator.h:127-141). */
static void hash_match_unique(Ctx& c, const std::vector<idmer>& list) {
    std::vector<int64_t> st(c.nseq, 0);
    for (const idmer& x : list) st[x.id] = (int64_t)x.position + 1;
    /* SetDirection */
    bool ref_forward = false;
    uint32_t g = 0;
    for (; g < c.nseq; ++g)
        if (st[g] != 0) { ref_forward = !(c.mers[g][(size_t)(st[g] - 1)] & 1); break; }
    for (++g; g < c.nseq; ++g)
        if (st[g] != 0 && ref_forward == (bool)(c.mers[g][(size_t)(st[g] - 1)] & 1)) st[g] = -st[g];
    uint32_t mult = 0;
    for (uint32_t i = 0; i < c.nseq; ++i) mult += st[i] != 0;
    if (mult < 2) return; /* "red flag" */
    add_hash_entry(c, st);
}

/* a7: UniqueMatchFinder::EnumerateMatches, src/UniqueMatchFinder.cpp:36-60
 * (and MemHash / MaskedMemHash / PairwiseMatchFinder variants). */
static void enumerate_unique(Ctx& c, std::vector<idmer>& bucket) {
    std::stable_sort(bucket.begin(), bucket.end(), [](const idmer& a, const idmer& b) { return a.id < b.id; });
    std::vector<idmer> unique_list;
    size_t i = 0, j = 0;
    unsigned cur_id_count = 1;
    while (j != bucket.size()) {
        ++j;
        if (j == bucket.size() || bucket[i].id != bucket[j].id) {
            if (cur_id_count == 1) unique_list.push_back(bucket[i]);
            else cur_id_count = 1;
        } else
            ++cur_id_count;
        ++i;
    }
    if (unique_list.size() < 2) return;
    if (c.mode == ORC_MODE_PAIRWISE) {
        for (size_t a = 0; a < unique_list.size(); ++a)
            for (size_t b = a + 1; b < unique_list.size(); ++b) {
                std::vector<idmer> pr{unique_list[a], unique_list[b]};
                hash_match_unique(c, pr);
            }
        return;
    }
    if (c.nway_mask) { /* MaskedMemHash: the surviving genome set must equal the mask (D9) */
        uint64_t present = 0;
        for (const idmer& x : unique_list) present |= 1ull << x.id;
        if (present != c.nway_mask) return;
    }
    hash_match_unique(c, unique_list);
}

/* a8: SeedMatchEnumerator::HashMatch + SetDirection, src/SeedMatchEnumerator.h:71-141 */
static void hash_match_enum(Ctx& c, std::vector<idmer>& bucket) {
    std::stable_sort(bucket.begin(), bucket.end(), [](const idmer& a, const idmer& b) { return a.position < b.position; });
    size_t m = bucket.size();
    std::vector<int64_t> st(m);
    for (size_t i = 0; i < m; ++i) st[i] = (int64_t)bucket[i].position + 1;
    /* SetDirection; GetSar(i) is always SML 0 (:54-57) */
    bool ref_forward = !(c.mers[0][(size_t)(st[0] - 1)] & 1);
    for (size_t i = 1; i < m; ++i)
        if (ref_forward == (bool)(c.mers[0][(size_t)(st[i] - 1)] & 1)) st[i] = -st[i];
    bool found_reverse = false;
    std::vector<size_t> component_map;
    if (c.direct_only)
        for (size_t i = 0; i < m; ++i) {
            if (st[i] > 0) component_map.push_back(i);
            else found_reverse = true;
        }
    if (m < 2) return;
    else if (m > c.max_multi || m < c.min_multi) return;
    else if (c.direct_only && found_reverse) {
        if (component_map.size() > 1) {
            MatchRec r; r.length = (uint32_t)c.seed.L;
            for (size_t k : component_map) r.start.push_back(st[k]);
            c.out.push_back(std::move(r));
        }
    } else {
        MatchRec r; r.length = (uint32_t)c.seed.L; r.start = st;
        c.out.push_back(std::move(r));
    }
}

/* RepeatHash (mauveAligner --repeats, src/mauveAligner.cpp:480-487) [LM-recall, SURVEY A.2]: one sequence; every bucket
 * of min_multi..max_multi occurrences (at most 255 here) becomes ONE entry with a component per occurrence in position
 * order, signs as SeedMatchEnumerator::SetDirection (src/SeedMatchEnumerator.h:127-141), then it goes through the same
 * AddHashEntry (containment de-dup D16 + ExtendMatch D14) as a MemHash entry. */
static void hash_match_repeat(Ctx& c, std::vector<idmer>& bucket) {
    std::stable_sort(bucket.begin(), bucket.end(), [](const idmer& a, const idmer& b) { return a.position < b.position; });
    size_t m = bucket.size();
    if (m < 2 || m < c.min_multi || m > c.max_multi || m > 255) return;
    std::vector<int64_t> st(m);
    for (size_t i = 0; i < m; ++i) st[i] = (int64_t)bucket[i].position + 1;
    bool ref_forward = !(c.mers[0][(size_t)(st[0] - 1)] & 1);
    for (size_t i = 1; i < m; ++i)
        if (ref_forward == (bool)(c.mers[0][(size_t)(st[i] - 1)] & 1)) st[i] = -st[i];
    add_hash_entry(c, st);
}

/* a6: MatchFinder::FindMatchSeeds — N-way merge of the SMLs, ascending masked mer. */
static void find_match_seeds(Ctx& c) {
    struct Head { uint64_t key; uint32_t id; uint64_t idx; };
    auto cmp = [](const Head& a, const Head& b) { return a.key != b.key ? a.key > b.key : a.id > b.id; };
    std::priority_queue<Head, std::vector<Head>, decltype(cmp)> heap(cmp);
    for (uint32_t g = 0; g < c.nseq; ++g)
        if (!c.sml[g].empty()) heap.push(Head{c.mers[g][c.sml[g][0]] >> 1, g, 0});
    std::vector<idmer> cur;
    uint64_t cur_key = 0;
    auto flush = [&]() {
        if (cur.size() > 1) {
            ++c.n_buckets;
            if (c.mode == ORC_MODE_SEED_ENUM) hash_match_enum(c, cur);
            else if (c.mode == ORC_MODE_REPEAT) hash_match_repeat(c, cur);
            else enumerate_unique(c, cur);
        }
        cur.clear();
    };
    while (!heap.empty()) {
        Head h = heap.top(); heap.pop();
        if (!cur.empty() && h.key != cur_key) flush();
        cur_key = h.key;
        uint32_t p = c.sml[h.id][h.idx];
        cur.push_back(idmer{p, c.mers[h.id][p], h.id});
        if (h.idx + 1 < c.sml[h.id].size()) {
            uint32_t q = c.sml[h.id][h.idx + 1];
            heap.push(Head{c.mers[h.id][q] >> 1, h.id, h.idx + 1});
        }
    }
    flush();
}

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

/* D18 canonical order */
static bool less_unique(const MatchRec& a, const MatchRec& b) {
    size_t n = a.start.size();
    for (size_t i = 0; i < n; ++i) {
        int64_t x = a.start[i] < 0 ? -a.start[i] : a.start[i], y = b.start[i] < 0 ? -b.start[i] : b.start[i];
        if (x != y) return x < y;
    }
    for (size_t i = 0; i < n; ++i) {
        bool x = a.start[i] < 0, y = b.start[i] < 0;
        if (x != y) return y;
    }
    return a.length < b.length;
}
static bool less_enum(const MatchRec& a, const MatchRec& b) {
    if (a.start[0] != b.start[0]) return a.start[0] < b.start[0];
    if (a.start.size() != b.start.size()) return a.start.size() < b.start.size();
    for (size_t i = 1; i < a.start.size(); ++i)
        if (a.start[i] != b.start[i]) return a.start[i] < b.start[i];
    return a.length < b.length; /* (extended RepeatHash matches; SeedMatchEnumerator matches all have the seed length) */
}

/* match records -> the CSR arrays of orc_result */
static void pack_result(const Ctx& c, int mode, orc_result* r) {
    r->n_buckets = c.n_buckets; r->n_candidates = c.n_candidates; r->n_contained = c.n_contained;
    r->n_matches = c.out.size();
    r->length = (uint32_t*)malloc(sizeof(uint32_t) * (r->n_matches + 1));
    r->comp_off = (uint64_t*)malloc(sizeof(uint64_t) * (r->n_matches + 1));
    uint64_t nc = 0;
    for (const MatchRec& m : c.out)
        for (int64_t s : m.start) nc += s != 0;
    r->n_comps = nc;
    r->comp_seq = (uint32_t*)malloc(sizeof(uint32_t) * (nc + 1));
    r->comp_start = (int64_t*)malloc(sizeof(int64_t) * (nc + 1));
    uint64_t k = 0;
    for (size_t i = 0; i < c.out.size(); ++i) {
        const MatchRec& m = c.out[i];
        r->length[i] = m.length;
        r->comp_off[i] = k;
        for (size_t g = 0; g < m.start.size(); ++g)
            if (m.start[g] != 0) {
                r->comp_seq[k] = (mode == ORC_MODE_SEED_ENUM || mode == ORC_MODE_REPEAT) ? 0u : (uint32_t)g;
                r->comp_start[k] = m.start[g];
                ++k;
            }
    }
    r->comp_off[r->n_matches] = k;
}

} // namespace

extern "C" {

int orc_seed_length(uint64_t p) { return seed_length(p); }
int orc_seed_weight(uint64_t p) { return seed_weight(p); }
int orc_seed_valid(uint64_t p) { return seed_valid(p); }

void orc_pack(const uint8_t* ascii, uint64_t len, uint64_t* out) {
    uint64_t nw = (len + 31) / 32;
    for (uint64_t i = 0; i < nw; ++i) out[i] = 0;
    for (uint64_t i = 0; i < len; ++i)
        out[i / 32] |= (uint64_t)base_code(ascii[i]) << (62 - 2 * (i % 32));
}

int64_t orc_mers(const uint8_t* ascii, uint64_t len, uint64_t pattern, uint64_t* out) {
    if (!seed_valid(pattern)) return -1;
    Seed s; seed_init(s, pattern);
    return compute_mers(ascii, len, s, out);
}

int64_t orc_sml(const uint8_t* ascii, uint64_t len, uint64_t pattern, uint32_t* out_pos) {
    if (!seed_valid(pattern)) return -1;
    Seed s; seed_init(s, pattern);
    if (len < (uint64_t)s.L) return 0;
    std::vector<uint64_t> mers(len - s.L + 1);
    compute_mers(ascii, len, s, mers.data());
    std::vector<uint32_t> pos;
    sort_sml(mers.data(), mers.size(), s, pos);
    memcpy(out_pos, pos.data(), pos.size() * sizeof(uint32_t));
    return (int64_t)pos.size();
}

int orc_find(uint32_t nseq, const uint8_t* const* ascii, const uint64_t* lens,
             uint64_t pattern, int mode,
             uint64_t min_multi, uint64_t max_multi, int direct_only, uint64_t nway_mask,
             orc_result** out) {
    if (!out) return -1;
    *out = nullptr;
    if (!seed_valid(pattern)) return -2;
    if (nseq == 0) return -3;
    if ((mode == ORC_MODE_UNIQUE || mode == ORC_MODE_PAIRWISE) && nseq > 64) return -4;
    for (uint32_t g = 0; g < nseq; ++g) if (lens[g] >= (1ull << 32)) return -5;
    Ctx c;
    c.nseq = nseq; c.lens.assign(lens, lens + nseq);
    seed_init(c.seed, pattern);
    c.mode = mode; c.min_multi = min_multi; c.max_multi = max_multi;
    c.direct_only = direct_only; c.nway_mask = nway_mask;

    orc_result* r = (orc_result*)calloc(1, sizeof(orc_result));
    r->nseq = nseq;
    r->unique_mers_per_seq = (uint64_t*)calloc(nseq, sizeof(uint64_t));

    double t0 = now_s();
    c.mers.resize(nseq); c.sml.resize(nseq);
    for (uint32_t g = 0; g < nseq; ++g) {
        uint64_t n = lens[g] >= (uint64_t)c.seed.L ? lens[g] - c.seed.L + 1 : 0;
        c.mers[g].resize(n);
        compute_mers(ascii[g], lens[g], c.seed, c.mers[g].data());
        r->n_seeds += n;
    }
    double t1 = now_s();
    for (uint32_t g = 0; g < nseq; ++g) {
        sort_sml(c.mers[g].data(), c.mers[g].size(), c.seed, c.sml[g]);
        /* a5 / D8: SortedMerList::UniqueMerCount */
        uint64_t u = 0;
        for (size_t i = 0; i < c.sml[g].size(); ++i)
            if (i == 0 || (c.mers[g][c.sml[g][i]] >> 1) != (c.mers[g][c.sml[g][i - 1]] >> 1)) ++u;
        r->unique_mers_per_seq[g] = u;
    }
    double t2 = now_s();

    if ((mode == ORC_MODE_SEED_ENUM || mode == ORC_MODE_REPEAT) && nseq != 1) {
        /* SeedMatchEnumerator::CreateMatches does nothing unless seq_count == 1 (:59-65); RepeatHash likewise */
    } else if (mode != ORC_MODE_UNIQUE_COUNT) {
        find_match_seeds(c);
    }
    /* distinct seeds over all genomes */
    {
        std::vector<uint64_t> all;
        if (nseq == 1) r->unique_mers = r->unique_mers_per_seq[0];
        else {
            for (uint32_t g = 0; g < nseq; ++g)
                for (size_t i = 0; i < c.sml[g].size(); ++i)
                    if (i == 0 || (c.mers[g][c.sml[g][i]] >> 1) != (c.mers[g][c.sml[g][i - 1]] >> 1))
                        all.push_back(c.mers[g][c.sml[g][i]] >> 1);
            std::sort(all.begin(), all.end());
            r->unique_mers = (uint64_t)(std::unique(all.begin(), all.end()) - all.begin());
        }
    }
    if (mode == ORC_MODE_SEED_ENUM || mode == ORC_MODE_REPEAT) std::sort(c.out.begin(), c.out.end(), less_enum);
    else std::sort(c.out.begin(), c.out.end(), less_unique);
    double t3 = now_s();

    pack_result(c, mode, r);
    r->t_mers = t1 - t0; r->t_sort = t2 - t1; r->t_match = t3 - t2; r->t_total = t3 - t0;
    *out = r;
    return 0;
}

/* Seed-family search (src/progressiveMauve.cpp:503-548): ONE UniqueMatchFinder runs FindMatches once per seed
 * pattern, in the order given (the reference sorts the family by seed length and searches the longest first, "so
 * that overlapping matches tend to get contained"), ClearSequences() in between and one GetMatchList at the end.
 * The MemHash table persists across the passes (only Clear() empties it), so a candidate of a later pattern is
 * dropped when an accepted match of ANY earlier pattern contains its seed (same group, D16), it is extended with
 * its own pattern's window predicate otherwise, and the result is the union of the accepted matches of all passes
 * in canonical order (D18).  [LM-recall for the persistence; the call sequence is in the tree.] */
int orc_find_family(uint32_t nseq, const uint8_t* const* ascii, const uint64_t* lens,
                    const uint64_t* patterns, uint32_t npat, uint64_t nway_mask, orc_result** out) {
    if (!out || !patterns || npat == 0) return -1;
    *out = nullptr;
    for (uint32_t p = 0; p < npat; ++p) if (!seed_valid(patterns[p])) return -2;
    if (nseq == 0) return -3;
    if (nseq > 64) return -4;
    for (uint32_t g = 0; g < nseq; ++g) if (lens[g] >= (1ull << 32)) return -5;
    Ctx c;
    c.nseq = nseq; c.lens.assign(lens, lens + nseq);
    c.mode = ORC_MODE_UNIQUE; c.min_multi = 2; c.max_multi = 1000; c.direct_only = 0; c.nway_mask = nway_mask;
    orc_result* r = (orc_result*)calloc(1, sizeof(orc_result));
    r->nseq = nseq;
    r->unique_mers_per_seq = (uint64_t*)calloc(nseq, sizeof(uint64_t));
    double t0 = now_s();
    for (uint32_t p = 0; p < npat; ++p) {
        seed_init(c.seed, patterns[p]);
        c.mers.assign(nseq, std::vector<uint64_t>());
        c.sml.assign(nseq, std::vector<uint32_t>());
        for (uint32_t g = 0; g < nseq; ++g) {
            uint64_t n = lens[g] >= (uint64_t)c.seed.L ? lens[g] - c.seed.L + 1 : 0;
            c.mers[g].resize(n);
            compute_mers(ascii[g], lens[g], c.seed, c.mers[g].data());
            r->n_seeds += n;
            sort_sml(c.mers[g].data(), c.mers[g].size(), c.seed, c.sml[g]);
        }
        find_match_seeds(c); /* c.table and c.out carry over to the next pattern */
    }
    std::sort(c.out.begin(), c.out.end(), less_unique);
    pack_result(c, ORC_MODE_UNIQUE, r);
    r->t_total = r->t_match = now_s() - t0;
    *out = r;
    return 0;
}

void orc_result_free(orc_result* r) {
    if (!r) return;
    free(r->length); free(r->comp_off); free(r->comp_seq); free(r->comp_start);
    free(r->unique_mers_per_seq);
    free(r);
}

} // extern "C"
