// synth.cpp — deterministic synthetic genomes for BASELINE.json configs C1..C5.
// Host only.  PRNG = xoshiro256** seeded by splitmix64(0x4D415556 + config).  Alphabet ACGT.
// Evolution along a branch: substitutions and geometric-length indels by skip sampling, then
// inversions (reverse complement in place) and translocations (cut + paste).
#include "mb_synth.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <new>
#include <string>
#include <vector>

namespace {

struct Rng {
    uint64_t s[4];
    explicit Rng(uint64_t seed) {
        for (int i = 0; i < 4; ++i) {
            seed += 0x9E3779B97F4A7C15ull;
            uint64_t z = seed;
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            s[i] = z ^ (z >> 31);
        }
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {
        uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    uint64_t below(uint64_t n) { return n ? (uint64_t)(uni() * (double)n) % n : 0; }
    uint64_t range(uint64_t lo, uint64_t hi) { return lo + below(hi - lo + 1); } // inclusive
    // number of failures before the first success, success probability p
    uint64_t geometric(double p) {
        if (p >= 1.0) return 0;
        double u = uni();
        if (u <= 0.0) u = 1e-300;
        return (uint64_t)(std::log(u) / std::log1p(-p));
    }
};

typedef std::string Seq;
const char ALPHA[4] = {'A', 'C', 'G', 'T'};

Seq random_seq(Rng& r, uint64_t n) {
    Seq s(n, 'A');
    uint64_t i = 0;
    while (i < n) {
        uint64_t x = r.next();
        for (int k = 0; k < 32 && i < n; ++k, ++i) { s[i] = ALPHA[x & 3]; x >>= 2; }
    }
    return s;
}
char comp(char c) { return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A'; }
void revcomp_inplace(Seq& s, uint64_t a, uint64_t b) { // [a, b)
    std::reverse(s.begin() + a, s.begin() + b);
    for (uint64_t i = a; i < b; ++i) s[i] = comp(s[i]);
}
char other_base(Rng& r, char c) {
    int k = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3;
    return ALPHA[(k + 1 + r.below(3)) & 3];
}

// point mutations + indels
Seq mutate(Rng& r, const Seq& p, double sub, double indel, double indel_mean, uint64_t indel_cap) {
    Seq out;
    out.reserve(p.size() + p.size() / 50 + 64);
    double rate = sub + indel;
    uint64_t i = 0, n = p.size();
    if (rate <= 0) return p;
    while (i < n) {
        uint64_t gap = r.geometric(rate);
        uint64_t j = std::min(n, i + gap);
        out.append(p, i, j - i);
        i = j;
        if (i >= n) break;
        if (r.uni() * rate < sub) {
            out.push_back(other_base(r, p[i]));
            ++i;
        } else {
            uint64_t len = 1 + std::min<uint64_t>(indel_cap - 1, r.geometric(1.0 / indel_mean));
            if (r.next() & 1) i = std::min(n, i + len);            // deletion
            else out += random_seq(r, len);                         // insertion
        }
    }
    return out;
}

void invert(Rng& r, Seq& s, int count, uint64_t lo, uint64_t hi) {
    for (int k = 0; k < count; ++k) {
        if (s.size() < 16) return;
        uint64_t len = std::min<uint64_t>(r.range(lo, hi), s.size() / 2);
        if (len < 4) len = 4;
        uint64_t a = r.below(s.size() - len);
        revcomp_inplace(s, a, a + len);
    }
}
void translocate(Rng& r, Seq& s, int count, uint64_t len) {
    for (int k = 0; k < count; ++k) {
        if (s.size() < 4 * len || len == 0) return;
        uint64_t a = r.below(s.size() - len);
        Seq seg = s.substr(a, len);
        s.erase(a, len);
        uint64_t b = r.below(s.size());
        s.insert(b, seg);
    }
}

// overwrite `copies` places with diverged copies of one unit
void plant_family(Rng& r, Seq& s, uint64_t unit_len, int copies, double div, bool both_strands) {
    if (s.size() < unit_len * 2 || unit_len == 0) return;
    Seq unit = random_seq(r, unit_len);
    for (int c = 0; c < copies; ++c) {
        Seq u = div > 0 ? mutate(r, unit, div, 0.0, 1.0, 1) : unit;
        if (both_strands && (r.next() & 1)) revcomp_inplace(u, 0, u.size());
        uint64_t a = r.below(s.size() - u.size());
        s.replace(a, u.size(), u);
    }
}
void plant_microsatellites(Rng& r, Seq& s, double fraction) {
    uint64_t target = (uint64_t)(fraction * (double)s.size()), done = 0;
    while (done < target && s.size() > 2000) {
        uint64_t ulen = r.range(1, 6), tract = r.range(50, 600);
        Seq unit = random_seq(r, ulen);
        uint64_t a = r.below(s.size() - tract);
        for (uint64_t i = 0; i < tract; ++i) s[a + i] = unit[i % ulen];
        done += tract;
    }
}

uint64_t scaled(uint64_t v, uint64_t scale, uint64_t floor_) { return std::max<uint64_t>(floor_, v / scale); }

// balanced binary tree: leaves in left-to-right order
void tree(Rng& r, const Seq& node, int depth, double sub, double indel, int inv, uint64_t inv_lo, uint64_t inv_hi, std::vector<Seq>& leaves) {
    if (depth == 0) { leaves.push_back(node); return; }
    for (int child = 0; child < 2; ++child) {
        Seq c = mutate(r, node, sub, indel, 3.0, 50);
        invert(r, c, inv, inv_lo, inv_hi);
        tree(r, c, depth - 1, sub, indel, inv, inv_lo, inv_hi, leaves);
    }
}

} // namespace

struct mb_synth { std::vector<Seq> seqs; };

extern "C" {

int mb_synth_create(int config, uint64_t scale, mb_synth** out) {
    if (!out || config < 1 || config > 5 || scale == 0) return -1;
    *out = nullptr;
    mb_synth* h = new (std::nothrow) mb_synth();
    if (!h) return -2;
    Rng r(0x4D415556ull + (uint64_t)config);
    try {
        if (config == 1 || config == 2 || config == 5) {
            uint64_t len = scaled(5000000, scale, 2000);
            Seq anc = random_seq(r, len);
            for (int f = 0; f < 20; ++f) plant_family(r, anc, scaled(1000, scale > 50 ? 10 : 1, 100), 5, 0.01, false);
            uint64_t inv_lo = scaled(50000, scale, 100), inv_hi = scaled(200000, scale, 400);
            if (config == 1) {
                Seq b = mutate(r, anc, 0.02, 1.0 / 2000.0, 3.0, 50);
                invert(r, b, 10, inv_lo, inv_hi);
                translocate(r, b, 2, scaled(100000, scale, 200));
                h->seqs.push_back(anc);
                h->seqs.push_back(b);
            } else if (config == 2) {
                tree(r, anc, 3, 0.01, 1.0 / 2000.0, 3, inv_lo, inv_hi, h->seqs);
            } else {
                tree(r, anc, 6, 0.005, 1.0 / 4000.0, 1, inv_lo, inv_hi, h->seqs);
            }
        } else if (config == 3) {
            uint64_t len = scaled(100000000, scale, 5000);
            Seq g = random_seq(r, len);
            uint64_t planted = 0;
            while (planted < len / 20) { // 5 % of the length in repeat families
                uint64_t unit = r.range(300, 3000);
                int copies = (int)r.range(2, 20);
                plant_family(r, g, std::min<uint64_t>(unit, len / 8), copies, 0.02, true);
                planted += unit * copies;
            }
            h->seqs.push_back(g);
        } else { // config 4
            uint64_t len = scaled(200000000, scale, 5000);
            Seq g = random_seq(r, len);
            uint64_t planted = 0;
            while (planted < len * 3 / 10) { // 30 % repeats, Zipf copy numbers over 2..5000
                double u = r.uni();
                uint64_t copies = (uint64_t)(2.0 * std::pow(2500.0, u * u * u)); // heavy tail towards 5000
                copies = std::min<uint64_t>(std::max<uint64_t>(copies, 2), 5000);
                uint64_t unit = std::min<uint64_t>(r.range(100, 6000), len / 8);
                if (copies * unit > len / 10) copies = std::max<uint64_t>(2, len / 10 / unit);
                plant_family(r, g, unit, (int)copies, r.uni() * 0.10, true);
                planted += unit * copies;
            }
            plant_microsatellites(r, g, 0.01);
            h->seqs.push_back(g);
        }
    } catch (const std::bad_alloc&) {
        delete h;
        return -2;
    }
    *out = h;
    return 0;
}

uint32_t mb_synth_nseq(const mb_synth* s) { return s ? (uint32_t)s->seqs.size() : 0; }
uint64_t mb_synth_len(const mb_synth* s, uint32_t i) { return (s && i < s->seqs.size()) ? s->seqs[i].size() : 0; }
const uint8_t* mb_synth_seq(const mb_synth* s, uint32_t i) {
    return (s && i < s->seqs.size()) ? reinterpret_cast<const uint8_t*>(s->seqs[i].data()) : nullptr;
}
void mb_synth_free(mb_synth* s) { delete s; }

} // extern "C"
