/*
 * mems_compat.h — host-side C++ mirror of the libMems types the seed-match path is reached through,
 * as thin wrappers over the C ABI of include/mauve_b200.h (libmauve_b200.so).
 *
 * A maintainer replaces `#include "libMems/MemHash.h"` etc. by this header for the hot path; the
 * downstream LCB / recursive-anchoring / gapped-alignment code keeps consuming `mems::MatchList`.
 * Names, argument meaning and error behaviour follow the reference call sites:
 *
 *   mems::Match            src/SeedMatchEnumerator.h:75-119, src/repeatoire.cpp:1926-1936
 *   mems::MatchList        src/mauveAligner.cpp:450-466,641-645 (seq_filename, sml_filename, seq_table, sml_table)
 *   MatchList::CreateMemorySMLs / LoadSMLs     src/mauveAligner.cpp:456,465; src/progressiveMauve.cpp:446-451; src/repeatoire.cpp:1850
 *   mems::MatchFinder::AddSequence / Clear / ClearSequences / LogProgress      src/SeedMatchEnumerator.h:25, src/progressiveMauve.cpp:492-495,542-547
 *   mems::MemHash::FindMatches / GetMatchList, MaskedMemHash::SetMask          src/mauveAligner.cpp:523-589, src/progressiveMauve.cpp:545
 *   UniqueMatchFinder      src/UniqueMatchFinder.h:21-32, src/UniqueMatchFinder.cpp:36-60
 *   SeedMatchEnumerator    src/SeedMatchEnumerator.h:14-141
 *   SortedMerList::UniqueMerCount / SeedLength / Seed                          src/uniqueMerCount.cpp:30-39, src/SeedMatchEnumerator.h:76
 *   WriteList              src/mauveAligner.cpp:603 (line shape: src/MatchRecord.h:349-355)
 *
 * The per-bucket virtual callbacks (EnumerateMatches / HashMatch) cannot cross to the device; each
 * finder class selects the corresponding device policy (mb_params.mode) instead.  There is no CPU
 * implementation behind these classes: without a B200 every FindMatches reports an error.
 */
#ifndef MEMS_COMPAT_H
#define MEMS_COMPAT_H

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../mauve_b200.h"
#include "../mauve_b200/seed_masks.h"

typedef bool boolean;
typedef uint32_t uint32;
typedef int64_t int64;
typedef uint64_t uint64;
typedef unsigned int uint;
typedef uint64_t gnSeqI;

namespace genome {

inline void ErrorMsg(const std::string& s) { std::cerr << s; }

class gnException : public std::runtime_error {
public:
    explicit gnException(const std::string& s) : std::runtime_error(s) {}
};

/* the little of libGenome's gnSequence the path touches: length() and ToString() (src/repeatoire.cpp:1870,1959) */
class gnSequence {
public:
    gnSequence() {}
    explicit gnSequence(const std::string& s) : seq_(s) {}
    gnSeqI length() const { return seq_.size(); }
    std::string ToString() const { return seq_; }
    const std::string& data() const { return seq_; }
private:
    std::string seq_;
};

} // namespace genome

namespace mems {

static const int64 NO_MATCH = 0;
static const int SOLID_SEED = MB_SOLID_SEED;
static const int CODING_SEED = MB_CODING_SEED;

/* getSeed(weight, rank): libMems' SeedMasks.h tables are not in the reference tree, so the patterns behind
 * (weight, rank) and CODING_SEED are this library's own palindromic table (include/mauve_b200/seed_masks.h).  The same
 * command line therefore searches with other patterns than stock Mauve and finds another — equally valid — match set, and
 * .sslist files are not interchangeable.  Said once per process on stderr (MAUVE_B200_QUIET=1 silences it); every API
 * also accepts a raw pattern for callers that hold the original tables. */
inline void seed_table_notice() {
    static bool said = false;
    if (said) return;
    said = true;
    const char* q = getenv("MAUVE_B200_QUIET");
    if (q && *q && *q != '0') return;
    std::cerr << "mauve_b200: seed patterns come from this library's own table, not from libMems' SeedMasks.h "
                 "(match sets differ from stock Mauve for the same --seed-weight; see INTEGRATION.md)\n";
}
inline int64 getSeed(int weight, int rank = 0) { seed_table_notice(); return (int64)mb_get_seed(weight, rank); }
inline uint32 getSeedLength(int64 seed) { return (uint32)mb_seed_length((uint64_t)seed); }
inline uint32 getDefaultSeedWeight(gnSeqI avg_len) { return (uint32)mb_default_seed_weight(avg_len); }

class Match {
public:
    explicit Match(uint seq_count = 0) : start_(seq_count, NO_MATCH), length_(0) {}
    void SetLength(gnSeqI len) { length_ = len; }
    void SetStart(uint seqI, int64 s) { start_[seqI] = s; }
    int64 operator[](uint seqI) const { return start_[seqI]; }
    int64 Start(uint seqI) const { return start_[seqI]; }
    gnSeqI Length(uint = 0) const { return length_; }
    uint SeqCount() const { return (uint)start_.size(); }
    uint Multiplicity() const { uint m = 0; for (int64 s : start_) m += s != NO_MATCH; return m; }
    /* 0 = forward, 1 = reverse, 2 = undefined */
    int Orientation(uint seqI) const { return start_[seqI] > 0 ? 0 : (start_[seqI] < 0 ? 1 : 2); }
    int64 LeftEnd(uint seqI) const { return start_[seqI] < 0 ? -start_[seqI] : start_[seqI]; }
    int64 RightEnd(uint seqI) const { return start_[seqI] == NO_MATCH ? NO_MATCH : LeftEnd(seqI) + (int64)length_ - 1; }
    /* Match::CropStart / CropEnd: drop `n` columns at the start / end of the match.  A forward component moves its
     * start with the match start; a reverse component, whose LEFT end is the match end, moves with the match end. */
    void CropStart(gnSeqI n) {
        for (int64& s : start_) if (s > 0) s += (int64)n;
        length_ -= n;
    }
    void CropEnd(gnSeqI n) {
        for (int64& s : start_) if (s < 0) s -= (int64)n;
        length_ -= n;
    }
    Match* Copy() const { return new Match(*this); }
    void Free() { delete this; }
    friend std::ostream& operator<<(std::ostream& os, const Match& m) {
        os << m.length_;
        for (int64 s : m.start_) os << '\t' << s;
        return os;
    }
private:
    std::vector<int64> start_;
    gnSeqI length_;
};

class MatchList;

/* Façade over the device-built sorted mer list of one sequence. */
class SortedMerList {
public:
    SortedMerList(const genome::gnSequence* seq, uint64 seed) : seq_(seq), seed_(seed), unique_(~0ull) {}
    uint64 Seed() const { return seed_; }
    uint32 SeedLength() const { return (uint32)mb_seed_length(seed_); }
    uint32 SeedWeight() const { return (uint32)mb_seed_weight(seed_); }
    gnSeqI Length() const { return seq_->length(); }
    const genome::gnSequence* Sequence() const { return seq_; }
    /* number of distinct seeds of this sequence (src/uniqueMerCount.cpp:39); computed on the device on first use */
    virtual ~SortedMerList() {}
    virtual gnSeqI UniqueMerCount() {
        if (unique_ != ~0ull) return unique_;
        mb_ctx* ctx = nullptr;
        int rc = mb_ctx_create(&ctx, 0);
        if (rc != MB_OK) throw genome::gnException(std::string("mb_ctx_create: ") + mb_strerror(rc));
        const std::string& s = seq_->data();
        mb_params p = {MB_MODE_UNIQUE_COUNT, 0, 2, 1000, 0};
        const mb_result* r = nullptr;
        rc = mb_add_sequence(ctx, (const uint8_t*)s.data(), s.size(), 0, nullptr);
        if (rc == MB_OK) rc = mb_set_seed(ctx, seed_);
        if (rc == MB_OK) rc = mb_find(ctx, &p, &r);
        if (rc == MB_OK) unique_ = r->unique_mers;
        std::string err = rc == MB_OK ? "" : std::string(mb_strerror(rc)) + " " + mb_last_cuda_error(ctx);
        mb_ctx_destroy(ctx);
        if (rc != MB_OK) throw genome::gnException(err);
        return unique_;
    }
protected:
    const genome::gnSequence* seq_;
    uint64 seed_;
    uint64 unique_;
};

/* mems::DNAFileSML — a sorted mer list kept in a file (".sslist"; src/uniqueMerCount.cpp:30-39, default naming
 * src/progressiveMauve.cpp:215-224).  libMems' own file layout is not available, so this is a
 * versioned format of this project: little-endian header { char magic[8] = "MBSSLIST"; u32 version = 1; u32 reserved;
 * u64 seed pattern; u64 sequence length; u64 positions; u64 unique mers }, then the 2-bit packed sequence (32 bases
 * per u64, first base in the top bits), then the positions (u32) sorted by (seed, position).  Create() builds the
 * list on the device (k_extract + k_onesweep + k_find_runs) through the C ABI. */
class DNAFileSML : public SortedMerList {
public:
    DNAFileSML() : SortedMerList(nullptr, 0) {}
    void Create(const genome::gnSequence& seq, uint64 seed) {
        own_ = seq; seq_ = &own_; seed_ = seed;
        if (!mb_seed_valid(seed)) throw genome::gnException("DNAFileSML::Create: invalid seed");
        mb_ctx* ctx = nullptr;
        int rc = mb_ctx_create(&ctx, 0);
        if (rc != MB_OK) throw genome::gnException(std::string("mb_ctx_create: ") + mb_strerror(rc));
        const std::string& s = own_.data();
        mb_params p = {MB_MODE_UNIQUE_COUNT, 0, 2, 1000, 0};
        const mb_result* r = nullptr;
        rc = mb_add_sequence(ctx, (const uint8_t*)s.data(), s.size(), 0, nullptr);
        if (rc == MB_OK) rc = mb_set_seed(ctx, seed);
        if (rc == MB_OK) rc = mb_find(ctx, &p, &r);
        if (rc == MB_OK) {
            unique_ = r->unique_mers;
            uint64_t n = s.size() >= (size_t)mb_seed_length(seed) ? s.size() - mb_seed_length(seed) + 1 : 0, got = 0;
            pos_.assign(n ? n : 1, 0);
            rc = mb_get_sml(ctx, 0, pos_.data(), n ? n : 1, &got);
            pos_.resize(n);
        }
        std::string err = rc == MB_OK ? "" : std::string(mb_strerror(rc)) + " " + mb_last_cuda_error(ctx);
        mb_ctx_destroy(ctx);
        if (rc != MB_OK) throw genome::gnException(err);
    }
    void WriteFile(const std::string& path) const {
        std::ofstream f(path.c_str(), std::ios::binary);
        if (!f) throw genome::gnException("cannot write " + path);
        const std::string& s = own_.data();
        uint64_t hdr[6] = {0, 1, seed_, (uint64_t)s.size(), (uint64_t)pos_.size(), unique_};
        memcpy(&hdr[0], "MBSSLIST", 8);
        f.write((const char*)hdr, sizeof(hdr));
        std::vector<uint64_t> words((s.size() + 31) / 32, 0);
        for (size_t i = 0; i < s.size(); ++i) {
            char c = s[i] & 0xDF;
            uint64_t code = c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 0;
            words[i / 32] |= code << (62 - 2 * (i % 32));
        }
        f.write((const char*)words.data(), words.size() * 8);
        f.write((const char*)pos_.data(), pos_.size() * 4);
        if (!f) throw genome::gnException("short write " + path);
    }
    void LoadFile(const std::string& path) {
        std::ifstream f(path.c_str(), std::ios::binary);
        if (!f) throw genome::gnException("cannot open " + path);
        uint64_t hdr[6];
        f.read((char*)hdr, sizeof(hdr));
        if (!f || memcmp(&hdr[0], "MBSSLIST", 8) != 0 || (uint32_t)hdr[1] != 1) throw genome::gnException(path + ": not a version-1 sorted mer list");
        seed_ = hdr[2]; unique_ = hdr[5];
        std::vector<uint64_t> words((hdr[3] + 31) / 32);
        f.read((char*)words.data(), words.size() * 8);
        pos_.resize(hdr[4]);
        f.read((char*)pos_.data(), pos_.size() * 4);
        if (!f) throw genome::gnException(path + ": truncated");
        std::string s(hdr[3], 'A');
        for (size_t i = 0; i < s.size(); ++i) s[i] = "ACGT"[(words[i / 32] >> (62 - 2 * (i % 32))) & 3];
        own_ = genome::gnSequence(s); seq_ = &own_;
    }
    virtual gnSeqI UniqueMerCount() { return unique_ != ~0ull ? unique_ : SortedMerList::UniqueMerCount(); }
    /* the sorted position array (libMems: operator[] / Read) */
    const std::vector<uint32_t>& Positions() const { return pos_; }
    gnSeqI SMLLength() const { return pos_.size(); }
private:
    genome::gnSequence own_;
    std::vector<uint32_t> pos_;
};

class MatchList : public std::vector<Match*> {
public:
    std::vector<std::string> seq_filename, sml_filename;
    std::vector<genome::gnSequence*> seq_table;
    std::vector<SortedMerList*> sml_table;

    /* MatchList::CreateMemorySMLs(mer_size, log, seed_rank): fixes the seed and creates the façades; the sorted
     * mer lists themselves are built on the device inside FindMatches. mer_size 0 = default weight. */
    void CreateMemorySMLs(uint32 mer_size, std::ostream* log_stream, int seed_rank = 0) {
        if (mer_size == 0) mer_size = GetDefaultMerSize(seq_table);
        seed_table_notice();
        uint64 seed = mb_get_seed((int)mer_size, seed_rank);
        if (!mb_seed_valid(seed)) throw genome::gnException("invalid seed weight / rank");
        for (SortedMerList* s : sml_table) delete s;
        sml_table.clear();
        for (genome::gnSequence* seq : seq_table) sml_table.push_back(new SortedMerList(seq, seed));
        if (log_stream) *log_stream << "Using weight " << mb_seed_weight(seed) << " mers, length " << mb_seed_length(seed) << "\n";
    }
    /* LoadSMLs(mer_size, log, seed_rank[, solid, force_recreate]) (src/mauveAligner.cpp:456, src/repeatoire.cpp:1850): with
     * sml_filename set, each list is loaded from its file when that exists and was built with the same seed, otherwise
     * (or with force_recreate) built on the device and written there; without file names it is CreateMemorySMLs. */
    void LoadSMLs(uint32 mer_size, std::ostream* log_stream, int seed_rank = 0, bool solid = false, bool force_recreate = false) {
        if (sml_filename.size() != seq_table.size()) { CreateMemorySMLs(mer_size, log_stream, solid ? SOLID_SEED : seed_rank); return; }
        if (mer_size == 0) mer_size = GetDefaultMerSize(seq_table);
        uint64 seed = mb_get_seed((int)mer_size, solid ? SOLID_SEED : seed_rank);
        if (!mb_seed_valid(seed)) throw genome::gnException("invalid seed weight / rank");
        for (SortedMerList* s : sml_table) delete s;
        sml_table.clear();
        for (size_t i = 0; i < seq_table.size(); ++i) {
            DNAFileSML* sml = new DNAFileSML();
            bool ok = false;
            if (!force_recreate) {
                try { sml->LoadFile(sml_filename[i]); ok = sml->Seed() == seed && sml->Length() == seq_table[i]->length(); }
                catch (const genome::gnException&) { ok = false; }
            }
            if (!ok) {
                if (log_stream) *log_stream << "Creating sorted mer list " << sml_filename[i] << "\n";
                sml->Create(*seq_table[i], seed);
                sml->WriteFile(sml_filename[i]);
            }
            sml_table.push_back(sml);
        }
    }
    /* MatchList::MultiplicityFilter(mult) (src/mauveAligner.cpp:600, src/transposeCoordinates.cpp:48): keep only the
     * matches present in exactly `mult` sequences; the others are released (the list owns what it drops). */
    void MultiplicityFilter(unsigned mult) {
        size_t kept = 0;
        for (size_t i = 0; i < size(); ++i) {
            Match* m = (*this)[i];
            if (m->Multiplicity() == mult) (*this)[kept++] = m;
            else m->Free();
        }
        resize(kept);
    }
    static uint32 GetDefaultMerSize(const std::vector<genome::gnSequence*>& seqs) {
        gnSeqI total = 0;
        for (auto* s : seqs) total += s->length();
        return getDefaultSeedWeight(seqs.empty() ? 0 : total / seqs.size());
    }
};

inline void WriteList(const MatchList& ml, std::ostream& os) {
    os << "FormatVersion\t3\nSequenceCount\t" << ml.seq_table.size() << "\n";
    for (size_t i = 0; i < ml.seq_table.size(); ++i) {
        os << "Sequence" << i << "File\t" << (i < ml.seq_filename.size() ? ml.seq_filename[i] : "") << "\n";
        os << "Sequence" << i << "Length\t" << ml.seq_table[i]->length() << "\n";
    }
    os << "MatchCount\t" << ml.size() << "\n";
    for (const Match* m : ml) os << *m << "\n";
}

/* EliminateOverlaps(MatchList&) (src/mauveAligner.cpp:594-596,611-612: "only count each base pair once"): after the call
 * no two matches of the list overlap in any sequence.  libMems' body is not in the tree; the rule here:
 * sequence by sequence, the matches present in it are swept by (left end, position in the list); the part of a match
 * that earlier matches of the sweep already cover — always a prefix in that sequence's coordinates — is cropped off
 * (CropStart for a forward component, CropEnd for a reverse one), a match covered entirely is released. */
inline void EliminateOverlaps(MatchList& ml) {
    const uint nseq = ml.empty() ? 0 : ml[0]->SeqCount();
    for (uint seqI = 0; seqI < nseq; ++seqI) {
        std::vector<size_t> order;
        for (size_t i = 0; i < ml.size(); ++i) if (ml[i] && ml[i]->Start(seqI) != NO_MATCH) order.push_back(i);
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return ml[a]->LeftEnd(seqI) < ml[b]->LeftEnd(seqI); });
        int64 covered = 0; /* right end of everything swept so far */
        for (size_t i : order) {
            Match* m = ml[i];
            const int64 l = m->LeftEnd(seqI), r = m->RightEnd(seqI);
            if (r <= covered) { m->Free(); ml[i] = nullptr; continue; }
            if (l <= covered) {
                const gnSeqI ov = (gnSeqI)(covered - l + 1);
                if (m->Orientation(seqI) == 0) m->CropStart(ov); else m->CropEnd(ov);
            }
            covered = r;
        }
    }
    size_t kept = 0;
    for (size_t i = 0; i < ml.size(); ++i) if (ml[i]) ml[kept++] = ml[i];
    ml.resize(kept);
}

/* transposeMatches(MatchList&, seqI, seq_regions) (src/mauveAligner.cpp:628-637, src/transposeCoordinates.cpp:46-65): the
 * sorted mer list of sequence seqI was built over a FILTERED sequence — the concatenation of the used regions
 * seq_regions = {first_0, last_0, first_1, last_1, ...} (1-based, inclusive, ascending) of the original — and the
 * matches carry filtered coordinates; put them back into the original coordinate system.  A match that runs across a
 * region boundary is split there (every piece keeps the columns of all its components).  libMems' body is not in the
 * tree; this is DESIGN.md D19. */
inline void transposeMatches(MatchList& ml, uint seqI, const std::vector<int64>& seq_regions) {
    if (seq_regions.size() < 2) return;
    const size_t nreg = seq_regions.size() / 2;
    std::vector<int64> cum(nreg + 1, 0); /* filtered coordinates before region k */
    for (size_t k = 0; k < nreg; ++k) cum[k + 1] = cum[k] + (seq_regions[2 * k + 1] - seq_regions[2 * k] + 1);
    std::vector<Match*> out;
    for (Match* m : ml) {
        if (m->Start(seqI) == NO_MATCH) { out.push_back(m); continue; }
        while (m) {
            const int64 l = m->LeftEnd(seqI);               /* filtered, 1-based */
            size_t k = (size_t)(std::upper_bound(cum.begin(), cum.end(), l - 1) - cum.begin()) - 1;
            if (k >= nreg) k = nreg - 1;                     /* beyond the last region: extrapolate from it */
            const int64 room = cum[k + 1] - (l - 1);         /* columns left in region k */
            Match* rest = nullptr;
            if (k + 1 < nreg && (int64)m->Length() > room) { /* split at the boundary, in seqI's coordinates */
                rest = m->Copy();
                const gnSeqI tail = (gnSeqI)((int64)m->Length() - room);
                if (m->Orientation(seqI) == 0) { m->CropEnd(tail); rest->CropStart((gnSeqI)room); }
                else { m->CropStart(tail); rest->CropEnd((gnSeqI)room); }
            }
            const int64 nl = seq_regions[2 * k] + (m->LeftEnd(seqI) - 1 - cum[k]);
            m->SetStart(seqI, m->Start(seqI) < 0 ? -nl : nl);
            out.push_back(m);
            m = rest;
        }
    }
    ml.assign(out.begin(), out.end());
}

/* ReadList: the inverse of WriteList (match lines only; the header's file names refill seq_filename).  Used for the
 * --match-input hand-off (src/mauveAligner.cpp:489-503). */
inline void ReadList(MatchList& ml, std::istream& is) {
    std::string key;
    size_t nseq = 0, nmatch = 0;
    for (Match* m : ml) m->Free();
    ml.clear();
    ml.seq_filename.clear();
    if (!(is >> key) || key != "FormatVersion") throw genome::gnException("ReadList: not a match list");
    int version; is >> version;
    is >> key >> nseq;
    for (size_t i = 0; i < nseq; ++i) {
        std::string line;
        is >> key; std::getline(is, line);
        ml.seq_filename.push_back(line.size() > 1 ? line.substr(1) : std::string());
        gnSeqI len; is >> key >> len;
    }
    is >> key >> nmatch;
    for (size_t k = 0; k < nmatch; ++k) {
        gnSeqI len; is >> len;
        Match* m = new Match((uint)nseq);
        m->SetLength(len);
        for (size_t i = 0; i < nseq; ++i) { int64 st; is >> st; m->SetStart((uint)i, st); }
        if (!is) { m->Free(); throw genome::gnException("ReadList: truncated match list"); }
        ml.push_back(m);
    }
}

/* mems::MatchFinder — holds the device context(s) and the sequences added so far.
 * One GPU by default (MAUVE_B200_DEVICE, default 0).  MAUVE_B200_DEVICES=0,1,2,... makes every finder hold one context
 * per listed device (a device may be listed more than once) and run its MB_MODE_UNIQUE searches — UniqueMatchFinder,
 * MemHash, MaskedMemHash — through mb_find_multi: one process, one host thread per GPU, exchanges over NVLink peer
 * access; the other policies keep running on the first device.  The match list is the same either way. */
class MatchFinder {
public:
    MatchFinder() : ctx_(nullptr), seq_count(0), seed_(0), log_(nullptr) {}
    MatchFinder(const MatchFinder& o) : ctx_(nullptr), seq_count(0), seed_(o.seed_), log_(o.log_) {}
    virtual ~MatchFinder() { destroy_ctxs(); }
    virtual MatchFinder* Clone() const = 0;

    virtual boolean AddSequence(SortedMerList* sar, genome::gnSequence* seq = nullptr) {
        if (!sar) return false;
        if (!ensure_ctx()) return false;
        if (seq_count == 0) seed_ = sar->Seed();
        else if (sar->Seed() != seed_) { genome::ErrorMsg("AddSequence: all sorted mer lists must use the same seed\n"); return false; }
        const genome::gnSequence* s = seq ? seq : sar->Sequence();
        int rc = mb_add_sequence(ctx_, (const uint8_t*)s->data().data(), s->length(), 0, nullptr);
        for (size_t k = 0; rc == MB_OK && k < more_.size(); ++k) rc = mb_add_sequence(more_[k], (const uint8_t*)s->data().data(), s->length(), 0, nullptr);
        if (rc != MB_OK) { report("AddSequence", rc); return false; }
        ++seq_count;
        return true;
    }
    virtual void Clear() {}
    virtual void ClearSequences() {
        if (ctx_) mb_clear_sequences(ctx_);
        for (mb_ctx* c : more_) mb_clear_sequences(c);
        seq_count = 0;
    }
    void LogProgress(std::ostream* os) { log_ = os; }
    uint32 SeqCount() const { return seq_count; }

protected:
    bool ensure_ctx() {
        if (ctx_) return true;
        std::vector<int> devs = devices_from_env();
        int rc = mb_ctx_create(&ctx_, devs[0]);
        if (rc != MB_OK) { ctx_ = nullptr; report("mb_ctx_create", rc); return false; }
        for (size_t k = 1; k < devs.size(); ++k) {
            mb_ctx* c = nullptr;
            rc = mb_ctx_create(&c, devs[k]);
            if (rc != MB_OK) { report("mb_ctx_create", rc); destroy_ctxs(); return false; }
            more_.push_back(c);
        }
        return true;
    }
    void destroy_ctxs() {
        for (mb_ctx* c : more_) mb_ctx_destroy(c);
        more_.clear();
        if (ctx_) mb_ctx_destroy(ctx_);
        ctx_ = nullptr;
    }
    static int device_from_env() { const char* e = getenv("MAUVE_B200_DEVICE"); return e ? atoi(e) : 0; }
    static std::vector<int> devices_from_env() {
        std::vector<int> devs;
        if (const char* e = getenv("MAUVE_B200_DEVICES")) {
            std::stringstream ss(e);
            std::string tok;
            while (std::getline(ss, tok, ',')) if (!tok.empty()) devs.push_back(atoi(tok.c_str()));
        }
        if (devs.empty()) devs.push_back(device_from_env());
        return devs;
    }
    void report(const char* what, int rc) const {
        genome::ErrorMsg(std::string(what) + ": " + mb_strerror(rc) + (ctx_ ? std::string(" ") + mb_last_cuda_error(ctx_) : "") + "\n");
    }
    /* (the compact result form: 5 bytes per component over PCIe instead of 12) */
    void append(const mb_result_compact* r, MatchList& out, bool dense) const {
        for (uint64_t i = 0; i < r->n_matches; ++i) {
            uint64_t a = r->comp_off[i], b = r->comp_off[i + 1];
            Match* m = new Match(dense ? seq_count : (uint)(b - a));
            m->SetLength(r->length[i]);
            for (uint64_t k = a; k < b; ++k) m->SetStart(dense ? r->comp_seq8[k] : (uint)(k - a), r->comp_start32[k]);
            out.push_back(m);
        }
    }
    /* run one search with the sequences added so far; fills `out` (dense = one column per sequence) */
    boolean run(const mb_params& p, MatchList& out, bool dense) {
        if (!ctx_ || seq_count == 0) return false;
        int rc = mb_set_seed(ctx_, seed_);
        if (!more_.empty() && p.mode == MB_MODE_UNIQUE) {
            // every GPU: the ranks' pieces are ascending ranges of the canonical order, appended in rank order
            std::vector<mb_ctx*> all(1, ctx_);
            all.insert(all.end(), more_.begin(), more_.end());
            for (size_t k = 1; rc == MB_OK && k < all.size(); ++k) rc = mb_set_seed(all[k], seed_);
            if (rc == MB_OK) rc = mb_find_multi(all.data(), (int)all.size(), &p);
            if (rc != MB_OK) { report("FindMatches (multi-GPU)", rc); return false; }
            uint64_t total = 0;
            for (mb_ctx* c : all) {
                const mb_result_compact* r = nullptr;
                rc = mb_fetch_result_compact(c, &r);
                if (rc != MB_OK) { report("mb_fetch_result", rc); return false; }
                append(r, out, dense);
                total += r->n_matches;
            }
            if (log_) *log_ << total << " matches\n";
            return true;
        }
        const mb_result_compact* r = nullptr;
        if (rc == MB_OK) rc = mb_find_compact(ctx_, &p, &r);
        if (rc != MB_OK) { report("FindMatches", rc); return false; }
        if (log_) *log_ << r->n_matches << " matches\n";
        append(r, out, dense);
        return true;
    }
    mb_ctx* ctx_;
    std::vector<mb_ctx*> more_;   /* ranks 1.. of the multi-GPU search */
    uint32 seq_count;
    uint64 seed_;
    std::ostream* log_;
};

/* mems::MemHash: multi-MUM search = unique-seed filter + extension + containment de-dup. */
class MemHash : public MatchFinder {
public:
    MemHash() : mask_(0) {}
    MemHash(const MemHash& o) : MatchFinder(o), mask_(o.mask_) {}
    virtual MemHash* Clone() const { return new MemHash(*this); }

    /* FindMatches(MatchList&): adds every sequence of the list, searches, and fills the list with the matches found so
     * far (GetMatchList semantics, src/progressiveMauve.cpp:538-547).  As in libMems the table persists until Clear():
     * a second FindMatches (the seed-family search, src/progressiveMauve.cpp:503-548, makes three with ClearSequences()
     * in between) drops every candidate that a match found so far contains — on the device, mb_accumulate — and
     * GetMatchList returns the union of all calls in canonical order. */
    virtual boolean FindMatches(MatchList& match_list) {
        for (size_t i = 0; i < match_list.seq_table.size(); ++i)
            if (!AddSequence(match_list.sml_table[i], match_list.seq_table[i])) {
                genome::ErrorMsg("Error adding " + (i < match_list.seq_filename.size() ? match_list.seq_filename[i] : std::string("sequence")) + "\n");
                return false;
            }
        if (!more_.empty() && !found_.empty()) {
            genome::ErrorMsg("FindMatches: a second search into the same table is not available with MAUVE_B200_DEVICES (call Clear() first)\n");
            return false;
        }
        if (ctx_ && more_.empty()) mb_accumulate(ctx_, 1);
        const size_t before = found_.size();
        mb_params p = {MB_MODE_UNIQUE, 0, 2, 1000, mask_};
        if (!run(p, found_, true)) return false;
        if (before) std::sort(found_.begin(), found_.end(), canonical_less);
        GetMatchList(match_list);
        return true;
    }
    /* FindMatchesFromPosition(match_list, start_offsets): the resumable entry of mauveAligner
     * (src/mauveAligner.cpp:585).  The device search takes milliseconds, so offsets are ignored. */
    boolean FindMatchesFromPosition(MatchList& match_list, const std::vector<gnSeqI>&) { return FindMatches(match_list); }
    virtual void GetMatchList(MatchList& ml) {
        ml.clear();
        for (Match* m : found_) ml.push_back(m->Copy());
    }
    virtual void Clear() {
        for (Match* m : found_) m->Free();
        found_.clear();
        if (ctx_) mb_accumulate(ctx_, 0);
    }
    virtual ~MemHash() { for (Match* m : found_) m->Free(); }
    /* canonical order of the match list: |start| per sequence, then signs, then length */
    static bool canonical_less(const Match* a, const Match* b) {
        const uint n = a->SeqCount();
        for (uint i = 0; i < n; ++i) {
            int64 x = a->Start(i) < 0 ? -a->Start(i) : a->Start(i), y = b->Start(i) < 0 ? -b->Start(i) : b->Start(i);
            if (x != y) return x < y;
        }
        for (uint i = 0; i < n; ++i) {
            bool x = a->Start(i) < 0, y = b->Start(i) < 0;
            if (x != y) return y;
        }
        return a->Length() < b->Length();
    }
    /* Many small searches in ONE pass of the device pipeline (recursive anchoring re-runs the search inside every gap
     * between anchors: `recursive` flag src/mauveAligner.cpp:94,698; SetRecursive src/progressiveMauve.cpp:661-664).
     * problems[i] = the MatchList of gap i with seq_table / sml_table filled (the same number of sequences and the same
     * seed everywhere; an empty sequence = the gap has none in that genome).  Every list receives exactly the matches its
     * own FindMatches + Clear() would have produced. */
    boolean FindMatchesBatch(std::vector<MatchList*>& problems) {
        if (problems.empty()) return true;
        if (!ensure_ctx()) return false;
        const size_t nseq = problems[0]->seq_table.size();
        std::vector<const uint8_t*> ptrs;
        std::vector<uint64_t> lens;
        uint64 seed = 0;
        for (MatchList* ml : problems) {
            if (ml->seq_table.size() != nseq || ml->sml_table.size() != nseq) { genome::ErrorMsg("FindMatchesBatch: every problem needs the same number of sequences\n"); return false; }
            for (size_t g = 0; g < nseq; ++g) {
                if (!seed) seed = ml->sml_table[g]->Seed();
                else if (ml->sml_table[g]->Seed() != seed) { genome::ErrorMsg("FindMatchesBatch: one seed for the whole batch\n"); return false; }
                ptrs.push_back((const uint8_t*)ml->seq_table[g]->data().data());
                lens.push_back(ml->seq_table[g]->length());
            }
        }
        int rc = mb_set_seed(ctx_, seed);
        mb_params p = {MB_MODE_UNIQUE, 0, 2, 1000, mask_};
        const mb_batch_result* r = nullptr;
        if (rc == MB_OK) rc = mb_find_batch(ctx_, &p, (uint32_t)problems.size(), (uint32_t)nseq, ptrs.data(), lens.data(), &r);
        if (rc != MB_OK) { report("FindMatchesBatch", rc); return false; }
        for (size_t i = 0; i < problems.size(); ++i) {
            MatchList& out = *problems[i];
            for (Match* m : out) m->Free();
            out.clear();
            for (uint64_t j = r->match_off[i]; j < r->match_off[i + 1]; ++j) {
                Match* m = new Match((uint)nseq);
                m->SetLength(r->length[j]);
                for (uint64_t k = r->comp_off[j]; k < r->comp_off[j + 1]; ++k) m->SetStart(r->comp_seq[k], r->comp_start[k]);
                out.push_back(m);
            }
        }
        seq_count = 0; /* the batch replaced the context's sequences */
        return true;
    }
protected:
    uint64 mask_;
    MatchList found_;
};

/* mems::PairwiseMatchFinder (src/progressiveMauve.cpp:496-501): MemHash's unique filter, then one two-genome match
 * per pair of the bucket's unique genomes. */
class PairwiseMatchFinder : public MemHash {
public:
    virtual PairwiseMatchFinder* Clone() const { return new PairwiseMatchFinder(*this); }
    virtual boolean FindMatches(MatchList& match_list) {
        for (size_t i = 0; i < match_list.seq_table.size(); ++i)
            if (!AddSequence(match_list.sml_table[i], match_list.seq_table[i])) return false;
        Clear(); /* pairs of one search only (no table across calls on this policy) */
        mb_params p = {MB_MODE_PAIRWISE, 0, 2, 1000, 0};
        if (!run(p, found_, true)) return false;
        GetMatchList(match_list);
        return true;
    }
};

/* mems::RepeatHash (mauveAligner --repeats, src/mauveAligner.cpp:480-487): repeats inside ONE sequence.  Every bucket of
 * 2..255 occurrences becomes one match with a column per occurrence (Match(multiplicity), as SeedMatchEnumerator's),
 * extended and de-duplicated like a MemHash entry. */
class RepeatHash : public MatchFinder {
public:
    virtual RepeatHash* Clone() const { return new RepeatHash(*this); }
    virtual boolean FindMatches(MatchList& match_list) {
        for (size_t i = 0; i < match_list.seq_table.size(); ++i)
            if (!AddSequence(match_list.sml_table[i], match_list.seq_table[i])) return false;
        if (seq_count != 1) { genome::ErrorMsg("RepeatHash: exactly one sequence expected\n"); return false; }
        for (Match* m : match_list) m->Free();
        match_list.clear();
        mb_params p = {MB_MODE_REPEAT, 0, 2, 255, 0};
        return run(p, match_list, false);
    }
};

class MaskedMemHash : public MemHash {
public:
    virtual MaskedMemHash* Clone() const { return new MaskedMemHash(*this); }
    void SetMask(uint64 mask) { mask_ = mask; }
};

} // namespace mems

/* src/UniqueMatchFinder.h:21-32 */
class UniqueMatchFinder : public mems::MemHash {
public:
    UniqueMatchFinder() {}
    ~UniqueMatchFinder() {}
    UniqueMatchFinder(const UniqueMatchFinder& mh) : mems::MemHash(mh) {}
    virtual UniqueMatchFinder* Clone() const { return new UniqueMatchFinder(*this); }
};

/* src/SeedMatchEnumerator.h:14-49 */
class SeedMatchEnumerator : public mems::MatchFinder {
public:
    virtual SeedMatchEnumerator* Clone() const { return new SeedMatchEnumerator(*this); }

    void FindMatches(mems::MatchList& match_list, size_t min_multi = 2, size_t max_multi = 1000, bool direct_repeats_only = false) {
        this->max_multiplicity = max_multi;
        this->min_multiplicity = min_multi;
        this->only_direct = direct_repeats_only;
        for (size_t seqI = 0; seqI < match_list.seq_table.size(); ++seqI) {
            if (!AddSequence(match_list.sml_table[seqI], match_list.seq_table[seqI])) {
                genome::ErrorMsg("Error adding " + (seqI < match_list.seq_filename.size() ? match_list.seq_filename[seqI] : std::string("sequence")) + "\n");
                return;
            }
        }
        mlist.clear(); /* the reference never clears mlist; fixed here */
        CreateMatches();
        match_list.clear();
        match_list.insert(match_list.end(), mlist.begin(), mlist.end());
    }
    /* does nothing unless exactly one sequence was added (src/SeedMatchEnumerator.h:59-65) */
    virtual boolean CreateMatches() {
        if (seq_count == 1) {
            mb_params p = {MB_MODE_SEED_ENUM, only_direct ? 1 : 0, (uint64_t)min_multiplicity, (uint64_t)max_multiplicity, 0};
            return run(p, mlist, false);
        }
        return false;
    }
    /* repeatoire part 3 (src/repeatoire.cpp:1944-1966): the match position lookup table over the one sequence — entry p
     * (1-based left end) = (match, component) that starts there, (NULL, 0) elsewhere — filled from the device-built index
     * arrays; `ml` must be the list the last FindMatches produced (its order is repeatoire's seed_sort_list order). */
    boolean GetMatchPositionLookupTable(const mems::MatchList& ml, std::vector<std::pair<mems::Match*, size_t> >& table) {
        const uint32_t *mo = nullptr, *co = nullptr;
        uint64_t n = 0;
        int rc = ctx_ ? mb_position_table(ctx_, &mo, &co, &n) : MB_E_STATE;
        if (rc != MB_OK) { report("GetMatchPositionLookupTable", rc); return false; }
        table.assign(n, std::make_pair((mems::Match*)nullptr, (size_t)0));
        for (uint64_t p = 0; p < n; ++p)
            if (mo[p] != 0xFFFFFFFFu && mo[p] < ml.size()) table[p] = std::make_pair(ml[mo[p]], (size_t)co[p]);
        return true;
    }
protected:
    mems::MatchList mlist;
private:
    size_t max_multiplicity = 1000;
    size_t min_multiplicity = 2;
    bool only_direct = false;
};

#endif
