// Reads sequences (one per line) from argv[2..] files' first line, runs the reference-shaped host API
// (include/mems_compat/mems_compat.h) and prints match lists in WriteList form.  Used by the GPU tests;
// the calls mirror src/progressiveMauve.cpp:437-558, src/repeatoire.cpp:1836-1867, src/uniqueMerCount.cpp:23-40.
#include <algorithm>
#include <fstream>
#include <iostream>
#include <string>

#include "mems_compat/mems_compat.h"

using namespace std;
using namespace genome;
using namespace mems;

static string read_seq(const char* path) {
    ifstream in(path);
    string s;
    getline(in, s);
    return s;
}

int main(int argc, char** argv) {
    if (argc < 3) { cerr << "usage: compat_driver <umf|pmf|sme|repeats|count|sml|family|batch> <weight> <rank> <seq files...> | smlcount <file.sslist> | readlist <file>\n"; return -1; }
    string what = argv[1];
    if (what == "smlcount") {
        // src/uniqueMerCount.cpp:29-39, line for line
        DNAFileSML file_sml;
        try { file_sml.LoadFile(argv[2]); }
        catch (gnException& gne) { cerr << gne.what() << endl; return -2; }
        cout << endl << file_sml.UniqueMerCount() << endl;
        for (uint32_t p : file_sml.Positions()) cout << p << "\n";
        return 0;
    }
    if (what == "readlist") {
        MatchList in;
        ifstream f(argv[2]);
        ReadList(in, f);
        for (const string& n : in.seq_filename) in.seq_table.push_back(new gnSequence(read_seq(n.c_str())));
        WriteList(in, cout);
        return 0;
    }
    if (what == "overlaps" || what == "transpose") {
        // the match-list post-filters of src/mauveAligner.cpp:594-596,628-637 / src/transposeCoordinates.cpp:46-65 (host only)
        MatchList in;
        ifstream f(argv[2]);
        ReadList(in, f);
        for (size_t i = 0; i < in.seq_filename.size(); ++i) in.seq_table.push_back(new gnSequence());
        if (what == "overlaps") EliminateOverlaps(in);
        else {
            vector<int64> coords;
            for (int i = 4; i < argc; ++i) coords.push_back(atoll(argv[i]));
            transposeMatches(in, (uint)atoi(argv[3]), coords);
        }
        for (const Match* m : in) cout << *m << "\n";
        return 0;
    }
    if (argc < 5) return -1;
    int weight = atoi(argv[2]), rank = atoi(argv[3]);
    MatchList ml;
    for (int i = 4; i < argc; ++i) {
        ml.seq_filename.push_back(argv[i]);
        ml.seq_table.push_back(new gnSequence(read_seq(argv[i])));
    }
    if (what == "sml") {
        // LoadSMLs with on-disk sorted mer lists (src/mauveAligner.cpp:450-456): <seq>.sslist next to every sequence
        for (int i = 4; i < argc; ++i) ml.sml_filename.push_back(string(argv[i]) + ".sslist");
        ml.LoadSMLs(weight, &cerr, rank);
        ml.LoadSMLs(weight, &cerr, rank); // second call must find the files and not rebuild
        cout << endl << ml.sml_table[0]->UniqueMerCount() << endl;
        return 0;
    }
    if (what == "batch") {
        // recursive anchoring in miniature: every sequence file is cut into pieces of 400 bases, piece i of every file = gap i;
        // all gaps searched in one pass (MemHash::FindMatchesBatch), printed gap by gap
        const size_t piece = 400;
        vector<string> full;
        size_t ngap = 0;
        for (int i = 4; i < argc; ++i) { full.push_back(read_seq(argv[i])); ngap = max(ngap, (full.back().size() + piece - 1) / piece); }
        vector<MatchList> gaps(ngap);
        vector<MatchList*> ptrs;
        for (size_t gI = 0; gI < ngap; ++gI) {
            for (const string& s : full) gaps[gI].seq_table.push_back(new gnSequence(gI * piece < s.size() ? s.substr(gI * piece, piece) : string()));
            gaps[gI].CreateMemorySMLs(weight, nullptr, rank);
            ptrs.push_back(&gaps[gI]);
        }
        UniqueMatchFinder umf;
        if (!umf.FindMatchesBatch(ptrs)) return -2;
        for (size_t gI = 0; gI < ngap; ++gI) {
            cout << "Gap\t" << gI << "\t" << gaps[gI].size() << "\n";
            for (const Match* m : gaps[gI]) cout << *m << "\n";
        }
        return 0;
    }
    if (what == "family") {
        // the seed-family search of src/progressiveMauve.cpp:503-548, call for call: the longest pattern first, ONE
        // UniqueMatchFinder, ClearSequences() between the patterns, one GetMatchList + Clear() at the end
        int mer_size = weight;
        vector<pair<int, int> > length_ranks(3);
        length_ranks[0] = make_pair(getSeedLength(getSeed(mer_size, 0)), 0);
        length_ranks[1] = make_pair(getSeedLength(getSeed(mer_size, 1)), 1);
        length_ranks[2] = make_pair(getSeedLength(getSeed(mer_size, 2)), 2);
        std::sort(length_ranks.begin(), length_ranks.end());
        UniqueMatchFinder umf;
        for (int seedI = 2; seedI >= 0; seedI--) {
            umf.LogProgress(nullptr);
            MatchList cur_list;
            cur_list.seq_filename = ml.seq_filename;
            cur_list.seq_table = ml.seq_table;
            cur_list.CreateMemorySMLs(mer_size, nullptr, length_ranks[seedI].second);
            umf.FindMatches(cur_list);
            umf.ClearSequences();
            for (size_t smlI = 0; smlI < cur_list.sml_table.size(); smlI++) delete cur_list.sml_table[smlI];
            for (size_t curI = 0; curI < cur_list.size(); curI++) cur_list[curI]->Free();
        }
        umf.GetMatchList(ml);
        umf.Clear();
        WriteList(ml, cout);
        for (Match* m : ml) m->Free();
        for (size_t i = 0; i < ml.seq_table.size(); ++i) delete ml.seq_table[i];
        return 0;
    }
    ml.CreateMemorySMLs(weight, nullptr, rank);
    if (what == "pmf") {
        PairwiseMatchFinder pmf;
        if (!pmf.FindMatches(ml)) return -2;
        pmf.Clear();
        WriteList(ml, cout);
    } else if (what == "umf") {
        UniqueMatchFinder umf;
        umf.LogProgress(nullptr);
        if (!umf.FindMatches(ml)) return -2;
        umf.Clear();
        WriteList(ml, cout);
    } else if (what == "repeats") {
        // src/mauveAligner.cpp:480-487
        RepeatHash repeat_finder;
        repeat_finder.LogProgress(nullptr);
        repeat_finder.FindMatches(ml);
        WriteList(ml, cout);
    } else if (what == "sme") {
        SeedMatchEnumerator sme;
        sme.FindMatches(ml, 2, 500, false);
        WriteList(ml, cout);
    } else if (what == "count") {
        cout << endl << ml.sml_table[0]->UniqueMerCount() << endl;
    }
    for (Match* m : ml) m->Free();
    for (size_t i = 0; i < ml.seq_table.size(); ++i) { delete ml.sml_table[i]; delete ml.seq_table[i]; }
    return 0;
}
