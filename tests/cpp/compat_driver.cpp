// Reads sequences (one per line) from argv[2..] files' first line, runs the reference-shaped host API
// (include/mems_compat/mems_compat.h) and prints match lists in WriteList form.  Used by the GPU tests;
// the calls mirror src/progressiveMauve.cpp:437-558, src/repeatoire.cpp:1836-1867, src/uniqueMerCount.cpp:23-40.
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
    if (argc < 5) { cerr << "usage: compat_driver <umf|sme|count> <weight> <rank> <seq files...>\n"; return -1; }
    string what = argv[1];
    int weight = atoi(argv[2]), rank = atoi(argv[3]);
    MatchList ml;
    for (int i = 4; i < argc; ++i) {
        ml.seq_filename.push_back(argv[i]);
        ml.seq_table.push_back(new gnSequence(read_seq(argv[i])));
    }
    ml.CreateMemorySMLs(weight, nullptr, rank);
    if (what == "umf") {
        UniqueMatchFinder umf;
        umf.LogProgress(nullptr);
        if (!umf.FindMatches(ml)) return -2;
        umf.Clear();
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
