// terefiner_main.cpp -- TERefiner_b200: the two modes of TERefiner_1 (TERefiner/main.cpp) that ARE the affine local aligner,
// with the reference's command line and output:
//   -M -r SEQ1 -s SEQ2     LocalAlignment::optAlign (main.cpp:207-213): "start_ref end_ref start_sgmt end_sgmt\n"
//   -A -r SEQ1 -s SEQ2     RepeatsClassifier::validateRepeats (main.cpp:202-206): one number
// and, because one pair per process leaves a GPU idle,
//   -M|-A --batch LIST     LIST holds one pair per line, "SEQ1<TAB>SEQ2"; one output line per pair, in order; every pair of
//                          the list goes through the same launches.
// The modes GAPPadder itself uses (-U, -P: the dedup stage) are `ContigsMerger_b200 --dedup`; the BAM classifiers have no
// alignment in them and are not provided.  No CPU fallback: without an sm_100 device the process exits with an error.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "gappadder_b200.h"
#include "local_alignment.hpp"

int main(int argc, char** argv)
{
    bool bm = false, ba = false;
    std::string ref, sgmt, batch;
    bool have_r = false, have_s = false;
    int gpu = 0;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "-M") bm = true;
        else if (a == "-A") ba = true;
        else if (a == "-r" && i + 1 < argc) { ref = argv[++i]; have_r = true; }
        else if (a == "-s" && i + 1 < argc) { sgmt = argv[++i]; have_s = true; }
        else if (a == "--batch" && i + 1 < argc) batch = argv[++i];
        else if (a == "--gpu" && i + 1 < argc) gpu = atoi(argv[++i]);
        else if (a.size() == 2 && a[0] == '-' && strchr("RCTOPKLSUGBE", a[1])) {
            fprintf(stderr, "TERefiner_b200: mode %s is not the alignment path (the dedup stage -U / -P is ContigsMerger_b200 --dedup)\n", a.c_str());
            return 2;
        } else { fprintf(stderr, "TERefiner_b200: unknown argument %s\n", a.c_str()); return 2; }
    }
    if (bm == ba || (batch.empty() && !(have_r && have_s))) {
        fprintf(stderr, "usage: TERefiner_b200 -M|-A -r SEQ1 -s SEQ2   |   TERefiner_b200 -M|-A --batch LIST\n");
        return 2;
    }
    std::vector<gpm::LaPair> pairs;
    if (batch.empty()) pairs.push_back({ref, sgmt});
    else {
        std::ifstream f(batch);
        if (!f) { fprintf(stderr, "TERefiner_b200: cannot read %s\n", batch.c_str()); return 2; }
        std::string line;
        while (std::getline(f, line)) {
            if (!line.empty() && line.back() == '\r') line.pop_back();
            const size_t tab = line.find('\t');
            if (tab == std::string::npos) { fprintf(stderr, "TERefiner_b200: %s: a line without a tab\n", batch.c_str()); return 2; }
            pairs.push_back({line.substr(0, tab), line.substr(tab + 1)});
        }
    }
    gp_ctx* ctx = nullptr;
    if (gp_create(gpu, &ctx) != GP_OK) { fprintf(stderr, "TERefiner_b200: %s\n", gp_last_error(nullptr)); return 3; }
    gpm::LocalAlignment la(ctx);
    std::string text;
    bool ok;
    if (bm) {
        std::vector<gpm::LaHit> h;
        ok = la.optAlignBatch(pairs, h);
        for (const gpm::LaHit& x : h)                                                       // main.cpp:212
            text += std::to_string(x.start_ref) + " " + std::to_string(x.end_ref) + " " + std::to_string(x.start_sgmt) + " " + std::to_string(x.end_sgmt) + "\n";
    } else {
        std::vector<int> v;
        ok = gpm::validate_repeats_batch(la, pairs, v);
        for (int x : v) text += std::to_string(x) + "\n";                                   // RepeatsClassifier.cpp:107-110
    }
    if (!ok) { fprintf(stderr, "TERefiner_b200: %s\n", la.error().c_str()); gp_destroy(ctx); return 3; }   // never partial output
    fwrite(text.data(), 1, text.size(), stdout);
    gp_destroy(ctx);
    return 0;
}
