// terefiner_main.cpp -- TERefiner_b200: the two modes of TERefiner_1 (TERefiner/main.cpp) that ARE the affine local aligner,
// with the reference's command line and output:
//   -M -r SEQ1 -s SEQ2     LocalAlignment::optAlign (main.cpp:207-213): "start_ref end_ref start_sgmt end_sgmt\n"
//   -A -r SEQ1 -s SEQ2     RepeatsClassifier::validateRepeats (main.cpp:202-206): one number
// and, because one pair per process leaves a GPU idle,
//   -M|-A --batch LIST     LIST holds one pair per line, "SEQ1<TAB>SEQ2"; one output line per pair, in order; every pair of
//                          the list goes through the same launches.
//   --align -r SEQ1 -s SEQ2   (no TERefiner_1 mode; for tests) LocalAlignment::align: twelve numbers, the best alignment and the
//                          ones before and after it, -1 where the reference leaves its outputs untouched
//   --gpus N               (with --batch) pairs are independent: they are dealt to N GPUs of the box, longest first, one host
//                          thread and one context per GPU, no exchange between them; output order is the list's.
// The modes GAPPadder itself uses (-U, -P: the dedup stage) are `ContigsMerger_b200 --dedup`; the BAM classifiers have no
// alignment in them and are not provided.  No CPU fallback: without an sm_100 device the process exits with an error.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <algorithm>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "gappadder_b200.h"
#include "local_alignment.hpp"

int main(int argc, char** argv)
{
    bool bm = false, ba = false, b_align = false;
    std::string ref, sgmt, batch;
    bool have_r = false, have_s = false;
    int gpu = 0, gpus = 1;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "-M") bm = true;
        else if (a == "-A") ba = true;
        else if (a == "--align") b_align = true;
        else if (a == "-r" && i + 1 < argc) { ref = argv[++i]; have_r = true; }
        else if (a == "-s" && i + 1 < argc) { sgmt = argv[++i]; have_s = true; }
        else if (a == "--batch" && i + 1 < argc) batch = argv[++i];
        else if (a == "--gpu" && i + 1 < argc) gpu = atoi(argv[++i]);
        else if (a == "--gpus" && i + 1 < argc) gpus = atoi(argv[++i]);
        else if (a.size() == 2 && a[0] == '-' && strchr("RCTOPKLSUGBE", a[1])) {
            fprintf(stderr, "TERefiner_b200: mode %s is not the alignment path (the dedup stage -U / -P is ContigsMerger_b200 --dedup)\n", a.c_str());
            return 2;
        } else { fprintf(stderr, "TERefiner_b200: unknown argument %s\n", a.c_str()); return 2; }
    }
    if (b_align) {
        if (!(have_r && have_s)) { fprintf(stderr, "usage: TERefiner_b200 --align -r SEQ1 -s SEQ2\n"); return 2; }
        gp_ctx* ctx = nullptr;
        if (gp_create(gpu, &ctx) != GP_OK) { fprintf(stderr, "TERefiner_b200: %s\n", gp_last_error(nullptr)); return 3; }
        gpm::LocalAlignment la(ctx);
        int v[12];
        for (int& x : v) x = -1;
        const bool ok = la.align(ref, sgmt, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], v[9], v[10], v[11]);
        if (!ok) fprintf(stderr, "TERefiner_b200: %s\n", la.error().c_str());
        else for (int k = 0; k < 12; ++k) printf("%d%c", v[k], k == 11 ? '\n' : ' ');
        gp_destroy(ctx);
        return ok ? 0 : 3;
    }
    if (bm == ba || (batch.empty() && !(have_r && have_s))) {
        fprintf(stderr, "usage: TERefiner_b200 -M|-A -r SEQ1 -s SEQ2   |   TERefiner_b200 -M|-A --batch LIST [--gpus N]\n");
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
    // one GPU: everything in this thread.  Several: deal the pairs out, longest first (cost ~ len1 * len2), and run the same
    // code on every share.
    auto run_share = [bm](int device, const std::vector<gpm::LaPair>& share, std::vector<std::string>& lines, std::string& err) -> bool {
        gp_ctx* ctx = nullptr;
        if (gp_create(device, &ctx) != GP_OK) { err = gp_last_error(nullptr); return false; }
        gpm::LocalAlignment la(ctx);
        bool ok;
        lines.clear();
        if (bm) {
            std::vector<gpm::LaHit> h;
            ok = la.optAlignBatch(share, h);
            for (const gpm::LaHit& x : h)                                                   // main.cpp:212
                lines.push_back(std::to_string(x.start_ref) + " " + std::to_string(x.end_ref) + " " + std::to_string(x.start_sgmt) + " " + std::to_string(x.end_sgmt) + "\n");
        } else {
            std::vector<int> v;
            ok = gpm::validate_repeats_batch(la, share, v);
            for (int x : v) lines.push_back(std::to_string(x) + "\n");                      // RepeatsClassifier.cpp:107-110
        }
        if (!ok) err = la.error();
        gp_destroy(ctx);
        return ok;
    };
    std::string text;
    if (gpus <= 1 || pairs.size() < 2) {
        std::vector<std::string> lines;
        std::string err;
        if (!run_share(gpu, pairs, lines, err)) { fprintf(stderr, "TERefiner_b200: %s\n", err.c_str()); return 3; }      // never partial output
        for (const std::string& l : lines) text += l;
    } else {
        const size_t n = pairs.size();
        std::vector<size_t> order(n);
        std::iota(order.begin(), order.end(), (size_t)0);
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) {
            return (uint64_t)pairs[a].ref.size() * pairs[a].sgmt.size() > (uint64_t)pairs[b].ref.size() * pairs[b].sgmt.size(); });
        std::vector<std::vector<size_t>> ids(gpus);
        std::vector<uint64_t> load(gpus, 0);
        for (size_t k : order) {                                                             // longest processing time first
            const int g = (int)(std::min_element(load.begin(), load.end()) - load.begin());
            ids[g].push_back(k);
            load[g] += (uint64_t)pairs[k].ref.size() * pairs[k].sgmt.size() + 1;
        }
        std::vector<std::vector<std::string>> lines(gpus);
        std::vector<std::string> errs(gpus);
        std::vector<char> oks(gpus, 1);
        std::vector<std::thread> th;
        for (int g = 0; g < gpus; ++g)
            th.emplace_back([&, g] {
                std::vector<gpm::LaPair> share;
                for (size_t k : ids[g]) share.push_back(pairs[k]);
                oks[g] = run_share(g, share, lines[g], errs[g]) ? 1 : 0;
            });
        for (auto& t : th) t.join();
        for (int g = 0; g < gpus; ++g)
            if (!oks[g]) { fprintf(stderr, "TERefiner_b200: GPU %d: %s\n", g, errs[g].c_str()); return 3; }
        std::vector<const std::string*> by_pair(n, nullptr);
        for (int g = 0; g < gpus; ++g)
            for (size_t s = 0; s < ids[g].size(); ++s) by_pair[ids[g][s]] = &lines[g][s];
        for (size_t k = 0; k < n; ++k) text += *by_pair[k];
    }
    fwrite(text.data(), 1, text.size(), stdout);
    return 0;
}
