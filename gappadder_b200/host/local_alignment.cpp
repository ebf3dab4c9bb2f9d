// local_alignment.cpp -- see local_alignment.hpp.
#include "local_alignment.hpp"

namespace gpm {

static LaHit to_hit(const gp_local_result& r)
{
    LaHit h;
    h.aligned = !(r.flags & (GP_LOCAL_NO_MATCH | GP_LOCAL_UNDEFINED));
    h.score = r.score;
    if (h.aligned) { h.start_ref = r.start1; h.end_ref = r.end1; h.start_sgmt = r.start2; h.end_sgmt = r.end2; }
    return h;
}

bool LocalAlignment::optAlignBatch(const std::vector<LaPair>& in, std::vector<LaHit>& out)
{
    out.assign(in.size(), LaHit());
    if (in.empty()) return true;
    std::vector<const char*> seqs(2 * in.size());
    std::vector<uint32_t> lens(2 * in.size());
    std::vector<gp_pair> pairs(in.size());
    for (size_t k = 0; k < in.size(); ++k) {
        seqs[2 * k] = in[k].ref.data(); lens[2 * k] = (uint32_t)in[k].ref.size();
        seqs[2 * k + 1] = in[k].sgmt.data(); lens[2 * k + 1] = (uint32_t)in[k].sgmt.size();
        pairs[k].row_seq = (uint32_t)(2 * k); pairs[k].col_seq = (uint32_t)(2 * k + 1);
    }
    std::vector<gp_local_result> res(in.size());
    const int rc = gp_local_affine_batch(ctx_, seqs.data(), lens.data(), (uint32_t)seqs.size(), pairs.data(), pairs.size(), nullptr, res.data());
    if (rc != GP_OK) { error_ = gp_last_error(ctx_); return false; }
    for (size_t k = 0; k < in.size(); ++k) out[k] = to_hit(res[k]);
    return true;
}

// local_alignment.cpp:1053-1090: the best alignment, then one in what lies before it on both sequences and one in what lies
// after it (each only if both sides are non-empty; coordinates relative to those substrings).
bool LocalAlignment::alignBatch(const std::vector<LaPair>& in, std::vector<LaAlign>& out)
{
    out.assign(in.size(), LaAlign());
    std::vector<LaHit> opt;
    if (!optAlignBatch(in, opt)) return false;
    std::vector<LaPair> sub;
    std::vector<size_t> owner;
    std::vector<int> side;
    for (size_t k = 0; k < in.size(); ++k) {
        out[k].opt = opt[k];
        const LaHit& o = opt[k];
        if (!o.aligned) continue;
        if (o.start_ref > 1 && o.start_sgmt > 1) {                                                        // :1067
            sub.push_back({in[k].ref.substr(0, o.start_ref - 1), in[k].sgmt.substr(0, o.start_sgmt - 1)});
            owner.push_back(k); side.push_back(0);
        }
        if (o.end_ref < (int)in[k].ref.size() && o.end_sgmt < (int)in[k].sgmt.size()) {                   // :1078
            sub.push_back({in[k].ref.substr(o.end_ref), in[k].sgmt.substr(o.end_sgmt)});
            owner.push_back(k); side.push_back(1);
        }
    }
    std::vector<LaHit> hits;
    if (!optAlignBatch(sub, hits)) return false;
    for (size_t s = 0; s < sub.size(); ++s) {
        LaAlign& a = out[owner[s]];
        if (side[s] == 0) { a.left = hits[s]; a.has_left = true; } else { a.right = hits[s]; a.has_right = true; }
    }
    return true;
}

// local_alignment.cpp:1097-1150: the best alignment, then one more on prefix + suffix of each sequence around it.
bool LocalAlignment::optAlignWithRestSecondOptBatch(const std::vector<LaPair>& in, std::vector<LaRest>& out)
{
    out.assign(in.size(), LaRest());
    std::vector<LaPair> first;
    std::vector<size_t> idx;
    for (size_t k = 0; k < in.size(); ++k)
        if (!in[k].ref.empty() && !in[k].sgmt.empty()) { first.push_back(in[k]); idx.push_back(k); }      // :1103-1109: 0 0 0 0 otherwise
    std::vector<LaHit> opt;
    if (!optAlignBatch(first, opt)) return false;
    for (size_t s = 0; s < first.size(); ++s) out[idx[s]].opt = opt[s];
    std::vector<LaPair> sub;
    std::vector<size_t> owner;
    for (size_t k = 0; k < in.size(); ++k) {
        const LaHit& o = out[k].opt;
        const std::string &ref = in[k].ref, &sg = in[k].sgmt;
        std::string lr, ls, rr, rs;
        if (o.start_ref >= 1 && o.start_sgmt >= 1) { lr = ref.substr(0, o.start_ref - 1); ls = sg.substr(0, o.start_sgmt - 1); }     // :1120-1125
        if (o.end_ref <= (int)ref.size() && o.end_sgmt <= (int)sg.size()) { rr = ref.substr(o.end_ref); rs = sg.substr(o.end_sgmt); }   // :1127-1131
        LaPair p{lr + rr, ls + rs};
        if (p.ref.empty() || p.sgmt.empty()) continue;                                                    // :1136-1142: 0 0 0 0
        sub.push_back(std::move(p));
        owner.push_back(k);
    }
    std::vector<LaHit> rest;
    if (!optAlignBatch(sub, rest)) return false;
    for (size_t s = 0; s < sub.size(); ++s) out[owner[s]].rest = rest[s];
    return true;
}

bool LocalAlignment::optAlign(const std::string& sref, const std::string& ssgmt, int& optm_start_ref, int& optm_end_ref, int& optm_start_sgmt, int& optm_end_sgmt)
{
    std::vector<LaHit> h;
    if (!optAlignBatch({{sref, ssgmt}}, h)) return false;
    optm_start_ref = h[0].start_ref; optm_end_ref = h[0].end_ref; optm_start_sgmt = h[0].start_sgmt; optm_end_sgmt = h[0].end_sgmt;
    return true;
}

bool LocalAlignment::align(const std::string& sref, const std::string& ssgmt, int& optm_start_ref, int& optm_end_ref, int& optm_start_sgmt, int& optm_end_sgmt,
                           int& start_ref1, int& end_ref1, int& start_sgmt1, int& end_sgmt1,
                           int& start_ref2, int& end_ref2, int& start_sgmt2, int& end_sgmt2)
{
    std::vector<LaAlign> a;
    if (!alignBatch({{sref, ssgmt}}, a)) return false;
    const LaAlign& r = a[0];
    optm_start_ref = r.opt.start_ref; optm_end_ref = r.opt.end_ref; optm_start_sgmt = r.opt.start_sgmt; optm_end_sgmt = r.opt.end_sgmt;
    if (r.has_left) { start_ref1 = r.left.start_ref; end_ref1 = r.left.end_ref; start_sgmt1 = r.left.start_sgmt; end_sgmt1 = r.left.end_sgmt; }       // untouched otherwise, as there
    if (r.has_right) { start_ref2 = r.right.start_ref; end_ref2 = r.right.end_ref; start_sgmt2 = r.right.start_sgmt; end_sgmt2 = r.right.end_sgmt; }
    return true;
}

bool LocalAlignment::optAlignWithRestSecondOpt(const std::string& sref, const std::string& ssgmt, int& optm_start_ref, int& optm_end_ref, int& optm_start_sgmt,
                                               int& optm_end_sgmt, int& start_ref1, int& end_ref1, int& start_sgmt1, int& end_sgmt1)
{
    std::vector<LaRest> a;
    if (!optAlignWithRestSecondOptBatch({{sref, ssgmt}}, a)) return false;
    const LaRest& r = a[0];
    optm_start_ref = r.opt.start_ref; optm_end_ref = r.opt.end_ref; optm_start_sgmt = r.opt.start_sgmt; optm_end_sgmt = r.opt.end_sgmt;
    start_ref1 = r.rest.start_ref; end_ref1 = r.rest.end_ref; start_sgmt1 = r.rest.start_sgmt; end_sgmt1 = r.rest.end_sgmt;
    return true;
}

static char supplementary(char c)            // StrOperation::getSupplementary, StrOperation.cpp:16-28
{
    switch (c) {
    case 'A': case 'a': return 'T';
    case 'T': case 't': return 'A';
    case 'C': case 'c': return 'G';
    case 'G': case 'g': return 'C';
    default: return 'N';
    }
}

bool validate_repeats_batch(LocalAlignment& la, const std::vector<LaPair>& in, std::vector<int>& out)
{
    std::vector<LaPair> both(2 * in.size());
    for (size_t k = 0; k < in.size(); ++k) {
        both[2 * k] = in[k];
        both[2 * k + 1].ref = in[k].ref;
        std::string& rs = both[2 * k + 1].sgmt;
        rs.resize(in[k].sgmt.size());
        for (size_t i = 0; i < rs.size(); ++i) rs[i] = supplementary(in[k].sgmt[rs.size() - 1 - i]);
    }
    std::vector<LaRest> r;
    if (!la.optAlignWithRestSecondOptBatch(both, r)) return false;
    auto span = [](const LaHit& h) { return h.end_ref > 0 ? h.end_ref - h.start_ref + 1 : 0; };            // RepeatsClassifier.cpp:59-69
    out.resize(in.size());
    for (size_t k = 0; k < in.size(); ++k) {
        const int iff = span(r[2 * k].opt) + span(r[2 * k].rest), ifr = span(r[2 * k + 1].opt) + span(r[2 * k + 1].rest);
        out[k] = iff > ifr ? iff : ifr;                                                                   // :106-111
    }
    return true;
}

} // namespace gpm
