// tests/shim_gp_oracle.cpp -- TEST INFRASTRUCTURE ONLY.
// Serves the DP entry points of include/gappadder_b200.h from the CPU oracle
// (oracle/overlap_oracle.c) so that the HOST logic of the drop-in binary (FASTA I/O, candidate
// filter, "Arrange error" rule, overlap graph, path search, merge chains, output formatting) can be
// checked against the reference's golden outputs on a machine without a GPU.  Linked only into
// build/ContigsMerger_hosttest by tests/test_contigsmerger_host.py; the product binary links
// libgappadder_b200.so and has no such path.
#include <cstdint>
#include <cstring>
#include <string>

#include "gappadder_b200.h"

extern "C" {
typedef struct { int32_t score, row_end, col_end, nclip, tb_row, tb_col, bcontained; } gpo_dp_result;
int gpo_evaluate(const char* s1, int m, const char* s2, int n, int mismatch, int indel, int maxclip, gpo_dp_result* r);
}

struct gp_ctx { std::string err; };

extern "C" {
int gp_create(int, gp_ctx** out) { *out = new gp_ctx(); return GP_OK; }
void gp_destroy(gp_ctx* c) { delete c; }
const char* gp_last_error(const gp_ctx* c) { return c ? c->err.c_str() : ""; }
int gp_overlap_batch(gp_ctx*, const char* const* seqs, const uint32_t* seq_len, uint32_t, const gp_pair* pairs,
                     uint64_t n_pairs, const gp_dp_params* p, gp_result* out)
{
    for (uint64_t k = 0; k < n_pairs; ++k) {
        gpo_dp_result r;
        const uint32_t a = pairs[k].row_seq, b = pairs[k].col_seq;
        if (gpo_evaluate(seqs[a], (int)seq_len[a], seqs[b], (int)seq_len[b], p->mismatch, p->indel, p->max_clip, &r) != 0) return GP_ERR_NOMEM;
        out[k].score = r.score; out[k].row_end = r.row_end; out[k].col_end = r.col_end; out[k].nclip = r.nclip;
        out[k].flags = (r.tb_row == 0 ? GP_FLAG_ROW0 : 0u) | (r.tb_col == 0 ? GP_FLAG_COL0 : 0u) | (r.bcontained ? GP_FLAG_CONTAINED : 0u);
    }
    return GP_OK;
}
// statistics the merger reads after a batch: nothing to report from the CPU shim
int gp_closed_form_stats(const gp_ctx*, uint64_t* pairs, uint64_t* cells) { if (pairs) *pairs = 0; if (cells) *cells = 0; return GP_OK; }
int gp_cert_stats(const gp_ctx*, uint64_t* a, uint64_t* b, uint64_t* c) { if (a) *a = 0; if (b) *b = 0; if (c) *c = 0; return GP_OK; }
int gp_last_team(const gp_ctx*) { return 0; }
int gp_last_timing(const gp_ctx*, double* out_ms, int n) { for (int i = 0; i < n; ++i) out_ms[i] = 0.0; return GP_OK; }
}
