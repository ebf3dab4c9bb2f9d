// tests/shim_gp_affine_oracle.cpp -- TEST INFRASTRUCTURE ONLY.
// Serves gp_local_affine_batch from the CPU oracle (oracle/local_affine_oracle.c) so that the HOST logic of
// TERefiner_b200 (command line, batch list, LocalAlignment's rest alignments, validateRepeats, output text) can be checked
// against the prebuilt TERefiner_1's golden outputs on a machine without a GPU.  Linked only into
// build/TERefiner_hosttest by tests/test_terefiner_host.py; the product binary links libgappadder_b200.so.
#include <cstdint>
#include <string>

#include "gappadder_b200.h"

extern "C" int lao_local_affine(const char* s1, int len1, const char* s2, int len2, int32_t* out);

struct gp_ctx { std::string err; };

extern "C" {
int gp_create(int, gp_ctx** out) { *out = new gp_ctx(); return GP_OK; }
void gp_destroy(gp_ctx* c) { delete c; }
const char* gp_last_error(const gp_ctx* c) { return c ? c->err.c_str() : ""; }
int gp_local_affine_batch(gp_ctx* c, const char* const* seqs, const uint32_t* seq_len, uint32_t, const gp_pair* pairs, uint64_t n_pairs,
                          const gp_affine_params* params, gp_local_result* out)
{
    if (params) { c->err = "the oracle shim knows aln_param_blast only"; return GP_ERR_INVALID; }
    for (uint64_t k = 0; k < n_pairs; ++k) {
        const uint32_t a = pairs[k].row_seq, b = pairs[k].col_seq;
        int32_t o[5];
        const int rc = lao_local_affine(seqs[a], (int)seq_len[a], seqs[b], (int)seq_len[b], o);
        out[k].score = o[0]; out[k].start1 = o[1]; out[k].end1 = o[2]; out[k].start2 = o[3]; out[k].end2 = o[4];
        out[k].flags = rc == 0 ? 0u : rc == 1 ? GP_LOCAL_NO_MATCH : GP_LOCAL_UNDEFINED;
        if (rc != 0) out[k].start1 = out[k].end1 = out[k].start2 = out[k].end2 = 0;
    }
    return GP_OK;
}
}
