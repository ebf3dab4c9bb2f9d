// tests/emulate_affine.cu -- CPU emulation of the affine local aligner (tests only).
//
// Runs the SAME __host__ __device__ functions the CUDA kernels use (gappadder_b200/csrc/affine_local.cuh):
// aff_lane_begin / aff_lane_step with 32 lanes in lock step (the warp shuffles and the shared-memory ring replaced by
// array reads, the strip schedule and the boundary line as in affine_forward_kernel), and aff_epilogue as it is.
// `-m "not gpu"` tests compare both with the reference's own aligner (oracle/_ref/libla_ref.so) and the golden vectors.
#include "../gappadder_b200/csrc/affine_local.cuh"
#include <cstring>
#include <vector>

using namespace gp;

namespace {

std::vector<uint32_t> pack_codes(const uint8_t* c, int len)
{
    std::vector<uint32_t> w((size_t)(len + 7) / 8 + 4, 0u);
    for (int p = 0; p < len; ++p) w[p >> 3] |= (uint32_t)(c[p] & 15u) << ((p & 7) * 4);
    return w;
}

} // namespace

extern "C" {

// params = {match, mismatch, nscore, open, ext, band}; out = {score, end1, end2}
int aff_emulate_forward(const uint8_t* c1, int m, const uint8_t* c2, int n, const int* params, int32_t* out)
{
    const AffParams P{params[0], params[1], params[2], params[3], params[4], params[5]};
    if (!aff_params_ok(P) || m <= 0 || n <= 0 || !aff_pair_ok((uint32_t)m, (uint32_t)n, P)) return 1;
    constexpr int R = AF_R, D = AF_SKEW;
    const uint32_t row0 = aff_pack(0, -(P.q + P.r));
    const int n_strips = (m + AF_STRIP - 1) / AF_STRIP;
    const int pad = n_strips * AF_STRIP - m;
    std::vector<uint32_t> bnd((size_t)n + 1, row0);
    AffLane<R> st[32];
    for (int l = 0; l < 32; ++l) { st[l].best = aff_key(0, 0, 0); st[l].bscore = 0; }
    for (int s = 0; s < n_strips; ++s) {
        const bool last = s == n_strips - 1;
        uint32_t bottom[32], mysym[32], recv_next[32], sym_next[32];
        for (int l = 0; l < 32; ++l) { aff_lane_begin<R>(st[l], P.q, P.r); bottom[l] = row0; mysym[l] = 0; recv_next[l] = row0; sym_next[l] = 0; }
        const int steps = n + D * 31;
        for (int t = 1; t <= steps; ++t) {
            uint32_t recv[32], csym[32], shb[32], shs[32];
            for (int l = 0; l < 32; ++l) { shb[l] = l ? bottom[l - 1] : bottom[0]; shs[l] = l ? mysym[l - 1] : mysym[0]; }
            for (int l = 0; l < 32; ++l) {
                recv[l] = recv_next[l]; csym[l] = sym_next[l];
                recv_next[l] = shb[l]; sym_next[l] = shs[l];
            }
            if (t <= n) { recv[0] = bnd[t]; csym[0] = c2[t - 1] > 4 ? 4u : c2[t - 1]; } else { recv[0] = row0; csym[0] = 0; }
            std::vector<uint32_t> wr;          // lane 31's boundary writes of this step (none is read in this strip)
            for (int l = 0; l < 32; ++l) {
                const int itop = s * AF_STRIP + l * R - pad;
                const int j = t - D * l;
                if (j >= 1 && j <= n) {
                    int inc[R];
                    for (int x = 0; x < R; ++x) inc[x] = itop + x >= 0 ? aff_sc(c1[itop + x], csym[l], P) : AF_NEG;
                    bottom[l] = aff_lane_step<R>(st[l], recv[l], inc, P.q, P.r, j, itop);
                    mysym[l] = csym[l];
                    if (l == 31 && !last) bnd[j] = bottom[l];
                }
            }
        }
    }
    long long k = st[0].best;
    for (int l = 1; l < 32; ++l) k = st[l].best > k ? st[l].best : k;
    out[0] = (int32_t)(k >> 40);
    out[1] = (int32_t)(AFF_MAX_LEN - (uint32_t)(k & 0xfffffll));
    out[2] = (int32_t)(AFF_MAX_LEN - (uint32_t)((k >> 20) & 0xfffffll));
    if (out[0] <= 0) out[1] = out[2] = 0;
    return 0;
}

// Passes 2 and 3 on the host: out = {score, start1, end1, start2, end2, flags}
int aff_host_epilogue(const uint8_t* c1, int m, const uint8_t* c2, int n, const int* params, int score, int end1, int end2, int32_t* out)
{
    const AffParams P{params[0], params[1], params[2], params[3], params[4], params[5]};
    if (!aff_params_ok(P) || score <= 0 || end1 < 1 || end1 > m || end2 < 1 || end2 > n) return 1;
    const std::vector<uint32_t> w1 = pack_codes(c1, m), w2 = pack_codes(c2, n);
    std::vector<int> work(aff_epilogue_words(end1));
    DevLocal res{};
    aff_epilogue(AffSeq{w1.data()}, AffSeq{w2.data()}, P, score, end1, end2, work.data(), &res);
    out[0] = res.score; out[1] = res.start1; out[2] = res.end1; out[3] = res.start2; out[4] = res.end2; out[5] = (int32_t)res.flags;
    return 0;
}

// The same two passes as ONE WARP computes them (aff_epilogue_warp; on the host every phase loops over the 32 lanes).
int aff_host_epilogue_warp(const uint8_t* c1, int m, const uint8_t* c2, int n, const int* params, int score, int end1, int end2, int32_t* out)
{
    const AffParams P{params[0], params[1], params[2], params[3], params[4], params[5]};
    if (!aff_params_ok(P) || score <= 0 || end1 < 1 || end1 > m || end2 < 1 || end2 > n) return 1;
    const std::vector<uint32_t> w1 = pack_codes(c1, m), w2 = pack_codes(c2, n);
    std::vector<int> work(aff_epilogue_words(end1), 0x5a5a5a5a);
    AffWarp area;
    memset(&area, 0x5a, sizeof area);
    DevLocal res{};
    aff_epilogue_warp(AffSeq{w1.data()}, AffSeq{w2.data()}, P, score, end1, end2, work.data(), &area, &res);
    out[0] = res.score; out[1] = res.start1; out[2] = res.end1; out[3] = res.start2; out[4] = res.end2; out[5] = (int32_t)res.flags;
    return 0;
}

} // extern "C"
