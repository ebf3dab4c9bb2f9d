// tests/emulate_wf16.cu -- CPU emulation of the packed 16-bit kernel (tests only).
//
// Runs the SAME __host__ __device__ per-lane functions the CUDA kernel uses
// (gappadder_b200/csrc/overlap_wf16.cuh: lane16_begin / lane16_step / lane16_send / lane16_scan,
// the strip schedule and the key bookkeeping), 32 lanes in lock step with the warp shuffles
// replaced by array reads.  `-m "not gpu"` tests compare it with the oracle, so the arithmetic
// (potential, clamping, tags, tie rule) is validated on a machine without a GPU.
#include "../gappadder_b200/csrc/overlap_wf16.cuh"
#include <cstring>
#include <vector>

using namespace gp;

namespace {

struct HostPair {
    std::vector<uint8_t> row, col;   // 4-bit codes
};

template <int K, bool SCAN_ALL>
long long strip_host(const HostPair& hp, const Wf16Pair& g, const Wf16Params& P, std::vector<uint32_t>& bnd,
                     int i0, bool store_bottom, long long lane_best[32])
{
    Lane16<K> st[32];
    uint32_t send[32];
    for (int lane = 0; lane < 32; ++lane) {
        uint32_t rcode[2 * K];
        for (int x = 0; x < 2 * K; ++x) { int idx = i0 + lane * 2 * K + x; rcode[x] = idx < g.m ? hp.row[idx] : 15u; }
        lane16_begin<K>(st[lane], g, i0 + lane * 2 * K, rcode);
        send[lane] = 0;
    }
    const int n = g.n;
    const int t_end = n + 1 + 62;
    const int t_scan = SCAN_ALL ? 1 : (n - g.C > 1 ? n - g.C : 1);
    uint16_t* bnd16 = reinterpret_cast<uint16_t*>(bnd.data());
    for (int t = 1; t <= t_end; ++t) {
        uint32_t recv[32];
        for (int lane = 0; lane < 32; ++lane)
            recv[lane] = lane == 0 ? (t <= n + 1 ? bnd[t] : 0u) : send[lane - 1];
        for (int lane = 0; lane < 32; ++lane) {
            const int itop = i0 + lane * 2 * K;
            const int j = t - 2 * lane;
            if (j >= 1 && j <= n + 1) {
                lane16_step<K>(st[lane], recv[lane], P, g.gup, g.gleft);
                if (j == 1) lane16_fix_first<K>(st[lane], g, itop);
                send[lane] = lane16_send<K>(st[lane]);
                if (store_bottom && lane == 31 && j >= 2) bnd16[2 * (j - 1)] = (uint16_t)(st[lane].W[K - 1] >> 16);
                if (t >= t_scan) lane_best[lane] = lane16_scan<K>(st[lane], g, itop, j, lane_best[lane]);
            }
        }
    }
    return 0;
}

} // namespace

// out: score, row_end, col_end, nclip, flags.  Returns 0, or -1 when the pair is outside the
// 16-bit kernel's domain (the library would route it to the 32-bit kernel).
extern "C" int wf16_emulate(const uint8_t* row_codes, int m, const uint8_t* col_codes, int n,
                            int mismatch, int indel, int max_clip, int32_t* out)
{
    if (!wf16_params_ok(mismatch, indel) || !wf16_pair_ok((uint32_t)m, (uint32_t)n)) return -1;
    const Wf16Params P = wf16_make_params(mismatch, indel, max_clip);
    const Wf16Pair g = wf16_make_pair(m, n, P);
    HostPair hp;
    hp.row.assign(row_codes, row_codes + m);
    hp.col.assign(col_codes, col_codes + n);
    std::vector<uint32_t> bnd((size_t)n + 66, 0u);
    for (int j = 1; j <= n + 1; ++j) {
        uint32_t c = j <= n ? hp.col[j - 1] : 0u;
        bnd[j] = g.v_row0(j <= n ? j : n) | (code11(c) << 16);
    }
    long long lane_best[32];
    for (int l = 0; l < 32; ++l) lane_best[l] = make_key(0, 0u, 1u | (n == 0 ? 2u : 0u));
    const int m_fast = wf16_fast_rows(m, g.C);
    int i0 = 0;
    while (m_fast - i0 >= 512) { strip_host<8, false>(hp, g, P, bnd, i0, true, lane_best); i0 += 512; }
    while (m_fast - i0 >= 128) { strip_host<2, false>(hp, g, P, bnd, i0, true, lane_best); i0 += 128; }
    while (m_fast - i0 >= 64)  { strip_host<1, false>(hp, g, P, bnd, i0, true, lane_best); i0 += 64; }
    while (i0 < m)             { strip_host<1, true>(hp, g, P, bnd, i0, i0 + 64 < m, lane_best); i0 += 64; }
    long long best = lane_best[0];
    for (int l = 1; l < 32; ++l) best = lane_best[l] > best ? lane_best[l] : best;
    DevResult r;
    store_result(&r, best, m, n, FLAG_KERNEL16);
    out[0] = r.score; out[1] = r.row_end; out[2] = r.col_end; out[3] = r.nclip; out[4] = (int32_t)r.flags;
    return 0;
}
