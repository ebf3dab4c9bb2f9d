// tests/emulate_wf16.cu -- CPU emulation of the packed 16-bit kernel (tests only).
//
// Runs the SAME __host__ __device__ per-lane functions the CUDA kernel uses
// (gappadder_b200/csrc/overlap_wf16.cuh: lane16_begin / lane16_step / lane16_send / lane16_scan,
// the strip schedule and the key bookkeeping), 32 lanes in lock step with the warp shuffles
// replaced by array reads.  `-m "not gpu"` tests compare it with the oracle, so the arithmetic
// (potential, clamping, tags, tie rule) is validated on a machine without a GPU.
#include "../gappadder_b200/csrc/overlap_wf16c.cuh"
#include <cstring>
#include <vector>

using namespace gp;

// instrumentation for tuning the candidate filter (slow-path calls, steps, threshold refreshes)
static long g_slow_calls = 0, g_steps = 0, g_s_changes = 0;
extern "C" void wf16_emulate_counters(long* out)
{
    out[0] = g_slow_calls; out[1] = g_steps; out[2] = g_s_changes;
    g_slow_calls = g_steps = g_s_changes = 0;
}

namespace {

struct HostPair {
    std::vector<uint8_t> row, col;   // 4-bit codes
};

struct HostWarp {
    const HostPair* hp;
    Wf16Pair g;
    std::vector<uint32_t>* bnd;
    int S;
};

template <int K, bool LAST, bool ROWSCAN, bool ADDSEL>
void strip_host(HostWarp& w, const Wf16Params& P, int i0, long long lane_best[32])
{
    const Wf16Pair& g = w.g;
    const HostPair& hp = *w.hp;
    std::vector<uint32_t>& bnd = *w.bnd;
    Lane16<K> st[32];
    uint32_t send[32];
    int mode[32];
    for (int lane = 0; lane < 32; ++lane) {
        const int itop = i0 + lane * 2 * K;
        uint32_t rcode[2 * K];
        for (int x = 0; x < 2 * K; ++x) { int idx = itop + x; rcode[x] = idx < g.m ? hp.row[idx] : 0u; }
        lane16_begin<K>(st[lane], g, itop, rcode);
        mode[lane] = ROWSCAN ? FILTER_ROWS : FILTER_NONE;
        lane16_set_filter<K>(st[lane], g, itop, 1, mode[lane]);
        send[lane] = 0;
    }
    {
        int s = -1000000000;
        for (int lane = 0; lane < 32; ++lane) { const int sc = (int)(lane_best[lane] >> 32); s = sc > s ? sc : s; }
        w.S = s;
    }
    uint32_t thrS[32];
    for (int lane = 0; lane < 32; ++lane) thrS[lane] = filter_thr(w.S);
    const int n = g.n;
    const int t_end = n + 1 + 62;
    g_steps += t_end;
    const int jswitch = n - g.C > 1 ? n - g.C : 1;
    uint16_t* bnd16 = reinterpret_cast<uint16_t*>(bnd.data());

    auto slow_path = [&](int lane, int j) {
        ++g_slow_calls;
        const int itop = i0 + lane * 2 * K;
        lane_best[lane] = lane16_scan<K>(st[lane], g, itop, j, lane_best[lane]);
        const int s = (int)(lane_best[lane] >> 32);
        thrS[lane] = filter_thr(s > w.S ? s : w.S);
    };
    auto generic_step = [&](int t) {
        uint32_t recv[32];
        for (int lane = 0; lane < 32; ++lane) recv[lane] = lane == 0 ? (t <= n + 1 ? bnd[t] : 0u) : send[lane - 1];
        for (int lane = 0; lane < 32; ++lane) {
            const int itop = i0 + lane * 2 * K, j = t - 2 * lane;
            if (j >= 1 && j <= n + 1) {
                lane16_step<K, ADDSEL>(st[lane], recv[lane], P, g.gup, g.gleft);
                if (j == 1) lane16_fix_first<K>(st[lane], g, itop);
                send[lane] = lane16_send<K>(st[lane]);
                if (!LAST && lane == 31 && j >= 2) bnd16[2 * (j - 1)] = (uint16_t)(st[lane].W[K - 1] >> 16);
                if (j == jswitch) { mode[lane] = FILTER_ALL; lane16_set_filter<K>(st[lane], g, itop, j, mode[lane]); }
                if (mode[lane] != FILTER_NONE) {
                    if (filter_fired(lane16_filter<K>(st[lane]), thrS[lane])) slow_path(lane, j);
                }
            }
        }
    };
    const int t_steady0 = 65;
    int t_steady1 = (jswitch > 65 ? jswitch : 65);
    t_steady1 = ((t_steady1 - 1) & ~31) + 1;
    int t = 1;
    for (; t <= t_end && t < t_steady0; ++t) generic_step(t);
    for (; t < t_steady1; t += 32) {
        for (int s = 0; s < 32; ++s) {
            uint32_t recv[32];
            for (int lane = 0; lane < 32; ++lane) recv[lane] = lane == 0 ? bnd[t + s] : send[lane - 1];
            for (int lane = 0; lane < 32; ++lane) {
                lane16_step<K, ADDSEL>(st[lane], recv[lane], P, g.gup, g.gleft);
                send[lane] = lane16_send<K>(st[lane]);
                const int j = t + s - 2 * lane;
                if (!LAST && lane == 31) bnd16[2 * (j - 1)] = (uint16_t)(st[lane].W[K - 1] >> 16);
                if (ROWSCAN) { if (filter_fired(lane16_filter<K>(st[lane]), thrS[lane])) slow_path(lane, j); }
            }
        }
    }
    for (; t <= t_end; ++t) generic_step(t);
}

} // namespace

template <bool ADDSEL>
static void run_pair(const HostPair& hp, const Wf16Pair& g, const Wf16Params& P, std::vector<uint32_t>& bnd, long long lane_best[32])
{
    const int m = g.m;
    HostWarp w{&hp, g, &bnd, 0};
    int i0 = 0;
    while (i0 < m) {
        const Wf16Strip st = wf16_next_strip(i0, m, g.C);
        if (!st.last) {
            if (st.rows == 512) { if (st.rowscan) strip_host<8, false, true, ADDSEL>(w, P, i0, lane_best); else strip_host<8, false, false, ADDSEL>(w, P, i0, lane_best); }
            else                { if (st.rowscan) strip_host<4, false, true, ADDSEL>(w, P, i0, lane_best); else strip_host<4, false, false, ADDSEL>(w, P, i0, lane_best); }
        } else {
            switch (st.rows) {
            case 64:  strip_host<1, true, true, ADDSEL>(w, P, i0, lane_best); break;
            case 128: strip_host<2, true, true, ADDSEL>(w, P, i0, lane_best); break;
            case 256: strip_host<4, true, true, ADDSEL>(w, P, i0, lane_best); break;
            default:  strip_host<8, true, true, ADDSEL>(w, P, i0, lane_best); break;
            }
        }
        i0 += st.rows;
    }
}

// out: score, row_end, col_end, nclip, flags.  Returns 0, or -1 when the pair is outside the
// 16-bit kernel's domain (the library would route it to the 32-bit kernel).  addsel selects the
// <= 4-symbol selector variant (codes must be 0..3).
extern "C" int wf16_emulate(const uint8_t* row_codes, int m, const uint8_t* col_codes, int n,
                            int mismatch, int indel, int max_clip, int addsel, int32_t* out)
{
    if (!wf16_params_ok(mismatch, indel) || !wf16_pair_ok((uint32_t)m, (uint32_t)n)) return -1;
    const Wf16Params P = wf16_make_params(mismatch, indel, max_clip, addsel != 0);
    const Wf16Pair g = wf16_make_pair(m, n, P);
    HostPair hp;
    hp.row.assign(row_codes, row_codes + m);
    hp.col.assign(col_codes, col_codes + n);
    std::vector<uint32_t> bnd((size_t)n + 66, 0u);
    for (int j = 1; j <= n + 1; ++j) {
        uint32_t c = j <= n ? hp.col[j - 1] : 0u;
        bnd[j] = g.v_row0(j <= n ? j : n) | (code11(c, P.addsel) << 16);
    }
    long long lane_best[32];
    for (int l = 0; l < 32; ++l) lane_best[l] = make_key(0, 0u, 1u | (n == 0 ? 2u : 0u));
    if (addsel) run_pair<true>(hp, g, P, bnd, lane_best); else run_pair<false>(hp, g, P, bnd, lane_best);
    long long best = lane_best[0];
    for (int l = 1; l < 32; ++l) best = lane_best[l] > best ? lane_best[l] : best;
    DevResult r;
    store_result(&r, best, m, n, FLAG_KERNEL16);
    out[0] = r.score; out[1] = r.row_end; out[2] = r.col_end; out[3] = r.nclip; out[4] = (int32_t)r.flags;
    return 0;
}

// ---- table kernel (overlap_wf16t.cuh) ---------------------------------------------------------------
// Same block structure as wf16t_strip: 32-step blocks, per-lane range checks, the column-potential
// candidate filter armed at column 1 (row lanes) or at the first tail column, the best score shared
// across the warp at block ends.  The shared-memory table and rings are plain arrays here.
namespace {

template <int K>
void strip_host_t(const HostPair& hp, const Wf16Pair& g, const Wf16tParams& P, std::vector<uint32_t>& bnd, int i0, bool rowscan,
                  bool store_bottom, long long lane_best[32])
{
    const int n = g.n, m = g.m;
    Lane16t<K> st[32];
    std::vector<uint32_t> tab((size_t)32 * 16 * K);
    for (int lane = 0; lane < 32; ++lane) {
        const int itop = i0 + lane * 2 * K;
        uint32_t rc[2 * K];
        for (int x = 0; x < 2 * K; ++x) rc[x] = itop + x < m ? hp.row[itop + x] : 0u;
        for (uint32_t combo = 0; combo < 16; ++combo)
            for (int k = 0; k < K; ++k) tab[((size_t)lane * 16 + combo) * K + k] = wf16t_table_word(rc[k], rc[K + k], combo & 3u, combo >> 2, P);
        lane16t_begin<K>(st[lane], g, itop);
    }
    int S0 = -1000000000;
    for (int lane = 0; lane < 32; ++lane) { const int sc = (int)(lane_best[lane] >> 32); S0 = sc > S0 ? sc : S0; }
    uint32_t thrS[32];
    for (int lane = 0; lane < 32; ++lane) thrS[lane] = filter_thr(S0);
    const int jswitch = n - g.C > 1 ? n - g.C : 1;
    constexpr int D = WF16T_SKEW;
    const int t_end = n + 1 + 31 * D;
    uint16_t* bnd16 = reinterpret_cast<uint16_t*>(bnd.data());
    uint32_t sent1[32], sent2[32];      // W[K-1] of every lane one and two steps ago (the shuffle is issued a step ahead)
    for (int lane = 0; lane < 32; ++lane) sent1[lane] = sent2[lane] = st[lane].W[K - 1];
    std::vector<uint32_t> top(bnd.begin(), bnd.end());      // the ring holds values read before they are overwritten
    for (int tb = 1; tb <= t_end; tb += 32) {
        const bool filt = rowscan || tb + 31 >= jswitch;
        const int cnt = t_end - tb + 1 < 32 ? ((t_end - tb + 2) & ~1) : 32;
        for (int s = 0; s < cnt; ++s) {
            const int t = tb + s;
            uint32_t recv[32];
            for (int lane = 0; lane < 32; ++lane) recv[lane] = lane == 0 ? (top[t <= n + 1 ? t : n + 1] << 16) : sent2[lane - 1];
            for (int lane = 0; lane < 32; ++lane) { sent2[lane] = sent1[lane]; }
            for (int lane = 0; lane < 32; ++lane) {
                const int itop = i0 + lane * 2 * K, j = t - D * lane;
                if (j < 1 || j > n + 1) { sent1[lane] = st[lane].W[K - 1]; continue; }
                const uint32_t combo = (top[j] >> 16) & 15u;
                uint32_t inc[K];
                for (int k = 0; k < K; ++k) inc[k] = tab[((size_t)lane * 16 + combo) * K + k];
                lane16t_step<K>(st[lane], recv[lane], inc, g.gup, g.gleft);
                if (j == 1) lane16t_fix_first<K>(st[lane], g);
                if (store_bottom && lane == 31 && j >= 2 && j - 1 <= n) bnd16[2 * (j - 1)] = (uint16_t)(st[lane].W[K - 1] >> 16);
                if (filt) {
                    const bool rowlane = rowscan && (itop + 2 * K >= m - g.C) && (itop + 1 <= m);
                    const uint32_t acc = p_add2(lane16t_max<K>(st[lane]), wf16t_nthr(g, j));
                    if (filter_fired(acc, j >= (rowlane ? 1 : jswitch) ? thrS[lane] : WF16T_UNARMED)) {
                        ++g_slow_calls;
                        Lane16<K> tmp;
                        for (int k = 0; k < K; ++k) tmp.W[k] = st[lane].W[k];
                        lane_best[lane] = lane16_scan<K>(tmp, g, itop, j, lane_best[lane]);
                        const int sc = (int)(lane_best[lane] >> 32);
                        thrS[lane] = filter_thr(sc > S0 ? sc : S0);
                    }
                }
                sent1[lane] = st[lane].W[K - 1];
            }
        }
        g_steps += cnt;
        if (filt) {
            for (int lane = 0; lane < 32; ++lane) { const int sc = (int)(lane_best[lane] >> 32); S0 = sc > S0 ? sc : S0; }
            for (int lane = 0; lane < 32; ++lane) thrS[lane] = filter_thr(S0);
        }
    }
}

} // namespace

extern "C" int wf16t_emulate(const uint8_t* row_codes, int m, const uint8_t* col_codes, int n,
                             int mismatch, int indel, int max_clip, int32_t* out)
{
    if (!wf16_params_ok(mismatch, indel) || !wf16t_pair_ok((uint32_t)m, (uint32_t)n)) return -1;
    const Wf16tParams P = wf16t_make_params(mismatch, indel, max_clip);
    const Wf16Pair g = wf16t_make_pair(m, n, P);
    HostPair hp;
    hp.row.assign(row_codes, row_codes + m);
    hp.col.assign(col_codes, col_codes + n);
    std::vector<uint32_t> bnd((size_t)n + 66, 0u);
    for (int j = 1; j <= n + 1; ++j) bnd[j] = wf16t_line_word(g, j, j <= n ? hp.col[j - 1] : 0u, j >= 2 ? hp.col[j - 2] : 0u);
    long long lane_best[32];
    for (int l = 0; l < 32; ++l) lane_best[l] = make_key(0, 0u, 1u | (n == 0 ? 2u : 0u));
    int i0 = 0;
    while (i0 < m) {
        const Wf16Strip s = wf16_next_strip(i0, m, g.C);
        switch (s.rows) {
        case 512: strip_host_t<8>(hp, g, P, bnd, i0, s.rowscan, !s.last, lane_best); break;
        case 256: strip_host_t<4>(hp, g, P, bnd, i0, s.rowscan, !s.last, lane_best); break;
        case 128: strip_host_t<2>(hp, g, P, bnd, i0, s.rowscan, !s.last, lane_best); break;
        default:  strip_host_t<1>(hp, g, P, bnd, i0, s.rowscan, !s.last, lane_best); break;
        }
        i0 += s.rows;
    }
    long long best = lane_best[0];
    for (int l = 1; l < 32; ++l) best = lane_best[l] > best ? lane_best[l] : best;
    DevResult r;
    store_result(&r, best, m, n, FLAG_KERNEL16);
    out[0] = r.score; out[1] = r.row_end; out[2] = r.col_end; out[3] = r.nclip; out[4] = (int32_t)r.flags;
    return 0;
}

// ---- certificate kernel (overlap_wf16c.cuh) -------------------------------------------------------------
// Same block structure as wf16c_strip; the pass logic (scan pass in one system, cell pass in the other on
// the sub-table that ends at the best cell) is the kernel's, with the starting system given by the caller.
namespace {

template <int K>
void strip_host_c(const HostPair& hp, const Wf16cPass& g, const Wf16cParams& P, std::vector<uint32_t>& bnd, int i0, bool rowscan,
                  bool store_bottom, long long lane_best[32])
{
    const int n = g.n, m = g.m;
    Lane16c<K> st[32];
    std::vector<uint32_t> tab((size_t)32 * 16 * K);
    for (int lane = 0; lane < 32; ++lane) {
        const int itop = i0 + lane * 2 * K;
        uint32_t rc[2 * K];
        for (int x = 0; x < 2 * K; ++x) rc[x] = itop + x < m ? hp.row[itop + x] : 0u;
        for (uint32_t combo = 0; combo < 16; ++combo)
            for (int k = 0; k < K; ++k) tab[((size_t)lane * 16 + combo) * K + k] = wf16c_table_word(rc[k], rc[K + k], combo & 3u, combo >> 2, P);
        lane16c_begin<K>(st[lane], g, itop, lane * 2 * K);
    }
    int S0 = -1000000000;
    for (int lane = 0; lane < 32; ++lane) { const int sc = (int)(lane_best[lane] >> 32); S0 = sc > S0 ? sc : S0; }
    uint32_t thrS[32];
    for (int lane = 0; lane < 32; ++lane) thrS[lane] = wf16c_filter_thr(S0);
    const int jswitch = n - g.C > 1 ? n - g.C : 1;
    // deferred exact scans, as wf16c_fire_cold / wf16c_flush_cold do them: the pending step of every lane
    uint32_t snapW[32][K];
    int snapj[32], snapA[32];
    for (int lane = 0; lane < 32; ++lane) { snapj[lane] = 0; snapA[lane] = WF16C_NO_SNAP; }
    auto scan_pending = [&](int lane) {
        ++g_slow_calls;
        lane_best[lane] = lane16c_scan_mn<K>(snapW[lane], g.m, g.n, g.C, false, i0 + lane * 2 * K, snapj[lane], S0, g.pot2 ? lane * 2 * K : -1, lane_best[lane], g.tr);
    };
    constexpr int D = WF16C_SKEW;
    const int t_end = n + 1 + 31 * D;
    uint16_t* bnd16 = reinterpret_cast<uint16_t*>(bnd.data());
    uint32_t sent1[32], sent2[32];
    for (int lane = 0; lane < 32; ++lane) sent1[lane] = sent2[lane] = st[lane].W[K - 1];
    std::vector<uint32_t> top(bnd.begin(), bnd.end());
    for (int tb = 1; tb <= t_end; tb += 32) {
        const bool filt = rowscan || tb + 31 >= jswitch;
        const int cnt = t_end - tb + 1 < 32 ? ((t_end - tb + 2) & ~1) : 32;
        for (int s = 0; s < cnt; ++s) {
            const int t = tb + s;
            uint32_t recv[32];
            for (int lane = 0; lane < 32; ++lane) recv[lane] = lane == 0 ? (top[t <= n + 1 ? t : n + 1] << 16) : sent2[lane - 1];
            for (int lane = 0; lane < 32; ++lane) { sent2[lane] = sent1[lane]; }
            bool any_fired = false;
            for (int lane = 0; lane < 32; ++lane) {
                const int itop = i0 + lane * 2 * K, j = t - D * lane;
                if (j < 1 || j > n + 1) { sent1[lane] = st[lane].W[K - 1]; continue; }
                const uint32_t combo = (top[j] >> 16) & 15u;
                uint32_t inc[K];
                for (int k = 0; k < K; ++k) inc[k] = tab[((size_t)lane * 16 + combo) * K + k];
                const int irel_top = lane * 2 * K, pot2 = g.pot2 ? irel_top : -1;
                lane16c_step<K>(st[lane], recv[lane], inc, g.gup, g.gleft, g.pot2);
                if (j == 1) lane16c_fix_first<K>(st[lane], g, itop, irel_top);
                if (store_bottom && lane == 31 && j >= 2 && j - 1 <= n) {
                    int vb = (int)(st[lane].W[K - 1] >> 16);
                    if (g.pot2) { vb -= 4 * 64 * K; vb = vb > 0 ? vb : 0; }       // row 0 of the next strip
                    bnd16[2 * (j - 1)] = (uint16_t)vb;
                }
                if (filt) {
                    const bool rowlane = rowscan && (itop + 2 * K >= m - g.C) && (itop + 1 <= m);
                    const uint32_t acc = p_add2(g.pot2 ? lane16c_max_pot2<K>(st[lane]) : lane16c_max<K>(st[lane]), wf16c_nthr<K>(g, j, irel_top));
                    if (filter_fired(acc, j >= (rowlane ? 1 : jswitch) ? thrS[lane] : WF16C_UNARMED)) {
                        any_fired = true;
                        if (wf16c_deferrable(n, g.C, g.cell, j)) {                      // wf16c_fire_cold
                            const int a = wf16c_exact_step_score<K>(st[lane].W, m, n, g.C, itop, j, pot2);
                            if (!(a == WF16C_NO_SNAP || wf16c_score_of(a) < (S0 > 1 ? S0 : 1))) {
                                if (snapA[lane] != WF16C_NO_SNAP && a <= snapA[lane] && wf16c_score_of(snapA[lane]) >= S0) scan_pending(lane);
                                for (int k = 0; k < K; ++k) snapW[lane][k] = st[lane].W[k];
                                snapj[lane] = j; snapA[lane] = a;
                            }
                        } else {
                            ++g_slow_calls;
                            lane_best[lane] = lane16c_scan_mn<K>(st[lane].W, g.m, g.n, g.C, g.cell, itop, j, S0, pot2, lane_best[lane], g.tr);
                        }
                    }
                }
                sent1[lane] = st[lane].W[K - 1];
            }
            if (any_fired) {                     // the kernel's share_floor(): one REDUX after a step in which a lane fired
                int s0 = -1000000000;
                for (int lane = 0; lane < 32; ++lane) {
                    int sc = (int)(lane_best[lane] >> 32);
                    if (snapA[lane] != WF16C_NO_SNAP && wf16c_score_of(snapA[lane]) > sc) sc = wf16c_score_of(snapA[lane]);
                    s0 = sc > s0 ? sc : s0;
                }
                S0 = s0;
                for (int lane = 0; lane < 32; ++lane) thrS[lane] = wf16c_filter_thr(S0);
            }
        }
        g_steps += cnt;
    }
    for (int lane = 0; lane < 32; ++lane)        // end of the strip: wf16c_flush_cold
        if (snapA[lane] != WF16C_NO_SNAP && wf16c_score_of(snapA[lane]) >= S0) scan_pending(lane);
}

long long pass_host_c(const HostPair& hp, const Wf16cPass& g, const Wf16cParams& P)
{
    const int m = g.m, n = g.n;
    std::vector<uint32_t> bnd((size_t)n + 66, 0u);
    for (int j = 1; j <= n + 1; ++j) bnd[j] = wf16c_line_word(g, j, j <= n ? hp.col[j - 1] : 0u, j >= 2 ? hp.col[j - 2] : 0u);
    long long lane_best[32];
    for (int l = 0; l < 32; ++l) lane_best[l] = wf16c_initial_best(g);
    int i0 = 0;
    while (i0 < m) {
        const Wf16Strip s = wf16c_next_strip(i0, m, g.C);
        const bool rs = s.rowscan && !g.cell;
        switch (s.rows) {
        case 512: strip_host_c<8>(hp, g, P, bnd, i0, rs, !s.last, lane_best); break;
        case 256: strip_host_c<4>(hp, g, P, bnd, i0, rs, !s.last, lane_best); break;
        case 128: strip_host_c<2>(hp, g, P, bnd, i0, rs, !s.last, lane_best); break;
        default:  strip_host_c<1>(hp, g, P, bnd, i0, rs, !s.last, lane_best); break;
        }
        i0 += s.rows;
    }
    long long best = lane_best[0];
    for (int l = 1; l < 32; ++l) best = lane_best[l] > best ? lane_best[l] : best;
    return best;
}

} // namespace

// out: score, row_end, col_end, nclip, flags.  Returns 0 when the first pass certified the origin, 1 when
// a later (sub-table) pass did, 2 when none did (origin flags in out[4] are then 0: the library recomputes the
// pair with an exact kernel), -1 outside the kernel's domain.  first_sys: 0 = U, 1 = L, 2 = C.
extern "C" int wf16c_emulate(const uint8_t* row_codes, int m, const uint8_t* col_codes, int n,
                             int mismatch, int indel, int max_clip, int first_sys, int32_t* out)
{
    if (!wf16_params_ok(mismatch, indel)) return -1;
    Wf16cParams P = wf16c_make_params(mismatch, indel, max_clip);
    bool tr = false;
    if (first_sys >= 8) { tr = true; first_sys -= 8; }  // bit 3: transposed (the computed table's rows are the column sequence)
    const int cm = tr ? n : m, cn = tr ? m : n;         // the computed table
    if (!wf16c_pair_ok((uint32_t)cm, (uint32_t)cn)) return -1;
    if (first_sys >= 4) {                              // bit 2: the free-moves layout (standard scores, short columns)
        if (!P.std_scores || (uint32_t)cn > WF16C_POT2_MAX_N) return -1;
        P.pot2 = 1;
        first_sys -= 4;
    }
    HostPair hp;
    if (tr) { hp.row.assign(col_codes, col_codes + n); hp.col.assign(row_codes, row_codes + m); }
    else { hp.row.assign(row_codes, row_codes + m); hp.col.assign(col_codes, col_codes + n); }
    if (first_sys < 0 || first_sys > 2) return -1;
    Wf16cPass g = wf16c_make_pass(cm, cn, P, first_sys, false, 0, tr);
    const long long key = pass_host_c(hp, g, P);
    uint32_t origin = wf16c_certified_origin(g, key);
    DevResult r;
    store_result(&r, key & ~3ll, m, n, FLAG_KERNEL16);
    int status = 0;
    for (int attempt = 1; attempt <= 2 && origin == 0u; ++attempt) {
        status = 1;
        g = wf16c_make_pass(tr ? r.col_end : r.row_end, tr ? r.row_end : r.col_end, P, wf16c_next_system(first_sys, attempt), true, r.score, tr);
        origin = wf16c_certified_origin(g, pass_host_c(hp, g, P));
    }
    if (origin == 0u) status = 2;
    store_result(&r, (key & ~3ll) | (long long)origin, m, n, FLAG_KERNEL16);
    out[0] = r.score; out[1] = r.row_end; out[2] = r.col_end; out[3] = r.nclip; out[4] = (int32_t)r.flags;
    return status;
}
