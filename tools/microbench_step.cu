// tools/microbench_step.cu -- issue-rate microbenchmark of whole wavefront STEPS (sm_100a).
//
// microbench_int.cu measures single instructions and pairs; this one measures candidate inner loops
// of the packed 16-bit overlap kernel with their real dependency structure (the `up` chain through
// the K registers of a lane, one shuffle per step, the substitution lookup) at the occupancies the
// kernel can have, and reports clocks per register-step per SM sub-partition and the cell rate the
// loop alone would sustain.  It is how the formulation in csrc/overlap_wf16.cuh was chosen
// (DESIGN.md "Choice of the inner step").
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o build/microbench_step tools/microbench_step.cu
// run  : build/microbench_step [json-out]
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(s));
    return d;
}

constexpr uint32_t ZCLR = 0xfffbfffbu;

enum Variant {
    V_PRMT = 0,      // today's kernel: VIADD sel, PRMT lookup, VIADD d, VIADDMNMX, VIADDMNMX.RELU, LOP3
    V_LDS,           // shared-memory increment table (LDS.128), VIADD d, VIADDMNMX, VIADDMNMX.RELU, LOP3
    V_LDS_NOZ,       // same without the tag clear (hypothetical lower bound)
    V_LDS_MAX3,      // LDS, 3x VIADD, VIMNMX3.RELU, LOP3
    V_LDS_SPLIT,     // LDS, VIADD d, VIADDMNMX(l,d), VIADD u, VIMNMX.RELU, LOP3
    V_LDS_ALLSPLIT,  // LDS, 3x VIADD, 2x VIMNMX, LOP3
    V_COUNT
};
static const char* VNAME[] = {"prmt-lookup (current)", "lds-table", "lds-table no-tag-clear", "lds-table max3", "lds-table split-up", "lds-table all-split"};

template <int K, int V>
__global__ void __launch_bounds__(128) step_kernel(uint32_t* out, const uint32_t* __restrict__ line, int steps, uint32_t gup, uint32_t gleft,
                                                   uint32_t tbl_lo, uint32_t tbl_hi)
{
    extern __shared__ uint4 smem4[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // per warp: 16 combos x (K/4) uint4 x 32 lanes
    uint4* tbl = smem4 + (size_t)warp * 16 * (K / 4) * 32;
    if (V != V_PRMT) {
        for (int c = 0; c < 16; ++c)
            for (int q = 0; q < K / 4; ++q) {
                uint4 v;
                v.x = ((c * 7 + q * 3 + lane) & 1) ? 0x00040004u : 0xffecffecu;
                v.y = ((c * 5 + q + lane) & 2) ? 0x00040004u : 0xffec0004u;
                v.z = ((c + q * 3 + lane) & 1) ? 0x0004ffecu : 0xffecffecu;
                v.w = ((c * 3 + q + lane) & 2) ? 0x00040004u : 0xffecffecu;
                tbl[(c * (K / 4) + q) * 32 + lane] = v;
            }
    }
    __syncwarp();
    uint32_t W[K], Rk[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { W[k] = 0x01000100u + lane * 8 + k * 16; Rk[k] = ((lane + k) & 3u) * 0x11u * 0x0101u | 0x80808080u; }
    uint32_t send = 0x00000100u, cvec = 0, up0_prev = 0x01000100u;
    const uint32_t tbl_base = (uint32_t)__cvta_generic_to_shared(tbl) + lane * 16;
    uint32_t chunk = 0;
    for (int tb = 0; tb < steps; tb += 32) {
        chunk = line[(tb + lane) & 1023];
#pragma unroll 1
        for (int s = 0; s < 32; ++s) {
            const uint32_t from_line = __shfl_sync(0xffffffffu, chunk, s);
            uint32_t recv = __shfl_up_sync(0xffffffffu, send, 1);
            if (lane == 0) recv = from_line;
            uint32_t inc[K];
            if (V == V_PRMT) {
                cvec = prmt(recv, cvec, 0x6542u);
            } else {
                // combo = (code(j), code(j-1)): 2 + 2 bits kept in cvec bits 0-3, scaled to the table stride
                cvec = ((cvec << 2) | ((recv >> 16) & 3u)) & 15u;
                const uint32_t addr = tbl_base + cvec * (K / 4) * 512;
#pragma unroll
                for (int q = 0; q < K / 4; ++q)
                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(inc[4 * q]), "=r"(inc[4 * q + 1]), "=r"(inc[4 * q + 2]), "=r"(inc[4 * q + 3]) : "r"(addr + q * 512));
            }
            const uint32_t up0 = prmt(recv, W[K - 1], 0x5410u);
            uint32_t diag = up0_prev;
            up0_prev = up0;
            uint32_t up = up0;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const uint32_t left = W[k];
                uint32_t w;
                if (V == V_PRMT) {
                    const uint32_t in = prmt(tbl_lo, tbl_hi, __vadd2(Rk[k], cvec));
                    const uint32_t d = __vadd2(diag, in);
                    const uint32_t t = __viaddmax_s16x2(left, gleft, d);
                    w = __viaddmax_s16x2_relu(up, gup, t) & ZCLR;
                } else if (V == V_LDS) {
                    const uint32_t d = __vadd2(diag, inc[k]);
                    const uint32_t t = __viaddmax_s16x2(left, gleft, d);
                    w = __viaddmax_s16x2_relu(up, gup, t) & ZCLR;
                } else if (V == V_LDS_NOZ) {
                    const uint32_t d = __vadd2(diag, inc[k]);
                    const uint32_t t = __viaddmax_s16x2(left, gleft, d);
                    w = __viaddmax_s16x2_relu(up, gup, t);
                } else if (V == V_LDS_MAX3) {
                    const uint32_t d = __vadd2(diag, inc[k]);
                    const uint32_t l = __vadd2(left, gleft);
                    const uint32_t u = __vadd2(up, gup);
                    w = __vimax3_s16x2_relu(d, l, u) & ZCLR;
                } else if (V == V_LDS_SPLIT) {
                    const uint32_t d = __vadd2(diag, inc[k]);
                    const uint32_t t = __viaddmax_s16x2(left, gleft, d);
                    const uint32_t u = __vadd2(up, gup);
                    w = __vimax_s16x2_relu(u, t) & ZCLR;
                } else {
                    const uint32_t d = __vadd2(diag, inc[k]);
                    const uint32_t l = __vadd2(left, gleft);
                    const uint32_t u = __vadd2(up, gup);
                    w = __vimax_s16x2_relu(u, __vmaxs2(d, l)) & ZCLR;
                }
                diag = left;
                up = w;
                W[k] = w;
            }
            send = prmt(W[K - 1], V == V_PRMT ? cvec : (cvec << 16), 0x7532u);
        }
    }
    uint32_t x = send;
#pragma unroll
    for (int k = 0; k < K; ++k) x ^= W[k];
    if (x == 0x12345u) out[0] = x;
}


// Prototype of the table kernel's steady loop: per-warp shared-memory increment table (16 column-code
// combinations x K registers), a shared-memory ring of per-column table offsets and top-boundary values
// (so nothing but the DP value travels through the shuffle), increments prefetched one step ahead.
template <int K, bool ZCLEAR, bool IMM, bool SKEW3 = false>
__global__ void __launch_bounds__(128) ring_kernel(uint32_t* out, const uint32_t* __restrict__ line, int steps, uint32_t gup_, uint32_t gleft_)
{
    const uint32_t gup = IMM ? 0xfff0fff0u : gup_, gleft = IMM ? 0xffe8ffe8u : gleft_;
    extern __shared__ uint4 smem4[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int TBL4 = 16 * (K / 4) * 32;            // uint4 per warp
    uint4* tbl = smem4 + (size_t)warp * (TBL4 + 64 + 16);
    uint32_t* ring = reinterpret_cast<uint32_t*>(tbl + TBL4);   // 256 words: [0,128) offsets+values, mirrored
    uint16_t* oring = reinterpret_cast<uint16_t*>(tbl + TBL4 + 64);   // 64 halfwords
    for (int c = 0; c < 16; ++c)
        for (int q = 0; q < K / 4; ++q) {
            uint4 v;
            v.x = ((c * 7 + q * 3 + lane) & 1) ? 0x00040004u : 0xffecffecu;
            v.y = ((c * 5 + q + lane) & 2) ? 0x00040004u : 0xffec0004u;
            v.z = ((c + q * 3 + lane) & 1) ? 0x0004ffecu : 0xffecffecu;
            v.w = ((c * 3 + q + lane) & 2) ? 0x00040004u : 0xffecffecu;
            tbl[(c * (K / 4) + q) * 32 + lane] = v;
        }
    for (int e = lane; e < 256; e += 32) ring[e] = 0x0100u | (((line[e & 127] >> 16) & 15u) * (K / 4) * 512u) << 16;
    __syncwarp();
    uint32_t W[K];
#pragma unroll
    for (int k = 0; k < K; ++k) W[k] = 0x01000100u + lane * 8 + k * 16;
    uint32_t up0_prev = 0x01000100u;
    const uint32_t tbl_base = (uint32_t)__cvta_generic_to_shared(tbl) + lane * 16;
    const uint32_t ring_base = (uint32_t)__cvta_generic_to_shared(ring);
    const uint32_t oring_base = (uint32_t)__cvta_generic_to_shared(oring);
    uint32_t incA[K], incB[K];
    auto load_inc = [&](uint32_t (&inc)[K], uint32_t word) {
        const uint32_t addr = tbl_base + (word >> 16);
#pragma unroll
        for (int q = 0; q < K / 4; ++q)
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(inc[4 * q]), "=r"(inc[4 * q + 1]), "=r"(inc[4 * q + 2]), "=r"(inc[4 * q + 3]) : "r"(addr + q * 512));
    };
    auto lds32 = [&](uint32_t addr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; };
    uint32_t recv_next = 0;
    auto step = [&](const uint32_t (&inc)[K], uint32_t word, int s) {
        uint32_t recv;
        if (SKEW3) {                       // lanes three columns apart: the value needed now left the lane above a step ago
            recv = recv_next;
            recv_next = __shfl_up_sync(0xffffffffu, W[K - 1], 1);
        } else {
            recv = __shfl_up_sync(0xffffffffu, W[K - 1], 1);
        }
        if (lane == 0) recv = word << 16;
        const uint32_t up0 = prmt(recv, W[K - 1], 0x5432u);
        uint32_t diag = up0_prev;
        up0_prev = up0;
        uint32_t up = up0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const uint32_t left = W[k];
            const uint32_t d = __vadd2(diag, inc[k]);
            const uint32_t t = __viaddmax_s16x2(left, gleft, d);
            uint32_t w = __viaddmax_s16x2_relu(up, gup, t);
            if (ZCLEAR) w &= ZCLR;
            diag = left;
            up = w;
            W[k] = w;
        }
        if (lane == 31) asm volatile("st.shared.u16 [%0], %1;" :: "r"(oring_base + 2 * s), "h"((uint16_t)(W[K - 1] >> 16)));
    };
    for (int tb = 0; tb < steps; tb += 32) {
        uint32_t p = ring_base + (((uint32_t)(tb - 2 * lane)) & 127u) * 4u;
        uint32_t wordA = lds32(p), wordB;
        load_inc(incA, wordA);
#pragma unroll 1
        for (int s = 0; s < 32; s += 2) {
            wordB = lds32(p + 4);
            load_inc(incB, wordB);
            step(incA, wordA, s);
            wordA = lds32(p + 8);
            load_inc(incA, wordA);
            step(incB, wordB, s + 1);
            p += 8;
        }
        // flush the bottom-row ring (coalesced) -- stands in for the boundary-line store
        __syncwarp();
        out[1024 + ((blockIdx.x * 4 + warp) * 64 + lane)] = oring[lane];
    }
    uint32_t x = up0_prev;
#pragma unroll
    for (int k = 0; k < K; ++k) x ^= W[k];
    if (x == 0x12345u) out[0] = x;
}

// The certificate kernel's steady loop as it is now (overlap_wf16c.cuh, run_block<false,false>): one register set
// of increments reloaded behind the diagonal adds, ring pointer as the only moving address, shuffle one step
// ahead.  LAG2: the hi row group runs TWO columns behind the lo group instead of one, so the head of a step's
// max chain needs the tail of the step before the previous one and consecutive steps can overlap.
template <int K, bool LAG2, int MIX = 0>
__global__ void __launch_bounds__(128, 3) cert_kernel(uint32_t* out, const uint32_t* __restrict__ line, int steps)
{
    const uint32_t gup = 0xfffcfffcu, gleft = 0xfffafffau;
    extern __shared__ uint4 smem4[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int TBL4 = 16 * (K / 4) * 32;
    uint4* tbl = smem4 + (size_t)warp * (TBL4 + 64 + 40);
    uint32_t* ring = reinterpret_cast<uint32_t*>(tbl + TBL4);          // 256 words, mirrored
    uint32_t* oring = reinterpret_cast<uint32_t*>(tbl + TBL4 + 64);    // 160 words
    for (int c = 0; c < 16; ++c)
        for (int q = 0; q < K / 4; ++q) {
            uint4 v;
            v.x = ((c * 7 + q * 3 + lane) & 1) ? 0u : 0xfffafffau;
            v.y = ((c * 5 + q + lane) & 2) ? 0u : 0xfffa0000u;
            v.z = ((c + q * 3 + lane) & 1) ? 0x0000fffau : 0xfffafffau;
            v.w = ((c * 3 + q + lane) & 2) ? 0u : 0xfffafffau;
            tbl[(c * (K / 4) + q) * 32 + lane] = v;
        }
    for (int e = lane; e < 256; e += 32) ring[e] = 0x0100u | (((line[e & 127] >> 16) & 15u) * (K / 4) * 512u) << 16;
    __syncwarp();
    uint32_t W[K];
#pragma unroll
    for (int k = 0; k < K; ++k) W[k] = 0x01000100u + lane * 8 + k * 16;
    uint32_t up0_prev = 0x01000100u, wlast_old = W[K - 1];
    const uint32_t my_tab = (uint32_t)__cvta_generic_to_shared(tbl) + lane * 16;
    const uint32_t ring_base = (uint32_t)__cvta_generic_to_shared(ring);
    constexpr uint32_t OR_OFF = 1024;
    auto lds32 = [&](uint32_t addr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; };
    auto load_inc = [&](uint32_t (&inc)[K], uint32_t word) {
        const uint32_t addr = my_tab + (word >> 16);
#pragma unroll
        for (int q = 0; q < K / 4; ++q)
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(inc[4 * q]), "=r"(inc[4 * q + 1]), "=r"(inc[4 * q + 2]), "=r"(inc[4 * q + 3]) : "r"(addr + q * 512));
    };
    uint32_t recv_next = 0;
    uint32_t inc[K];
    const bool do_store = lane == 31;
    for (int tb = 0; tb < steps; tb += 32) {
        uint32_t p = ring_base + (((uint32_t)(tb - 3 * lane)) & 127u) * 4u;
        uint32_t p_end = p + 128u;
        asm volatile("" : "+r"(p_end));
        uint32_t w0 = lds32(p), w1 = lds32(p + 4);
        load_inc(inc, w0);
        auto step = [&](uint32_t word, uint32_t next_word, uint32_t oaddr) {
            uint32_t d[K];
            d[0] = __vadd2(up0_prev, inc[0]);
#pragma unroll
            for (int k = 1; k < K; ++k) d[k] = __vadd2(W[k - 1], inc[k]);
            load_inc(inc, next_word);
            uint32_t recv = recv_next;
            recv_next = __shfl_up_sync(0xffffffffu, W[K - 1], 1);
            if (lane == 0) recv = word << 16;
            const uint32_t up0 = prmt(recv, LAG2 ? wlast_old : W[K - 1], 0x5432u);
            if (LAG2) wlast_old = W[K - 1];
            up0_prev = up0;
            uint32_t up = up0;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                uint32_t w;
                if (MIX == 0) {                 // today: two 3-input add-max per cell pair
                    const uint32_t t = __viaddmax_s16x2(W[k], gleft, d[k]);
                    w = __viaddmax_s16x2_relu(up, gup, t);
                } else if (MIX == 1) {          // potential with a free left move: 2-input max + one add-max
                    const uint32_t t = __vmaxs2(W[k], d[k]);
                    w = __viaddmax_s16x2_relu(up, gup, t);
                } else {                        // potential with free left and up moves: one 3-input max
                    w = __vimax3_s16x2_relu(W[k], up, d[k]);
                }
                up = w;
                W[k] = w;
            }
            if (do_store) asm volatile("st.shared.u32 [%0], %1;" :: "r"(oaddr), "r"(W[K - 1]) : "memory");
        };
#pragma unroll 1
        do {
            const uint32_t w2 = lds32(p + 8);
            step(w0, w1, p + OR_OFF);
            const uint32_t w3 = lds32(p + 12);
            step(w1, w2, p + OR_OFF + 4);
            w0 = w2; w1 = w3;
            p += 8;
        } while (p != p_end);
        __syncwarp();
        out[1024 + ((blockIdx.x * 4 + warp) * 64 + lane)] = oring[lane];
    }
    uint32_t x = up0_prev ^ wlast_old;
#pragma unroll
    for (int k = 0; k < K; ++k) x ^= W[k];
    if (x == 0x12345u) out[0] = x;
}

// hi-lag 2 with the two steps of an iteration interleaved BY HAND (ptxas keeps each step's chain together): register k
// of step A, then register k-1 of step B, so that dependent max instructions are two apart in program order.
template <int K>
__global__ void __launch_bounds__(128, 3) cert2_kernel(uint32_t* out, const uint32_t* __restrict__ line, int steps)
{
    const uint32_t gup = 0xfffcfffcu, gleft = 0xfffafffau;
    extern __shared__ uint4 smem4[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int TBL4 = 16 * (K / 4) * 32;
    uint4* tbl = smem4 + (size_t)warp * (TBL4 + 64 + 40);
    uint32_t* ring = reinterpret_cast<uint32_t*>(tbl + TBL4);
    uint32_t* oring = reinterpret_cast<uint32_t*>(tbl + TBL4 + 64);
    for (int c = 0; c < 16; ++c)
        for (int q = 0; q < K / 4; ++q) {
            uint4 v;
            v.x = ((c * 7 + q * 3 + lane) & 1) ? 0u : 0xfffafffau;
            v.y = ((c * 5 + q + lane) & 2) ? 0u : 0xfffa0000u;
            v.z = ((c + q * 3 + lane) & 1) ? 0x0000fffau : 0xfffafffau;
            v.w = ((c * 3 + q + lane) & 2) ? 0u : 0xfffafffau;
            tbl[(c * (K / 4) + q) * 32 + lane] = v;
        }
    for (int e = lane; e < 256; e += 32) ring[e] = 0x0100u | (((line[e & 127] >> 16) & 15u) * (K / 4) * 512u) << 16;
    __syncwarp();
    uint32_t W[K];
#pragma unroll
    for (int k = 0; k < K; ++k) W[k] = 0x01000100u + lane * 8 + k * 16;
    uint32_t up0_prev = 0x01000100u, wlast_old = W[K - 1];
    const uint32_t my_tab = (uint32_t)__cvta_generic_to_shared(tbl) + lane * 16;
    const uint32_t ring_base = (uint32_t)__cvta_generic_to_shared(ring);
    constexpr uint32_t OR_OFF = 1024;
    auto lds32 = [&](uint32_t addr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; };
    auto load_inc = [&](uint32_t (&inc)[K], uint32_t word) {
        const uint32_t addr = my_tab + (word >> 16);
#pragma unroll
        for (int q = 0; q < K / 4; ++q)
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(inc[4 * q]), "=r"(inc[4 * q + 1]), "=r"(inc[4 * q + 2]), "=r"(inc[4 * q + 3]) : "r"(addr + q * 512));
    };
    uint32_t recvA = 0, recvB = 0;            // shuffles issued an iteration ahead
    uint32_t incA[K], incB[K];
    const bool do_store = lane == 31;
    for (int tb = 0; tb < steps; tb += 32) {
        uint32_t p = ring_base + (((uint32_t)(tb - 4 * lane)) & 127u) * 4u;
        uint32_t p_end = p + 128u;
        asm volatile("" : "+r"(p_end));
        uint32_t w0 = lds32(p), w1 = lds32(p + 4);
        load_inc(incA, w0);
        load_inc(incB, w1);
#pragma unroll 1
        do {
            const uint32_t w2 = lds32(p + 8), w3 = lds32(p + 12);
            // heads of both chains: neither needs a value of this iteration
            uint32_t rA = recvA, rB = recvB;
            if (lane == 0) { rA = w0 << 16; rB = w1 << 16; }
            const uint32_t wl1 = W[K - 1];                       // tail of the previous step
            const uint32_t upA0 = prmt(rA, wlast_old, 0x5432u);  // hi head of A: tail of two steps ago
            const uint32_t upB0 = prmt(rB, wl1, 0x5432u);        // hi head of B: tail of the previous step
            uint32_t a[K], b[K];
            uint32_t upA = upA0, upB = upB0;
            uint32_t diagA = up0_prev, diagB = upA0;
#pragma unroll
            for (int k = 0; k <= K; ++k) {
                if (k < K) {
                    const uint32_t d = __vadd2(diagA, incA[k]);
                    const uint32_t t = __viaddmax_s16x2(W[k], gleft, d);
                    a[k] = __viaddmax_s16x2_relu(upA, gup, t);
                    diagA = W[k];
                    upA = a[k];
                }
                if (k >= 1) {
                    const uint32_t d = __vadd2(diagB, incB[k - 1]);
                    const uint32_t t = __viaddmax_s16x2(a[k - 1], gleft, d);
                    b[k - 1] = __viaddmax_s16x2_relu(upB, gup, t);
                    diagB = a[k - 1];
                    upB = b[k - 1];
                }
            }
            wlast_old = a[K - 1];
            up0_prev = upB0;
#pragma unroll
            for (int k = 0; k < K; ++k) W[k] = b[k];
            recvA = __shfl_up_sync(0xffffffffu, a[K - 1], 1);
            recvB = __shfl_up_sync(0xffffffffu, b[K - 1], 1);
            load_inc(incA, w2);
            load_inc(incB, w3);
            if (do_store) {
                asm volatile("st.shared.u32 [%0], %1;" :: "r"(p + OR_OFF), "r"(a[K - 1]) : "memory");
                asm volatile("st.shared.u32 [%0], %1;" :: "r"(p + OR_OFF + 4), "r"(b[K - 1]) : "memory");
            }
            w0 = w2; w1 = w3;
            p += 8;
        } while (p != p_end);
        __syncwarp();
        out[1024 + ((blockIdx.x * 4 + warp) * 64 + lane)] = oring[lane];
    }
    uint32_t x = up0_prev ^ wlast_old;
#pragma unroll
    for (int k = 0; k < K; ++k) x ^= W[k];
    if (x == 0x12345u) out[0] = x;
}

struct Row { std::string name; int K, warps; double clk_per_regstep, tcups; float ms; };

template <int K, int V>
static Row run(int warps_per_sm, uint32_t* dout, const uint32_t* dline, int sms, double ghz)
{
    const int block = 128, blocks_per_sm = warps_per_sm / 4;
    const int grid = sms * blocks_per_sm;
    const size_t smem = (V == V_PRMT) ? 0 : (size_t)4 * 16 * (K / 4) * 32 * sizeof(uint4);
    auto kern = step_kernel<K, V>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, block, smem));
    const int steps = 1 << 15;
    kern<<<grid, block, smem>>>(dout, dline, 1024, 0xffe8ffe8u, 0xfff0fff0u, 0xecececfcu, 0xecececfcu);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        kern<<<grid, block, smem>>>(dout, dline, steps, 0xffe8ffe8u, 0xfff0fff0u, 0xecececfcu, 0xecececfcu);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    // register-steps per SM sub-partition: warps_per_sm/4 warps x steps x K
    const double clocks = best * 1e-3 * ghz * 1e9;
    const double regsteps_per_smsp = (double)warps_per_sm / 4 * steps * K;
    const double cells = (double)grid * 4 * 32 * (double)steps * K * 2;
    Row r{VNAME[V], K, warps_per_sm, clocks / regsteps_per_smsp, cells / (best * 1e-3) / 1e12, best};
    printf("%-26s K=%d warps/SM=%2d (occ %d blk)  %7.3f ms  %6.2f clk/reg-step/SMSP  %6.2f TCUPS\n", r.name.c_str(), K, warps_per_sm, occ, best, r.clk_per_regstep, r.tcups);
    return r;
}

template <int K, bool ZCLEAR, bool IMM, bool SKEW3 = false>
static Row run_ring(int warps_per_sm, uint32_t* dout, const uint32_t* dline, int sms, double ghz)
{
    const int block = 128, blocks_per_sm = warps_per_sm / 4;
    const int grid = sms * blocks_per_sm;
    const size_t smem = (size_t)4 * (16 * (K / 4) * 32 + 64 + 16) * sizeof(uint4);
    auto kern = ring_kernel<K, ZCLEAR, IMM, SKEW3>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, block, smem));
    const int steps = 1 << 15;
    kern<<<grid, block, smem>>>(dout, dline, 1024, 0xffe8ffe8u, 0xfff0fff0u);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        kern<<<grid, block, smem>>>(dout, dline, steps, 0xffe8ffe8u, 0xfff0fff0u);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    const double clocks = best * 1e-3 * ghz * 1e9;
    const double regsteps_per_smsp = (double)warps_per_sm / 4 * steps * K;
    const double cells = (double)grid * 4 * 32 * (double)steps * K * 2;
    Row r{std::string(ZCLEAR ? "ring+table" : "ring+table no-tag-clear") + (IMM ? " imm" : " reg") + (SKEW3 ? " skew3" : ""), K, warps_per_sm, clocks / regsteps_per_smsp, cells / (best * 1e-3) / 1e12, best};
    printf("%-26s K=%d warps/SM=%2d (occ %d blk)  %7.3f ms  %6.2f clk/reg-step/SMSP  %6.2f TCUPS\n", r.name.c_str(), K, warps_per_sm, occ, best, r.clk_per_regstep, r.tcups);
    return r;
}

template <int K, bool LAG2, int MIX = 0>
static Row run_cert(int warps_per_sm, uint32_t* dout, const uint32_t* dline, int sms, double ghz)
{
    const int block = 128, blocks_per_sm = warps_per_sm / 4;
    const int grid = sms * blocks_per_sm;
    const size_t smem = (size_t)4 * (16 * (K / 4) * 32 + 64 + 40) * sizeof(uint4);
    auto kern = cert_kernel<K, LAG2, MIX>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, block, smem));
    const int steps = 1 << 15;
    kern<<<grid, block, smem>>>(dout, dline, 1024);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        kern<<<grid, block, smem>>>(dout, dline, steps);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    const double clocks = best * 1e-3 * ghz * 1e9;
    const double regsteps_per_smsp = (double)warps_per_sm / 4 * steps * K;
    const double cells = (double)grid * 4 * 32 * (double)steps * K * 2;
    Row r{std::string("certificate loop") + (LAG2 ? " hi-lag 2" : " hi-lag 1 (current)") + (MIX == 1 ? ", free-left potential (VIMNMX + VIADDMNMX)" : MIX == 2 ? ", free-left-and-up potential (VIMNMX3)" : ""), K, warps_per_sm, clocks / regsteps_per_smsp, cells / (best * 1e-3) / 1e12, best};
    printf("%-80s K=%d warps/SM=%2d (occ %d blk)  %7.3f ms  %6.2f clk/reg-step/SMSP  %6.2f TCUPS\n", r.name.c_str(), K, warps_per_sm, occ, best, r.clk_per_regstep, r.tcups);
    return r;
}

template <int K>
static Row run_cert2(int warps_per_sm, uint32_t* dout, const uint32_t* dline, int sms, double ghz)
{
    const int block = 128, blocks_per_sm = warps_per_sm / 4;
    const int grid = sms * blocks_per_sm;
    const size_t smem = (size_t)4 * (16 * (K / 4) * 32 + 64 + 40) * sizeof(uint4);
    auto kern = cert2_kernel<K>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, block, smem));
    const int steps = 1 << 15;
    kern<<<grid, block, smem>>>(dout, dline, 1024);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        kern<<<grid, block, smem>>>(dout, dline, steps);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    const double clocks = best * 1e-3 * ghz * 1e9;
    const double regsteps_per_smsp = (double)warps_per_sm / 4 * steps * K;
    const double cells = (double)grid * 4 * 32 * (double)steps * K * 2;
    Row r{"certificate loop hi-lag 2, steps interleaved", K, warps_per_sm, clocks / regsteps_per_smsp, cells / (best * 1e-3) / 1e12, best};
    printf("%-44s K=%d warps/SM=%2d (occ %d blk)  %7.3f ms  %6.2f clk/reg-step/SMSP  %6.2f TCUPS\n", r.name.c_str(), K, warps_per_sm, occ, best, r.clk_per_regstep, r.tcups);
    return r;
}

int main(int argc, char** argv)
{
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    int khz = 0; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    const double ghz = khz * 1e-6;
    printf("device %s, %d SMs, %.3f GHz nominal\n", p.name, sms, ghz);
    uint32_t* dout; CK(cudaMalloc(&dout, 4096 + 148 * 16 * 64 * 4 * 4));
    uint32_t* dline; CK(cudaMalloc(&dline, 4096));
    std::vector<uint32_t> h(1024);
    for (int i = 0; i < 1024; ++i) h[i] = 0x0100u | ((uint32_t)(rand() & 3) * 0x11u << 16);
    CK(cudaMemcpy(dline, h.data(), 4096, cudaMemcpyHostToDevice));
    std::vector<Row> rows;
    rows.push_back(run_cert<8, false>(12, dout, dline, sms, ghz));
    rows.push_back(run_cert<8, true>(12, dout, dline, sms, ghz));
    rows.push_back(run_cert<8, false, 1>(12, dout, dline, sms, ghz));
    rows.push_back(run_cert<8, false, 2>(12, dout, dline, sms, ghz));
    rows.push_back(run_cert<8, true, 2>(12, dout, dline, sms, ghz));
    rows.push_back(run_cert<8, false, 2>(16, dout, dline, sms, ghz));
    rows.push_back(run_cert2<8>(12, dout, dline, sms, ghz));
    rows.push_back(run_cert2<8>(8, dout, dline, sms, ghz));
    rows.push_back(run_cert<8, false>(8, dout, dline, sms, ghz));
    rows.push_back(run_cert<8, true>(8, dout, dline, sms, ghz));
    rows.push_back(run_cert<4, false>(12, dout, dline, sms, ghz));
    rows.push_back(run_cert<4, true>(12, dout, dline, sms, ghz));
    if (argc > 2 && std::string(argv[2]) == "cert-only") goto write_out;
    rows.push_back(run<8, V_PRMT>(16, dout, dline, sms, ghz));
    rows.push_back(run<8, V_PRMT>(12, dout, dline, sms, ghz));
    rows.push_back(run<8, V_LDS>(12, dout, dline, sms, ghz));
    rows.push_back(run<8, V_LDS>(8, dout, dline, sms, ghz));
    rows.push_back(run<8, V_LDS_NOZ>(12, dout, dline, sms, ghz));
    rows.push_back(run<8, V_LDS_MAX3>(12, dout, dline, sms, ghz));
    rows.push_back(run<8, V_LDS_SPLIT>(12, dout, dline, sms, ghz));
    rows.push_back(run<8, V_LDS_ALLSPLIT>(12, dout, dline, sms, ghz));
    rows.push_back(run<4, V_LDS>(16, dout, dline, sms, ghz));
    rows.push_back(run<4, V_LDS>(24, dout, dline, sms, ghz));
    rows.push_back(run<4, V_LDS_SPLIT>(24, dout, dline, sms, ghz));
    rows.push_back(run<4, V_LDS_ALLSPLIT>(24, dout, dline, sms, ghz));
    rows.push_back(run<12, V_LDS>(8, dout, dline, sms, ghz));
    rows.push_back(run<12, V_LDS_SPLIT>(8, dout, dline, sms, ghz));
    rows.push_back(run_ring<8, true, false>(12, dout, dline, sms, ghz));
    rows.push_back(run_ring<8, true, false>(8, dout, dline, sms, ghz));
    rows.push_back(run_ring<8, false, false>(12, dout, dline, sms, ghz));
    rows.push_back(run_ring<4, true, false>(20, dout, dline, sms, ghz));
    rows.push_back(run_ring<4, true, true>(24, dout, dline, sms, ghz));
    rows.push_back(run_ring<8, true, true>(12, dout, dline, sms, ghz));
    rows.push_back(run_ring<8, false, true>(12, dout, dline, sms, ghz));
    rows.push_back(run_ring<8, true, true, true>(12, dout, dline, sms, ghz));
    rows.push_back(run_ring<8, true, true, true>(8, dout, dline, sms, ghz));
    rows.push_back(run_ring<8, false, true, true>(12, dout, dline, sms, ghz));
    rows.push_back(run_ring<4, true, true, true>(24, dout, dline, sms, ghz));
write_out:
    if (argc > 1) {
        FILE* f = fopen(argv[1], "w");
        if (f) {
            fprintf(f, "{\"device\": \"%s\", \"sms\": %d, \"nominal_sm_ghz\": %.3f, \"rows\": [\n", p.name, sms, ghz);
            for (size_t i = 0; i < rows.size(); ++i)
                fprintf(f, "  {\"variant\": \"%s\", \"K\": %d, \"warps_per_sm\": %d, \"ms\": %.4f, \"clk_per_regstep_per_smsp\": %.3f, \"loop_tcups\": %.3f}%s\n",
                        rows[i].name.c_str(), rows[i].K, rows[i].warps, rows[i].ms, rows[i].clk_per_regstep, rows[i].tcups, i + 1 < rows.size() ? "," : "");
            fprintf(f, "]}\n"); fclose(f);
        }
    }
    return 0;
}
