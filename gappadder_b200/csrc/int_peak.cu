// int_peak.cu -- live measurement of the integer issue-rate ceiling the DP kernels are judged
// against (DESIGN.md "Roofline").  Same chains as tools/microbench_int.cu, reduced to the two
// numbers bench.py needs:
//   alu   : VIADDMNMX.S16x2 alone (the ALU pipe: 2 warp-instructions / clk / SM)
//   dual  : VIMNMX.S16x2 + VIADD.16x2 interleaved (ALU pipe + FMA-side pipe: 4 / clk / SM)
// Rates are thread-level instructions per second for the whole chip; each instruction works on
// two 16-bit lanes.
#include "gappadder_b200.h"
#include <cuda_runtime.h>
#include <cstdint>

namespace {

constexpr int ILP = 8;

template <int MODE>
__global__ void __launch_bounds__(256) peak_kernel(unsigned* out, unsigned c, int iters)
{
    unsigned acc[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) acc[k] = threadIdx.x * 2654435761u + k * 40503u + blockIdx.x;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int k = 0; k < ILP; ++k) {
                unsigned b = acc[(k + 3) % ILP];
                if (MODE == 0) acc[k] = __viaddmax_s16x2(acc[k], b, c);
                else acc[k] = __vimax_s16x2_relu(__vadd2(acc[k], c), b);
            }
        }
    }
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s ^= acc[k];
    if (s == 0x12345u) out[0] = s;
}

template <int MODE>
double run_peak(cudaStream_t st, int sms, unsigned* dout, int inst_per_op)
{
    const int grid = sms * 8, block = 256, iters = 2048;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    peak_kernel<MODE><<<grid, block, 0, st>>>(dout, 0xfffe0005u, 64);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0, st);
        peak_kernel<MODE><<<grid, block, 0, st>>>(dout, 0xfffe0005u, iters);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        if (ms > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    const double ops = (double)grid * block * (double)iters * 4 * ILP * inst_per_op;
    return ops / (best * 1e-3);
}

} // namespace

extern "C" int gp_int_peak(gp_ctx* ctx, double* alu_inst_per_s, double* dual_inst_per_s)
{
    if (!ctx) return GP_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)gp_stream(ctx);
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return GP_ERR_CUDA;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return GP_ERR_CUDA;
    unsigned* dout = nullptr;
    if (cudaMalloc(&dout, 256) != cudaSuccess) return GP_ERR_CUDA;
    double a = run_peak<0>(st, sms, dout, 1);
    double d = run_peak<1>(st, sms, dout, 2);
    cudaFree(dout);
    if (cudaGetLastError() != cudaSuccess) return GP_ERR_CUDA;
    if (alu_inst_per_s) *alu_inst_per_s = a;
    if (dual_inst_per_s) *dual_inst_per_s = d;
    return GP_OK;
}
