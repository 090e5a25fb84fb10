// Micro-benchmark: peak issue rate of VABSDIFF4.U8.ACC (the SAD primitive of the ME kernel) on sm_100a.
// The roofline denominator for me_kernel is measured with this, not assumed.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(unsigned *out, unsigned a0, unsigned b0, int iters)
{
    unsigned acc[8], a = a0 + threadIdx.x, b = b0 + blockIdx.x;
#pragma unroll
    for (int i = 0; i < 8; i++)
        acc[i] = i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < 8; i++)
                asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a + i), "r"(b + u));
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++)
        s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int blocks = p.multiProcessorCount * 8, threads = 256, iters = 4096;
    unsigned *out;
    cudaMalloc(&out, (size_t)blocks * threads * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<<<blocks, threads>>>(out, 1, 2, 16);
    cudaDeviceSynchronize();
    float best = 1e9;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        k<<<blocks, threads>>>(out, 1, 2, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best)
            best = ms;
    }
    double instr = (double)blocks * threads * iters * 32.0;
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"sms\": %d, \"vabsdiff4_lane_instr_per_s\": %.4e, \"absdiff_per_s\": %.4e, \"ms\": %.3f, "
           "\"per_sm_per_clk_at_max_clock\": %.2f, \"max_clock_khz\": %d}\n",
           p.multiProcessorCount, instr / (best * 1e-3), 4 * instr / (best * 1e-3), best,
           instr / (best * 1e-3) / p.multiProcessorCount / (clk * 1e3), clk);
    return 0;
}
