// Micro-benchmark: FP64 vector FMA vs FP64 tensor-core MMA (m8n8k4) throughput on this GPU, alone and mixed.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NMMA, int NFMA>
__global__ void k(double* out, int iters)
{
    double c[8][2], f[8];
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = 1.0; f[i] = threadIdx.x * 1e-9 + i; }
    double a = 1.0000001, b = 1e-7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NMMA; ++i) dmma(c[i][0], c[i][1], a, b);
#pragma unroll
        for (int i = 0; i < NFMA; ++i) f[i] = fma(f[i], a, b);
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NMMA, int NFMA>
void run(const char* name, int warps_per_sm)
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int threads = 32 * warps_per_sm, blocks = p.multiProcessorCount, iters = 1 << 14;
    double* d; cudaMalloc(&d, sizeof(double) * threads * blocks);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0); k<NMMA, NFMA><<<blocks, threads>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
    }
    double nw = (double)blocks * warps_per_sm * iters;
    double mma_tf = nw * NMMA * 512.0 / (best * 1e-3) / 1e12, fma_tf = nw * NFMA * 64.0 / (best * 1e-3) / 1e12;
    double cyc = best * 1e-3 * p.clockRate * 1e3 / iters;   // cycles per iteration per warp slot (nominal clock)
    printf("%-28s warps/SM %2d  %.3f ms  DMMA %.2f TF  DFMA %.2f TF  cycles/iter %.1f\n", name, warps_per_sm, best, mma_tf, fma_tf, cyc);
    cudaFree(d);
}
int main()
{
    for (int w : {4, 8, 16, 32}) {
        run<8, 0>("dmma x8 independent", w);
        run<1, 0>("dmma x1 dependent chain", w);
        run<0, 8>("dfma x8 independent", w);
        run<0, 1>("dfma x1 dependent chain", w);
        run<4, 8>("dmma x4 + dfma x8", w);
        run<1, 8>("dmma x1 + dfma x8", w);
    }
    return 0;
}
