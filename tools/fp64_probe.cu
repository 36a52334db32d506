// FP64 pipe probe for B200: issue rate as a function of resident warps per SM and independent
// chains per thread (no FMA: DADD/DMUL alternating, as in the render path).  Prints one line per
// configuration: warps/SM, chains, T lane-ops/s.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
// -O3 --fmad=false -o fp64_probe tools/fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int C>
__global__ void probe(double* sink, int iters, double m, double a) {
    double v[C];
#pragma unroll
    for (int c = 0; c < C; c++) v[c] = threadIdx.x * 1e-9 + c;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
#pragma unroll
            for (int c = 0; c < C; c++) v[c] = (u & 1) ? __dadd_rn(v[c], a) : __dmul_rn(v[c], m);
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < C; c++) s += v[c];
    if (s == 123.456) sink[0] = s;
}

template <int C>
void run(double* sink, int sms, int warps_per_sm) {
    int threads = 32 * warps_per_sm;      // one block per SM
    int block = threads > 1024 ? 1024 : threads;
    int blocks = sms * (threads / block);
    int iters = 2048;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        probe<C><<<blocks, block>>>(sink, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double ops = double(blocks) * block * iters * 16.0 * C;
    printf("warps_per_sm=%2d chains=%d  %.2f Tlaneop/s  (%.3f ms)\n", warps_per_sm, C, ops / (best * 1e-3) / 1e12, best);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* sink; cudaMalloc(&sink, 64);
    int sms = p.multiProcessorCount;
    printf("%s, %d SMs\n", p.name, sms);
    for (int w : {4, 8, 16, 32, 64}) {
        run<1>(sink, sms, w); run<2>(sink, sms, w); run<4>(sink, sms, w); run<8>(sink, sms, w);
    }
    return 0;
}
