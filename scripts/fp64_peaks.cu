// scripts/fp64_peaks.cu -- measured FP64 peaks of the box: vector DFMA and tensor DMMA (mma.sync.m8n8k4.f64), the denominators of
// the contraction-bound configurations ("the builder must measure FP64 FMA and DMMA peaks").
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/fp64_peaks scripts/fp64_peaks.cu && gpurun_out/fp64_peaks
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double *out, int iters, double a, double b)
{
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma_kernel(double *out, int iters, double a, double b)
{
    double c[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-3;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 20000;
    double *out;
    cudaMalloc(&out, (size_t)blocks * threads * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    double best_fma = 0, best_mma = 0;
    for (int rep = 0; rep < 5; ++rep) {
        dfma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e0);
        dfma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        double tf = 2.0 * 8 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
        if (tf > best_fma) best_fma = tf;
        dmma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e0);
        dmma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        // one m8n8k4 per warp = 8*8*4 multiply-adds
        tf = 2.0 * 256 * 4 * iters * (double)blocks * (threads / 32) / (ms * 1e-3) / 1e12;
        if (tf > best_mma) best_mma = tf;
    }
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"fp64_dfma_tflops\": %.2f, \"fp64_dmma_m8n8k4_tflops\": %.2f, "
           "\"how\": \"8 independent DFMA chains per thread / 4 independent mma.sync.m8n8k4.f64 accumulators per warp, %d blocks x %d threads, "
           "best of 5, CUDA events\"}\n", p.name, p.multiProcessorCount, best_fma, best_mma, blocks, threads);
    return cudaGetLastError() != cudaSuccess;
}
