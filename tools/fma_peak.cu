// FMA-chain microbenchmark: peak FP32 FFMA and FP64 DFMA throughput of the device (SURVEY.md §8d asks for these
// denominators, MEASURED_PEAKS.json only carries HBM and bf16 tensor numbers).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fma_peak tools/fma_peak.cu && tools/fma_peak
// Prints one JSON line: {"fp32_tflops": ..., "fp64_tflops": ..., "sm_count": ..., "clock_mhz": ...}
#include <cstdio>
#include <cuda_runtime.h>

template <typename T>
__global__ void fma_chain(T* out, int iters) {
    T a0 = (T)threadIdx.x * (T)1e-3, a1 = a0 + (T)1, a2 = a0 + (T)2, a3 = a0 + (T)3, a4 = a0 + (T)4, a5 = a0 + (T)5,
      a6 = a0 + (T)6, a7 = a0 + (T)7;
    const T m = (T)0.999999, c = (T)1e-6;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            a0 = a0 * m + c; a1 = a1 * m + c; a2 = a2 * m + c; a3 = a3 * m + c;
            a4 = a4 * m + c; a5 = a5 * m + c; a6 = a6 * m + c; a7 = a7 * m + c;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

template <typename T>
double measure(int sms, int iters) {
    const int blocks = sms * 8, threads = 256;
    T* out;
    cudaMalloc(&out, sizeof(T) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 6; rep++) {
        cudaEventRecord(e0);
        fma_chain<T><<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 64.0 * (double)iters * blocks * threads;      // 64 FMAs per loop iteration
        if (rep > 0) best = flops / (ms * 1e-3) > best ? flops / (ms * 1e-3) : best;
    }
    cudaFree(out);
    return best * 1e-12;
}

int main() {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) { std::fprintf(stderr, "no CUDA device\n"); return 1; }
    const double f32 = measure<float>(p.multiProcessorCount, 20000);
    const double f64 = measure<double>(p.multiProcessorCount, 4000);
    std::printf("{\"gpu\": \"%s\", \"sm_count\": %d, \"clock_mhz\": %d, \"fp32_tflops\": %.2f, \"fp64_tflops\": %.2f, "
                "\"how\": \"8 independent FMA chains per thread, 256 threads x 8 blocks per SM, best of 5, CUDA events\"}\n",
                p.name, p.multiProcessorCount, p.clockRate / 1000, f32, f64);
    return 0;
}
