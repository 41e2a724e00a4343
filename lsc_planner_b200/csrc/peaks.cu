// Roofline denominators measured on the device the engine runs on (instrumentation; SURVEY.md §8d: FP32 / FP64 FMA
// peaks are not in MEASURED_PEAKS.json, so the bench measures them in the same run, with the same clocks, as the
// number they are set against) and the dependent-issue latencies that bound the QP's one-warp chain.
#include <cuda_runtime.h>

#include <algorithm>

#include "../../include/lscgpu.h"

namespace {

template <typename T>
__global__ void k_fma_chain(T* out, int iters) {
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
double measure_peak(int sms, int iters) {
    const int blocks = sms * 8, threads = 256;
    T* out = nullptr;
    if (cudaMalloc(&out, sizeof(T) * blocks * threads) != cudaSuccess) return 0.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        k_fma_chain<T><<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 64.0 * (double)iters * blocks * threads;      // 64 FMAs per loop iteration
        if (rep > 0 && ms > 0) best = std::max(best, flops / (ms * 1e-3));
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    return best * 1e-12;
}

// one warp, one dependent chain of 2048 operations: cycles per operation
__global__ void k_latency(double* out) {
    __shared__ double sm[64];
    const int lane = threadIdx.x;
    sm[lane] = 1.0 + lane * 1e-9; sm[lane + 32] = 0.5;
    __syncwarp();
    double x = 1.0 + lane * 1e-9;
    const double m = 0.999999, c = 1e-6;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < 2048; i++) x = x * m + c;
    long long t1 = clock64();
    const double dfma = (double)(t1 - t0) / 2048.0;
    float y = (float)x;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < 2048; i++) y = y * 0.999999f + 1e-6f;
    t1 = clock64();
    const double ffma = (double)(t1 - t0) / 2048.0;
    // shared-memory pointer chase
    int idx = lane;
    __shared__ int nxt[32];
    nxt[lane] = (lane * 7 + 3) & 31;
    __syncwarp();
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < 2048; i++) idx = nxt[idx];
    t1 = clock64();
    const double lds = (double)(t1 - t0) / 2048.0;
    // shuffle + add chain (one round of a double warp reduction: two 32-bit shuffles and one DADD)
    double s = x;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < 512; i++) s += __shfl_xor_sync(0xffffffffu, s, 1 + (i & 15));
    t1 = clock64();
    const double shfl = (double)(t1 - t0) / 512.0;
    double r = x + 2.0;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < 256; i++) r = rsqrt(r) + 1.5;
    t1 = clock64();
    const double rsq = (double)(t1 - t0) / 256.0;
    double dv = x + 2.0;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < 256; i++) dv = 1.7 / dv + 1.1;
    t1 = clock64();
    const double div = (double)(t1 - t0) / 256.0;
    unsigned u = (unsigned)lane;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < 512; i++) u = __reduce_min_sync(0xffffffffu, u + (unsigned)i) + (unsigned)lane;
    t1 = clock64();
    const double redux = (double)(t1 - t0) / 512.0;
    if (lane == 0) {
        out[0] = dfma; out[1] = ffma; out[2] = lds; out[3] = shfl; out[4] = rsq; out[5] = div; out[6] = redux;
        out[7] = (double)y + idx + s + r + dv + u;      // keep everything alive
    }
}

}  // namespace

extern "C" int lscgpu_measure_fma_peaks(int device, double* fp32_tflops, double* fp64_tflops) {
    if (!fp32_tflops || !fp64_tflops) return LSCGPU_ERR_ARG;
    cudaDeviceProp p;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&p, device) != cudaSuccess) return LSCGPU_ERR_CUDA;
    *fp32_tflops = measure_peak<float>(p.multiProcessorCount, 20000);
    *fp64_tflops = measure_peak<double>(p.multiProcessorCount, 4000);
    return cudaGetLastError() == cudaSuccess ? LSCGPU_OK : LSCGPU_ERR_CUDA;
}

extern "C" int lscgpu_measure_latencies(int device, double cycles_out[7]) {
    if (!cycles_out) return LSCGPU_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return LSCGPU_ERR_CUDA;
    double* d = nullptr;
    if (cudaMalloc(&d, sizeof(double) * 8) != cudaSuccess) return LSCGPU_ERR_CUDA;
    k_latency<<<1, 32>>>(d);
    double h[8];
    const cudaError_t rc = cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (rc != cudaSuccess) return LSCGPU_ERR_CUDA;
    for (int i = 0; i < 7; i++) cycles_out[i] = h[i];
    return LSCGPU_OK;
}
