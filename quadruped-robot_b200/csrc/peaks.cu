// peaks.cu -- measures the FP64 / FP32 FMA-pipe peaks of the device (roofline denominators that
// MEASURED_PEAKS.json does not carry: SURVEY.md section 8d asks the builder to measure them).
// Built into libqr_peaks.so; used by bench.py only.
#include <cuda_runtime.h>

template <typename T>
__global__ void fma_chain_kernel(T* out, int iters, T a, T b) {
    T x0 = a + threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b;
            x4 = x4 * a + b; x5 = x5 * a + b; x6 = x6 * a + b; x7 = x7 * a + b;
        }
    }
    T s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == (T)12345.678) out[0] = s;   // never true; keeps the chains alive
}

template <typename T>
static double measure(int sm_count, int iters) {
    T* out = nullptr;
    if (cudaMalloc(&out, sizeof(T)) != cudaSuccess) return -1.0;
    const int blocks = sm_count * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    fma_chain_kernel<T><<<blocks, threads>>>(out, iters / 8, (T)0.999999, (T)1e-6);
    cudaDeviceSynchronize();
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        fma_chain_kernel<T><<<blocks, threads>>>(out, iters, (T)0.999999, (T)1e-6);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 64.0 * (double)iters * (double)blocks * threads;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return cudaGetLastError() == cudaSuccess ? best : -1.0;
}

extern "C" int qr_peak_fma(double* fp64_tflops, double* fp32_tflops) {
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return -5;
    *fp64_tflops = measure<double>(prop.multiProcessorCount, 4096);
    *fp32_tflops = measure<float>(prop.multiProcessorCount, 8192);
    return (*fp64_tflops > 0 && *fp32_tflops > 0) ? 0 : -5;
}
