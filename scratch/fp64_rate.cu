// DFMA / FFMA / F2F throughput per SM (B200): is fp64 a first-class pipe here?
#include <cstdio>
#include <cuda_runtime.h>
template <typename T> __global__ void fma_chain(T* out, int iters) {
    T a[8]; for (int k = 0; k < 8; ++k) a[k] = (T)(threadIdx.x + k) * (T)1e-3;
    const T b = (T)1.0000001, c = (T)1e-7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = a[k] * b + c;
    }
    T s = 0; for (int k = 0; k < 8; ++k) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void cvt_chain(double* out, int iters) {
    float x = threadIdx.x * 1e-3f; double acc = 0.0;
    for (int i = 0; i < iters; ++i) { acc += (double)x; x += 1.0f; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
    void* buf; cudaMalloc(&buf, 148 * 8 * 256 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096, blocks = 148 * 8, thr = 256;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); fma_chain<double><<<blocks, thr>>>((double*)buf, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fma = (double)blocks * thr * iters * 8;
        printf("DFMA: %.3f ms  %.2f TFLOP/s  (%.1f FMA/clk/SM at 1.965 GHz)\n", ms, 2 * fma / ms / 1e9, fma / (ms * 1e-3) / 148 / 1.965e9);
        cudaEventRecord(e0); fma_chain<float><<<blocks, thr>>>((float*)buf, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("FFMA: %.3f ms  %.2f TFLOP/s  (%.1f FMA/clk/SM)\n", ms, 2 * fma / ms / 1e9, fma / (ms * 1e-3) / 148 / 1.965e9);
        cudaEventRecord(e0); cvt_chain<<<blocks, thr>>>((double*)buf, iters * 8); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("F2F.F64.F32 + DADD: %.3f ms  (%.1f pairs/clk/SM)\n", ms, fma / (ms * 1e-3) / 148 / 1.965e9);
    }
    return 0;
}
