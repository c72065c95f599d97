// Micro-benchmark: read bandwidth of buffers that fit the L2 (B200: 126 MB) against an HBM-sized one, and a mix of the two
// (half of the bytes streamed from HBM, half re-read from an L2-resident window) — the access mix of a symmetric
// (upper-triangle) block SpMV.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l2bw tools/l2bw.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_read(const double2* __restrict__ p, size_t n, int reps, double* out) {
    double s = 0.0;
    for (int r = 0; r < reps; ++r)
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
            const double2 v = p[i];
            s += v.x + v.y;
        }
    if (s == 1.2345) *out = s;
}
// streams `big` once while re-reading a sliding window of `big` that was touched `lag` elements earlier
__global__ void k_mix(const double2* __restrict__ big, size_t n, size_t lag, double* out) {
    double s = 0.0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double2 v = big[i];
        const double2 w = i >= lag ? big[i - lag] : v;
        s += v.x + v.y + w.x + w.y;
    }
    if (s == 1.2345) *out = s;
}

int main() {
    double* out;
    cudaMalloc(&out, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const size_t MB = 1 << 20;
    double2* buf;
    const size_t big = 4096 * MB;
    cudaMalloc(&buf, big);
    cudaMemset(buf, 0, big);
    for (size_t mb : {8, 16, 32, 48, 64, 96, 128, 256, 4096}) {
        const size_t n = mb * MB / sizeof(double2);
        const int reps = mb >= 4096 ? 2 : (int)(8192 / mb);
        k_read<<<148 * 8, 512>>>(buf, n, 1, out);
        cudaEventRecord(e0);
        k_read<<<148 * 8, 512>>>(buf, n, reps, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("read %5zu MB x %4d: %8.1f GB/s\n", mb, reps, (double)mb * MB * reps / (ms * 1e-3) / 1e9);
    }
    for (size_t lag_mb : {1, 4, 16, 32, 64}) {
        const size_t n = big / sizeof(double2), lag = lag_mb * MB / sizeof(double2);
        k_mix<<<148 * 8, 512>>>(buf, n, lag, out);
        cudaEventRecord(e0);
        k_mix<<<148 * 8, 512>>>(buf, n, lag, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("stream 4096 MB + re-read at lag %3zu MB: %8.1f GB/s from HBM (%.1f GB/s into the SMs), %.3f ms\n", lag_mb,
               (double)big / (ms * 1e-3) / 1e9, 2.0 * big / (ms * 1e-3) / 1e9, ms);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
