// microbench.cu -- the two device ceilings the accumulate kernel is measured against that
// MEASURED_PEAKS.json does not hold: the fp64 FMA pipe and L2-resident read bandwidth.
// Used only by bench.py to state roofline fractions "of measured".
#include <vector>

#include "../../include/pb200_lbl.h"
#include "common.cuh"

namespace pb200 {

__global__ void __launch_bounds__(256) fp64_fma_kernel(double *out, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void __launch_bounds__(256) l2_read_kernel(const double2 *__restrict__ buf, size_t n,
                                                      int passes, double *out) {
    double acc = 0.0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int p = 0; p < passes; p++) {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
            const double2 v = __ldcg(buf + i);  // cache in L2 only: every pass re-reads L2
            acc += v.x + v.y;
        }
    }
    if (acc == 123.456) out[0] = acc;
}

}  // namespace pb200

using namespace pb200;

extern "C" {

// fp64 FMA throughput in TFLOP/s (2 flops per FMA), best of `reps`.
int pb200_bench_fp64(int device, int reps, double *tflops) {
    if (!tflops) return PB200_EINVAL;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        set_error("no usable CUDA device");
        return PB200_ENODEVICE;
    }
    PB_CUDA(cudaSetDevice(device));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int blocks = sms * 8, threads = 256, iters = 1 << 15;
    double *d_out = nullptr;
    PB_CUDA(cudaMalloc((void **)&d_out, sizeof(double) * blocks * threads));
    cudaEvent_t a, b;
    PB_CUDA(cudaEventCreate(&a));
    PB_CUDA(cudaEventCreate(&b));
    double best = 0.0;
    for (int r = 0; r < reps + 1; r++) {
        PB_CUDA(cudaEventRecord(a));
        fp64_fma_kernel<<<blocks, threads>>>(d_out, iters, 1.0 + r);
        PB_CUDA(cudaEventRecord(b));
        PB_CUDA(cudaEventSynchronize(b));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        const double tf = 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
        if (r > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d_out);
    *tflops = best;
    return 0;
}

// Read bandwidth (GB/s) of a buffer of `mbytes` MiB that stays resident in L2.
int pb200_bench_l2(int device, int mbytes, int reps, double *gbs) {
    if (!gbs || mbytes < 1) return PB200_EINVAL;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        set_error("no usable CUDA device");
        return PB200_ENODEVICE;
    }
    PB_CUDA(cudaSetDevice(device));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const size_t bytes = (size_t)mbytes << 20;
    const size_t nel = bytes / sizeof(double2);
    double2 *buf = nullptr;
    double *d_out = nullptr;
    PB_CUDA(cudaMalloc((void **)&buf, bytes));
    PB_CUDA(cudaMalloc((void **)&d_out, sizeof(double)));
    PB_CUDA(cudaMemset(buf, 0, bytes));
    cudaEvent_t a, b;
    PB_CUDA(cudaEventCreate(&a));
    PB_CUDA(cudaEventCreate(&b));
    const int passes = 20;
    double best = 0.0;
    for (int r = 0; r < reps + 1; r++) {
        PB_CUDA(cudaEventRecord(a));
        l2_read_kernel<<<sms * 8, 256>>>(buf, nel, passes, d_out);
        PB_CUDA(cudaEventRecord(b));
        PB_CUDA(cudaEventSynchronize(b));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        const double g = (double)bytes * passes / (ms * 1e-3) / 1e9;
        if (r > 0 && g > best) best = g;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(buf);
    cudaFree(d_out);
    *gbs = best;
    return 0;
}

}  // extern "C"
