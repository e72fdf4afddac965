// microbench.cu -- the two device ceilings the accumulate kernel is measured against that
// MEASURED_PEAKS.json does not hold: the fp64 FMA pipe and L2-resident read bandwidth.
// Used only by bench.py to state roofline fractions "of measured".
#include <cmath>
#include <cstring>
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

// Self-test of the exactness claims of common.cuh: quotient_rn(a, b, RN(1/b)) against the IEEE
// division, and nearest_index_thr against nearest_index, on pseudo-random operands of the
// shapes the accumulate kernel sees (a = wn - own0 in [0, 4e4), b = ownstep * divisor).
__global__ void __launch_bounds__(256)
selftest_kernel(long long n, unsigned long long seed, const double *__restrict__ steps, int nsteps,
                const double *__restrict__ grid, const double *__restrict__ thr, int ngrid,
                int hi0, float inv_step, unsigned long long *__restrict__ bad /* [2] */) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long x = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
    const double u = (double)(x >> 11) * (1.0 / 9007199254740992.0);          // [0, 1)
    unsigned long long y = x * 0xD6E8FEB86659FD93ull; y ^= y >> 32;
    const double b = steps[(int)(y % (unsigned long long)nsteps)];
    // operands near multiples of b (where the truncated quotient can flip) and generic ones
    double a = u * 4.0e4;
    if (i & 1) a = __dmul_rn(b, floor(u * 4.0e7)) + ((i & 2) ? 0.0 : __dmul_rn(b, 1e-9 * (u - 0.5)));
    if (a < 0.0) a = 0.0;
    const double q = quotient_rn(a, b, __ddiv_rn(1.0, b));
    if (q != __ddiv_rn(a, b)) atomicAdd(&bad[0], 1ull);
    if (ngrid >= 2) {
        const double span = grid[ngrid - 1] * 1.2;
        double v = u * span;
        if (i & 4) v = thr[1 + (int)(y % (unsigned long long)(ngrid - 1))];     // exactly on a step
        if ((i & 12) == 12) v = __longlong_as_double(__double_as_longlong(v) - 1);  // just below it
        if (nearest_index_thr(thr, ngrid, v, hi0, inv_step) != nearest_index(grid, ngrid, v))
            atomicAdd(&bad[1], 1ull);
    }
}

}  // namespace pb200

using namespace pb200;

extern "C" {

// Device self-test (tests/test_gpu_parity.py): mismatches[0] = operands on which the FMA-only
// quotient differs from the IEEE division, mismatches[1] = widths on which the threshold-table
// nearest index differs from the bisection.  steps[nsteps]: divisors (dynamic steps) to draw
// from; grid/thr[ngrid]: a Doppler grid and its threshold table (pb200_nearest_thresholds).
int pb200_selftest_exact(int device, int64_t n, uint64_t seed, const double *steps, int nsteps,
                         const double *grid, const double *thr, int ngrid,
                         uint64_t mismatches[2]) {
    if (!steps || nsteps < 1 || !mismatches || n < 1 || (ngrid >= 2 && (!grid || !thr)))
        return PB200_EINVAL;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        set_error("no usable CUDA device");
        return PB200_ENODEVICE;
    }
    PB_CUDA(cudaSetDevice(device));
    double *d_steps = nullptr, *d_grid = nullptr, *d_thr = nullptr;
    unsigned long long *d_bad = nullptr;
    PB_CUDA(cudaMalloc((void **)&d_steps, sizeof(double) * nsteps));
    PB_CUDA(cudaMalloc((void **)&d_grid, sizeof(double) * (ngrid > 0 ? ngrid : 1)));
    PB_CUDA(cudaMalloc((void **)&d_thr, sizeof(double) * (ngrid > 0 ? ngrid : 1)));
    PB_CUDA(cudaMalloc((void **)&d_bad, 2 * sizeof(unsigned long long)));
    PB_CUDA(cudaMemcpy(d_steps, steps, sizeof(double) * nsteps, cudaMemcpyHostToDevice));
    int hi0 = 0;
    float inv_step = 0.f;
    if (ngrid >= 2) {
        PB_CUDA(cudaMemcpy(d_grid, grid, sizeof(double) * ngrid, cudaMemcpyHostToDevice));
        PB_CUDA(cudaMemcpy(d_thr, thr, sizeof(double) * ngrid, cudaMemcpyHostToDevice));
        long long bits;
        memcpy(&bits, &grid[0], sizeof(bits));
        hi0 = (int)(bits >> 32);
        inv_step = (float)((ngrid - 1) / log2(grid[ngrid - 1] / grid[0]) / 1048576.0);
    }
    PB_CUDA(cudaMemset(d_bad, 0, 2 * sizeof(unsigned long long)));
    selftest_kernel<<<(unsigned)((n + 255) / 256), 256>>>(n, seed, d_steps, nsteps, d_grid, d_thr,
                                                          ngrid, hi0, inv_step, d_bad);
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess) err = cudaDeviceSynchronize();
    unsigned long long h[2] = {0, 0};
    if (err == cudaSuccess) err = cudaMemcpy(h, d_bad, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(d_steps); cudaFree(d_grid); cudaFree(d_thr); cudaFree(d_bad);
    if (err != cudaSuccess) return cuda_fail(err, "selftest_kernel", __FILE__, __LINE__);
    mismatches[0] = h[0];
    mismatches[1] = h[1];
    return 0;
}

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
