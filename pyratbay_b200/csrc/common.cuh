// common.cuh -- shared helpers for the pb200 line-by-line engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace pb200 {

// Physical / numerical constants with the values of the reference's C layer
// (src_c/include/constants.h:5-21).  These deliberately differ from the CODATA values
// the reference's Python layer uses for densities.
constexpr double kPi = 3.141592653589793;
constexpr double kSqrtLn2 = 0.83255461115769775635;
constexpr double kTwoOverSqrtPi = 1.12837916709551257389;
constexpr double kSqrtLn2OverPi = 0.46971863934982566689;
constexpr double kLightSpeed = 2.99792458e10;
constexpr double kBoltzmann = 1.380658e-16;
constexpr double kAmu = 1.66053886e-24;
constexpr double kPlanck = 6.6260755e-27;
constexpr double kECharge = 4.8032068e-10;
constexpr double kEMass = 9.1093897e-28;
// SIGCTE, EXPCTE (constants.h:20-21), folded left to right in IEEE double.
constexpr double kSigCte = kPi * kECharge * kECharge / kLightSpeed / kLightSpeed / kEMass;
constexpr double kExpCte = kPlanck * kLightSpeed / kBoltzmann;

void set_error(const std::string &msg);
int cuda_fail(cudaError_t err, const char *what, const char *file, int line);

#define PB_CUDA(call)                                                        \
    do {                                                                     \
        cudaError_t _e = (call);                                             \
        if (_e != cudaSuccess) return pb200::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

// IEEE round-to-nearest primitives that the compiler may not contract into FMAs.
// Every value that feeds a discrete decision of the reference (an index, a comparison)
// is computed with these so that it rounds exactly like the reference's C expression.
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dadd_rn(a, -b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }

// RN(a/b) from y = RN(1/b) with fused multiply-adds only (no reciprocal seed, no special-case
// path): q0 = RN(a*y) is within 2 ulp, one residual correction makes it faithful and the
// second one (Markstein's theorem) correctly rounded, i.e. bit-identical to the IEEE
// division a/b for normal, finite operands.  6 fp64 instructions instead of ~20.
__device__ __forceinline__ double quotient_rn(double a, double b, double y) {
    const double q0 = __dmul_rn(a, y);
    const double q1 = __fma_rn(__fma_rn(-q0, b, a), y, q0);
    return __fma_rn(__fma_rn(-q1, b, a), y, q1);
}

// Nearest index from the threshold table of a strictly increasing grid (StaticView::dop_thr),
// identical to nearest_index(grid, n, v); the bracket guess is the one of nearest_index_log.
__device__ __forceinline__ int nearest_index_thr(const double *thr, int n, double v, int hi0,
                                                 float inv_step) {
    int lo = (int)((float)(__double2hiint(v) - hi0) * inv_step);
    lo = max(0, min(lo, n - 1));
    while (lo < n - 1 && !(v < thr[lo + 1])) lo++;
    while (lo > 0 && v < thr[lo]) lo--;
    return lo;
}

// Nearest grid index with the semantics of pyramidsearch (src_c/include/utils.h:44-72)
// for a monotonically increasing query sequence: clamp outside the grid, otherwise the
// closer of the two bracketing samples, ties to the lower index.
__device__ __forceinline__ int nearest_index(const double *grid, int n, double v) {
    if (v < grid[0]) return 0;
    if (grid[n - 1] < v) return n - 1;
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        int mid = (hi + lo) >> 1;
        if (grid[mid] < v) lo = mid; else hi = mid;
    }
    return (fabs(dsub(grid[hi], v)) < fabs(dsub(grid[lo], v))) ? hi : lo;
}

// Same result as nearest_index for a (roughly) log-spaced grid, in O(1): the bracket is
// guessed from the IEEE exponent/mantissa bits (piecewise-linear log2, error < 0.09) and then
// corrected exactly, so the returned index is identical for ANY increasing grid.
//   hi0      = high 32 bits of grid[0]
//   inv_step = (n-1) / log2(grid[n-1]/grid[0]) / 2^20
__device__ __forceinline__ int nearest_index_log(const double *grid, int n, double v, int hi0,
                                                 float inv_step) {
    int lo = (int)((float)(__double2hiint(v) - hi0) * inv_step);
    lo = max(0, min(lo, n - 2));
    while (lo < n - 2 && grid[lo + 1] < v) lo++;
    while (lo > 0 && !(grid[lo] < v)) lo--;
    const int hi = lo + 1;
    return (fabs(dsub(grid[hi], v)) < fabs(dsub(grid[lo], v))) ? hi : lo;
}

}  // namespace pb200
