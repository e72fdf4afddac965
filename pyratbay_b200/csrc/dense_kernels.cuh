// dense_kernels.cuh -- dense-convolution accumulate path for isotopes whose lines fill a large
// fraction of the fine grid (dense_kernels.cu).
#pragma once
#include "engine.cuh"

namespace pb200 {

constexpr int kDenseJ = 16;                              // outputs per half-warp (register tile)
constexpr int kDenseLanes = 16;                          // sub-cell offsets per half-warp / tile row
constexpr int kDenseWarps = 8;                           // warps per CTA
constexpr int kDenseTile = 2 * kDenseJ * kDenseWarps;    // outputs per CTA
constexpr int kDenseSpanMax = 192;   // widest line footprint (outputs) the shared tiles hold
constexpr int kDenseMaxStride = 1024;  // fine samples per output sample (window table size)

// Dynamic shared memory of accumulate_dense_kernel.
size_t dense_smem_bytes();

// bits[cell] = 1 for every group of [gbeg, gend) whose dynamic index at this ofactor is one
// below (nearest fine index)/ofactor ("anomalous" cells, see dense_kernels.cu); *err = 2 if a
// group deviates in any other way.  `U` is any unit with the ofactor in question.
int launch_anomaly_bits(cudaStream_t st, const StaticView &V, long long gbeg, long long gend,
                        const UnitParams &U, unsigned *bits, int *err);

// kd[fine cell] = group strength (0 for groups below the ethresh cut); kd must be zeroed.
int launch_densify(cudaStream_t st, const StaticView &V, long long gbeg, long long gend,
                   const double *ksum_tp, const unsigned long long *kmax_entry, double ethresh,
                   double *kd);

// bounds[j], j = 0..ndop: fine cells [bounds[j], bounds[j+1]) hold the groups whose nearest
// Doppler sample is j (Doppler HWHM per unit wavenumber `adop`).
int launch_segment_bounds(cudaStream_t st, const StaticView &V, long long gbeg, long long gend,
                          double adop, int *bounds);

// out[unit, row, :] += dense convolution of kd with the windowed Voigt profiles of every unit.
int launch_accumulate_dense(cudaStream_t st, const StaticView &V, int nunits,
                            const UnitParams *units, const IsoUnit *iso_units, int iso, int row,
                            int nrows, const double *kd, const int *bounds, const unsigned *abits,
                            long long abits_words, double cutoff, double *out, int *err);

}  // namespace pb200
