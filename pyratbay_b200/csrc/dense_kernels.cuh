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

constexpr int kMaxMerge = 8;    // isotopes sharing one dense plane (main + merged minors)

// The isotopes one dense pass evaluates.  [0] is the main isotope (its own dense array kd[0],
// its Doppler segments `bounds`); [1..n) are minor isotopes of the same species merged into
// `kd_all` = kd[0] + sum of kd[i], where kd[i] holds only the groups that select the main
// isotope's Doppler sample on their cell (merge_minor_kernel).  Units whose isotopes do not
// all select the same Lorentz sample read kd[0] alone (UnitParams::aslot bit 30 clear).
struct DenseSet {
    int n;
    int iso[kMaxMerge];
    const double *kd[kMaxMerge];
    const unsigned *abits[kMaxMerge];   // anomaly bitmasks of each isotope, slot-major
    const double *kd_all;
    const int *bounds;
};
constexpr int kMergedUnitBit = 1 << 30;   // in UnitParams::aslot: the unit reads kd_all

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

// kd[cell] = strength of the minor isotope's groups of [gbeg, gend) that the dense plane takes
// (not below the ethresh cut, selecting the main isotope's Doppler sample on their cell), and
// kd_all[cell] += the same.  kd must be zeroed, kd_all hold the main plane.
int launch_merge_minor(cudaStream_t st, const StaticView &V, long long gbeg, long long gend,
                       const double *ksum_tp, const unsigned long long *kmax_entry,
                       double ethresh, double adop, const int *main_bounds, double *kd,
                       double *kd_all);

// out[unit, row, :] += dense convolution of the set's strengths with the windowed Voigt profiles
// of every unit (`row`, ilor, dens: the main isotope's).
int launch_accumulate_dense(cudaStream_t st, const StaticView &V, int nunits,
                            const UnitParams *units, const IsoUnit *iso_units,
                            const DenseSet &set, int row, int nrows, long long abits_words,
                            double cutoff, double *out, int *err);

}  // namespace pb200
