// lbl_kernels.cuh -- launch wrappers of the line-by-line kernels (lbl_kernels.cu).
#pragma once
#include "engine.cuh"

namespace pb200 {

int launch_strengths(cudaStream_t st, const StaticView &V, int ntp, const double *tp_temp,
                     const double *tp_isoz, const int *iso_row, int nrows, double *ksum,
                     unsigned long long *kmax, const int *multi, int nmulti);

// List the groups with more than one member line (set_lines; `count` must be zeroed).
int launch_multi_list(cudaStream_t st, const unsigned int *g_start, long long ngroups,
                      int *multi, unsigned int *count);

int launch_accumulate(cudaStream_t st, const StaticView &V, int nunits,
                      const UnitParams *units, const IsoUnit *iso_units, const int *iso_row,
                      const double *ksum, const unsigned long long *kmax, int nrows,
                      double ethresh, double cutoff, int mode, double *out, int ksplit,
                      double *partial, int chunked);

// mode values of launch_accumulate
constexpr int kModeStrided = 0, kModeLinterp = 1, kModeTransposed = 2;

// Build the output-stride copy of the Voigt table (one-time; synchronises `st`).
int launch_transpose(cudaStream_t st, int nprof, const long long *src, const long long *dst,
                     const int *nbin, const int *rowlen, long long total, int stride,
                     const double *profile, double *tprofile);

int launch_counters(cudaStream_t st, const StaticView &V, int nunits, const UnitParams *units,
                    const IsoUnit *iso_units, const int *iso_row, const double *ksum,
                    const unsigned long long *kmax, int nrows, double ethresh, double cutoff,
                    int linterp, unsigned long long *counters);

int launch_interp_ec(cudaStream_t st, double *ext, const double *table, const int *tlo,
                     const double *w_lo, const double *w_hi, const double *density, int nspec,
                     int ntemp, int nlayers, int nwave, int lay1, int lay2, int per_mol);

}  // namespace pb200
