// lbl_kernels.cuh -- launch wrappers of the line-by-line kernels (lbl_kernels.cu).
#pragma once
#include "engine.cuh"

namespace pb200 {

// tp_t[ntp] = {T, RN(1/T)} and tp_z[ntp, niso] = {Z, RN(1/Z)} of every strengths pass.
int launch_strengths(cudaStream_t st, const StaticView &V, int ntp, const double2 *tp_t,
                     const double2 *tp_z, const int *iso_row, int nrows, double *ksum,
                     unsigned long long *kmax, const int *l_group,
                     const unsigned short *l_iso, long long nlines);

// l_group[line] (group id for head lines, ~id for absorbed members) and l_iso[line]; set_lines.
int launch_line_groups(cudaStream_t st, const unsigned int *g_start,
                       const unsigned short *g_iso, long long ngroups, int *l_group,
                       unsigned short *l_iso);

// Minor isotopes merged into the main isotope's dense plane (dense_kernels.cu): the gather
// kernels evaluate only their groups on cells where the two select different Doppler samples.
// Doppler segments per strengths pass: main_bounds[tpass][ndop+1],
// minor_bounds[tpass][nminor][ndop+1].
constexpr int kMaxMinor = 7;
struct MergeView {
    int nminor = 0;
    int iso[kMaxMinor] = {};
    const int *main_bounds = nullptr;
    const int *minor_bounds = nullptr;
};
constexpr int kMaxEntries = 512;   // work-list entries of a gather tile: niso + 2*nminor*ndop

int launch_accumulate(cudaStream_t st, const StaticView &V, int nunits,
                      const UnitParams *units, const IsoUnit *iso_units, const int *iso_row,
                      const double *ksum, const unsigned long long *kmax, int nrows,
                      double ethresh, double cutoff, int mode, double *out, int ksplit,
                      double *partial, int chunked, const MergeView &M = MergeView());

// Interpolate dynamic-grid spectra ktmp[nunits, nrows, dn] onto the output grid (units[].aslot =
// output index of the unit, units[].out_index = its row in ktmp).
int launch_linterp_rows(cudaStream_t st, const StaticView &V, int nunits, const UnitParams *units,
                        const double *ktmp, int dn, int nrows, double *out);

// mode values of launch_accumulate
constexpr int kModeStrided = 0, kModeLinterp = 1, kModeTransposed = 2;
// host only: constant-R unit evaluated on its dynamic grid (chunk kernel) + linterp_rows_kernel
constexpr int kModeDynGrid = 3;

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
                     int ntemp, int nlayers, int nwave, int lay1, int lay2, int per_mol,
                     int overwrite = 0);

}  // namespace pb200
